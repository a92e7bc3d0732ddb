mkdir -p gpurun_out
{
python tools/prof_chain.py 592 6 telemetry
python tools/prof_chain.py 592 6 mixed
python tools/prof_chain.py 592 9 mixed
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"zs_lz_kernel" --launch-skip 1 -c 1 -f -o gpurun_out/chain_l6 python tools/prof_chain.py 296 6 telemetry 2>&1 | tail -2
} > gpurun_out/exp15.log 2>&1
