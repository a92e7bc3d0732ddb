mkdir -p gpurun_out
{
python tools/prof_chain.py 592 6 telemetry
python tools/prof_chain.py 592 6 mixed
python tools/prof_chain.py 592 9 mixed
python tools/prof_chain.py 592 9 telemetry
timeout 900 python -m pytest tests/test_gpu.py -x -q -m gpu -k "deflate or ratio or matrix or level or model" 2>&1 | tail -3
} > gpurun_out/exp16.log 2>&1
