cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_gpu_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02s_gpu_tests.log
tail -4 gpurun_out/r02s_gpu_tests.log
python bench.py > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err; echo bench rc $?
python bench.py --impl reference > gpurun_out/r02s_bench_reference.json 2>/dev/null; echo ref rc $?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo smoke rc $?; tail -2 gpurun_out/r02s_smoke.log
