#!/bin/bash
# experiment 1: checksum kernel, inflate variants
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python tools/prof_sums.py 4096
timeout 600 python -m pytest tests/test_gpu.py -x -q -m gpu -k "checksum or inflate or uncompress or kat or corrupt" 2>&1 | tail -5
timeout 400 python tools/prof_inflate.py 512 128 0,16,8,4
for v in l9 l9r80 l10r80; do
  ZSC_B200_LIB=build/var/$v/libzsc_b200.so timeout 300 python tools/prof_inflate.py 512 128 16,8,4
done
timeout 200 python tools/prof_inflate.py 512 8 32,16
ZSC_B200_LIB=build/var/l10r80/libzsc_b200.so timeout 200 python tools/prof_inflate.py 512 8 32
} > gpurun_out/exp1.log 2>&1
