#!/bin/bash
mkdir -p gpurun_out
{
timeout 400 python tools/prof_inflate.py 512 128 16,8,4
ZSC_B200_LIB=build/var/ownb/libzsc_b200.so timeout 300 python tools/prof_inflate.py 512 128 16,8,4
ZSC_B200_LIB=build/var/xglob/libzsc_b200.so timeout 300 python tools/prof_inflate.py 512 128 16,8
timeout 200 python tools/prof_inflate.py 512 8 32,16
ZSC_B200_LIB=build/var/ownb/libzsc_b200.so timeout 200 python tools/prof_inflate.py 512 8 32
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
ZSC_B200_INFLATE_G=16 timeout 600 ncu --set full --import-source on --clock-control none -k regex:zs_inflate_group_kernel -c 1 -f -o gpurun_out/inf_g16 python tools/prof_inflate.py 512 32 16 2>&1 | tail -3
ZSC_B200_INFLATE_G=32 timeout 600 ncu --set full --import-source on --clock-control none -k regex:zs_inflate_group_kernel -c 1 -f -o gpurun_out/inf_g32 python tools/prof_inflate.py 512 8 32 2>&1 | tail -3
} > gpurun_out/exp2.log 2>&1
