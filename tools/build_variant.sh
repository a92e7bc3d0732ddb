#!/bin/bash
# tools/build_variant.sh NAME "NVCC FLAGS" — a tuning build of the engine under build/var/NAME/ (travels to the GPU
# box; select it with ZSC_B200_LIB=build/var/NAME/libzsc_b200.so).
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
d=build/var/$name
mkdir -p $d
for f in deflate_lz deflate_chain deflate_huff checksum inflate engine; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Izsc_b200/csrc/cuda $flags -c zsc_b200/csrc/cuda/$f.cu -o $d/$f.o &
done
wait
gcc -std=gnu11 -O2 -fPIC -Iinclude -c zsc_b200/csrc/host/zsc_api.c -o $d/zsc_api.o
gcc -std=gnu11 -O2 -fPIC -Iinclude -c zsc_b200/csrc/host/zsc_stream.c -o $d/zsc_stream.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/libzsc_b200.so $d/*.o -Xlinker -Bsymbolic -cudart static -lpthread
echo built $d/libzsc_b200.so
