# zsc-b200 build: host C with gcc, kernels with nvcc for sm_100a, one in-tree shared library.
#   make            -> zsc_b200/libzsc_b200.so  (the product: zsc_pub.h surface + zscgpu.h C-ABI)
#   make testlibs   -> tests/libzsc_cpuharness.so (product decode/huffman logic compiled for the host,
#                      used only by -m "not gpu" unit tests) and tools/libzscgen.so (synthetic data)
#   make oracle     -> oracle/libzsc_oracle.so and, where /root/reference exists, oracle/_ref/libzsc_ref.so
NVCC ?= nvcc
CC ?= gcc
CXX ?= g++
ARCH = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Izsc_b200/csrc/cuda $(EXTRA_NVFLAGS)
CFLAGS = -std=gnu11 -O2 -fPIC -Wall -Wextra -Iinclude

CU = deflate_lz deflate_chain deflate_huff checksum inflate engine
CU_OBJS = $(foreach f,$(CU),build/$(f).o)
LIB = zsc_b200/libzsc_b200.so

all: $(LIB) testlibs

build:
	mkdir -p build

build/%.o: zsc_b200/csrc/cuda/%.cu zsc_b200/csrc/cuda/inflate_group.inc zsc_b200/csrc/cuda/inflate_spec.inc zsc_b200/csrc/cuda/inflate_spec.h zsc_b200/csrc/cuda/common.cuh zsc_b200/csrc/cuda/lz_common.cuh zsc_b200/csrc/cuda/huff_build.h zsc_b200/csrc/cuda/inflate_core.h include/zscgpu.h | build
	$(NVCC) $(NVFLAGS) -c $< -o $@

build/zsc_%.o: zsc_b200/csrc/host/zsc_%.c include/zsc/zsc_pub.h include/zsc/zlib.h include/zscgpu.h | build
	$(CC) $(CFLAGS) -c $< -o $@

$(LIB): $(CU_OBJS) build/zsc_api.o build/zsc_stream.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -Xlinker -Bsymbolic -cudart static -lpthread

testlibs: tests/libzsc_cpuharness.so tests/libzsc_cpuharness_n.so tests/libzsc_cpuharness_w.so tools/libzscgen.so

tests/libzsc_cpuharness.so: tests/cpu_harness.cpp zsc_b200/csrc/cuda/huff_build.h zsc_b200/csrc/cuda/inflate_core.h zsc_b200/csrc/cuda/inflate_spec.h
	$(CXX) -O2 -fPIC -shared -std=c++17 -Izsc_b200/csrc/cuda -o $@ tests/cpu_harness.cpp

# the same harness with the decode tables of narrow batches (10 / 8 root bits, inflate.cu namespace zn)
tests/libzsc_cpuharness_n.so: tests/cpu_harness.cpp zsc_b200/csrc/cuda/huff_build.h zsc_b200/csrc/cuda/inflate_core.h zsc_b200/csrc/cuda/inflate_spec.h
	$(CXX) -O2 -fPIC -shared -std=c++17 -DZI_LBITS=10 -DZI_DBITS=8 -DZI_POOL=256 -Izsc_b200/csrc/cuda -o $@ tests/cpu_harness.cpp

# ... and with a second geometry of the speculative decoder (regions of 320 bits; the product uses 512 everywhere, the lane
# functions must not depend on it)
tests/libzsc_cpuharness_w.so: tests/cpu_harness.cpp zsc_b200/csrc/cuda/huff_build.h zsc_b200/csrc/cuda/inflate_core.h zsc_b200/csrc/cuda/inflate_spec.h
	$(CXX) -O2 -fPIC -shared -std=c++17 -DZI_LBITS=10 -DZI_DBITS=8 -DZI_POOL=256 -DZP_R=320u -DZP_CAP=80u -Izsc_b200/csrc/cuda -o $@ tests/cpu_harness.cpp

tools/libzscgen.so: tools/datagen.c
	$(CC) -O2 -fPIC -shared -o $@ tools/datagen.c -lm -lpthread

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf build $(LIB) tests/libzsc_cpuharness.so tests/libzsc_cpuharness_n.so tests/libzsc_cpuharness_w.so tools/libzscgen.so
	$(MAKE) -C oracle clean

.PHONY: all testlibs oracle clean
