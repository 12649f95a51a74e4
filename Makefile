# Builds libmosdepth_b200.so (sm_100a only), the host CLI stand-in, the BAM generator and the test oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v --fmad=false
LIB := mosdepth_b200/libmosdepth_b200.so
CSRC := $(wildcard mosdepth_b200/csrc/*.cu mosdepth_b200/csrc/*.cuh) include/mosdepth_b200.h

all: $(LIB) mosdepth_b200/bin/mosdepth-b200 tools/gen_bam tools/inflate_model tools/bedgz_model tools/h2d_probe oracle

$(LIB): $(CSRC)
	$(NVCC) $(NVFLAGS) -shared -o $@ mosdepth_b200/csrc/msd_api.cu 2> build_ptxas.log || (cat build_ptxas.log; exit 1)
	@grep -E "error|warning" build_ptxas.log | grep -v "ptxas info" || true

mosdepth_b200/bin/mosdepth-b200: mosdepth_b200/host/main.cpp $(wildcard mosdepth_b200/host/*.hpp) include/mosdepth_b200.h $(LIB)
	mkdir -p mosdepth_b200/bin
	$(CXX) -O2 -g -std=c++17 -Wall -pthread -o $@ mosdepth_b200/host/main.cpp -Iinclude -Lmosdepth_b200 -lmosdepth_b200 -lz -Wl,-rpath,'$$ORIGIN/..'

tools/gen_bam: tools/gen_bam.cpp
	$(CXX) -O3 -std=c++17 -Wall -pthread -o $@ $< -lz

# host replay of the inflate kernels' arithmetic against zlib (CPU test of csrc/inflate_core.cuh)
tools/inflate_model: tools/inflate_model.cpp mosdepth_b200/csrc/inflate_core.cuh mosdepth_b200/csrc/common.cuh
	$(CXX) -O2 -std=c++17 -Wall -I/usr/local/cuda/include -o $@ $< -lz

# host replay of K10's per-line arithmetic against gzip (CPU test of csrc/bedgz_core.cuh)
tools/bedgz_model: tools/bedgz_model.cpp mosdepth_b200/csrc/bedgz_core.cuh mosdepth_b200/csrc/common.cuh
	$(CXX) -O2 -std=c++17 -Wall -I/usr/local/cuda/include -o $@ $< -lz

# host->device copy ceiling of the box at 1..N GPUs (pinned / registered mmap / bounce), profiles/*h2d_probe*
tools/h2d_probe: tools/h2d_probe.cu
	$(NVCC) $(ARCH) -O2 -std=c++17 -o $@ $< -Xcompiler -pthread

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -f $(LIB) mosdepth_b200/bin/mosdepth-b200 tools/gen_bam tools/inflate_model tools/bedgz_model tools/h2d_probe build_ptxas.log
	$(MAKE) -C oracle clean
.PHONY: all oracle clean
