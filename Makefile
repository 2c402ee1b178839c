# Build of libsce.so (the C-ABI engine), the standalone GEMM self-test and the oracle's C pieces.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
CSRC := sparse_coding_b200/csrc
LIB := sparse_coding_b200/libsce.so

all: $(LIB)

$(LIB): $(CSRC)/sce_engine.cu $(CSRC)/*.cuh $(CSRC)/sce_tmap.h include/sce.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/sce_engine.cu

selftest: build/gemm_selftest
build/gemm_selftest: tests/csrc/gemm_selftest.cu $(CSRC)/*.cuh $(CSRC)/sce_tmap.h
	mkdir -p build
	$(NVCC) $(NVFLAGS) -o $@ tests/csrc/gemm_selftest.cu

clean:
	rm -f $(LIB) build/gemm_selftest
