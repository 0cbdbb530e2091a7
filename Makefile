# Builds the sm_100a C-ABI library in-tree (the .so travels to the GPU box with the gpurun snapshot).
NVCC ?= nvcc
PKG := gaussian_process_transportation_b200
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC
SRC := $(PKG)/csrc/gptb200.cu
HDR := $(wildcard $(PKG)/csrc/*.cuh) include/gptb200.h
LIB := $(PKG)/lib/libgptb200.so

all: $(LIB)

$(LIB): $(SRC) $(HDR)
	mkdir -p $(PKG)/lib
	$(NVCC) $(NVFLAGS) -o $@ $(SRC)

clean:
	rm -f $(LIB)
.PHONY: all clean
