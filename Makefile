# Builds the sm_100a C-ABI library in-tree (the .so travels to the GPU box with the gpurun snapshot).
NVCC ?= nvcc
PKG := gaussian_process_transportation_b200
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC --split-compile 0
SRC := $(PKG)/csrc/gptb200.cu
HDR := $(wildcard $(PKG)/csrc/*.cuh) include/gptb200.h
LIB := $(PKG)/lib/libgptb200.so
# developer build with the what-if switches of the product kernel compiled in (tools/whatif.py; never loaded by the package itself)
LIB_WHATIF := $(PKG)/lib/libgptb200_whatif.so

all: $(LIB)

$(LIB): $(SRC) $(HDR)
	mkdir -p $(PKG)/lib
	$(NVCC) $(NVFLAGS) -o $@ $(SRC)

whatif: $(LIB_WHATIF)
$(LIB_WHATIF): $(SRC) $(HDR)
	mkdir -p $(PKG)/lib
	$(NVCC) $(NVFLAGS) -DGPTB_OZ_WHATIF -o $@ $(SRC)

clean:
	rm -f $(LIB) $(LIB_WHATIF)
.PHONY: all clean whatif
