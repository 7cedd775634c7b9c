# Builds the C-ABI shared library (sm_100a only) and the C oracle.  No torch dependency.
PKG      := lk-s-2022-estimacija-pokreta_b200
CSRC     := $(PKG)/csrc
NVCC     ?= /usr/local/cuda/bin/nvcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr -Iinclude
SOURCES  := $(wildcard $(CSRC)/*.cu)
OBJECTS  := $(patsubst $(CSRC)/%.cu,$(CSRC)/build/%.o,$(SOURCES))
LIB      := $(PKG)/libflowb200.so

all: $(LIB) oracle

$(CSRC)/build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/flowb200.h
	@mkdir -p $(CSRC)/build
	$(NVCC) $(NVFLAGS) $(EXTRA) -c $< -o $@

$(LIB): $(OBJECTS)
	$(NVCC) $(ARCH) -shared -o $@ $^

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(CSRC)/build $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
