# Builds librlmd_b200.so (sm_100a) in-tree, and the C oracle.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v -fmad=false
SRCS := $(wildcard rlmd_b200/csrc/*.cu)
OBJS := $(patsubst rlmd_b200/csrc/%.cu,build/%.o,$(SRCS))
LIB := rlmd_b200/librlmd_b200.so

all: $(LIB)

build/%.o: rlmd_b200/csrc/%.cu $(wildcard rlmd_b200/csrc/*.cuh) include/rlmd_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart_static -ldl -lrt -lpthread

clean:
	rm -rf build $(LIB)
.PHONY: all clean
