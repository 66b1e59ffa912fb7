# Builds the C-ABI library without Python (flope_b200/build.py runs the same command).
NVCC ?= nvcc
LIB  := flope_b200/libflope_b200.so
SRC  := flope_b200/csrc/engine.cu
DEPS := $(wildcard flope_b200/csrc/*.cuh) include/flope_b200.h

all: $(LIB)

$(LIB): $(SRC) $(DEPS)
	$(NVCC) -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -o $@ $(SRC)

c_abi_smoke: $(LIB) tests/c_abi/c_abi_smoke.c
	gcc -std=c99 -O1 -Wall -I include -I $(CUDA_HOME)/include tests/c_abi/c_abi_smoke.c -o $@ \
	    -L flope_b200 -lflope_b200 -L $(CUDA_HOME)/lib64 -lcudart -Wl,-rpath,$(CURDIR)/flope_b200

CUDA_HOME ?= /usr/local/cuda
clean:
	rm -f $(LIB) c_abi_smoke
.PHONY: all clean
