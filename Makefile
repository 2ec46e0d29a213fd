# spmv-b200 build.  Everything is built in-tree so the binaries travel with the
# repo snapshot to the GPU box:
#   spmv_scpa_b200/lib/libspmv_b200.so   CUDA kernels + C ABI (sm_100a only)
#   spmv_scpa_b200/lib/libspmv_host.so   C host layer (loader, packer, CSV, generators)
#   bin/spmv                             the reference-compatible CLI
#   bin/kbench                           kernel sweep tool (roofline tables)
#   bin/dist_check                       multi-GPU iterated SpMV through the C ABI only (self-check + timing)
#   oracle/                              test-only checker (see oracle/Makefile)
#
# /usr/bin/gcc explicitly: the image's $CC (/opt/gcc) cannot link libgomp.

CC      := /usr/bin/gcc
CXX     := /usr/bin/g++
NVCC    := /usr/local/cuda/bin/nvcc
ARCH    := -gencode arch=compute_100a,code=sm_100a

PKG     := spmv_scpa_b200
LIBDIR  := $(PKG)/lib
INC     := -Iinclude

CFLAGS  := -std=gnu99 -O3 -march=x86-64-v3 -fopenmp -fPIC -Wall -Wextra -Wno-unused-parameter $(INC)
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -ccbin $(CXX) -Xcompiler -fPIC,-Wall,-fopenmp \
           -Xptxas -v --expt-relaxed-constexpr $(INC) -I$(PKG)/csrc

HOST_SRC := $(addprefix $(PKG)/host/,mmio.c support.c logger.c csr.c hll.c gen.c)
CUDA_SRC := $(PKG)/csrc/spmv_b200.cu $(PKG)/csrc/entry.cu $(PKG)/csrc/dist.cu
CUDA_OBJ := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(CUDA_SRC))
CUDA_DEP := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard include/*.h)

all: $(LIBDIR)/libspmv_b200.so $(LIBDIR)/libspmv_host.so bin/spmv bin/kbench bin/dist_check oracle

build/%.o: $(PKG)/csrc/%.cu $(CUDA_DEP)
	@mkdir -p build $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)
	@grep -E "error|warning" build/$*.ptxas.log | grep -v "Wno-" || true

$(LIBDIR)/libspmv_b200.so: $(CUDA_OBJ)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(CUDA_OBJ) -Xcompiler -fopenmp -ldl
	@cat build/*.ptxas.log > $(LIBDIR)/ptxas.log

$(LIBDIR)/libspmv_host.so: $(HOST_SRC) $(wildcard include/*.h) $(LIBDIR)/libspmv_b200.so
	$(CC) $(CFLAGS) -shared -o $@ $(HOST_SRC) -L$(LIBDIR) -lspmv_b200 -Wl,-rpath,'$$ORIGIN' -lm

bin/spmv: $(PKG)/host/main.c $(LIBDIR)/libspmv_host.so
	@mkdir -p bin
	$(CC) $(CFLAGS) -o $@ $< -L$(LIBDIR) -lspmv_host -lspmv_b200 -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)' -lm

bin/kbench: tools/kbench.c $(LIBDIR)/libspmv_host.so
	@mkdir -p bin
	$(CC) $(CFLAGS) -o $@ $< -L$(LIBDIR) -lspmv_host -lspmv_b200 -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)' -lm

bin/dist_check: tools/dist_check.c $(LIBDIR)/libspmv_host.so
	@mkdir -p bin
	$(CC) $(CFLAGS) -o $@ $< -L$(LIBDIR) -lspmv_host -lspmv_b200 -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)' -lm

oracle: $(LIBDIR)/libspmv_b200.so
	$(MAKE) -C oracle

sass: $(LIBDIR)/libspmv_b200.so
	cuobjdump -sass $< > profiles/libspmv_b200.sass

clean:
	rm -rf $(LIBDIR)/*.so $(LIBDIR)/ptxas.log bin/spmv bin/kbench bin/dist_check build
	$(MAKE) -C oracle clean

.PHONY: all oracle clean sass
