# Top-level build: the product library (CUDA, sm_100a only), the host helper library (plain C),
# the stand-alone C driver, and the test-only oracle libraries.
NVCC    ?= /usr/local/cuda/bin/nvcc
HOSTCC  := $(shell [ -x /usr/bin/gcc ] && echo /usr/bin/gcc || echo gcc)
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude
CSRC    := nbodysim_b200/csrc
CU      := $(CSRC)/nbody_gpu.cu $(CSRC)/force_f32.cu $(CSRC)/force_f64.cu $(CSRC)/layout.cu $(CSRC)/barnes_hut.cu $(CSRC)/collide.cu
HDR     := $(CSRC)/common.cuh $(CSRC)/kernels.h $(CSRC)/force_f32_fast.cuh $(CSRC)/radix_sort.cuh $(CSRC)/collide.cuh $(CSRC)/nccl_dyn.h include/nbody_gpu.h include/nbody_body.h
OBJ     := $(patsubst $(CSRC)/%.cu,build/obj/%.o,$(CU))

all: lib host oracle tools

lib: nbodysim_b200/libnbody_gpu.so
build/obj/%.o: $(CSRC)/%.cu $(HDR)
	@mkdir -p build/obj
	$(NVCC) $(NVFLAGS) -c $< -o $@
nbodysim_b200/libnbody_gpu.so: $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -ldl

host: nbodysim_b200/libnbody_host.so host/_build/nbody_run host/_build/nbody_viewer_feed
nbodysim_b200/libnbody_host.so: host/nbody_ic.c include/nbody_host.h include/nbody_body.h
	$(HOSTCC) -std=c11 -O2 -ffp-contract=off -Wall -Wextra -fPIC -shared -Iinclude -o $@ host/nbody_ic.c -lm
host/_build/nbody_run: host/nbody_main.c host/nbody_ic.c include/nbody_host.h include/nbody_gpu.h nbodysim_b200/libnbody_gpu.so
	@mkdir -p host/_build
	$(HOSTCC) -std=c11 -O2 -ffp-contract=off -Wall -Wextra -Iinclude -o $@ host/nbody_main.c host/nbody_ic.c \
	    -Lnbodysim_b200 -lnbody_gpu -Wl,-rpath,'$$ORIGIN/../../nbodysim_b200' -lm

host/_build/nbody_viewer_feed: host/nbody_viewer_feed.c host/nbody_ic.c include/nbody_host.h include/nbody_gpu.h nbodysim_b200/libnbody_gpu.so
	@mkdir -p host/_build
	$(HOSTCC) -std=c11 -O2 -ffp-contract=off -Wall -Wextra -pthread -Iinclude -o $@ host/nbody_viewer_feed.c host/nbody_ic.c \
	    -Lnbodysim_b200 -lnbody_gpu -Wl,-rpath,'$$ORIGIN/../../nbodysim_b200' -lm

oracle:
	$(MAKE) -C oracle all

# tuning / microbenchmark harnesses (not part of the product path)
tools: build/kbench build/ubench build/sort_check build/kbench_sym
build/kbench_sym: tools/kbench_sym.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -Iinclude -o $@ $<
build/sort_check: tools/sort_check.cu $(CSRC)/radix_sort.cuh
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -std=c++17 -o $@ $<
build/kbench: tools/kbench.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -Iinclude -o $@ $<
build/ubench: tools/ubench.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -std=c++17 -o $@ $<

clean:
	rm -rf build/obj nbodysim_b200/*.so host/_build
	$(MAKE) -C oracle clean
.PHONY: all lib host oracle tools clean
