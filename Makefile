# In-tree build of the C-ABI library (sm_100a only).  `make -j8`
NVCC ?= /usr/local/cuda/bin/nvcc
# EXPERIMENTAL=1 adds the non-default kernel variants under csrc/experimental/ (exp_variant 2/3/5/6, bwd_variant 2/3/4/5,
# chain_variant 2): the lab notebook of round 1, kept buildable, not part of the product library.
ifeq ($(EXPERIMENTAL),1)
XFLAGS = -DDPGP_EXPERIMENTAL
endif
NVFLAGS = -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --cudart shared -Xcompiler -fPIC $(XFLAGS) $(EXTRA)
SRC = dp_gp_lvm_b200/csrc
BUILD = build/obj
QPS = 2 4 6 8 10 12 16 20 24 28 32
OBJS = $(BUILD)/dpgp_api.o $(foreach q,$(QPS),$(BUILD)/qp_kernels_$(q).o)
HDRS = $(wildcard $(SRC)/*.cuh) $(wildcard $(SRC)/experimental/*.cuh) include/dpgp.h
LIB = dp_gp_lvm_b200/libdpgp.so

all: $(LIB) $(SRC)/microbench/fp64_peaks $(SRC)/microbench/umma_i8_probe

$(LIB): $(OBJS)
	$(NVCC) -shared --cudart shared -gencode arch=compute_100a,code=sm_100a -o $@ $(OBJS)

$(BUILD)/dpgp_api.o: $(SRC)/dpgp_api.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(BUILD)/qp_kernels_%.o: $(SRC)/qp_kernels.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -DDPGP_QP=$* -c $< -o $@

$(SRC)/microbench/fp64_peaks: $(SRC)/microbench/fp64_peaks.cu $(SRC)/fast_exp.cuh
	$(NVCC) -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo $< -o $@

$(SRC)/microbench/umma_i8_probe: $(SRC)/microbench/umma_i8_probe.cu $(SRC)/umma.cuh $(SRC)/common.cuh $(SRC)/fast_exp.cuh
	$(NVCC) -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo $< -o $@

clean:
	rm -rf build $(LIB) $(SRC)/microbench/fp64_peaks $(SRC)/microbench/umma_i8_probe
