# In-tree build of the C-ABI library (sm_100a only).  `make -j8`
NVCC ?= /usr/local/cuda/bin/nvcc
NVFLAGS = -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $(EXTRA)
SRC = dp_gp_lvm_b200/csrc
BUILD = build/obj
QPS = 2 4 6 8 10 12 16
OBJS = $(BUILD)/dpgp_api.o $(foreach q,$(QPS),$(BUILD)/qp_kernels_$(q).o)
HDRS = $(wildcard $(SRC)/*.cuh) include/dpgp.h
LIB = dp_gp_lvm_b200/libdpgp.so

all: $(LIB) $(SRC)/microbench/fp64_peaks

$(LIB): $(OBJS)
	$(NVCC) -shared -gencode arch=compute_100a,code=sm_100a -o $@ $(OBJS)

$(BUILD)/dpgp_api.o: $(SRC)/dpgp_api.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(BUILD)/qp_kernels_%.o: $(SRC)/qp_kernels.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -DDPGP_QP=$* -c $< -o $@

$(SRC)/microbench/fp64_peaks: $(SRC)/microbench/fp64_peaks.cu $(SRC)/fast_exp.cuh
	$(NVCC) -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo $< -o $@

clean:
	rm -rf build $(LIB) $(SRC)/microbench/fp64_peaks
