#!/bin/bash
# Builds the C-ABI shared library in-tree (the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=dp_gp_lvm_b200/libdpgp.so
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr \
  -Xcompiler -fPIC -shared ${DPGP_NVCC_EXTRA} -o $OUT dp_gp_lvm_b200/csrc/dpgp_api.cu
echo "built $OUT"
