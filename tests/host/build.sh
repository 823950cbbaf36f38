#!/bin/bash
# Builds the standalone (no Python) GPU harnesses of the experimental kernels into tests/host/bin/
# (git-ignored; they travel to the GPU box with the snapshot).  Needs libkfac_b200.so (build()).
#   bash tests/host/build.sh && gpurun -- 'export LD_LIBRARY_PATH=kfac-pytorch_b200/csrc; tests/host/bin/jsys_gpu 64 72; tests/host/bin/ema_gpu'
set -e
cd "$(dirname "$0")/../.."
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
mkdir -p tests/host/bin
for t in jacobi_systolic_gpu:jsys_gpu; do
  src=${t%%:*}; out=${t##*:}
  $NVCC -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include tests/host/$src.cu \
        -L kfac-pytorch_b200/csrc -lkfac_b200 -o tests/host/bin/$out
done
echo "built tests/host/bin/{jsys_gpu,ema_gpu}"
