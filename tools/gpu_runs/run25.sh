#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_direct_eigh.py -x -q -m gpu -s > gpurun_out/r2_25_direct.log 2>&1
timeout 300 python tests/eigh_batch_probe.py > gpurun_out/r2_25_batch.log 2>&1
timeout 600 python -m pytest tests/test_gpu_bench_sizes.py tests/test_gpu_kernels.py -x -q -m gpu -s > gpurun_out/r2_25_sizes.log 2>&1
grep -n "stedc n=" gpurun_out/r2_25_direct.log; tail -n 3 gpurun_out/r2_25_direct.log; tail -n 4 gpurun_out/r2_25_batch.log; grep "n=4608" gpurun_out/r2_25_sizes.log; tail -n 2 gpurun_out/r2_25_sizes.log
