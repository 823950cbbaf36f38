#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_direct_eigh.py tests/test_gpu_bench_sizes.py tests/test_gpu_kernels.py -x -q -m gpu -s > gpurun_out/r2_34_direct.log 2>&1
timeout 300 python tests/eigh_batch_probe.py > gpurun_out/r2_34_batch.log 2>&1
grep "n=4608\|n=2049\|n=2304 m" gpurun_out/r2_34_direct.log | head; tail -n 3 gpurun_out/r2_34_direct.log; tail -n 4 gpurun_out/r2_34_batch.log
