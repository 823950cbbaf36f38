#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_direct_eigh.py -x -q -m gpu -s > gpurun_out/r2_03_direct.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k "eigh" > gpurun_out/r2_03_eigh.log 2>&1
timeout 300 python -m pytest tests/test_gpu_bench_sizes.py -x -q -m gpu -s -k "eigh_bench" > gpurun_out/r2_03_sizes.log 2>&1
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_03_bench.json 2> gpurun_out/r2_03_bench.err
tail -n 4 gpurun_out/r2_03_*.log
