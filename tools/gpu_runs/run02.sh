#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_direct_eigh.py -x -q -m gpu -s -k "sytrd" > gpurun_out/r2_02_sytrd.log 2>&1
timeout 600 python -m pytest tests/test_gpu_direct_eigh.py -x -q -m gpu -s -k "stedc" > gpurun_out/r2_02_stedc.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k "eigh" > gpurun_out/r2_02_eigh.log 2>&1
timeout 300 python -m pytest tests/test_gpu_bench_sizes.py -x -q -m gpu -s -k "eigh_bench" > gpurun_out/r2_02_sizes.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tc_gemm.py -x -q -m gpu -k "not eigh" > gpurun_out/r2_02_kernels.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_02_parity.log 2>&1
tail -n 5 gpurun_out/r2_02_*.log
