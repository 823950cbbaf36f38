#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bench_sizes.py tests/test_gpu_direct_eigh.py -x -q -m gpu -s -k "mixed or stedc or sytrd" > gpurun_out/r2_38_mixed.log 2>&1
grep "mixed batch" gpurun_out/r2_38_mixed.log | cut -c1-200; tail -n 3 gpurun_out/r2_38_mixed.log
