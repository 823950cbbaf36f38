#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_direct_eigh.py -x -q -m gpu -s -k sytrd > gpurun_out/r2_05_direct.log 2>&1
timeout 300 python -m pytest tests/test_gpu_bench_sizes.py -x -q -m gpu -s -k "eigh_bench" > gpurun_out/r2_05_sizes.log 2>&1
python tests/eigh_batch_probe.py 3 > gpurun_out/r2_05_probe.log 2>&1
tail -n 4 gpurun_out/r2_05_*.log
