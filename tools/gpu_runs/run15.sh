#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_direct_eigh.py -x -q -m gpu -s -k sytrd > gpurun_out/r2_15_direct.log 2>&1
for cfg in "4608 148" "4608 36" "1024 12" "256 1"; do timeout 120 python tests/sytrd_probe.py $cfg prof; done > gpurun_out/r2_15_phases.log 2>&1
timeout 300 python -m pytest tests/test_gpu_bench_sizes.py -x -q -m gpu -s -k "eigh_bench" > gpurun_out/r2_15_sizes.log 2>&1
timeout 200 python tests/eigh_batch_probe.py 3 > gpurun_out/r2_15_probe.log 2>&1
tail -n 4 gpurun_out/r2_15_direct.log gpurun_out/r2_15_probe.log
