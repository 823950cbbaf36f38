#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 300 python tests/eigh_batch_probe.py > gpurun_out/r2_22_batch.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_22_batch_launches.csv python tests/eigh_batch_probe.py 1 > gpurun_out/r2_22_ncu.log 2>&1
timeout 600 python -m pytest tests/test_gpu_bench_sizes.py tests/test_gpu_direct_eigh.py -x -q -m gpu -s > gpurun_out/r2_22_sizes.log 2>&1
tail -n 4 gpurun_out/r2_22_batch.log; tail -n 3 gpurun_out/r2_22_sizes.log
