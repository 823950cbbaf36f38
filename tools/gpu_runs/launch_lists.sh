#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-stationary > gpurun_out/r2_ncu_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_batch_launches.csv python tests/eigh_batch_probe.py 1 > gpurun_out/r2_ncu.log 2>&1
wc -l gpurun_out/r2_bench_launches.csv gpurun_out/r2_batch_launches.csv
