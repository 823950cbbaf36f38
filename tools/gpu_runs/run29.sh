#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_29_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-stationary > gpurun_out/r2_29_ncu_bench.log 2>&1
timeout 300 python tests/eigh_batch_probe.py > gpurun_out/r2_29_batch.log 2>&1
tail -n 3 gpurun_out/r2_29_ncu_bench.log | cut -c1-300; tail -n 4 gpurun_out/r2_29_batch.log; wc -l gpurun_out/r2_29_bench_launches.csv
