#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_31_batch_launches.csv python tests/eigh_batch_probe.py 1 > gpurun_out/r2_31_ncu.log 2>&1
tail -n 3 gpurun_out/r2_31_ncu.log; wc -l gpurun_out/r2_31_batch_launches.csv
