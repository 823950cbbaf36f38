#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_08_launches.csv python tests/eigh_batch_probe.py 1 > gpurun_out/r2_08_ncu.log 2>&1
python tests/eigh_batch_probe.py 2 4608 > gpurun_out/r2_08_probe.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_08_launches4608.csv python tests/eigh_batch_probe.py 1 4608 >> gpurun_out/r2_08_ncu.log 2>&1
tail -n 3 gpurun_out/r2_08_probe.log
