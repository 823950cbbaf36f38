#!/bin/bash
cd $GRAFT_REPO_ROOT
python tests/sytrd_probe.py 1024 12 > gpurun_out/r2_04_probe.log 2>&1
python tests/sytrd_probe.py 4608 148 >> gpurun_out/r2_04_probe.log 2>&1
python tests/eigh_batch_probe.py 3 >> gpurun_out/r2_04_probe.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sytrd_kernel -s 1 -c 1 -o gpurun_out/r2_04_sytrd1024 python tests/sytrd_probe.py 1024 12 > gpurun_out/r2_04_ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_04_launches.csv python tests/eigh_batch_probe.py 1 > gpurun_out/r2_04_ncu2.log 2>&1
tail -n 6 gpurun_out/r2_04_probe.log gpurun_out/r2_04_ncu1.log
