#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sytrd_kernel -s 1 -c 1 -o gpurun_out/r2_10_sytrd256 python tests/sytrd_probe.py 256 1 > gpurun_out/r2_10_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sytrd_kernel -s 1 -c 1 -o gpurun_out/r2_10_sytrd4608 python tests/sytrd_probe.py 4608 148 > gpurun_out/r2_10_ncu2.log 2>&1
tail -n 3 gpurun_out/r2_10_ncu1.log
