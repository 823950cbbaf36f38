#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sytrd_kernel -c 1 -f -o gpurun_out/r2_sytrd2304c1 python tests/sytrd_probe.py 2304 1 > gpurun_out/r2_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sytrd_kernel -c 1 -f -o gpurun_out/r2_sytrd4608 python tests/sytrd_probe.py 4608 148 > gpurun_out/r2_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 3 gpurun_out/r2_ncu1.log gpurun_out/r2_ncu2.log
