#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pipeline_kernel --launch-skip 4 -c 3 -f -o gpurun_out/r2_grouped_merges python tests/eigh_batch_probe.py 1 > gpurun_out/r2_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pipeline_kernel --launch-skip 61 -c 3 -f -o gpurun_out/r2_grouped_bt python tests/eigh_batch_probe.py 1 > gpurun_out/r2_ncu2.log 2>&1
ls -la gpurun_out/r2_41*; tail -n 2 gpurun_out/r2_ncu1.log gpurun_out/r2_ncu2.log
