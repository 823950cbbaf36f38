#!/bin/bash
cd $GRAFT_REPO_ROOT
for cfg in "1024 1" "2304 4" "4608 31" "256 1"; do
  timeout 300 python tests/sytrd_probe.py $cfg prof >> gpurun_out/r2_23_phases.log 2>&1
done
cat gpurun_out/r2_23_phases.log
