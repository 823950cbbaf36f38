#!/bin/bash
cd $GRAFT_REPO_ROOT
for cfg in "4608 148" "4608 36" "2304 24" "1024 12" "256 1"; do python tests/sytrd_probe.py $cfg prof; done > gpurun_out/r2_14_phases.log 2>&1
tail -n 60 gpurun_out/r2_14_phases.log
