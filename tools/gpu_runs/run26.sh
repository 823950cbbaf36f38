#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_direct_eigh.py -x -q -m gpu -s -k "sytrd or direct" > gpurun_out/r2_27_direct.log 2>&1
timeout 300 python tests/sytrd_probe.py 4608 148 prof > gpurun_out/r2_27_prof.log 2>&1
timeout 300 python tests/sytrd_probe.py 4608 31 prof >> gpurun_out/r2_27_prof.log 2>&1
timeout 300 python tests/sytrd_probe.py 2304 4 prof >> gpurun_out/r2_27_prof.log 2>&1
timeout 600 python tests/sytrd_sweep.py > gpurun_out/sytrd_sweep3.csv 2> gpurun_out/sytrd_sweep3.err
timeout 300 python tests/eigh_batch_probe.py > gpurun_out/r2_27_batch.log 2>&1
tail -n 3 gpurun_out/r2_27_direct.log; grep "sytrd n=\|tiles\|total" gpurun_out/r2_27_prof.log; tail -n 4 gpurun_out/r2_27_batch.log
