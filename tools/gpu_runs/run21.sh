#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_direct_eigh.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r2_21_direct.log 2>&1
timeout 300 python tests/sytrd_probe.py 4608 148 prof > gpurun_out/r2_21_prof4608.log 2>&1
timeout 300 python tests/sytrd_probe.py 2304 36 prof > gpurun_out/r2_21_prof2304.log 2>&1
timeout 600 python tests/sytrd_sweep.py > gpurun_out/sytrd_sweep.csv 2> gpurun_out/sytrd_sweep.err
timeout 300 python tests/eigh_batch_probe.py > gpurun_out/r2_21_batch.log 2>&1
tail -n 3 gpurun_out/r2_21_direct.log; head -3 gpurun_out/r2_21_prof4608.log; tail -3 gpurun_out/r2_21_batch.log
