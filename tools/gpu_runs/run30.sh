#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_30_pytest.log 2>&1
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_30_bench.json 2> gpurun_out/r2_30_bench.err
tail -n 3 gpurun_out/r2_30_pytest.log; tail -c 1500 gpurun_out/r2_30_bench.json
