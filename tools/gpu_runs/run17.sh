#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_17_pytest.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_17_bench.json 2> gpurun_out/r2_17_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_17_bench_ref.json 2> gpurun_out/r2_17_bench_ref.err
tail -n 5 gpurun_out/r2_17_pytest.log
