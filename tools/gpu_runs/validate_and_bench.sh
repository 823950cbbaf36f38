#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -n 3 gpurun_out/r2_pytest.log gpurun_out/r2_smoke.log; tail -c 600 gpurun_out/r2_bench_ref.json; tail -c 1200 gpurun_out/r2_bench.json
