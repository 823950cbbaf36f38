#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_20_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_20_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_20_bench.json 2> gpurun_out/r2_20_bench.err
tail -n 3 gpurun_out/r2_20_pytest.log gpurun_out/r2_20_smoke.log
