#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_40_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_40_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_40_bench.json 2> gpurun_out/r2_40_bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:GroupedPolicy --launch-skip 8 -c 3 -f -o gpurun_out/r2_40_grouped python tests/eigh_batch_probe.py 1 > gpurun_out/r2_40_ncu.log 2>&1
tail -n 3 gpurun_out/r2_40_pytest.log gpurun_out/r2_40_smoke.log; tail -c 800 gpurun_out/r2_40_bench.json; tail -n 3 gpurun_out/r2_40_ncu.log
