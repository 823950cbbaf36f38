#!/bin/bash
# baseline of the round-1 solver under the new (honest) protocol
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bench_sizes.py -x -q -m gpu -s -k "eigh_bench" > gpurun_out/r2_01_eigh_sizes.log 2>&1
KFAC_EIGH_JOPT=16 python -m pytest tests/test_gpu_bench_sizes.py -x -q -m gpu -s -k "eigh_bench and 4608" > gpurun_out/r2_01_eigh_sizes_jopt16.log 2>&1
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err
python -m pytest tests/test_gpu_bench_sizes.py -x -q -m gpu -s -k "resnet50 or neox" > gpurun_out/r2_01_full_parity.log 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden_replay" > gpurun_out/r2_01_golden.log 2>&1
tail -3 gpurun_out/r2_01_*.log
