#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "exit code $?" >> gpurun_out/r2_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tests/dist_parity.py > gpurun_out/r2_dist8.log 2>&1
tail -c 300 gpurun_out/r2_bench_n8.err; tail -c 900 gpurun_out/r2_bench_n8.json; tail -n 6 gpurun_out/r2_dist8.log
