#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_parity.py > gpurun_out/r2_32_dist2.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2_32_bench_n2.json 2> gpurun_out/r2_32_bench_n2.err
echo "exit code $?" >> gpurun_out/r2_32_bench_n2.err
tail -n 14 gpurun_out/r2_32_dist2.log; tail -c 300 gpurun_out/r2_32_bench_n2.err; tail -c 1200 gpurun_out/r2_32_bench_n2.json
