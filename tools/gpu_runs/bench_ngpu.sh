#!/bin/bash
cd $GRAFT_REPO_ROOT
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "exit code $?" >> gpurun_out/r2_bench_n$N.err
tail -c 200 gpurun_out/r2_bench_n$N.err; tail -c 600 gpurun_out/r2_bench_n$N.json
