#!/bin/bash
# the driver's multi-GPU line on the final code (N = $1): every config in one line, e2e included
set +e
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-eager --no-cpu > gpurun_out/r2_bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n=$N rc=$?"; cut -c1-700 gpurun_out/r2_bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
