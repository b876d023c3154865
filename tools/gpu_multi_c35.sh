#!/bin/bash
# N-GPU bench lines of the multi-light configs only (refresh after a kernel change)
set +e
N=${NGPU:-2}
mkdir -p gpurun_out
for c in ${CFGS:-c3 c5}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${c}_n$N.log 2>gpurun_out/bench_${c}_n$N.err; echo "bench $c n=$N rc=$?"
  tail -1 gpurun_out/bench_${c}_n$N.log | cut -c1-200
done
