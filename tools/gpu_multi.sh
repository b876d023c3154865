#!/bin/bash
# N-GPU bench lines exactly as the driver launches them (torchrun, one rank per GPU, NCCL).
set +e
N=${NGPU:-2}
mkdir -p gpurun_out
for c in c2 c3 c5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps 10 --warmup 3 > gpurun_out/bench_${c}_n$N.log 2>gpurun_out/bench_${c}_n$N.err; echo "bench $c n=$N rc=$?"
  tail -1 gpurun_out/bench_${c}_n$N.log; tail -3 gpurun_out/bench_${c}_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref n=$N rc=$?"; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-300
