#!/bin/bash
# multi-GPU line exactly as the driver launches it (N = $1)
set +e
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_c2_n$N.log 2> gpurun_out/bench_c2_n$N.err; echo "bench n=$N rc=$?"; cat gpurun_out/bench_c2_n$N.log; tail -5 gpurun_out/bench_c2_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref n=$N rc=$?"; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --config c5 --fit one-launch-sync --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_c5sync_n$N.log 2>&1; echo "c5 sync n=$N rc=$?"; tail -1 gpurun_out/bench_c5sync_n$N.log | cut -c1-400
