#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
for n in 1 2 4 $N; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n tools/h2d_scaling.py 2>/dev/null | tail -1 | tee -a gpurun_out/r2_h2d_scaling.jsonl; done
BIND=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29559 tools/h2d_scaling.py 2>/dev/null | tail -1 | tee -a gpurun_out/r2_h2d_scaling.jsonl
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1; nproc >> gpurun_out/r2_topo_n$N.txt; free -g | head -2 >> gpurun_out/r2_topo_n$N.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/r2_topo_n$N.txt
