#!/bin/bash
# Round-2 check on the GPU box: smoke, parity tests, the driver's bench line, the other config lines.
set +e
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.log 2>gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; cat gpurun_out/bench_c2.log; tail -5 gpurun_out/bench_c2.err
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/bench_c4.log 2>gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; cat gpurun_out/bench_c4.log; tail -5 gpurun_out/bench_c4.err
nproc > gpurun_out/nproc.txt; free -g | head -2 >> gpurun_out/nproc.txt; nvidia-smi topo -m >> gpurun_out/nproc.txt 2>&1
