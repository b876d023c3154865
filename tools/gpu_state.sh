#!/bin/bash
# State check: smoke, parity tests, bench on c2 / c3 / c5, reference arm.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.log 2>gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
cat gpurun_out/bench_c2.log; tail -5 gpurun_out/bench_c2.err
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c3.log 2>gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
cat gpurun_out/bench_c3.log; tail -5 gpurun_out/bench_c3.err
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c5.log 2>gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
cat gpurun_out/bench_c5.log; tail -5 gpurun_out/bench_c5.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "bench ref rc=$?"
cat gpurun_out/bench_ref.log; tail -5 gpurun_out/bench_ref.err
