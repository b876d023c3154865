#!/bin/bash
set +e
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== C2 shape (B=32, L=1)"; TUNE_B=32 TUNE_L=1 python tools/tune.py 2>&1 | tail -8
echo "== C3 shape (B=16, L=16 accumulate)"; TUNE_B=16 TUNE_L=16 python tools/tune.py 2>&1 | tail -8
echo "== C5 shape (B=32, L=8 per-light fused loss)"; TUNE_B=32 TUNE_L=8 TUNE_PER_LIGHT=1 python tools/tune.py 2>&1 | tail -8
echo "== L=12 accumulate"; TUNE_B=16 TUNE_L=12 python tools/tune.py 2>&1 | tail -8
echo "== L=6 accumulate"; TUNE_B=16 TUNE_L=6 python tools/tune.py 2>&1 | tail -8
for c in c3 c5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-e2e --no-cpu --no-eager | tee gpurun_out/bench_$c.log | cut -c1-400; done
