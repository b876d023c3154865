#!/bin/bash
# Round 2, second session: parity suite on the kFast / per-warp streamed kernels, the C2-shape variants
# (tools/build_variants.py r3_*) and the driver-style bench line.
set +e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== C2 shape (B=64, L=1)"; TUNE_B=64 TUNE_L=1 python tools/tune.py 2>&1 | tail -14
cp gpurun_out/tune_B64_L1.json gpurun_out/r2_tune_c2_perwarp_variants.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_default.log
