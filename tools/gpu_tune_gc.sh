#!/bin/bash
# A/B of the libraries in pypbr_b200/lib/variants/ (python tools/build_variants.py <names>) on the multi-light shapes and on C2
# through the bare C ABI (tools/tune.py).  Usage: gpurun -- 'bash tools/gpu_tune_gc.sh'
set +e
mkdir -p gpurun_out
echo "== L=16 accumulate B=16 1024^2 (C3 shape / 4)"; TUNE_B=16 TUNE_L=16 timeout 600 python tools/tune.py 2>&1 | tail -8
echo "== L=8 per-light fused loss B=32 1024^2 (C5 shape)"; TUNE_B=32 TUNE_L=8 TUNE_PER_LIGHT=1 timeout 600 python tools/tune.py 2>&1 | tail -8
echo "== L=4 accumulate B=16"; TUNE_B=16 TUNE_L=4 timeout 600 python tools/tune.py 2>&1 | tail -8
echo "== L=1 B=64 1024^2 (C2 shape, streamed kernels)"; TUNE_B=64 TUNE_L=1 timeout 600 python tools/tune.py 2>&1 | tail -8
