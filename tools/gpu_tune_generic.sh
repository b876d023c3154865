#!/bin/bash
# light-loop unrolling / register budgets on top of the plain-case flavour of the generic kernels (tools/build_variants.py r3_*)
set +e
mkdir -p gpurun_out
echo "== C3 shape (B=16, L=16 accumulate, 1024^2)"; TUNE_B=16 TUNE_L=16 python tools/tune.py 2>&1 | tail -8
cp gpurun_out/tune_B16_L16.json gpurun_out/r2_tune_c3shape_plain_unroll.json
echo "== L=8 accumulate"; TUNE_B=16 TUNE_L=8 python tools/tune.py 2>&1 | tail -8
echo "== L=4 accumulate"; TUNE_B=16 TUNE_L=4 python tools/tune.py 2>&1 | tail -8
