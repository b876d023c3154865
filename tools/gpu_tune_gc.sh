#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== L=16 accumulate B=16 1024^2"; TUNE_B=16 TUNE_L=16 timeout 600 python tools/tune.py 2>&1 | tail -3
echo "== L=8 per-light fused loss B=32"; TUNE_B=32 TUNE_L=8 TUNE_PER_LIGHT=1 timeout 600 python tools/tune.py 2>&1 | tail -3
echo "== L=4 accumulate B=16"; TUNE_B=16 TUNE_L=4 timeout 600 python tools/tune.py 2>&1 | tail -3
echo "== L=2 accumulate B=16"; TUNE_B=16 TUNE_L=2 timeout 600 python tools/tune.py 2>&1 | tail -3
