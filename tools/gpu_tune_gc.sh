#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | grep -E "^E   |passed|failed" | cut -c1-260 | head -6
echo "== L=1 B=64 1024^2"; TUNE_B=64 TUNE_L=1 timeout 600 python tools/tune.py 2>&1 | tail -6
