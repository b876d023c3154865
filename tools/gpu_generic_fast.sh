#!/bin/bash
# Plain-case flavours of the generic multi-light kernels (kFmAccum / kFmFit): parity suite, then A/B against the general flavour
# (PBR_DISABLE_GENERIC_FAST=1) on C3 and C5, bench.py's own timing.
set +e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
b() { local name=$1; shift; timeout 600 python bench.py "$@" --steps 10 --warmup 3 --no-e2e --no-eager --no-cpu > gpurun_out/r2_ab_$name.json 2>gpurun_out/ab_$name.err; echo "bench $name rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_ab_$name.json').read().strip().splitlines()[-1])
print('$name', d['value'], d['ms_per_step'], json.dumps(d.get('roofline'))[:600])
PY
}
b c3_fast --config c3
PBR_DISABLE_GENERIC_FAST=1 b c3_general --config c3
b c5_fast --config c5
PBR_DISABLE_GENERIC_FAST=1 b c5_general --config c5
b c5_two_kernel_fast --config c5 --fit two-kernel
