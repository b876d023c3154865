#!/bin/bash
# compute-sanitizer memcheck over smoke() (every kernel family, small shapes) and the C2-shape kFast kernels at a reduced batch
set +e
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_smoke.log 2>&1; echo "memcheck smoke rc=$?"; grep -E "ERROR SUMMARY|smoke ok|Invalid|error" gpurun_out/sanitize_smoke.log | head -10
TUNE_B=3 TUNE_L=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/tune.py > gpurun_out/sanitize_c2shape.log 2>&1; echo "memcheck kFast 3x1024^2 rc=$?"; grep -E "ERROR SUMMARY|Invalid|default" gpurun_out/sanitize_c2shape.log | head -6
TUNE_B=3 TUNE_L=5 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/tune.py > gpurun_out/sanitize_l5.log 2>&1; echo "memcheck plain-case L=5 rc=$?"; grep -E "ERROR SUMMARY|Invalid|default" gpurun_out/sanitize_l5.log | head -6
if [ "$1" = full ]; then
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q -k "not c3_shape and not c5_shape and not full_size and not no_host_round_trip and not no_host_staging" > gpurun_out/sanitize_pytest.log 2>&1; echo "memcheck pytest rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize_pytest.log | head -8
fi
