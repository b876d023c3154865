#!/bin/bash
# all GPU parity tests, no early exit; plus optional extra command line
set +e
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1800 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
