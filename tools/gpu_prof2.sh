#!/bin/bash
# ncu full captures (with source) of the Cook-Torrance kernels on c2 (L=1, streamed) and c3 (L=16, generic)
set +e
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ct_ -s 6 -c 2 -o gpurun_out/prof_c2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ct_ -s 2 -c 2 -o gpurun_out/prof_c3 python bench.py --config c3 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
ls -la gpurun_out
