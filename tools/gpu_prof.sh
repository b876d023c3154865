#!/bin/bash
# parity tests + bench + ncu (launch list and full capture with source) of the default library
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ct_ -s 6 -c 2 -o gpurun_out/prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
