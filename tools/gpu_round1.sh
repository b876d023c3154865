#!/bin/bash
# First GPU pass: smoke, parity tests, variant timing, bench, ncu launch list + full capture.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python tools/tune.py > gpurun_out/tune.log 2>&1; echo "tune rc=$?"
cat gpurun_out/tune.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ct_ -s 6 -c 2 -o gpurun_out/prof_r1 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
