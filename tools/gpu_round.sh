#!/bin/bash
# Round evidence: smoke, parity tests, bench on every config + reference arm, ncu launch list and full captures.
set +e
mkdir -p gpurun_out
R=${ROUND:-r1}
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.log 2>gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; cat gpurun_out/bench_c2.log
for c in c3 c5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_$c.log 2>gpurun_out/bench_$c.err; echo "bench $c rc=$?"; cat gpurun_out/bench_$c.log; done
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/bench_c4.log 2>gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; cat gpurun_out/bench_c4.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${R}_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ct_ -s 6 -c 2 -o gpurun_out/prof_c2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ct_ -s 2 -c 2 -o gpurun_out/prof_c3 python bench.py --config c3 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ct_ -s 2 -c 1 -o gpurun_out/prof_c5 python bench.py --config c5 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'convert_|blend_|normal_' -s 40 -c 8 -o gpurun_out/prof_c4 python bench.py --config c4 --steps 1 --warmup 3 > gpurun_out/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
ls -la gpurun_out | head -40
