#!/bin/bash
# Round evidence: smoke, parity tests, bench on every config + reference arm, ncu launch list and full captures.
# The .ncu-rep files are converted to CSV on the box (gpurun_out/ may carry 64 MiB back) and only c2's is kept.
set +e
mkdir -p gpurun_out
R=${ROUND:-r1}
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${R}_clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.log 2>gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; cat gpurun_out/bench_c2.log
kill $SMI
for c in c3 c5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_$c.log 2>gpurun_out/bench_$c.err; echo "bench $c rc=$?"; cat gpurun_out/bench_$c.log; done
timeout 600 python bench.py --config c5 --fit two-kernel --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_c5_two_kernel.log 2>gpurun_out/bench_c5_two_kernel.err; echo "bench c5 two-kernel rc=$?"; cat gpurun_out/bench_c5_two_kernel.log
timeout 600 python bench.py --config aux > gpurun_out/bench_aux.log 2>gpurun_out/bench_aux.err; echo "bench aux rc=$?"; cat gpurun_out/bench_aux.log
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/bench_c4.log 2>gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; cat gpurun_out/bench_c4.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${R}_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
prof() {  # name skip count kernel-regex bench-args...
  local name=$1 skip=$2 count=$3 re=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c $count -o gpurun_out/prof_$name python bench.py "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/${R}_prof_$name.raw.csv
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv > gpurun_out/prof_$name.sass.csv
  [ "$name" = c2 ] || rm -f gpurun_out/prof_$name.ncu-rep
}
prof c2 6 2 ct_ --steps 2 --warmup 3 --no-e2e --no-cpu
prof c3 2 2 ct_ --config c3 --steps 1 --warmup 1 --no-e2e --no-cpu
prof c5 2 1 ct_ --config c5 --steps 1 --warmup 1 --no-e2e --no-cpu
prof c4 40 8 'convert_|blend_|normal_' --config c4 --steps 1 --warmup 3
rm -f gpurun_out/prof_c3.sass.csv gpurun_out/prof_c4.sass.csv
du -sh gpurun_out
