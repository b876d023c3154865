#!/bin/bash
# Round-2 evidence, one run on the final code: smoke, parity tests, the driver's bench line + the reference arm, the other
# config lines, ncu launch lists of the bench commands, `--set full` captures of the dominant kernels (CSV on the box), the hash
# of the kernel sources everything was taken from.
set +e
mkdir -p gpurun_out
R=r2
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-300
python -c "import bench; print(bench.csrc_sha16())" > gpurun_out/csrc_sha16.txt; cat gpurun_out/csrc_sha16.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log | cut -c1-200
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${R}_clocks.csv &
SMI=$!
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${R}_bench_default.json 2>gpurun_out/bench_default.err; echo "bench default rc=$?"; cut -c1-400 gpurun_out/${R}_bench_default.json
kill $SMI
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/${R}_bench_ref.json 2>gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/${R}_bench_ref.json
b() { local name=$1; shift; timeout 600 python bench.py "$@" --steps 10 --warmup 3 --no-e2e --no-eager > gpurun_out/${R}_bench_$name.json 2>gpurun_out/bench_$name.err; echo "bench $name rc=$?"; cut -c1-260 gpurun_out/${R}_bench_$name.json; }
b c5_sync --config c5 --fit one-launch-sync --no-cpu
b c5_two_kernel --config c5 --fit two-kernel --no-cpu
b c5_two_kernel_shared_grads --config c5 --fit two-kernel --shared-grads --no-cpu
b c4 --config c4
b aux --config aux
b ref_c4 --impl reference --config c4
launches() { local name=$1; shift; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${R}_launches_$name.csv python bench.py "$@" --steps 3 --warmup 3 --no-e2e --no-cpu --no-eager --no-configs > /dev/null 2>&1; echo "ncu launches $name rc=$?"; }
launches c2
launches c3 --config c3
launches c5 --config c5
prof() {  # name skip count kernel-regex bench-args...
  local name=$1 skip=$2 count=$3 re=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c $count -o gpurun_out/prof_$name python bench.py "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/${R}_prof_$name.raw.csv
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv > gpurun_out/prof_$name.sass.csv 2>/dev/null
  python tools/ncu_by_opcode.py < gpurun_out/prof_$name.sass.csv > gpurun_out/${R}_${name}_executed_opcodes.txt 2>&1
  [ "$name" = c2 ] || rm -f gpurun_out/prof_$name.ncu-rep gpurun_out/prof_$name.sass.csv
}
prof c2 6 2 ct_ --steps 2 --warmup 3 --no-e2e --no-cpu --no-eager --no-configs
prof c3 2 2 ct_ --config c3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-eager
prof c5 2 1 ct_ --config c5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-eager
prof c4 18 6 'convert_|blend_|normal_' --config c4 --steps 1 --warmup 3 --no-cpu
du -sh gpurun_out
