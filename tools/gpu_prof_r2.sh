#!/bin/bash
# Round-2 evidence on the GPU box: ncu launch list of the driver's bench command, `--set full` captures of the dominant kernels
# (converted to CSV on the box; gpurun_out/ carries at most 64 MiB back), the hash of the kernel sources they were taken from.
set +e
mkdir -p gpurun_out
R=${ROUND:-r2}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
python -c "import bench; print(bench.csrc_sha16())" > gpurun_out/csrc_sha16.txt; cat gpurun_out/csrc_sha16.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-eager > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
prof() {  # name skip count kernel-regex bench-args...
  local name=$1 skip=$2 count=$3 re=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c $count -o gpurun_out/prof_$name python bench.py "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/${R}_prof_$name.raw.csv
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv > gpurun_out/prof_$name.sass.csv 2>/dev/null
  [ "$name" = c2 ] || rm -f gpurun_out/prof_$name.ncu-rep
}
prof c2 6 2 ct_ --steps 2 --warmup 3 --no-e2e --no-cpu --no-eager --no-configs
prof c3 2 2 ct_ --config c3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-eager
prof c5 2 1 ct_ --config c5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-eager
prof c4 40 8 'convert_|blend_|normal_' --config c4 --steps 1 --warmup 3
python tools/ncu_by_opcode.py < gpurun_out/prof_c2.sass.csv > gpurun_out/${R}_c2_executed_opcodes.txt 2>&1
python tools/ncu_by_opcode.py < gpurun_out/prof_c5.sass.csv > gpurun_out/${R}_c5_executed_opcodes.txt 2>&1
python tools/ncu_by_opcode.py < gpurun_out/prof_c3.sass.csv > gpurun_out/${R}_c3_executed_opcodes.txt 2>&1
rm -f gpurun_out/prof_c3.sass.csv gpurun_out/prof_c4.sass.csv gpurun_out/prof_c5.sass.csv
du -sh gpurun_out
