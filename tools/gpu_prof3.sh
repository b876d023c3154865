#!/bin/bash
# ncu full capture of one bench config; keeps the report only if small, always exports the raw + source CSVs
set +e
mkdir -p gpurun_out
C=${CFG:-c5}; S=${SKIP:-2}; N=${COUNT:-1}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ct_ -s $S -c $N -o gpurun_out/prof_$C python bench.py --config $C --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_$C.log 2>&1; echo "ncu $C rc=$?"
ncu -i gpurun_out/prof_$C.ncu-rep --page raw --csv > gpurun_out/prof_$C.raw.csv
ncu -i gpurun_out/prof_$C.ncu-rep --page source --csv > gpurun_out/prof_$C.sass.csv
ncu -i gpurun_out/prof_$C.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_$C.src.csv
rm -f gpurun_out/prof_$C.ncu-rep
ls -la gpurun_out
