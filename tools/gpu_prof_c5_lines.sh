#!/bin/bash
# per-source-line instruction counts of the C5 fit kernel (ncu source view with -lineinfo)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ct_backward_kernel -s 2 -c 1 -o gpurun_out/prof_c5l python bench.py --config c5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-eager > gpurun_out/ncu_c5l.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_c5l.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_c5_lines.csv
rm -f gpurun_out/prof_c5l.ncu-rep
ls -la gpurun_out/prof_c5_lines.csv
