#!/bin/bash
# bench lines of the multi-light configs (c3, c5 one-launch / two-kernel) + c2 as a regression check
set +e
mkdir -p gpurun_out
for a in "--config c3" "--config c5" "--config c5 --fit two-kernel" "--config c2"; do
  n=$(echo $a | tr -d ' -' )
  timeout 600 python bench.py $a --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_$n.log 2>gpurun_out/bench_$n.err; echo "bench $a rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$n.log").read().strip().splitlines()[-1])
r=d["roofline"]
print("  value %.2f %s  ms/step %.3f  kernel %.3f ms frac %.3f" % (d["value"], d["unit"], d["ms_per_step"], r["kernel_ms"], r["frac"]), {k:(round(v["kernel_ms"],3)) for k,v in r.items() if isinstance(v,dict) and "kernel_ms" in v})
PY
done
