"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K | python tools/ncu_by_line.py [top]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rows = list(csv.reader(sys.stdin))
agg = {}
fpath = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        ss = hdr.index("# Samples")
        continue
    if hdr is None or r[0] in ("", "Function Name"):
        continue
    try:
        inst = int(r[ie]); samp = int(r[ss])
    except ValueError:
        continue
    key = (fpath, int(r[0]), r[1].strip()[:110])
    a = agg.setdefault(key, [0, 0])
    a[0] += inst; a[1] += samp
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total warp-instructions {tot_i}  samples {tot_s}")
for (f, ln, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*i/tot_i:5.1f}% inst {100*s/max(tot_s,1):5.1f}% samp  {f}:{ln:<4d} {src}")
