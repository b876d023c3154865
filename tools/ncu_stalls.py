"""Top SASS instructions by a stall reason from `ncu --page source --csv` (stdin).
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K | python tools/ncu_stalls.py stall_long_sb [top]"""
import csv, sys
col = sys.argv[1] if len(sys.argv) > 1 else "stall_long_sb"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(sys.stdin))
hdr = None
out = []
tot = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        ic, isamp = hdr.index(col), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ic:
        continue
    try:
        v = int(r[ic]); s = int(r[isamp])
    except ValueError:
        continue
    tot += v
    out.append((v, s, r[0], r[1].strip()))
print(f"total {col}: {tot}")
for v, s, a, src in sorted(out, reverse=True)[:top]:
    print(f"{v:8d} ({100*v/max(tot,1):4.1f}%) samples {s:7d}  {a}  {src[:100]}")
