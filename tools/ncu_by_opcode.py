"""Dynamic opcode mix from `ncu --page source --csv` (SASS view) on stdin: warp-instructions executed per opcode."""
import csv, sys, re
rows = list(csv.reader(sys.stdin))
hdr = None; agg = {}; tot = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r; ie = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= ie: continue
    try: n = int(r[ie])
    except ValueError: continue
    s = re.sub(r"^\s*@!?U?P\d+\s+", "", r[1].strip())
    op = s.split()[0].rstrip(";") if s else "?"
    op = op.split(".")[0] if len(sys.argv) < 2 else op
    agg[op] = agg.get(op, 0) + n; tot += n
print("total", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:45]:
    print(f"{100*v/tot:5.1f}%  {v/ (tot) * float(sys.argv[2]) if len(sys.argv)>2 else v:12.1f}  {k}")
