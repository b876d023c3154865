"""Write profiles/traffic.json: per-launch DRAM bytes of the dominant kernels from `ncu --page raw --csv` dumps.
usage: python tools/ncu_traffic.py c2=gpurun_out/prof_c2.raw.csv c3=gpurun_out/prof_c3.raw.csv ..."""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out_path = os.path.join(ROOT, "profiles", "traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    cfg, path = arg.split("=")
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        rd = float(r[col["dram__bytes_read.sum"]]) * UNIT[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * UNIT[units[col["dram__bytes_write.sum"]]]
        kind = "backward" if "backward" in name else ("forward" if "forward" in name else name.split("(")[0].replace("void ", ""))
        def num(metric):
            return float(r[col[metric]]) if metric in col and r[col[metric]] not in ("", "n/a") else None
        out[f"{cfg}:{kind}"] = {"kernel": name, "bytes": rd + wr, "read": rd, "write": wr,
                                "fma_pipe_busy_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                                "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                "warp_instructions": num("smsp__inst_executed.sum"),
                                "source": f"ncu --set full, {os.path.basename(path)} (profiles/)"}
# the hash of the kernel sources the captures were taken from, written on the GPU box by tools/gpu_prof_r2.sh at capture time:
# bench.py only reports `roofline.traffic` while the sources it is built from still hash to this
stamp = os.path.join(ROOT, "gpurun_out", "csrc_sha16.txt")
if os.path.exists(stamp):
    out["_csrc_sha16"] = open(stamp).read().strip()
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
