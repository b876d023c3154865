"""Key metrics of every kernel in an `ncu --page raw --csv` dump (stdin or file argument)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_xu.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")])
    for w in want:
        if w in hdr:
            print(f"  {w}: {r[hdr.index(w)]} [{rows[1][hdr.index(w)]}]")
    st = sorted(((float(r[hdr.index(h)]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for h in stalls), reverse=True)
    print("  stalls: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))
