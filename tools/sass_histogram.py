"""Per-kernel SASS opcode histogram of libpbrcuda.so (static counts from `cuobjdump -sass`): the Blackwell-native evidence
(UBLKCP = TMA 1-D bulk copy, SYNCS = mbarrier, FFMA2 / FMUL2 / FADD2 = packed FP32, LDGSTS = cp.async, STG.E.EF = evict-first
stores) in a tracked file.  Runs without a GPU.
usage: python tools/sass_histogram.py [kernel-regex ...] > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pypbr_b200", "lib", "libpbrcuda.so")
want = [re.compile(p) for p in (sys.argv[1:] or [r"ct_forward_stream<0, 2, true>", r"ct_backward_stream<0, 2, 0, true>", r"ct_backward_stream<0, 2, 1, true>",
                                                 r"ct_backward_stream<0, 2, 0, false>",
                                                 r"ct_forward_kernel<0, 3, true, 1>", r"ct_backward_kernel<0, 3, false, true, 1>",
                                                 r"ct_backward_kernel<0, 3, false, true, 0>", r"ct_backward_kernel<0, 5, false, true, 0>",
                                                 r"convert_kernel", r"convert_bwd_kernel",
                                                 r"blend_kernel", r"blend_bwd_kernel", r"ingest_kernel<3, 8>", r"index_transform_kernel",
                                                 r"adam_kernel", r"normal_op_kernel", r"normal_ingest"])]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("pbr::", "").replace("(CtKParams)", "")
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        key = {"UBLKCP": "UBLKCP (TMA bulk copy)", "SYNCS": "SYNCS (mbarrier)", "LDGSTS": "LDGSTS (cp.async)"}.get(base, base)
        if base == "STG" and ".EF" in op:
            key = "STG.EF (evict-first store)"
        kernels[cur][key] += 1
digest = subprocess.run(["sha256sum", LIB], capture_output=True, text=True).stdout.split()[0][:16]
print(f"# SASS opcode histogram of libpbrcuda.so (sm_100a), static instruction counts per kernel\n")
print(f"`cuobjdump -sass pypbr_b200/lib/libpbrcuda.so` -> tools/sass_histogram.py; library sha256 {digest}...; {len(kernels)} kernels in the library.\n")
print("Packed FP32 (`FFMA2 / FMUL2 / FADD2`), the TMA bulk copy (`UBLKCP`), mbarrier operations (`SYNCS`), `cp.async` (`LDGSTS`) and\n"
      "evict-first stores are what marks these kernels as written for sm_100a; there is no tensor-core instruction because the path has\n"
      "no contraction (DESIGN.md §3).\n")
for name, c in kernels.items():
    if not any(p.search(name) for p in want):
        continue
    total = sum(c.values())
    top = ", ".join(f"{k} {v}" for k, v in c.most_common(14))
    marks = {k: c[k] for k in c if any(t in k for t in ("UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "LDGSTS", "STG.EF", "MUFU"))}
    print(f"## `{name}` — {total} instructions\n")
    print("markers: " + ", ".join(f"**{k} {v}**" for k, v in sorted(marks.items())) + "\n")
    print("top opcodes: " + top + "\n")
