"""Build tuning variants of libpbrcuda.so (different -D knobs) into pypbr_b200/lib/variants/ — run on the CPU box."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

def v(texels, fg, bg, fcta, bcta, hoist=8, threads=256):
    return [f"-DPBR_TEXELS={texels}", f"-DPBR_FWD_GROUP={fg}", f"-DPBR_BWD_GROUP={bg}", f"-DPBR_FWD_MIN_CTAS={fcta}",
            f"-DPBR_BWD_MIN_CTAS={bcta}", f"-DPBR_HOIST_MATS={hoist}", f"-DPBR_THREADS={threads}"]


VARIANTS = {
    "strict": v(4, 2, 1, 3, 2) + ["-DPBR_STRICT_IEEE"],
    "t4_f2c3_b1c2": v(4, 2, 1, 3, 2),
    "t4_f2c3_b1c2_h16": v(4, 2, 1, 3, 2, hoist=16),
    "t4_f1c4_b1c3": v(4, 1, 1, 4, 3),
    "t2_f2c4_b1c2": v(2, 2, 1, 4, 2),
    "t2_f2c4_b1c3": v(2, 2, 1, 4, 3),
    "t2_f1c5_b1c3": v(2, 1, 1, 5, 3),
    "t1_c3_c2": v(1, 1, 1, 3, 2),
    "t1_c4_c2": v(1, 1, 1, 4, 2),
    "t1_c4_c3": v(1, 1, 1, 4, 3),
    "t1_c5_c3": v(1, 1, 1, 5, 3),
    "t1_c6_c4": v(1, 1, 1, 6, 4),
    "t1_c4_c3_h16": v(1, 1, 1, 4, 3, hoist=16),
    "t1_c4_c3_h4": v(1, 1, 1, 4, 3, hoist=4),
    "t1_128_c8_c6": v(1, 1, 1, 8, 6, threads=128),
    "t2_128_c8_c5": v(2, 2, 1, 8, 5, threads=128),
}


def one(item):
    name, flags = item
    out = os.path.join(ROOT, "pypbr_b200", "lib", "variants", f"libpbrcuda_{name}.so")
    g.build_cuda(force=True, extra_flags=flags, out=out)
    return name


if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    with ThreadPoolExecutor(8) as ex:
        for n in ex.map(one, [(n, VARIANTS[n]) for n in names]):
            print("built", n, flush=True)
