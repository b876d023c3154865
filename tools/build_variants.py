"""Build tuning variants of libpbrcuda.so (different -D knobs) into pypbr_b200/lib/variants/ — run on the CPU box.
usage: python tools/build_variants.py [name ...]   (default: all)"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def v(**kw):
    return [f"-DPBR_{k.upper()}={val}" for k, val in kw.items()]


VARIANTS = {
    "strict": ["-DPBR_STRICT_IEEE"],
    "default": [],
    # streamed kernels: stages / CTAs per SM (register cap) / texels shaded together / materials per CTA walk
    "f2g2": v(stream_fwd_min_ctas=2),
    "f2g4": v(stream_fwd_min_ctas=2, stream_fwd_group=4),
    "f3g1": v(stream_fwd_group=1),
    "f4g1": v(stream_fwd_min_ctas=4, stream_fwd_group=1),
    "f2g2_s3": v(stream_fwd_min_ctas=2, stream_stages=3),
    "b2g2": v(stream_bwd_group=2),
    "b3g1": v(stream_bwd_min_ctas=3),
    "h8": v(hoist_mats=8),
    "h32": v(hoist_mats=32),
    "h64": v(hoist_mats=64),
    "t128_f6b4": v(threads=128, stream_fwd_min_ctas=6, stream_bwd_min_ctas=4),
    "t128_f4b4_s3": v(threads=128, stream_fwd_min_ctas=4, stream_bwd_min_ctas=4, stream_stages=3),
    "t512_f1b1": v(threads=512, stream_fwd_min_ctas=1, stream_bwd_min_ctas=1),
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
