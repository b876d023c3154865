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
    "scalar": ["-DPBR_SCALAR_LANES", "-DPBR_FWD_GROUP=2", "-DPBR_CT_THREADS=256", "-DPBR_CT_TEXELS=4", "-DPBR_FWD_MIN_CTAS=3", "-DPBR_BWD_MIN_CTAS=2",
               "-DPBR_STREAM_THREADS=256", "-DPBR_STREAM_FWD_MIN_CTAS=2", "-DPBR_STREAM_BWD_MIN_CTAS=2"],   # one texel per lane (V = float): accuracy / speed reference
    "default": [],
    # generic (multi-light) kernels: threads per CTA / texels per thread / CTAs per SM (register cap)
    "g_t128x2_f5b4": v(fwd_min_ctas=5, bwd_min_ctas=4),
    "g_t128x2_f4b2": v(fwd_min_ctas=4, bwd_min_ctas=2),
    "g_t128x4_f4b3": v(ct_texels=4),
    "g_t256x2_f2b2": v(ct_threads=256, fwd_min_ctas=2, bwd_min_ctas=2),
    "g_t256x2_f3b1": v(ct_threads=256, fwd_min_ctas=3, bwd_min_ctas=1),
    "g_t64x2_f8b6": v(ct_threads=64, fwd_min_ctas=8, bwd_min_ctas=6),
    "g_t64x2_f10b5": v(ct_threads=64, fwd_min_ctas=10, bwd_min_ctas=5),
    # geometry cache of the multi-light kernels (point lights, L > 1): off / backward register budget / light-loop unrolling
    "gc_off": ["-DPBR_GC_MAX_BYTES=0"],
    "gc_b3": v(bwd_cached_min_ctas=3),
    "s_cta": v(stream_per_warp_fwd=0, stream_per_warp_bwd=0),
    "s_pw": v(stream_per_warp_fwd=1, stream_per_warp_bwd=1),
    "s_pw_late": v(stream_per_warp_fwd=1, stream_per_warp_bwd=1, stream_bwd_late_refill=1),
    "s_pw_late3": v(stream_per_warp_fwd=1, stream_per_warp_bwd=1, stream_bwd_late_refill=1, stream_stages=3),
    "s_pw3": v(stream_stages=3),
    "s_pw4": v(stream_stages=4),
    "x_nopf": v(prefetch_next=0),
    "x_noregpf": v(fwd_reg_prefetch=0),
    "x_nolate": v(bwd_late_prefetch=0),
    "x_cold1": v(cold_unroll=1),
    "x_pl0": v(packed_loss=0),
    "u_f2": v(fwd_unroll=2),
    "u_p2": v(p1_unroll=2),
    "u_f2p2": v(fwd_unroll=2, p1_unroll=2),
    "u_f2_f3": v(fwd_unroll=2, fwd_min_ctas=3),
    "u_f4_f2": v(fwd_unroll=4, fwd_min_ctas=2),
    # round 2
    "r2_noenc": v(encode_diet=0),               # sRGB encode on the shading path with the clamps / guards that cannot bind
    "r2_nobr": v(boundary_recompute=0),         # one-pass accumulate backward without the boundary recompute
    "r2_nobig": v(gc_big_max_lights=0),         # backward, 4 < L <= 8: 6-field cache instead of all 8 fields
    "r2_bpw": v(stream_per_warp_bwd=1),                                # streamed backward: per-warp copy pipelines, refill after the last read
    "r2_bpw_late": v(stream_per_warp_bwd=1, stream_bwd_late_refill=1), # ... refill after the tile
    "r2_st3": v(stream_stages=3),                                      # three stages (CTA-level backward, per-warp forward)
    "r2_bpw_st3": v(stream_per_warp_bwd=1, stream_stages=3),
    "r2_hm22": v(hoist_mats=22),                                       # 22 materials per CTA walk: 3 chunks of a 64-material batch
    "r2_hm32": v(hoist_mats=32),
    "r2_f2": v(fwd_min_ctas=2),                 # generic forward with the full register budget (the cache's shared memory caps it at 2 CTAs for L >= 11 anyway)
    "r2_f2u2": v(fwd_min_ctas=2, fwd_unroll=2),
    "r2_f3u2": v(fwd_min_ctas=3, fwd_unroll=2),
    "r2_u2": v(fwd_unroll=2),
    "r2_late": v(bwd_late_prefetch=1),          # backward: next material's maps requested when the light loop is over
    "r2_sb4": v(stream_bwd_min_ctas=4),         # streamed backward capped at 128 registers (4 CTAs per SM)
    "r2_sf5": v(stream_fwd_min_ctas=5),         # streamed forward capped at 102 registers (5 CTAs per SM)
    # round 2, second session: kFast flavour of the streamed kernels + try_wait with a suspend-time hint
    "r3_nofast": v(stream_fast=0),                                     # hint only
    "r3_nohint": v(wait_hint_ns=0),                                    # kFast only
    "r3_old": v(stream_fast=0, wait_hint_ns=0),                        # neither: the code of the first session
    "r3_bpw": v(stream_per_warp_bwd=1),                                # kFast + per-warp backward pipelines (the default since)
    "r3_bcta": v(stream_per_warp_bwd=0),                               # kFast + CTA-level backward pipeline
    "r3_bpw_late": v(stream_bwd_late_refill=1),
    "r3_bpw_st3": v(stream_stages=3),
    "r3_bpw_hm32": v(hoist_mats=32),
    "r3_bpw_hm8": v(hoist_mats=8),
    "r3_last": v(stream_bwd_last_refill=1),                            # kFast + the last reader of a stage refills it
    "r3_st3": v(stream_stages=3),
    "r3_last_st3": v(stream_bwd_last_refill=1, stream_stages=3),
    "r3_sb4": v(stream_bwd_min_ctas=4),
    "r3_hm32": v(hoist_mats=32),
    # light-loop unrolling on top of the plain-case (single basic block) flavour of the generic kernels
    "r3_fu2": v(fwd_unroll=2),
    "r3_fu4": v(fwd_unroll=4),
    "r3_bu2": v(bwd_unroll=2),
    "r3_fu2_f3": v(fwd_unroll=2, fwd_cached_min_ctas=3),
    "r3_gbu2": v(bwd_unroll=2),               # general flavour of the backward (the per-light loss / fit kernels), light loop unrolled by two
    "r3_f3": v(fwd_cached_min_ctas=3),
    "r3_b3": v(bwd_cached_min_ctas=3),
    # memory pipeline only (no shading math): the floor of the TMA-in / STG-out design
    "nomath": ["-DPBR_DBG_NOMATH"],
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
