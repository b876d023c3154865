// pbr_grad_kernels.cuh — adjoints of the streaming kernels around the shading path, so that a fit which goes through a
// workflow conversion, a blend or a normal-map ingestion back-propagates like the reference's plain torch ops do
// (pypbr/materials/metallic.py:103-109, diffuse.py:129-147, blending/functional.py:104-145, materials/base.py:215-242):
//   convert_bwd_kernel<kM2S>  : pbr_convert_m2s_backward / pbr_convert_s2m_backward
//   blend_bwd_kernel          : pbr_blend_backward (all maps of a blend in one pass, d_mask / d_prop reduced per texel)
//   normal_ingest_bwd_kernel  : pbr_normal_ingest_backward
// The adjoints of the index transforms are index transforms themselves (pbr_index_transform, reduce_y / reduce_x for tile).
// All HBM-bound, same thread mapping as the forward kernels.  Included by pbr_kernels.cu.
#pragma once

namespace pbr {

// zero when the plane is absent (a NULL gradient is a zero gradient)
template <int N>
__device__ __forceinline__ void load_or_zero(const PbrPlane& pl, const Where& w, int c, float (&dst)[N]) {
  if (pl.ptr) {
    load_seg<N>(pl.ptr + plane_off(pl, w.b, c, w.row, w.col0), w.vec, w.valid, dst);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = 0.0f;
  }
}

template <bool kM2S>
__global__ void __launch_bounds__(kThreads) convert_bwd_kernel(const __grid_constant__ ConvKParams p) {
  const Where w = locate(p.H, p.W, p.vec_ok != 0);
  if (!w.active) return;
  float a[3][kTexels], m[3][kTexels], g0[3][kTexels], g1[3][kTexels], da[3][kTexels], dm[3][kTexels];
  const int mc = kM2S ? p.met_channels : 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    load_seg<kTexels>(p.albedo.ptr + plane_off(p.albedo, w.b, c, w.row, w.col0), w.vec, w.valid, a[c]);
    if (c < mc) load_seg<kTexels>(p.metspec.ptr + plane_off(p.metspec, w.b, c, w.row, w.col0), w.vec, w.valid, m[c]);
    load_or_zero<kTexels>(p.g0, w, c, g0[c]);
    load_or_zero<kTexels>(p.g1, w, c, g1[c]);
  }
#pragma unroll
  for (int i = 0; i < kTexels; ++i) {
    if (kM2S) {
      const float a3[3] = {a[0][i], a[1][i], a[2][i]};
      const float m3[3] = {m[0][i], mc == 3 ? m[1][i] : m[0][i], mc == 3 ? m[2][i] : m[0][i]};
      const float gd[3] = {g0[0][i], g0[1][i], g0[2][i]}, gs[3] = {g1[0][i], g1[1][i], g1[2][i]};
      float da3[3], dm3[3];
      convert_m2s_bwd(a3, m3, p.albedo_is_srgb != 0, gd, gs, da3, dm3);
#pragma unroll
      for (int c = 0; c < 3; ++c) { da[c][i] = da3[c]; dm[c][i] = dm3[c]; }
      if (mc == 1) dm[0][i] = (dm3[0] + dm3[1]) + dm3[2];   // a 1-channel map broadcasts: its gradient sums the three
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) convert_s2m_bwd(a[c][i], m[c][i], p.albedo_is_srgb != 0, g0[c][i], g1[c][i], &da[c][i], &dm[c][i]);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (p.d_albedo.ptr) store_seg<kTexels>(p.d_albedo.ptr + plane_off(p.d_albedo, w.b, c, w.row, w.col0), w.vec, w.valid, da[c]);
    if (p.d_metspec.ptr && c < mc) store_seg<kTexels>(p.d_metspec.ptr + plane_off(p.d_metspec, w.b, c, w.row, w.col0), w.vec, w.valid, dm[c]);
  }
}

struct BlendBwdKParams {
  PbrBlendDesc d;
  PbrBlendGrads g;
  int vec_ok;
  int need_dmask;      // some gradient w.r.t. the mask (or the maps it was built from) is wanted
  float width_eps;     // blend_width + 1e-6 (fp32)
};

__global__ void __launch_bounds__(kThreads) blend_bwd_kernel(const __grid_constant__ BlendBwdKParams p) {
  const PbrBlendDesc& d = p.d;
  const PbrBlendGrads& g = p.g;
  const Where w = locate(d.H, d.W, p.vec_ok != 0);
  if (!w.active) return;
  float mask[kTexels], dmask[kTexels];
  load_seg<kTexels>(g.mask.ptr + plane_off(g.mask, w.b, 0, w.row, w.col0), w.vec, w.valid, mask);
  load_or_zero<kTexels>(g.g_mask_out, w, 0, dmask);
  for (int m = 0; m < d.n_maps; ++m) {
    const PbrBlendMap& bm = d.maps[m];
    const PbrBlendGradMap& gm = g.maps[m];
    if (!gm.g_out.ptr) continue;   // zero gradient: contributes nothing (the host zero-fills d_a / d_b it still wants)
    if (bm.is_normal) {
      float a[3][kTexels], b[3][kTexels], go[3][kTexels], da[3][kTexels], db[3][kTexels];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        load_seg<kTexels>(bm.a.ptr + plane_off(bm.a, w.b, c, w.row, w.col0), w.vec, w.valid, a[c]);
        load_seg<kTexels>(bm.b.ptr + plane_off(bm.b, w.b, c, w.row, w.col0), w.vec, w.valid, b[c]);
        load_seg<kTexels>(gm.g_out.ptr + plane_off(gm.g_out, w.b, c, w.row, w.col0), w.vec, w.valid, go[c]);
      }
#pragma unroll
      for (int i = 0; i < kTexels; ++i) {
        const float a3[3] = {a[0][i], a[1][i], a[2][i]}, b3[3] = {b[0][i], b[1][i], b[2][i]}, g3[3] = {go[0][i], go[1][i], go[2][i]};
        float da3[3], db3[3];
        dmask[i] += blend_normal_bwd(mask[i], a3, b3, g3, da3, db3);
#pragma unroll
        for (int c = 0; c < 3; ++c) { da[c][i] = da3[c]; db[c][i] = db3[c]; }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (gm.d_a.ptr) store_seg<kTexels>(gm.d_a.ptr + plane_off(gm.d_a, w.b, c, w.row, w.col0), w.vec, w.valid, da[c]);
        if (gm.d_b.ptr) store_seg<kTexels>(gm.d_b.ptr + plane_off(gm.d_b, w.b, c, w.row, w.col0), w.vec, w.valid, db[c]);
      }
    } else {
      for (int c = 0; c < bm.channels; ++c) {
        float go[kTexels], o[kTexels];
        load_seg<kTexels>(gm.g_out.ptr + plane_off(gm.g_out, w.b, c, w.row, w.col0), w.vec, w.valid, go);
        if (p.need_dmask) {
          float a[kTexels], b[kTexels];
          load_seg<kTexels>(bm.a.ptr + plane_off(bm.a, w.b, c, w.row, w.col0), w.vec, w.valid, a);
          load_seg<kTexels>(bm.b.ptr + plane_off(bm.b, w.b, c, w.row, w.col0), w.vec, w.valid, b);
#pragma unroll
          for (int i = 0; i < kTexels; ++i) dmask[i] += go[i] * (a[i] - b[i]);
        }
        if (gm.d_a.ptr) {
#pragma unroll
          for (int i = 0; i < kTexels; ++i) o[i] = mask[i] * go[i];
          store_seg<kTexels>(gm.d_a.ptr + plane_off(gm.d_a, w.b, c, w.row, w.col0), w.vec, w.valid, o);
        }
        if (gm.d_b.ptr) {
#pragma unroll
          for (int i = 0; i < kTexels; ++i) o[i] = (1.0f - mask[i]) * go[i];
          store_seg<kTexels>(gm.d_b.ptr + plane_off(gm.d_b, w.b, c, w.row, w.col0), w.vec, w.valid, o);
        }
      }
    }
  }
  if (d.mask_mode == PBR_MASK_GIVEN) {
    if (g.d_mask.ptr) store_seg<kTexels>(g.d_mask.ptr + plane_off(g.d_mask, w.b, 0, w.row, w.col0), w.vec, w.valid, dmask);
  } else if (d.mask_mode == PBR_MASK_SIGMOID) {
    // mask = sigmoid(x), x = ((p1 [+ shift]) - p2) / (blend_width + 1e-6): d mask / d p1 = mask (1 - mask) / (width + 1e-6)
    float t[kTexels];
#pragma unroll
    for (int i = 0; i < kTexels; ++i) t[i] = dmask[i] * mask[i] * (1.0f - mask[i]) / p.width_eps;
    if (g.d_prop1.ptr) store_seg<kTexels>(g.d_prop1.ptr + plane_off(g.d_prop1, w.b, 0, w.row, w.col0), w.vec, w.valid, t);
    if (g.d_prop2.ptr) {
#pragma unroll
      for (int i = 0; i < kTexels; ++i) t[i] = -t[i];
      store_seg<kTexels>(g.d_prop2.ptr + plane_off(g.d_prop2, w.b, 0, w.row, w.col0), w.vec, w.valid, t);
    }
  }
}

__global__ void __launch_bounds__(kThreads) normal_ingest_bwd_kernel(const __grid_constant__ NormalKParams p) {
  const Where w = locate(p.H, p.W, p.vec_ok != 0);
  if (!w.active) return;
  float v[3][kTexels], go[3][kTexels], o[3][kTexels];
  for (int c = 0; c < p.channels; ++c) load_seg<kTexels>(p.in.ptr + plane_off(p.in, w.b, c, w.row, w.col0), w.vec, w.valid, v[c]);
#pragma unroll
  for (int c = 0; c < 3; ++c) load_seg<kTexels>(p.g_out.ptr + plane_off(p.g_out, w.b, c, w.row, w.col0), w.vec, w.valid, go[c]);
#pragma unroll
  for (int i = 0; i < kTexels; ++i) {
    const float g3[3] = {go[0][i], go[1][i], go[2][i]};
    if (p.channels == 3) {
      const float v3[3] = {v[0][i], v[1][i], v[2][i]};
      float d3[3];
      ingest_normal3_bwd(v3, g3, d3);
      o[0][i] = d3[0]; o[1][i] = d3[1]; o[2][i] = d3[2];
    } else {
      const float v2[2] = {v[0][i], v[1][i]};
      float d2[2];
      ingest_normal2_bwd(v2, g3, d2);
      o[0][i] = d2[0]; o[1][i] = d2[1];
    }
  }
  for (int c = 0; c < p.channels; ++c) store_seg<kTexels>(p.d_in.ptr + plane_off(p.d_in, w.b, c, w.row, w.col0), w.vec, w.valid, o[c]);
}

}  // namespace pbr
