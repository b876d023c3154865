// pbr_aux_kernels.cuh — the streaming kernels either side of the shading path:
//   ingest_kernel          : 8/16-bit interleaved image -> float32 planar map, with the normal-map remap
//                            (the step BEFORE the path: MaterialBase._to_tensor + _process_normal_map)
//   index_transform_kernel : flip / roll / tile / crop of every map of a material in one pass, with the
//                            normal's sign flips (MaterialBase.flip_horizontal ... crop)
//   adam_kernel            : Adam update of every parameter map of the fit + projection onto the valid
//                            range (the step AFTER the path: the optimiser of the inverse-rendering fit)
// All HBM-bound: one thread owns kTexels consecutive texels of a row, 128-bit accesses where the
// layout allows.  Included by pbr_kernels.cu (uses its Where / load_seg / store_seg helpers).
#pragma once

namespace pbr {

// ------------------------------------------------------------------------------------------------
// ingestion: pypbr/materials/base.py:122-168 (TF.to_tensor: /255; 16-bit: /65535) + :191-242
// ------------------------------------------------------------------------------------------------
struct IngestKParams {
  PbrIngestDesc d;
  int vec_ok;
  int src_words_ok;   // source base and strides are 4-byte aligned: a thread's 4 pixels are read as whole words
};

// CS interleaved source channels of BITS bits: the 4 pixels of a thread are 4*CS*BITS/8 contiguous bytes, i.e.
// CS*BITS/8 32-bit words when the row is word aligned; otherwise (and at a ragged edge) element loads.
template <int CS, int BITS>
__global__ void __launch_bounds__(kThreads) ingest_kernel(const __grid_constant__ IngestKParams p) {
  const PbrIngestDesc& d = p.d;
  // x / 255 for every byte value, computed once per CTA with the IEEE division torch uses: exact by construction
  __shared__ float lut[256];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (BITS == 8)
    for (int i = tid; i < 256; i += kThreads) lut[i] = xdiv((float)i, 255.0f);
  __syncthreads();
  const Where w = locate(d.H, d.W, p.vec_ok != 0);
  if (!w.active) return;
  constexpr int ESZ = BITS / 8;
  constexpr int NW = CS * ESZ;   // words per 4 pixels
  const unsigned char* row = static_cast<const unsigned char*>(d.src) + (int64_t)w.b * d.src_batch_stride + (int64_t)w.row * d.src_row_stride;
  unsigned raw[kTexels][CS];
  if (p.src_words_ok && w.valid == kTexels && kTexels == 4) {
    unsigned words[NW];
    const unsigned* wp = reinterpret_cast<const unsigned*>(row + (int64_t)w.col0 * CS * ESZ);
#pragma unroll
    for (int k = 0; k < NW; ++k) words[k] = __ldcs(wp + k);
#pragma unroll
    for (int i = 0; i < kTexels; ++i)
#pragma unroll
      for (int c = 0; c < CS; ++c) {
        const int byte = (i * CS + c) * ESZ;
        raw[i][c] = BITS == 8 ? (words[byte >> 2] >> ((byte & 3) * 8)) & 0xffu : (words[byte >> 2] >> ((byte & 3) * 8)) & 0xffffu;
      }
  } else {
#pragma unroll
    for (int i = 0; i < kTexels; ++i) {
      const int col = w.col0 + (i < w.valid ? i : w.valid - 1);
#pragma unroll
      for (int c = 0; c < CS; ++c)
        raw[i][c] = BITS == 8 ? row[(int64_t)col * CS + c] : reinterpret_cast<const unsigned short*>(row)[(int64_t)col * CS + c];
    }
  }
  float v[CS][kTexels];
#pragma unroll
  for (int c = 0; c < CS; ++c)
#pragma unroll
    for (int i = 0; i < kTexels; ++i) v[c][i] = BITS == 8 ? lut[raw[i][c]] : xdiv((float)raw[i][c], 65535.0f);
  if (d.mode == PBR_INGEST_PLAIN) {
#pragma unroll
    for (int c = 0; c < CS; ++c)
      if (c < d.channels) store_seg<kTexels>(d.out.ptr + plane_off(d.out, w.b, c, w.row, w.col0), w.vec, w.valid, v[c]);
    return;
  }
  float o[3][kTexels];
#pragma unroll
  for (int i = 0; i < kTexels; ++i) {
    float o3[3];
    if (d.mode == PBR_INGEST_NORMAL3) {
      const float v3[3] = {v[0][i], v[CS > 1 ? 1 : 0][i], v[CS > 2 ? 2 : 0][i]};
      ingest_normal3(v3, o3);
    } else {
      const float v2[2] = {v[0][i], v[CS > 1 ? 1 : 0][i]};
      ingest_normal2(v2, o3);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c][i] = o3[c];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) store_seg<kTexels>(d.out.ptr + plane_off(d.out, w.b, c, w.row, w.col0), w.vec, w.valid, o[c]);
}

// ------------------------------------------------------------------------------------------------
// index transforms: pypbr/materials/base.py:490-537 (crop, tile), :605-655 (flips, roll)
//   out[c, y, x] = sign_c * in[c, fy(y), fx(x)],  f(i) = origin + step * i, wrapped into [0, n_in) when `wrap`,
//   else 0 where it falls outside (TF.crop pads with zeros)
// ------------------------------------------------------------------------------------------------
struct IndexKParams {
  PbrIndexDesc d;
  int vec_ok;
};

__device__ __forceinline__ int wrap_index(int i, int n) {
  i %= n;
  return i < 0 ? i + n : i;
}

__global__ void __launch_bounds__(kThreads) index_transform_kernel(const __grid_constant__ IndexKParams p) {
  const PbrIndexDesc& d = p.d;
  const Where w = locate(d.H_out, d.W_out, p.vec_ok != 0);
  if (!w.active) return;
  const int ry = d.reduce_y > 1 ? d.reduce_y : 1, rx = d.reduce_x > 1 ? d.reduce_x : 1;
  if (ry * rx > 1) {
    // adjoint of a tile: every output texel sums the ry x rx source texels that were copies of it
    for (int m = 0; m < d.n_maps; ++m) {
      const PbrIndexMap& im = d.maps[m];
      for (int c = 0; c < im.channels; ++c) {
        const bool neg = (im.negate_mask >> c) & 1;
        float v[kTexels];
#pragma unroll
        for (int i = 0; i < kTexels; ++i) v[i] = 0.0f;
        for (int ty = 0; ty < ry; ++ty) {
          int sy = d.origin_y + d.step_y * (w.row + ty * d.H_out);
          bool row_in = true;
          if (d.wrap) sy = wrap_index(sy, d.H_in);
          else row_in = sy >= 0 && sy < d.H_in;
          if (!row_in) continue;
          const float* src = im.in.ptr + plane_off(im.in, w.b, c, sy, 0);
          for (int tx = 0; tx < rx; ++tx) {
#pragma unroll
            for (int i = 0; i < kTexels; ++i) {
              int x = d.origin_x + d.step_x * (w.col0 + i + tx * d.W_out);
              bool in = true;
              if (d.wrap) x = wrap_index(x, d.W_in);
              else in = x >= 0 && x < d.W_in;
              if (i < w.valid && in) v[i] += __ldg(src + x);
            }
          }
        }
        if (neg) {
#pragma unroll
          for (int i = 0; i < kTexels; ++i) v[i] = -v[i];
        }
        store_seg<kTexels>(im.out.ptr + plane_off(im.out, w.b, c, w.row, w.col0), w.vec, w.valid, v);
      }
    }
    return;
  }
  int sy = d.origin_y + d.step_y * w.row;
  bool row_in = true;
  if (d.wrap) sy = wrap_index(sy, d.H_in);
  else row_in = sy >= 0 && sy < d.H_in;
  int sx[kTexels];
  bool in[kTexels];
#pragma unroll
  for (int i = 0; i < kTexels; ++i) {
    int x = d.origin_x + d.step_x * (w.col0 + i);
    if (d.wrap) { x = wrap_index(x, d.W_in); in[i] = true; }
    else in[i] = row_in && x >= 0 && x < d.W_in;
    sx[i] = in[i] ? x : 0;
  }
  if (!row_in) sy = 0;
  for (int m = 0; m < d.n_maps; ++m) {
    const PbrIndexMap& im = d.maps[m];
    for (int c = 0; c < im.channels; ++c) {
      const float* src = im.in.ptr + plane_off(im.in, w.b, c, sy, 0);
      const bool neg = (im.negate_mask >> c) & 1;
      float v[kTexels];
#pragma unroll
      for (int i = 0; i < kTexels; ++i) {
        float t = (i < w.valid && in[i]) ? __ldg(src + sx[i]) : 0.0f;
        v[i] = neg ? -t : t;
      }
      store_seg<kTexels>(im.out.ptr + plane_off(im.out, w.b, c, w.row, w.col0), w.vec, w.valid, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Adam + projection (no reference code: the tutorial stops at "perform backpropagation and
// optimization steps here", docs/source/tutorials/06_advanced.rst:136-137).  Same update as
// torch.optim.Adam (single-tensor path): m = lerp(m, g, 1-b1); v = v*b2 + (1-b2)*g*g;
// p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps); then clamp to [lo, hi] or renormalise.
// ------------------------------------------------------------------------------------------------
struct AdamKParams {
  PbrAdamDesc d;
  int vec_ok;
};

__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, const PbrAdamDesc& d) {
  const AdamCoef a{d.step_size, d.one_minus_beta1, d.beta2, d.one_minus_beta2, d.bias2_sqrt, d.eps};
  return adam_update(p, g, m, v, a);   // defined next to the fused fit epilogue in pbr_kernels.cu
}

__global__ void __launch_bounds__(kThreads) adam_kernel(const __grid_constant__ AdamKParams p) {
  const PbrAdamDesc& d = p.d;
  const Where w = locate(d.H, d.W, p.vec_ok != 0);
  if (!w.active) return;
  for (int mi = 0; mi < d.n_maps; ++mi) {
    const PbrAdamMap& am = d.maps[mi];
    if (am.project == PBR_PROJECT_NORMALIZE) {
      float pn[3][kTexels];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float pv[kTexels], g[kTexels], m[kTexels], v[kTexels];
        const int64_t o = plane_off(am.param, w.b, c, w.row, w.col0);
        load_seg<kTexels>(am.param.ptr + o, w.vec, w.valid, pv);
        load_seg<kTexels>(am.grad.ptr + plane_off(am.grad, w.b, c, w.row, w.col0), w.vec, w.valid, g);
        load_seg<kTexels>(am.exp_avg.ptr + plane_off(am.exp_avg, w.b, c, w.row, w.col0), w.vec, w.valid, m);
        load_seg<kTexels>(am.exp_avg_sq.ptr + plane_off(am.exp_avg_sq, w.b, c, w.row, w.col0), w.vec, w.valid, v);
#pragma unroll
        for (int i = 0; i < kTexels; ++i) pn[c][i] = adam_update(pv[i], g[i] * d.grad_scale, m[i], v[i], d);
        store_seg<kTexels>(am.exp_avg.ptr + plane_off(am.exp_avg, w.b, c, w.row, w.col0), w.vec, w.valid, m);
        store_seg<kTexels>(am.exp_avg_sq.ptr + plane_off(am.exp_avg_sq, w.b, c, w.row, w.col0), w.vec, w.valid, v);
      }
#pragma unroll
      for (int i = 0; i < kTexels; ++i) {
        float o3[3];
        normalize3(pn[0][i], pn[1][i], pn[2][i], o3);   // F.normalize(dim=channel), eps 1e-12
        pn[0][i] = o3[0]; pn[1][i] = o3[1]; pn[2][i] = o3[2];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) store_seg<kTexels>(am.param.ptr + plane_off(am.param, w.b, c, w.row, w.col0), w.vec, w.valid, pn[c]);
    } else {
      for (int c = 0; c < am.channels; ++c) {
        float pv[kTexels], g[kTexels], m[kTexels], v[kTexels];
        load_seg<kTexels>(am.param.ptr + plane_off(am.param, w.b, c, w.row, w.col0), w.vec, w.valid, pv);
        load_seg<kTexels>(am.grad.ptr + plane_off(am.grad, w.b, c, w.row, w.col0), w.vec, w.valid, g);
        load_seg<kTexels>(am.exp_avg.ptr + plane_off(am.exp_avg, w.b, c, w.row, w.col0), w.vec, w.valid, m);
        load_seg<kTexels>(am.exp_avg_sq.ptr + plane_off(am.exp_avg_sq, w.b, c, w.row, w.col0), w.vec, w.valid, v);
#pragma unroll
        for (int i = 0; i < kTexels; ++i) {
          float np = adam_update(pv[i], g[i] * d.grad_scale, m[i], v[i], d);
          if (am.project == PBR_PROJECT_CLAMP) np = fminf(fmaxf(np, am.lo), am.hi);
          pv[i] = np;
        }
        store_seg<kTexels>(am.param.ptr + plane_off(am.param, w.b, c, w.row, w.col0), w.vec, w.valid, pv);
        store_seg<kTexels>(am.exp_avg.ptr + plane_off(am.exp_avg, w.b, c, w.row, w.col0), w.vec, w.valid, m);
        store_seg<kTexels>(am.exp_avg_sq.ptr + plane_off(am.exp_avg_sq, w.b, c, w.row, w.col0), w.vec, w.valid, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// per-texel normal utilities of an augmentation pipeline
//   ROTATE     : utils/functions.py:69-108 rotate_normals - (x, y) @ R(angle)^T, then F.normalize over the channels
//   FROM_HEIGHT: utils/functions.py:123-177 compute_normal_from_height - one-sided differences of the zero-padded
//                height map, (-gx*scale, -+gy*scale, 1), F.normalize
//   DIVERGENCE : utils/functions.py:211-283, the per-texel half of compute_height_from_normal - the gradient field
//                g = (-Nx, -+Ny) / (Nz + 1e-8) * scale and its forward-difference divergence with replicate padding
//                (the Poisson solve that follows is two cuFFT calls on the host side)
// ------------------------------------------------------------------------------------------------
struct NormalOpKParams {
  PbrNormalOpDesc d;
  int vec_ok;
};

__global__ void __launch_bounds__(kThreads) normal_op_kernel(const __grid_constant__ NormalOpKParams p) {
  const PbrNormalOpDesc& d = p.d;
  const Where w = locate(d.H, d.W, p.vec_ok != 0);
  if (!w.active) return;
  float o[3][kTexels];
  if (d.op == PBR_NORMAL_OP_DIVERGENCE) {
    // div[y][x] = (gx[y][x+1] - gx[y][x]) + (gy[y+1][x] - gy[y][x]); the replicated last column / row differences are 0
    const float* nx = d.in.ptr + plane_off(d.in, w.b, 0, w.row, 0);
    const float* ny = d.in.ptr + plane_off(d.in, w.b, 1, w.row, 0);
    const float* nz = d.in.ptr + plane_off(d.in, w.b, 2, w.row, 0);
    const bool has_dn = w.row + 1 < d.H;
    float gx[kTexels + 1], div[kTexels];
#pragma unroll
    for (int i = 0; i < kTexels + 1; ++i) {
      const int x = min(w.col0 + i, d.W - 1);
      gx[i] = xmul(xdiv(-__ldg(nx + x), xadd(__ldg(nz + x), 1e-8f)), d.scale);
    }
#pragma unroll
    for (int i = 0; i < kTexels; ++i) {
      const int x = min(w.col0 + i, d.W - 1);
      const float y0 = __ldg(ny + x), z0 = xadd(__ldg(nz + x), 1e-8f);
      const float y1 = has_dn ? __ldg(ny + d.in.sh + x) : y0, z1 = has_dn ? xadd(__ldg(nz + d.in.sh + x), 1e-8f) : z0;
      const float gy0 = xmul(xdiv(d.flip_y ? y0 : -y0, z0), d.scale), gy1 = xmul(xdiv(d.flip_y ? y1 : -y1, z1), d.scale);
      div[i] = xadd(xsub(gx[i + 1], gx[i]), xsub(gy1, gy0));
    }
    store_seg<kTexels>(d.out.ptr + plane_off(d.out, w.b, 0, w.row, w.col0), w.vec, w.valid, div);
    return;
  }
  if (d.op == PBR_NORMAL_OP_FROM_HEIGHT_BWD) {
    // d_h[y][x] = scale * (a[y][x+1] - a[y][x-1] + b[y+1][x] - b[y-1][x]); a = -g_u.x, b = (flip_y ? +1 : -1) * g_u.y at the
    // neighbour, g_u = the neighbour's normal gradient pushed through ITS normalisation (recomputed from the height map)
    const float* hbase = d.in.ptr + plane_off(d.in, w.b, 0, 0, 0);
    auto hat = [&](int y, int x) -> float { return (y >= 0 && y < d.H && x >= 0 && x < d.W) ? __ldg(hbase + (int64_t)y * d.in.sh + x) : 0.0f; };
    // (a, b) of texel (y, x); zero outside the image
    auto ab = [&](int y, int x, float& a, float& b) {
      a = 0.0f; b = 0.0f;
      if (y < 0 || y >= d.H || x < 0 || x >= d.W) return;
      const float gx = (hat(y, x - 1) - hat(y, x + 1)) * d.scale, gy = (hat(y - 1, x) - hat(y + 1, x)) * d.scale;
      const float u[3] = {-gx, d.flip_y ? gy : -gy, 1.0f};
      float g[3], gu[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) g[c] = __ldg(d.aux.ptr + plane_off(d.aux, w.b, c, y, x));
      normalize3_bwd(u, g, gu);
      a = -gu[0];
      b = d.flip_y ? gu[1] : -gu[1];
    };
    float dh[kTexels];
    float a_prev, a_cur, b_dummy;
    ab(w.row, w.col0 - 1, a_prev, b_dummy);
    ab(w.row, w.col0, a_cur, b_dummy);
#pragma unroll
    for (int i = 0; i < kTexels; ++i) {
      const int x = w.col0 + i;
      float a_next, b_up, b_dn, t;
      ab(w.row, x + 1, a_next, t);
      ab(w.row - 1, x, t, b_up);
      ab(w.row + 1, x, t, b_dn);
      dh[i] = d.scale * ((a_next - a_prev) + (b_dn - b_up));
      a_prev = a_cur;
      a_cur = a_next;
    }
    store_seg<kTexels>(d.out.ptr + plane_off(d.out, w.b, 0, w.row, w.col0), w.vec, w.valid, dh);
    return;
  }
  if (d.op == PBR_NORMAL_OP_ROTATE_BWD) {
    // y = normalize(R (x, y), z): d_in = R^T-applied gradient pushed through the normalisation (recomputed from `in`, the map
    // BEFORE the rotation).  cos_a = strength, sin_a = 0 is the adjoint of adjust_normal_strength.
    float v[3][kTexels], g[3][kTexels];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      load_seg<kTexels>(d.in.ptr + plane_off(d.in, w.b, c, w.row, w.col0), w.vec, w.valid, v[c]);
      load_seg<kTexels>(d.aux.ptr + plane_off(d.aux, w.b, c, w.row, w.col0), w.vec, w.valid, g[c]);
    }
#pragma unroll
    for (int i = 0; i < kTexels; ++i) {
      const float u[3] = {xfma(v[1][i], -d.sin_a, xmul(v[0][i], d.cos_a)), xfma(v[1][i], d.cos_a, xmul(v[0][i], d.sin_a)), v[2][i]};
      const float gi[3] = {g[0][i], g[1][i], g[2][i]};
      float gu[3];
      normalize3_bwd(u, gi, gu);
      o[0][i] = gu[0] * d.cos_a + gu[1] * d.sin_a;
      o[1][i] = gu[1] * d.cos_a - gu[0] * d.sin_a;
      o[2][i] = gu[2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) store_seg<kTexels>(d.out.ptr + plane_off(d.out, w.b, c, w.row, w.col0), w.vec, w.valid, o[c]);
    return;
  }
  if (d.op == PBR_NORMAL_OP_DIVERGENCE_BWD) {
    // div[y][x] = gx[y][min(x+1, W-1)] - gx[y][x] + gy[min(y+1, H-1)][x] - gy[y][x] with gx = -Nx r s, gy = -+Ny r s, r = 1/(Nz + 1e-8):
    //   A = d/d gx[y][x] = G[y][x-1] (x >= 1) + G[y][W-1] (x == W-1) - G[y][x],  B = d/d gy[y][x] likewise along y  (G = `aux`)
    //   d_Nx = -r s A,  d_Ny = sy r s B (sy = +1 DirectX, -1 OpenGL),  d_Nz = r^2 s (Nx A - sy Ny B)
    const float* G = d.aux.ptr + plane_off(d.aux, w.b, 0, 0, 0);
    const float sy = d.flip_y ? 1.0f : -1.0f;
    float n[3][kTexels];
#pragma unroll
    for (int c = 0; c < 3; ++c) load_seg<kTexels>(d.in.ptr + plane_off(d.in, w.b, c, w.row, w.col0), w.vec, w.valid, n[c]);
#pragma unroll
    for (int i = 0; i < kTexels; ++i) {
      const int x = min(w.col0 + i, d.W - 1), y = w.row;
      const float* Gr = G + (int64_t)y * d.aux.sh;
      const float g0 = __ldg(Gr + x);
      const float A = (x >= 1 ? __ldg(Gr + x - 1) : 0.0f) + (x == d.W - 1 ? g0 : 0.0f) - g0;
      const float B = (y >= 1 ? __ldg(Gr - d.aux.sh + x) : 0.0f) + (y == d.H - 1 ? g0 : 0.0f) - g0;
      const float r = 1.0f / (n[2][i] + 1e-8f);
      const float rs = r * d.scale;
      o[0][i] = -rs * A;
      o[1][i] = sy * rs * B;
      o[2][i] = r * rs * (n[0][i] * A - sy * n[1][i] * B);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) store_seg<kTexels>(d.out.ptr + plane_off(d.out, w.b, c, w.row, w.col0), w.vec, w.valid, o[c]);
    return;
  }
  if (d.op == PBR_NORMAL_OP_ROTATE) {
    float v[3][kTexels];
#pragma unroll
    for (int c = 0; c < 3; ++c) load_seg<kTexels>(d.in.ptr + plane_off(d.in, w.b, c, w.row, w.col0), w.vec, w.valid, v[c]);
#pragma unroll
    for (int i = 0; i < kTexels; ++i) {
      // row-vector times R^T with K = 2: a multiply and one fused multiply-add per output, as the GEMM computes it
      const float rx = xfma(v[1][i], -d.sin_a, xmul(v[0][i], d.cos_a));
      const float ry = xfma(v[1][i], d.cos_a, xmul(v[0][i], d.sin_a));
      float o3[3];
      normalize3(rx, ry, v[2][i], o3);
      o[0][i] = o3[0]; o[1][i] = o3[1]; o[2][i] = o3[2];
    }
  } else {
    // grad_x[x] = h[x-1] - h[x+1], grad_y[y] = h[y-1] - h[y+1], zero outside the image (F.pad)
    const float* base = d.in.ptr + plane_off(d.in, w.b, 0, 0, 0);
    const float* row = base + (int64_t)w.row * d.in.sh;
    float c[kTexels + 2], up[kTexels], dn[kTexels];
#pragma unroll
    for (int i = 0; i < kTexels + 2; ++i) {
      const int x = w.col0 + i - 1;
      c[i] = (x >= 0 && x < d.W) ? __ldg(row + x) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < kTexels; ++i) {
      const int x = w.col0 + i;
      const bool in = x < d.W;
      up[i] = (in && w.row > 0) ? __ldg(row - d.in.sh + x) : 0.0f;
      dn[i] = (in && w.row + 1 < d.H) ? __ldg(row + d.in.sh + x) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < kTexels; ++i) {
      const float gx = xmul(xsub(c[i], c[i + 2]), d.scale);
      const float gy = xmul(xsub(up[i], dn[i]), d.scale);
      float o3[3];
      normalize3(-gx, d.flip_y ? gy : -gy, 1.0f, o3);
      o[0][i] = o3[0]; o[1][i] = o3[1]; o[2][i] = o3[2];
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) store_seg<kTexels>(d.out.ptr + plane_off(d.out, w.b, c, w.row, w.col0), w.vec, w.valid, o[c]);
}

}  // namespace pbr
