// pbr_ct_stream.cuh — the streamed (TMA-fed) Cook-Torrance kernels: the fast path of
// pbr_ct_forward / pbr_ct_backward / pbr_ct_loss_fwd_bwd for ONE light (L == 1) on 16-byte aligned,
// W % 4 == 0 maps.  Included by pbr_kernels.cu after the generic kernels.
//
// Why: the shading math costs a few hundred instructions per texel, so a thread that issues its own
// global loads and then computes keeps too few bytes in flight at the occupancy the register budget
// allows (ncu: 24 % warps active, 64 % issue-slot use, 55 % of HBM peak for the generic backward).
// Here the copy engine does the loading:
//
//   tile    = blockDim.y rows x (blockDim.x * 4) texels of ONE material: every input plane of the tile
//             (albedo x3, roughness, metallic | specular, normal x3, and grad_out | target x3 for the
//             backward) is kTileFloats contiguous floats of shared memory per stage;
//   producer= warp 0: one cp.async.bulk per (plane, row) -> stage buffer, completion counted in bytes
//             on the stage's mbarrier (pbr_async.cuh);
//   consumer= all threads: wait for the stage, pull their 4 texels of every plane into registers with
//             LDS.128 at immediate offsets, one arrival per warp on the stage's "empty" mbarrier, shade
//             from registers, write results with STG.128; warp 0 then waits for "empty" (already
//             complete in practice) and refills the stage with the tile that is kStages ahead.
//             There is no CTA-wide barrier inside the loop.
//
// A CTA walks over `mats_per_cta` materials at a fixed image position, so for a point light the
// per-texel light geometry (material independent) is computed once and reused - as in the generic
// kernels - and the pipeline stays full for the whole walk.
#pragma once

#include "pbr_async.cuh"

namespace pbr {

#ifndef PBR_STREAM_STAGES
#define PBR_STREAM_STAGES 2
#endif
#ifndef PBR_STREAM_FWD_MIN_CTAS
#define PBR_STREAM_FWD_MIN_CTAS 2
#endif
#ifndef PBR_STREAM_BWD_MIN_CTAS
#define PBR_STREAM_BWD_MIN_CTAS 2
#endif
#ifndef PBR_STREAM_FWD_GROUP
#define PBR_STREAM_FWD_GROUP 4
#endif
#ifndef PBR_STREAM_BWD_GROUP
#define PBR_STREAM_BWD_GROUP 1
#endif

constexpr int kStages = PBR_STREAM_STAGES;
constexpr int kTileFloats = kThreads * 4;   // floats of one plane in one stage (4 KB at 256 threads)
constexpr int kMaxSrcPlanes = 13;           // albedo 3 + roughness 1 + metallic|specular 3 + normal 3 + grad_out|target 3

// backward flavours (compile-time, so the plain backward carries no reduction code)
constexpr int kModeLoss = 1;      // gsrc is the target image: grad_out = 2*scale*(render - target), loss reduced
constexpr int kModeIntGrad = 2;   // reduce d loss / d light intensity

struct StreamSrc {
  const float* base;   // plane pointer at (b = 0, channel, first row of the tile, first column of the tile)
  int64_t sb, sh;      // batch and row strides in elements
  int slot;            // plane slot inside a stage buffer
  int pad;
};

struct StreamShared {
  uint64_t full[kStages];    // producer -> consumers: the copies of the stage have landed (transaction bytes)
  uint64_t empty[kStages];   // consumers -> producer: one arrival per warp once its threads hold the stage in registers
  StreamSrc src[kMaxSrcPlanes];
  int n_src;
};

// plane slots of a stage: fixed at compile time so that the consumer's LDS use immediate offsets
template <int WF, bool kBwd>
struct Slots {
  static constexpr int mc = WF == 0 ? 1 : 3;
  static constexpr int albedo = 0, rough = 3, met = 4;
  static constexpr int gsrc = 4 + mc;                       // backward only
  static constexpr int normal = kBwd ? 7 + mc : 4 + mc;     // last, so a missing normal map shortens the stage
  static constexpr int count_no_normal = normal;
  static constexpr int count = normal + 3;
};

template <int WF, bool kBwd>
__device__ __forceinline__ void stream_build_table(const CtKParams& p, int row0, int col0, StreamShared& sh) {
  using SL = Slots<WF, kBwd>;
  int n = 0;
  auto add = [&](const PbrPlane& pl, int c, int slot) {
    sh.src[n].base = pl.ptr + (int64_t)c * pl.sc + (int64_t)row0 * pl.sh + col0;
    sh.src[n].sb = pl.sb;
    sh.src[n].sh = pl.sh;
    sh.src[n].slot = slot;
    ++n;
  };
  for (int c = 0; c < 3; ++c) add(p.albedo, c, SL::albedo + c);
  add(p.roughness, 0, SL::rough);
  for (int c = 0; c < SL::mc; ++c) add(p.metspec, c, SL::met + c);
  if (kBwd)
    for (int c = 0; c < 3; ++c) add(p.gsrc, c, SL::gsrc + c);
  if (p.normal.ptr)
    for (int c = 0; c < 3; ++c) add(p.normal, c, SL::normal + c);
  sh.n_src = n;
}

// warp 0: enqueue the copies of material `b` into `stage` (floats) and arm its barrier
__device__ __forceinline__ void stream_issue(const StreamShared& sh, uint64_t* bar, float* stage, int b, int rows_valid,
                                             int seg_bytes, int row_floats, uint64_t policy) {
  const int lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
  const int items = sh.n_src * rows_valid;
  if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(items * seg_bytes));
  __syncwarp();
  for (int it = lane; it < items; it += 32) {
    const int pl = it / rows_valid, r = it - pl * rows_valid;
    const StreamSrc& s = sh.src[pl];
    bulk_g2s(stage + s.slot * kTileFloats + r * row_floats, s.base + (int64_t)b * s.sb + (int64_t)r * s.sh,
             (uint32_t)seg_bytes, bar, policy);
  }
}

__device__ __forceinline__ void lds4(const float* p, float (&dst)[4]) {
  float4 v = *reinterpret_cast<const float4*>(p);
  dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
}
__device__ __forceinline__ void stg4(float* p, const float (&src)[4]) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(src[0], src[1], src[2], src[3]));
}
__device__ __forceinline__ float* plane_at(const PbrPlane& pl, int b, int c, int toff) {
  return pl.ptr + ((int64_t)b * pl.sb + (int64_t)c * pl.sc) + toff;
}

// Geometry of the tile this CTA owns and of the thread inside it.
struct StreamWhere {
  int row0, col0;       // tile origin
  int row, col;         // this thread's first texel
  int rows_valid, seg_bytes, row_floats;
  bool active;
};
__device__ __forceinline__ StreamWhere stream_locate(int H, int W) {
  StreamWhere w;
  w.row_floats = blockDim.x * 4;
  w.row0 = blockIdx.y * blockDim.y;
  w.col0 = blockIdx.x * w.row_floats;
  w.row = w.row0 + threadIdx.y;
  w.col = w.col0 + threadIdx.x * 4;
  w.rows_valid = min((int)blockDim.y, H - w.row0);
  w.seg_bytes = min(w.row_floats, W - w.col0) * 4;
  w.active = w.row < H && w.col < W;   // W % 4 == 0: a thread's 4 texels are all inside or all outside
  return w;
}

template <int kLight>
__device__ __forceinline__ void stream_coords(const CtStage& S, const StreamWhere& w, float (&x)[4], float& y,
                                              LightGeom (&hg)[4]) {
  const int row = w.active ? w.row : 0, col = w.active ? w.col : 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = linspace_at(S.lsx, col + i);
  y = linspace_at(S.lsy, row);
  if (kLight == kLightPointHoisted) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      point_light_geom(S.light[0].p[0], S.light[0].p[1], S.light[0].p[2], x[i], y, S.vx, S.vy, S.vz, hg[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int WF, int kLight>
__global__ void __launch_bounds__(kThreads, PBR_STREAM_FWD_MIN_CTAS) ct_forward_stream(const __grid_constant__ CtKParams p) {
  using SL = Slots<WF, false>;
  constexpr int G = PBR_STREAM_FWD_GROUP;
  extern __shared__ __align__(128) float stream_smem[];
  __shared__ CtStage S;
  __shared__ StreamShared sh;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const StreamWhere w = stream_locate(p.H, p.W);
  if (tid == 0) {
    stream_build_table<WF, false>(p, w.row0, w.col0, sh);
#pragma unroll
    for (int s = 0; s < kStages; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], kThreads / 32); }
    mbar_init_fence();
  }
  stage_params(p, S);  // ends with __syncthreads()

  const int b0 = blockIdx.z * p.mats_per_cta;
  const int ntiles = min(b0 + p.mats_per_cta, p.B) - b0;
  const int stage_floats = (p.normal.ptr ? SL::count : SL::count_no_normal) * kTileFloats;
  uint64_t policy = 0;
  if (tid < 32) {
    policy = l2_evict_first_policy();
    for (int k = 0; k < kStages && k < ntiles; ++k)
      stream_issue(sh, &sh.full[k], stream_smem + k * stage_floats, b0 + k, w.rows_valid, w.seg_bytes, w.row_floats, policy);
  }

  float x[4], y;
  LightGeom hg[4];
  stream_coords<kLight>(S, w, x, y, hg);
  const int toff = w.row * (int)p.out.sh + w.col;
  const bool has_normal = p.normal.ptr != nullptr;

  for (int k = 0; k < ntiles; ++k) {
    const int s = k % kStages;
    const float* st = stream_smem + s * stage_floats + tid * 4;
    mbar_wait(&sh.full[s], (k / kStages) & 1);
    float araw[3][4], nraw[3][4], rough[4], mraw[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) lds4(st + (SL::albedo + c) * kTileFloats, araw[c]);
    lds4(st + SL::rough * kTileFloats, rough);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c < SL::mc) lds4(st + (SL::met + c) * kTileFloats, mraw[c]);
      else mraw[c][0] = mraw[c][1] = mraw[c][2] = mraw[c][3] = 0.0f;
    }
    if (has_normal) {
#pragma unroll
      for (int c = 0; c < 3; ++c) lds4(st + (SL::normal + c) * kTileFloats, nraw[c]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) { nraw[0][i] = 0.0f; nraw[1][i] = 0.0f; nraw[2][i] = 1.0f; }
    }
    // this warp holds its texels in registers: tell the producer (no CTA-wide barrier - warps drift freely)
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&sh.empty[s]);
    if (w.active) {

    const int b = b0 + k;
    float outv[3][4];
#pragma unroll
    for (int g0 = 0; g0 < 4; g0 += G) {
      float a[3][G], n[3][G], r[G], m[3][G], xs[G];
      LightGeom hgs[G];
      slice3<G>(araw, g0, a); slice3<G>(nraw, g0, n); slice3<G>(mraw, g0, m);
#pragma unroll
      for (int i = 0; i < G; ++i) { r[i] = rough[g0 + i]; xs[i] = x[g0 + i]; hgs[i] = hg[g0 + i]; }
      auto emit = [&](int, const float(&v)[3][G]) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int i = 0; i < G; ++i) outv[c][g0 + i] = v[c][i];
      };
      ct_forward_group<WF, kLight, G>(S, p.flags, a, n, r, m, xs, y, hgs, emit);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) stg4(plane_at(p.out, b, c, toff), outv[c]);
    }
    // producer: by the time warp 0 has shaded tile k every warp has long since read stage s
    if (tid < 32 && k + kStages < ntiles) {
      mbar_wait(&sh.empty[s], (k / kStages) & 1);
      stream_issue(sh, &sh.full[s], stream_smem + s * stage_floats, b0 + k + kStages, w.rows_valid, w.seg_bytes,
                   w.row_floats, policy);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward / fused loss
// ------------------------------------------------------------------------------------------------
template <int WF, int kLight, int kMode>
__global__ void __launch_bounds__(kThreads, PBR_STREAM_BWD_MIN_CTAS) ct_backward_stream(const __grid_constant__ CtKParams p) {
  using SL = Slots<WF, true>;
  constexpr int G = PBR_STREAM_BWD_GROUP;
  constexpr bool kLoss = (kMode & kModeLoss) != 0;
  constexpr bool kIntGrad = (kMode & kModeIntGrad) != 0;
  extern __shared__ __align__(128) float stream_smem[];
  __shared__ CtStage S;
  __shared__ StreamShared sh;
  __shared__ float s_red[kThreads / 32][4];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const StreamWhere w = stream_locate(p.H, p.W);
  if (tid == 0) {
    stream_build_table<WF, true>(p, w.row0, w.col0, sh);
#pragma unroll
    for (int s = 0; s < kStages; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], kThreads / 32); }
    mbar_init_fence();
  }
  stage_params(p, S);  // ends with __syncthreads()

  const int b0 = blockIdx.z * p.mats_per_cta;
  const int ntiles = min(b0 + p.mats_per_cta, p.B) - b0;
  const int stage_floats = (p.normal.ptr ? SL::count : SL::count_no_normal) * kTileFloats;
  uint64_t policy = 0;
  if (tid < 32) {
    policy = l2_evict_first_policy();
    for (int k = 0; k < kStages && k < ntiles; ++k)
      stream_issue(sh, &sh.full[k], stream_smem + k * stage_floats, b0 + k, w.rows_valid, w.seg_bytes, w.row_floats, policy);
  }

  float x[4], y;
  LightGeom hg[4];
  stream_coords<kLight>(S, w, x, y, hg);
  const int off_a = w.row * (int)p.d_albedo.sh + w.col;
  const int off_n = w.row * (int)p.d_normal.sh + w.col;
  const int off_r = w.row * (int)p.d_roughness.sh + w.col;
  const int off_m = w.row * (int)p.d_metspec.sh + w.col;
  const bool has_normal = p.normal.ptr != nullptr;
  float loss_local = 0.0f, gi_local[3] = {0.0f, 0.0f, 0.0f};

  for (int k = 0; k < ntiles; ++k) {
    const int s = k % kStages;
    const float* st = stream_smem + s * stage_floats + tid * 4;
    mbar_wait(&sh.full[s], (k / kStages) & 1);
    float araw[3][4], nraw[3][4], rough[4], mraw[3][4], gs[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) lds4(st + (SL::albedo + c) * kTileFloats, araw[c]);
    lds4(st + SL::rough * kTileFloats, rough);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c < SL::mc) lds4(st + (SL::met + c) * kTileFloats, mraw[c]);
      else mraw[c][0] = mraw[c][1] = mraw[c][2] = mraw[c][3] = 0.0f;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) lds4(st + (SL::gsrc + c) * kTileFloats, gs[c]);
    if (has_normal) {
#pragma unroll
      for (int c = 0; c < 3; ++c) lds4(st + (SL::normal + c) * kTileFloats, nraw[c]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) { nraw[0][i] = 0.0f; nraw[1][i] = 0.0f; nraw[2][i] = 1.0f; }
    }
    // this warp holds its texels in registers: tell the producer (no CTA-wide barrier - warps drift freely)
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&sh.empty[s]);
    if (w.active) {

    const int b = b0 + k;
    float d_albedo[3][4], d_normal[3][4], d_rough[4], d_met[3][4];
#pragma unroll
    for (int g0 = 0; g0 < 4; g0 += G) {
      float a[3][G], n[3][G], r[G], m[3][G], xs[G];
      LightGeom hgs[G];
      slice3<G>(araw, g0, a); slice3<G>(nraw, g0, n); slice3<G>(mraw, g0, m);
#pragma unroll
      for (int i = 0; i < G; ++i) { r[i] = rough[g0 + i]; xs[i] = x[g0 + i]; hgs[i] = hg[g0 + i]; }
      auto gout = [&](int, const float(&outv)[3][G], float(&g)[3][G]) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int i = 0; i < G; ++i) {
            if (kLoss) {
              float diff = outv[c][i] - gs[c][g0 + i];
              loss_local += diff * diff;
              g[c][i] = 2.0f * p.loss_scale * diff;
            } else {
              g[c][i] = gs[c][g0 + i];
            }
          }
      };
      auto int_sink = [&](int, const float(&gi)[3]) {
        if (kIntGrad) {
#pragma unroll
          for (int c = 0; c < 3; ++c) gi_local[c] += gi[c];
        }
      };
      float da[3][G], dn[3][G], dr[G], dm[3][G];
      ct_backward_group<WF, kLight, G>(S, p.flags, a, n, r, m, xs, y, hgs, gout, int_sink, da, dn, dr, dm);
#pragma unroll
      for (int i = 0; i < G; ++i) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { d_albedo[c][g0 + i] = da[c][i]; d_normal[c][g0 + i] = dn[c][i]; d_met[c][g0 + i] = dm[c][i]; }
        d_rough[g0 + i] = dr[i];
      }
    }
    if (p.d_albedo.ptr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) stg4(plane_at(p.d_albedo, b, c, off_a), d_albedo[c]);
    }
    if (has_normal && p.d_normal.ptr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) stg4(plane_at(p.d_normal, b, c, off_n), d_normal[c]);
    }
    if (p.d_roughness.ptr) stg4(plane_at(p.d_roughness, b, 0, off_r), d_rough);
    if (p.d_metspec.ptr) {
#pragma unroll
      for (int c = 0; c < SL::mc; ++c) stg4(plane_at(p.d_metspec, b, c, off_m), d_met[c]);
    }
    }
    // producer: by the time warp 0 has shaded tile k every warp has long since read stage s
    if (tid < 32 && k + kStages < ntiles) {
      mbar_wait(&sh.empty[s], (k / kStages) & 1);
      stream_issue(sh, &sh.full[s], stream_smem + s * stage_floats, b0 + k + kStages, w.rows_valid, w.seg_bytes,
                   w.row_floats, policy);
    }
  }

  // CTA-level reductions: warp shuffle -> shared -> ONE atomic per CTA and value
  if (kLoss || kIntGrad) {
    float v[4] = {kLoss ? loss_local : 0.0f, gi_local[0], gi_local[1], gi_local[2]};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if ((j == 0 && !kLoss) || (j > 0 && !kIntGrad)) continue;
      float sum = warp_sum(v[j]);
      if ((tid & 31) == 0) s_red[tid >> 5][j] = sum;
    }
    __syncthreads();
    if (tid < 4 && ((tid == 0 && kLoss) || (tid > 0 && kIntGrad))) {
      float sum = 0.0f;
#pragma unroll
      for (int i = 0; i < kThreads / 32; ++i) sum += s_red[i][tid];
      if (tid == 0) atomicAdd(p.loss_sum, sum);
      else atomicAdd(&p.d_intensity[tid - 1], sum);
    }
  }
}

}  // namespace pbr
