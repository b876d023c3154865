// tests/hostsim/hostsim.cpp — TEST INFRASTRUCTURE.  Compiles the device math headers
// (pypbr_b200/csrc/pbr_math.cuh, pbr_shade.cuh) as plain C++ so the CPU suite can replay the golden
// vectors through the very expressions the CUDA kernels execute (only MUFU pow differs: powf here).
// Build: g++ -O2 -ffp-contract=off -shared -fPIC   (tests/conftest.py does it).
// Never loaded by the product package.
#include <cstdint>
#include <cstring>
#include <vector>

#include "pbr_shade.cuh"

using namespace pbr;

namespace {

struct Dims {
  int B, H, W, L;
};

template <int WF, bool HASN, int LIGHT>
void fwd_impl(const Dims& d, const CtStage& S, const CtFlags& F, const float* albedo, const float* normal,
              const float* rough, const float* metspec, float* out) {
  const int mc = WF == 0 ? 1 : 3;  // WF 2: metallic with 3 channels
  const int64_t HW = (int64_t)d.H * d.W;
  for (int b = 0; b < d.B; ++b)
    for (int yy = 0; yy < d.H; ++yy)
      for (int xx = 0; xx < d.W; ++xx) {
        const int64_t o = (int64_t)yy * d.W + xx;
        float a[3][1], n[3][1] = {{0}, {0}, {1}}, r[1], m[3][1] = {{0}, {0}, {0}}, x[1];
        for (int c = 0; c < 3; ++c) a[c][0] = albedo[(b * 3 + c) * HW + o];
        if (HASN)
          for (int c = 0; c < 3; ++c) n[c][0] = normal[(b * 3 + c) * HW + o];
        r[0] = rough[b * HW + o];
        for (int c = 0; c < mc; ++c) m[c][0] = metspec[(b * mc + c) * HW + o];
        x[0] = linspace_at(S.lsx, xx);
        float y = linspace_at(S.lsy, yy);
        auto emit = [&](int l, const float(&v)[3][1]) {
          for (int c = 0; c < 3; ++c) {
            int64_t idx = F.per_light ? (((int64_t)b * d.L + l) * 3 + c) * HW + o : ((int64_t)b * 3 + c) * HW + o;
            out[idx] = v[c][0];
          }
        };
        LightGeom hg[1];
        if (LIGHT == kLightPointHoisted)
          point_light_geom(S.light[0].p[0], S.light[0].p[1], S.light[0].p[2], x[0], y, S.vx, S.vy, S.vz, hg[0]);
        ct_forward_group<WF, LIGHT, 1>(S, F, a, n, r, m, x, y, hg, emit);
      }
}

template <int WF, bool HASN, int LIGHT>
void bwd_impl(const Dims& d, const CtStage& S, const CtFlags& F, const float* albedo, const float* normal,
              const float* rough, const float* metspec, const float* grad_out, const float* target,
              float loss_scale, double* loss_sum, float* d_albedo, float* d_normal, float* d_rough, float* d_met,
              double* d_int) {
  const int mc = WF == 0 ? 1 : 3;  // WF 2: metallic with 3 channels
  const int64_t HW = (int64_t)d.H * d.W;
  for (int b = 0; b < d.B; ++b)
    for (int yy = 0; yy < d.H; ++yy)
      for (int xx = 0; xx < d.W; ++xx) {
        const int64_t o = (int64_t)yy * d.W + xx;
        float a[3][1], n[3][1] = {{0}, {0}, {1}}, r[1], m[3][1] = {{0}, {0}, {0}}, x[1];
        for (int c = 0; c < 3; ++c) a[c][0] = albedo[(b * 3 + c) * HW + o];
        if (HASN)
          for (int c = 0; c < 3; ++c) n[c][0] = normal[(b * 3 + c) * HW + o];
        r[0] = rough[b * HW + o];
        for (int c = 0; c < mc; ++c) m[c][0] = metspec[(b * mc + c) * HW + o];
        x[0] = linspace_at(S.lsx, xx);
        float y = linspace_at(S.lsy, yy);
        auto gout = [&](int l, const float(&outv)[3][1], float(&g)[3][1]) {
          for (int c = 0; c < 3; ++c) {
            int64_t idx = F.per_light ? (((int64_t)b * d.L + l) * 3 + c) * HW + o : ((int64_t)b * 3 + c) * HW + o;
            if (target) {
              float diff = outv[c][0] - target[idx];
              *loss_sum += (double)diff * diff;
              g[c][0] = 2.0f * loss_scale * diff;
            } else {
              g[c][0] = grad_out[idx];
            }
          }
        };
        auto sink = [&](int l, const float(&gi)[3]) {
          if (d_int)
            for (int c = 0; c < 3; ++c) d_int[3 * l + c] += gi[c];
        };
        float da[3][1], dn[3][1], dr[1], dm[3][1];
        LightGeom hg[1];
        if (LIGHT == kLightPointHoisted)
          point_light_geom(S.light[0].p[0], S.light[0].p[1], S.light[0].p[2], x[0], y, S.vx, S.vy, S.vz, hg[0]);
        ct_backward_group<WF, LIGHT, 1>(S, F, a, n, r, m, x, y, hg, gout, sink, da, dn, dr, dm);
        for (int c = 0; c < 3; ++c) d_albedo[(b * 3 + c) * HW + o] = da[c][0];
        if (HASN && d_normal)
          for (int c = 0; c < 3; ++c) d_normal[(b * 3 + c) * HW + o] = dn[c][0];
        d_rough[b * HW + o] = dr[0];
        for (int c = 0; c < mc; ++c) d_met[(b * mc + c) * HW + o] = dm[c][0];
      }
}

void make_stage(const Dims& d, int light_type, float light_size, const float* view, const float* lights,
                const float* inten, CtStage& S) {
  stage_view(view, S.vx, S.vy, S.vz);
  float s = light_size > 0.0f ? light_size : 1.0f;
  S.lsx = make_linspace(-s / 2, s / 2, d.W);
  S.lsy = make_linspace(-s / 2, s / 2, d.H);
  for (int l = 0; l < d.L; ++l) stage_light(l, lights, inten, light_type == 1, S.vx, S.vy, S.vz, S.light[l]);
}

// light mode: directional / point / point-hoisted (the kernels pick the hoisted variant for L == 1)
#define DISPATCH(fn, ...)                                                                   \
  do {                                                                                      \
    int lm = light_type == 1 ? ((L == 1 && !force_generic) ? 2 : 1) : 0;                    \
    int key = (workflow * 2 + (normal != nullptr)) * 3 + lm;                                \
    switch (key) {                                                                          \
      case 0: fn<0, false, 0>(__VA_ARGS__); break;                                          \
      case 1: fn<0, false, 1>(__VA_ARGS__); break;                                          \
      case 2: fn<0, false, 2>(__VA_ARGS__); break;                                          \
      case 3: fn<0, true, 0>(__VA_ARGS__); break;                                           \
      case 4: fn<0, true, 1>(__VA_ARGS__); break;                                           \
      case 5: fn<0, true, 2>(__VA_ARGS__); break;                                           \
      case 6: fn<1, false, 0>(__VA_ARGS__); break;                                          \
      case 7: fn<1, false, 1>(__VA_ARGS__); break;                                          \
      case 8: fn<1, false, 2>(__VA_ARGS__); break;                                          \
      case 9: fn<1, true, 0>(__VA_ARGS__); break;                                           \
      case 10: fn<1, true, 1>(__VA_ARGS__); break;                                          \
      case 11: fn<1, true, 2>(__VA_ARGS__); break;                                          \
      case 12: fn<2, false, 0>(__VA_ARGS__); break;                                         \
      case 13: fn<2, false, 1>(__VA_ARGS__); break;                                         \
      case 14: fn<2, false, 2>(__VA_ARGS__); break;                                         \
      case 15: fn<2, true, 0>(__VA_ARGS__); break;                                          \
      case 16: fn<2, true, 1>(__VA_ARGS__); break;                                          \
      case 17: fn<2, true, 2>(__VA_ARGS__); break;                                          \
    }                                                                                       \
  } while (0)

}  // namespace

extern "C" {

int hs_ct_forward(int B, int H, int W, int L, int workflow, int light_type, int albedo_is_srgb, int specular_is_srgb,
                  int return_srgb, int per_light, float light_size, const float* albedo, const float* normal,
                  const float* rough, const float* metspec, const float* view, const float* lights,
                  const float* inten, float* out, int force_generic) {
  if (L > PBR_MAX_LIGHTS) return -4;
  Dims d{B, H, W, L};
  static CtStage S;
  make_stage(d, light_type, light_size, view, lights, inten, S);
  CtFlags F{L, light_type == 1, albedo_is_srgb != 0, specular_is_srgb != 0, return_srgb != 0, per_light != 0};
  DISPATCH(fwd_impl, d, S, F, albedo, normal, rough, metspec, out);
  return 0;
}

int hs_ct_backward(int B, int H, int W, int L, int workflow, int light_type, int albedo_is_srgb, int specular_is_srgb,
                   int return_srgb, int per_light, float light_size, const float* albedo, const float* normal,
                   const float* rough, const float* metspec, const float* view, const float* lights,
                   const float* inten, const float* grad_out, const float* target, float loss_scale,
                   double* loss_sum, float* d_albedo, float* d_normal, float* d_rough, float* d_met, double* d_int,
                   int force_generic) {
  if (L > PBR_MAX_LIGHTS) return -4;
  Dims d{B, H, W, L};
  static CtStage S;
  make_stage(d, light_type, light_size, view, lights, inten, S);
  CtFlags F{L, light_type == 1, albedo_is_srgb != 0, specular_is_srgb != 0, return_srgb != 0, per_light != 0};
  DISPATCH(bwd_impl, d, S, F, albedo, normal, rough, metspec, grad_out, target, loss_scale, loss_sum, d_albedo,
           d_normal, d_rough, d_met, d_int);
  return 0;
}

void hs_convert_m2s(int64_t n_texels, int albedo_is_srgb, const float* albedo, const float* met, float* diffuse,
                    float* specular) {
  // single material (3, n_texels) planar
  for (int64_t i = 0; i < n_texels; ++i) {
    float a[3] = {albedo[i], albedo[n_texels + i], albedo[2 * n_texels + i]}, dd[3], ss[3];
    convert_m2s(a, met[i], albedo_is_srgb != 0, dd, ss);
    for (int c = 0; c < 3; ++c) {
      diffuse[c * n_texels + i] = dd[c];
      specular[c * n_texels + i] = ss[c];
    }
  }
}

void hs_convert_s2m(int64_t n, int albedo_is_srgb, const float* diffuse, const float* specular, float* basecolor,
                    float* metallic) {
  for (int64_t i = 0; i < n; ++i) convert_s2m(diffuse[i], specular[i], albedo_is_srgb != 0, &basecolor[i], &metallic[i]);
}

void hs_blend(int64_t n_texels, int channels, int is_normal, const float* mask, const float* a, const float* b,
              float* out) {
  for (int64_t i = 0; i < n_texels; ++i) {
    if (is_normal) {
      float aa[3] = {a[i], a[n_texels + i], a[2 * n_texels + i]};
      float bb[3] = {b[i], b[n_texels + i], b[2 * n_texels + i]};
      float o[3];
      blend_normal(mask[i], aa, bb, o);
      for (int c = 0; c < 3; ++c) out[c * n_texels + i] = o[c];
    } else {
      for (int c = 0; c < channels; ++c) out[c * n_texels + i] = blend_lerp(mask[i], a[c * n_texels + i], b[c * n_texels + i]);
    }
  }
}

void hs_sigmoid_mask(int64_t n, const float* p1, const float* p2, float shift, int apply_shift, float blend_width,
                     float* out) {
  float w = blend_width + 1e-6f;
  for (int64_t i = 0; i < n; ++i) out[i] = sigmoid_mask(p1[i], p2[i], shift, apply_shift != 0, w);
}

void hs_linspace(float start, float end, int n, float* out) {
  Linspace ls = make_linspace(start, end, n);
  for (int i = 0; i < n; ++i) out[i] = linspace_at(ls, i);
}

void hs_srgb(int64_t n, int to_linear, const float* in, float* out) {
  for (int64_t i = 0; i < n; ++i)
    out[i] = to_linear ? srgb_decode<false>(in[i], nullptr) : srgb_encode<false>(in[i], nullptr);
}

void hs_ingest_normal(int64_t n_texels, int channels, const float* in, float* out) {
  for (int64_t i = 0; i < n_texels; ++i) {
    float o[3];
    if (channels == 3) {
      float v[3] = {in[i], in[n_texels + i], in[2 * n_texels + i]};
      ingest_normal3(v, o);
    } else {
      float v[2] = {in[i], in[n_texels + i]};
      ingest_normal2(v, o);
    }
    for (int c = 0; c < 3; ++c) out[c * n_texels + i] = o[c];
  }
}

// Markstein division with a correctly rounded reciprocal vs IEEE division: returns mismatches.
int64_t hs_div_check(int64_t n, const float* a, const float* b) {
  int64_t bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    float q = xdiv_r(a[i], b[i], xrcp(b[i]));
    float e = a[i] / b[i];
    if (memcmp(&q, &e, 4) != 0) ++bad;
  }
  return bad;
}

}  // extern "C"

extern "C" {
// seeded sqrt / reciprocal recurrences vs IEEE: mismatch counts (sqrt, reciprocal-of-sqrt-based division)
void hs_seeded_check(int64_t n, const float* ss, const float* num, int64_t* bad_sqrt, int64_t* bad_div) {
  *bad_sqrt = 0;
  *bad_div = 0;
  for (int64_t i = 0; i < n; ++i) {
    float seed;
    float len = sqrt_seeded(ss[i], &seed);
    float ref = sqrtf(ss[i]);
    if (memcmp(&len, &ref, 4) != 0) ++*bad_sqrt;
    float r = refine_rcp(ref, seed);
    float q = xdiv_r(num[i], ref, r);
    float e = num[i] / ref;
    if (memcmp(&q, &e, 4) != 0) ++*bad_div;
  }
}
}
