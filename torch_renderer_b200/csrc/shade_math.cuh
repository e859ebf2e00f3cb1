// Device-side Phong lighting + blend helpers shared by the stand-alone shade kernels (shade.cu)
// and the fused render kernels (render.cu).  Semantics: SURVEY.md Appendix A6-A8 (PyTorch3D
// renderer/mesh/shading.py, lighting.py, blending.py).
#pragma once
#include "raster_math.cuh"

namespace trb {

struct ViewParams {
  float lv[3], amb[3], dif[3], spec[3], shin, cam[3], znear, zfar;
};

__device__ __forceinline__ ViewParams load_view_params(const float* __restrict__ vp, int n) {
  const float* p = vp + (size_t)n * TRB_VIEW_PARAM_STRIDE;
  ViewParams o;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o.lv[i] = __ldg(p + i); o.amb[i] = __ldg(p + 3 + i); o.dif[i] = __ldg(p + 6 + i);
    o.spec[i] = __ldg(p + 9 + i); o.cam[i] = __ldg(p + 13 + i);
  }
  o.shin = __ldg(p + 12); o.znear = __ldg(p + 16); o.zfar = __ldg(p + 17);
  return o;
}

struct F3 {
  float x, y, z;
};
__device__ __forceinline__ F3 ld3(const float* __restrict__ p, int i) {
  const float* q = p + 3 * (size_t)i;
  return {__ldg(q), __ldg(q + 1), __ldg(q + 2)};
}
__device__ __forceinline__ float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 interp3(float b0, float b1, float b2, F3 a0, F3 a1, F3 a2) {
  return {b0 * a0.x + b1 * a1.x + b2 * a2.x, b0 * a0.y + b1 * a1.y + b2 * a2.y,
          b0 * a0.z + b1 * a1.z + b2 * a2.z};
}
// F.normalize(x, eps=1e-6): x / max(|x|, eps)
__device__ __forceinline__ F3 normalize3(F3 v, float& len_clamped, bool& clamped) {
  const float len = sqrtf(dot3(v, v));
  clamped = !(len > 1e-6f);
  len_clamped = clamped ? 1e-6f : len;
  const float inv = 1.0f / len_clamped;
  return {v.x * inv, v.y * inv, v.z * inv};
}
__device__ __forceinline__ F3 normalize3_bwd(F3 unit, F3 g_unit, float len_clamped, bool clamped) {
  const float inv = 1.0f / len_clamped;
  if (clamped) return {g_unit.x * inv, g_unit.y * inv, g_unit.z * inv};
  const float d = dot3(unit, g_unit);
  return {(g_unit.x - unit.x * d) * inv, (g_unit.y - unit.y * d) * inv, (g_unit.z - unit.z * d) * inv};
}

struct Lit {
  // forward intermediates kept for the backward
  F3 nh, lh, vh, refl;
  float nlen, llen, vlen, cosv, dotvr, a;
  bool nclamp, lclamp, vclamp;
  float diffuse_s;  // relu(cos)
  float pw;         // a^shininess
};

// colour = (amb + dif*relu(cos)) * tex + spec * a^shin        (A6)
// FAST (backward passes only: the image was produced by the precise version, gradients are held to 1e-3): the
// specular power goes through the SFU (__powf) instead of the ~70-instruction powf.
template <int LIGHT, bool FAST = false>
__device__ __forceinline__ F3 phong_color(const ViewParams& vp, F3 P, F3 nrm, F3 tex, Lit& s) {
  if (LIGHT == TRB_LIGHT_AMBIENT) return {vp.amb[0] * tex.x, vp.amb[1] * tex.y, vp.amb[2] * tex.z};
  s.nh = normalize3(nrm, s.nlen, s.nclamp);
  F3 l;
  if (LIGHT == TRB_LIGHT_POINT) l = {vp.lv[0] - P.x, vp.lv[1] - P.y, vp.lv[2] - P.z};
  else l = {vp.lv[0], vp.lv[1], vp.lv[2]};
  s.lh = normalize3(l, s.llen, s.lclamp);
  s.cosv = dot3(s.nh, s.lh);
  s.diffuse_s = fmaxf(s.cosv, 0.0f);
  const F3 vd = {vp.cam[0] - P.x, vp.cam[1] - P.y, vp.cam[2] - P.z};
  s.vh = normalize3(vd, s.vlen, s.vclamp);
  s.refl = {-s.lh.x + 2.0f * s.cosv * s.nh.x, -s.lh.y + 2.0f * s.cosv * s.nh.y,
            -s.lh.z + 2.0f * s.cosv * s.nh.z};
  s.dotvr = dot3(s.vh, s.refl);
  s.a = (s.cosv > 0.0f) ? fmaxf(s.dotvr, 0.0f) : 0.0f;
  s.pw = (s.a > 0.0f) ? (FAST ? __powf(s.a, vp.shin) : powf(s.a, vp.shin)) : (vp.shin == 0.0f ? 1.0f : 0.0f);
  return {(vp.amb[0] + vp.dif[0] * s.diffuse_s) * tex.x + vp.spec[0] * s.pw,
          (vp.amb[1] + vp.dif[1] * s.diffuse_s) * tex.y + vp.spec[1] * s.pw,
          (vp.amb[2] + vp.dif[2] * s.diffuse_s) * tex.z + vp.spec[2] * s.pw};
}

// Backward of phong_color: g = dL/dcolour -> g_tex, g_P, g_nrm, g_light_vec, g_cam.
template <int LIGHT, bool FAST = false>
__device__ __forceinline__ void phong_color_bwd(const ViewParams& vp, F3 tex, const Lit& s, F3 g,
                                                F3& g_tex, F3& g_P, F3& g_nrm, F3& g_lv, F3& g_cam) {
  g_P = {0, 0, 0}; g_nrm = {0, 0, 0}; g_lv = {0, 0, 0}; g_cam = {0, 0, 0};
  if (LIGHT == TRB_LIGHT_AMBIENT) {
    g_tex = {g.x * vp.amb[0], g.y * vp.amb[1], g.z * vp.amb[2]};
    return;
  }
  g_tex = {g.x * (vp.amb[0] + vp.dif[0] * s.diffuse_s), g.y * (vp.amb[1] + vp.dif[1] * s.diffuse_s),
           g.z * (vp.amb[2] + vp.dif[2] * s.diffuse_s)};
  const float g_diff = g.x * tex.x * vp.dif[0] + g.y * tex.y * vp.dif[1] + g.z * tex.z * vp.dif[2];
  float g_cos = (s.cosv > 0.0f) ? g_diff : 0.0f;
  const float g_pw = g.x * vp.spec[0] + g.y * vp.spec[1] + g.z * vp.spec[2];
  float g_dot = 0.0f;
  // d a^shin / d a = shin a^(shin-1); FAST: = shin * (a^shin) / a with the forward's power (no second pow)
  if (s.a > 0.0f && s.dotvr > 0.0f && s.cosv > 0.0f)
    g_dot = g_pw * vp.shin * (FAST ? s.pw * __frcp_rn(s.a) : powf(s.a, vp.shin - 1.0f));
  F3 g_vh = {g_dot * s.refl.x, g_dot * s.refl.y, g_dot * s.refl.z};
  const F3 g_refl = {g_dot * s.vh.x, g_dot * s.vh.y, g_dot * s.vh.z};
  F3 g_lh = {-g_refl.x, -g_refl.y, -g_refl.z};
  g_cos += 2.0f * dot3(g_refl, s.nh);
  F3 g_nh = {2.0f * s.cosv * g_refl.x, 2.0f * s.cosv * g_refl.y, 2.0f * s.cosv * g_refl.z};
  g_nh.x += g_cos * s.lh.x; g_nh.y += g_cos * s.lh.y; g_nh.z += g_cos * s.lh.z;
  g_lh.x += g_cos * s.nh.x; g_lh.y += g_cos * s.nh.y; g_lh.z += g_cos * s.nh.z;
  g_nrm = normalize3_bwd(s.nh, g_nh, s.nlen, s.nclamp);
  const F3 g_l = normalize3_bwd(s.lh, g_lh, s.llen, s.lclamp);
  const F3 g_vd = normalize3_bwd(s.vh, g_vh, s.vlen, s.vclamp);
  g_lv = g_l;
  if (LIGHT == TRB_LIGHT_POINT) { g_P.x -= g_l.x; g_P.y -= g_l.y; g_P.z -= g_l.z; }
  g_P.x -= g_vd.x; g_P.y -= g_vd.y; g_P.z -= g_vd.z;
  g_cam = g_vd;
}

// ---- TexturesUV (A7): bilinear lookup, align_corners=True, border padding, y-flipped map ---------------
// What PyTorch3D expresses as grid_sample(flip(maps, H), 2*uv - 1, align_corners=True, padding_mode="border"):
// column = u (Wt-1), row of the FLIPPED map = v (Ht-1), both clamped to the map; a clamped coordinate
// has zero gradient.
struct UvTex {
  const float* map;        // f32 [Ht, Wt, 3]; nullptr = per-vertex colours
  const float* verts_uvs;  // f32 [Vt, 2]
  const int* faces_uvs;    // i32 [F, 3]
  int h, w;
};

struct UvTap {
  int o00, o01, o10, o11;      // element offsets of the four texels (x3 channels), -1 = outside the map
  float w00, w01, w10, w11;
  float wx, wy, du, dv;        // fractions; d(column)/du and d(flipped row)/dv (0 when clamped)
};

__device__ __forceinline__ UvTap uv_tap(const UvTex& t, float u, float v) {
  UvTap k;
  const float xmax = (float)(t.w - 1), ymax = (float)(t.h - 1);
  float ix = u * xmax, iy = v * ymax;
  k.du = (ix > 0.0f && ix < xmax) ? xmax : 0.0f;
  k.dv = (iy > 0.0f && iy < ymax) ? ymax : 0.0f;
  ix = fminf(fmaxf(ix, 0.0f), xmax);
  iy = fminf(fmaxf(iy, 0.0f), ymax);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  k.wx = ix - fx; k.wy = iy - fy;
  k.w00 = (1.0f - k.wx) * (1.0f - k.wy); k.w01 = k.wx * (1.0f - k.wy);
  k.w10 = (1.0f - k.wx) * k.wy; k.w11 = k.wx * k.wy;
  const bool xin = x1 < t.w, yin = y1 < t.h;
  const int r0 = t.h - 1 - y0, r1 = t.h - 1 - y1;  // rows of the stored (un-flipped) map
  k.o00 = 3 * (r0 * t.w + x0);
  k.o01 = xin ? 3 * (r0 * t.w + x1) : -1;
  k.o10 = yin ? 3 * (r1 * t.w + x0) : -1;
  k.o11 = (xin && yin) ? 3 * (r1 * t.w + x1) : -1;
  return k;
}

__device__ __forceinline__ F3 uv_ld(const float* __restrict__ m, int o) {
  if (o < 0) return {0.0f, 0.0f, 0.0f};
  return {__ldg(m + o), __ldg(m + o + 1), __ldg(m + o + 2)};
}

// texel = sum_ij w_ij M_ij; optionally d texel / du and d texel / dv
__device__ __forceinline__ F3 uv_sample(const UvTex& t, const UvTap& k, F3* d_du, F3* d_dv) {
  const F3 a = uv_ld(t.map, k.o00), b = uv_ld(t.map, k.o01), c = uv_ld(t.map, k.o10), d = uv_ld(t.map, k.o11);
  if (d_du) {
    const float s = k.du, q = 1.0f - k.wy;
    *d_du = {s * (q * (b.x - a.x) + k.wy * (d.x - c.x)), s * (q * (b.y - a.y) + k.wy * (d.y - c.y)),
             s * (q * (b.z - a.z) + k.wy * (d.z - c.z))};
  }
  if (d_dv) {
    const float s = k.dv, q = 1.0f - k.wx;
    *d_dv = {s * (q * (c.x - a.x) + k.wx * (d.x - b.x)), s * (q * (c.y - a.y) + k.wx * (d.y - b.y)),
             s * (q * (c.z - a.z) + k.wx * (d.z - b.z))};
  }
  return {k.w00 * a.x + k.w01 * b.x + k.w10 * c.x + k.w11 * d.x,
          k.w00 * a.y + k.w01 * b.y + k.w10 * c.y + k.w11 * d.y,
          k.w00 * a.z + k.w01 * b.z + k.w10 * c.z + k.w11 * d.z};
}

// scatter of d loss / d texel into the gradient of the map
__device__ __forceinline__ void uv_scatter(float* __restrict__ g_map, const UvTap& k, F3 g) {
  const int o[4] = {k.o00, k.o01, k.o10, k.o11};
  const float w[4] = {k.w00, k.w01, k.w10, k.w11};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (o[i] < 0 || w[i] == 0.0f) continue;
    atomicAdd(g_map + o[i], w[i] * g.x); atomicAdd(g_map + o[i] + 1, w[i] * g.y);
    atomicAdd(g_map + o[i] + 2, w[i] * g.z);
  }
}

struct F2 {
  float x, y;
};
__device__ __forceinline__ F2 ld2(const float* __restrict__ p, int i) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p) + i);
  return {v.x, v.y};
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }
// backward passes: SFU exponential and reciprocal (~2e-7 relative)
__device__ __forceinline__ float sigmoid_fast(float x) { return __frcp_rn(1.0f + __expf(-x)); }

struct FaceIds {
  int i0, i1, i2;
};
__device__ __forceinline__ FaceIds face_ids(const int* __restrict__ faces, const trb_view& vd,
                                            long long f) {
  const size_t r = (size_t)(vd.face_start + (int)(f - vd.p2f_base));
  return {__ldg(faces + 3 * r), __ldg(faces + 3 * r + 1), __ldg(faces + 3 * r + 2)};
}

}  // namespace trb
