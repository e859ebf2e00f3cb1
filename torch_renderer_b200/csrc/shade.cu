// Fused attribute interpolation + Phong lighting + blending, forward and backward, plus the
// stand-alone interpolate_face_attributes twin.  Replaces what PyTorch3D runs as three
// interp_face_attrs launches and ~40 ATen elementwise/reduction ops per render
// (renderer/mesh/shading.py::phong_shading, lighting.py, blending.py -- SURVEY.md A6-A8; reference
// call sites: renderer.py:87-101, torch_renderer.py:102-108,144-158, camera_pose_optimizer.py:130-158,
// mesh_deformer.py:142-145, myrenderer.py:88,105).
//
// One thread per pixel, K layers in the inner loop; Fragments are read once, nothing of size
// N*H*W*K*3 is ever materialised.  The backward evaluates the lighting model once per sample and
// scatters vertex-attribute gradients with warp-aggregated atomics.
#include "shade_math.cuh"

namespace trb {

// ------------------------------------------------------------------------------------------
template <int SHADER, int LIGHT, int TEX>
__global__ void __launch_bounds__(256)
shade_forward_kernel(trb_shade_config cfg, const trb_view* __restrict__ views,
                     const float* __restrict__ view_params, const long long* __restrict__ p2f,
                     const float* __restrict__ bary, const float* __restrict__ zbuf,
                     const float* __restrict__ dists, const int* __restrict__ faces,
                     const float* __restrict__ verts, const float* __restrict__ normals,
                     const float* __restrict__ colors, const float* __restrict__ texels,
                     float* __restrict__ images) {
  const int K = cfg.K;
  const long long npix = (long long)cfg.N * cfg.H * cfg.W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int n = (int)(pix / ((long long)cfg.H * cfg.W));
  const long long s0 = pix * K;
  float4 out;
  if (SHADER == TRB_SHADER_SOFT_SILHOUETTE) {
    float alpha = 1.0f;
    for (int k = 0; k < K; ++k) {
      if (p2f[s0 + k] < 0) break;  // layers are front-to-back then -1
      alpha *= 1.0f - sigmoidf(-dists[s0 + k] / cfg.sigma);
    }
    out = make_float4(1.0f, 1.0f, 1.0f, 1.0f - alpha);
    st_cs(reinterpret_cast<float4*>(images) + pix, out);
    return;
  }
  const trb_view vd = views[n];
  const ViewParams vp = load_view_params(view_params, n);
  if (SHADER == TRB_SHADER_HARD_PHONG) {
    const long long f = p2f[s0];
    if (f < 0) {
      out = make_float4(cfg.background[0], cfg.background[1], cfg.background[2], 0.0f);
    } else {
      const FaceIds id = face_ids(faces, vd, f);
      const float b0 = bary[s0 * 3], b1 = bary[s0 * 3 + 1], b2 = bary[s0 * 3 + 2];
      const F3 P = interp3(b0, b1, b2, ld3(verts, id.i0), ld3(verts, id.i1), ld3(verts, id.i2));
      const F3 nr = interp3(b0, b1, b2, ld3(normals, id.i0), ld3(normals, id.i1), ld3(normals, id.i2));
      F3 tex;
      if (TEX == TRB_TEX_VERTEX) tex = interp3(b0, b1, b2, ld3(colors, id.i0), ld3(colors, id.i1), ld3(colors, id.i2));
      else tex = {texels[s0 * 3], texels[s0 * 3 + 1], texels[s0 * 3 + 2]};
      Lit lit;
      const F3 c = phong_color<LIGHT>(vp, P, nr, tex, lit);
      out = make_float4(c.x, c.y, c.z, 1.0f);
    }
    st_cs(reinterpret_cast<float4*>(images) + pix, out);
    return;
  }
  // soft phong: softmax_rgb_blend (A8)
  const float eps = 1e-10f;
  const float zrange = vp.zfar - vp.znear;
  float zmax = eps;
  for (int k = 0; k < K; ++k) {
    if (p2f[s0 + k] < 0) break;
    zmax = fmaxf(zmax, (vp.zfar - zbuf[s0 + k]) / zrange);
  }
  float alpha = 1.0f, wsum = 0.0f;
  F3 acc = {0, 0, 0};
  for (int k = 0; k < K; ++k) {
    const long long f = p2f[s0 + k];
    if (f < 0) break;
    const long long s = s0 + k;
    const float prob = sigmoidf(-dists[s] / cfg.sigma);
    alpha *= 1.0f - prob;
    const float zinv = (vp.zfar - zbuf[s]) / zrange;
    const float w = prob * expf((zinv - zmax) / cfg.gamma);
    const FaceIds id = face_ids(faces, vd, f);
    const float b0 = bary[s * 3], b1 = bary[s * 3 + 1], b2 = bary[s * 3 + 2];
    const F3 P = interp3(b0, b1, b2, ld3(verts, id.i0), ld3(verts, id.i1), ld3(verts, id.i2));
    const F3 nr = interp3(b0, b1, b2, ld3(normals, id.i0), ld3(normals, id.i1), ld3(normals, id.i2));
    F3 tex;
    if (TEX == TRB_TEX_VERTEX) tex = interp3(b0, b1, b2, ld3(colors, id.i0), ld3(colors, id.i1), ld3(colors, id.i2));
    else tex = {texels[s * 3], texels[s * 3 + 1], texels[s * 3 + 2]};
    Lit lit;
    const F3 c = phong_color<LIGHT>(vp, P, nr, tex, lit);
    wsum += w;
    acc.x += w * c.x; acc.y += w * c.y; acc.z += w * c.z;
  }
  const float delta = fmaxf(expf((eps - zmax) / cfg.gamma), eps);
  const float inv = 1.0f / (wsum + delta);
  out = make_float4((acc.x + delta * cfg.background[0]) * inv, (acc.y + delta * cfg.background[1]) * inv,
                    (acc.z + delta * cfg.background[2]) * inv, 1.0f - alpha);
  st_cs(reinterpret_cast<float4*>(images) + pix, out);
}

// ------------------------------------------------------------------------------------------
// Backward.  Pass A (no colours): probabilities, softmax weights, denominators.  Pass B: the
// lighting model is evaluated once per sample, forward and backward; its contribution to the
// blend-weight gradient (g . colour_k) is parked in grad_dists.  Pass C turns the parked values
// into grad_dists / grad_zbuf.
template <int SHADER, int LIGHT, int TEX>
__global__ void __launch_bounds__(256)
shade_backward_kernel(trb_shade_config cfg, const trb_view* __restrict__ views,
                      const float* __restrict__ view_params, const long long* __restrict__ p2f,
                      const float* __restrict__ bary, const float* __restrict__ zbuf,
                      const float* __restrict__ dists, const int* __restrict__ faces,
                      const float* __restrict__ verts, const float* __restrict__ normals,
                      const float* __restrict__ colors, const float* __restrict__ texels,
                      const float* __restrict__ grad_images, float* __restrict__ grad_bary,
                      float* __restrict__ grad_zbuf, float* __restrict__ grad_dists,
                      float* __restrict__ grad_verts, float* __restrict__ grad_normals,
                      float* __restrict__ grad_colors, float* __restrict__ grad_texels,
                      float* __restrict__ grad_view_params) {
  const int K = cfg.K;
  const long long npix = (long long)cfg.N * cfg.H * cfg.W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = pix < npix;
  const long long pixc = live ? pix : npix - 1;
  const int n = (int)(pixc / ((long long)cfg.H * cfg.W));
  const long long s0 = pixc * K;
  const float4 g = live ? __ldcs(reinterpret_cast<const float4*>(grad_images) + pixc)
                        : make_float4(0, 0, 0, 0);

  // ---- alpha channel: A = 1 - prod_k (1 - p_k); shared by all three shaders except hard.
  int nk = 0;  // number of valid layers of this pixel
  if (live) {
    while (nk < K && p2f[s0 + nk] >= 0) ++nk;
  }
  if (SHADER == TRB_SHADER_SOFT_SILHOUETTE) {
    if (!live) return;
    float prod_nz = 1.0f; int zeros = 0;
    for (int k = 0; k < nk; ++k) {
      const float q = 1.0f - sigmoidf(-dists[s0 + k] / cfg.sigma);
      if (q == 0.0f) ++zeros; else prod_nz *= q;
    }
    for (int k = 0; k < K; ++k) {
      float gd = 0.0f;
      if (k < nk) {
        const float p = sigmoidf(-dists[s0 + k] / cfg.sigma);
        const float q = 1.0f - p;
        const float others = zeros == 0 ? prod_nz / q : (zeros == 1 && q == 0.0f ? prod_nz : 0.0f);
        gd = g.w * others * p * q * (-1.0f / cfg.sigma);
      }
      if (grad_dists) grad_dists[s0 + k] = gd;
      if (grad_zbuf) grad_zbuf[s0 + k] = 0.0f;
      if (grad_bary) { grad_bary[(s0 + k) * 3] = 0.0f; grad_bary[(s0 + k) * 3 + 1] = 0.0f; grad_bary[(s0 + k) * 3 + 2] = 0.0f; }
    }
    return;
  }

  const trb_view vd = views[n];
  const ViewParams vp = load_view_params(view_params, n);
  const float eps = 1e-10f;
  const float zrange = vp.zfar - vp.znear;
  const bool hard = (SHADER == TRB_SHADER_HARD_PHONG);
  const int nloop = hard ? min(nk, 1) : nk;

  // ---- pass A
  float zmax = eps; int kmax = -1;
  float prod_nz = 1.0f; int zeros = 0;
  float wsum = 0.0f, delta = 0.0f, den = 1.0f;
  if (!hard) {
    for (int k = 0; k < nk; ++k) {
      const float zinv = (vp.zfar - zbuf[s0 + k]) / zrange;
      if (zinv > zmax) { zmax = zinv; kmax = k; }
      const float q = 1.0f - sigmoidf(-dists[s0 + k] / cfg.sigma);
      if (q == 0.0f) ++zeros; else prod_nz *= q;
    }
    for (int k = 0; k < nk; ++k) {
      const float zinv = (vp.zfar - zbuf[s0 + k]) / zrange;
      wsum += sigmoidf(-dists[s0 + k] / cfg.sigma) * expf((zinv - zmax) / cfg.gamma);
    }
    delta = fmaxf(expf((eps - zmax) / cfg.gamma), eps);
    den = wsum + delta;
  }
  const float inv_den = 1.0f / den;

  // ---- pass B: colours
  F3 acc = {0, 0, 0};  // sum_k w_k c_k
  F3 g_lv_acc = {0, 0, 0}, g_cam_acc = {0, 0, 0};
  const int nloop_warp = __reduce_max_sync(0xffffffffu, nloop);
  for (int k = 0; k < nloop_warp; ++k) {
    const bool on = k < nloop;
    const long long s = s0 + k;
    int key = -1;
    FaceIds id = {0, 0, 0};
    float b0 = 0, b1 = 0, b2 = 0;
    F3 gP = {0, 0, 0}, gN = {0, 0, 0}, gT = {0, 0, 0};
    if (on) {
      const long long f = p2f[s];
      id = face_ids(faces, vd, f);
      key = (int)f;
      b0 = bary[s * 3]; b1 = bary[s * 3 + 1]; b2 = bary[s * 3 + 2];
      const F3 X0 = ld3(verts, id.i0), X1 = ld3(verts, id.i1), X2 = ld3(verts, id.i2);
      const F3 N0 = ld3(normals, id.i0), N1 = ld3(normals, id.i1), N2 = ld3(normals, id.i2);
      const F3 P = interp3(b0, b1, b2, X0, X1, X2);
      const F3 nr = interp3(b0, b1, b2, N0, N1, N2);
      F3 C0 = {0, 0, 0}, C1 = {0, 0, 0}, C2 = {0, 0, 0}, tex;
      if (TEX == TRB_TEX_VERTEX) {
        C0 = ld3(colors, id.i0); C1 = ld3(colors, id.i1); C2 = ld3(colors, id.i2);
        tex = interp3(b0, b1, b2, C0, C1, C2);
      } else {
        tex = {texels[s * 3], texels[s * 3 + 1], texels[s * 3 + 2]};
      }
      Lit lit;
      const F3 c = phong_color<LIGHT, true>(vp, P, nr, tex, lit);
      float w = 1.0f;  // d rgb / d colour_k
      if (!hard) {
        const float zinv = (vp.zfar - zbuf[s]) / zrange;
        w = sigmoidf(-dists[s] / cfg.sigma) * expf((zinv - zmax) / cfg.gamma);
        acc.x += w * c.x; acc.y += w * c.y; acc.z += w * c.z;
        if (grad_dists) grad_dists[s] = g.x * c.x + g.y * c.y + g.z * c.z;  // parked for pass C
      }
      const float wn = hard ? 1.0f : w * inv_den;
      const F3 gc = {g.x * wn, g.y * wn, g.z * wn};
      F3 g_lv, g_cam;
      phong_color_bwd<LIGHT, true>(vp, tex, lit, gc, gT, gP, gN, g_lv, g_cam);
      g_lv_acc.x += g_lv.x; g_lv_acc.y += g_lv.y; g_lv_acc.z += g_lv.z;
      g_cam_acc.x += g_cam.x; g_cam_acc.y += g_cam.y; g_cam_acc.z += g_cam.z;
      if (grad_bary) {
        float gb0 = dot3(gP, X0) + dot3(gN, N0), gb1 = dot3(gP, X1) + dot3(gN, N1),
              gb2 = dot3(gP, X2) + dot3(gN, N2);
        if (TEX == TRB_TEX_VERTEX) { gb0 += dot3(gT, C0); gb1 += dot3(gT, C1); gb2 += dot3(gT, C2); }
        grad_bary[s * 3] = gb0; grad_bary[s * 3 + 1] = gb1; grad_bary[s * 3 + 2] = gb2;
      }
      if (TEX == TRB_TEX_TEXELS && grad_texels) {
        grad_texels[s * 3] = gT.x; grad_texels[s * 3 + 1] = gT.y; grad_texels[s * 3 + 2] = gT.z;
      }
    }
    // vertex-attribute scatters (whole warp participates)
    if (TEX == TRB_TEX_VERTEX && grad_colors) {
      const float v[9] = {b0 * gT.x, b0 * gT.y, b0 * gT.z, b1 * gT.x, b1 * gT.y, b1 * gT.z,
                          b2 * gT.x, b2 * gT.y, b2 * gT.z};
      float* const d[9] = {grad_colors + 3 * (size_t)id.i0, grad_colors + 3 * (size_t)id.i0 + 1,
                           grad_colors + 3 * (size_t)id.i0 + 2, grad_colors + 3 * (size_t)id.i1,
                           grad_colors + 3 * (size_t)id.i1 + 1, grad_colors + 3 * (size_t)id.i1 + 2,
                           grad_colors + 3 * (size_t)id.i2, grad_colors + 3 * (size_t)id.i2 + 1,
                           grad_colors + 3 * (size_t)id.i2 + 2};
      warp_aggregated_add<9>(key, v, d);
    }
    if (LIGHT != TRB_LIGHT_AMBIENT && grad_verts) {
      const float v[9] = {b0 * gP.x, b0 * gP.y, b0 * gP.z, b1 * gP.x, b1 * gP.y, b1 * gP.z,
                          b2 * gP.x, b2 * gP.y, b2 * gP.z};
      float* const d[9] = {grad_verts + 3 * (size_t)id.i0, grad_verts + 3 * (size_t)id.i0 + 1,
                           grad_verts + 3 * (size_t)id.i0 + 2, grad_verts + 3 * (size_t)id.i1,
                           grad_verts + 3 * (size_t)id.i1 + 1, grad_verts + 3 * (size_t)id.i1 + 2,
                           grad_verts + 3 * (size_t)id.i2, grad_verts + 3 * (size_t)id.i2 + 1,
                           grad_verts + 3 * (size_t)id.i2 + 2};
      warp_aggregated_add<9>(key, v, d);
    }
    if (LIGHT != TRB_LIGHT_AMBIENT && grad_normals) {
      const float v[9] = {b0 * gN.x, b0 * gN.y, b0 * gN.z, b1 * gN.x, b1 * gN.y, b1 * gN.z,
                          b2 * gN.x, b2 * gN.y, b2 * gN.z};
      float* const d[9] = {grad_normals + 3 * (size_t)id.i0, grad_normals + 3 * (size_t)id.i0 + 1,
                           grad_normals + 3 * (size_t)id.i0 + 2, grad_normals + 3 * (size_t)id.i1,
                           grad_normals + 3 * (size_t)id.i1 + 1, grad_normals + 3 * (size_t)id.i1 + 2,
                           grad_normals + 3 * (size_t)id.i2, grad_normals + 3 * (size_t)id.i2 + 1,
                           grad_normals + 3 * (size_t)id.i2 + 2};
      warp_aggregated_add<9>(key, v, d);
    }
  }
  if (LIGHT != TRB_LIGHT_AMBIENT && grad_view_params) {
    // one atomic per warp and component; a warp never straddles two views when H*W % 32 == 0,
    // otherwise fall back to per-lane atomics.
    const int n0 = __shfl_sync(0xffffffffu, n, 0);
    const bool uniform = __all_sync(0xffffffffu, n == n0);
    float vals[6] = {g_lv_acc.x, g_lv_acc.y, g_lv_acc.z, g_cam_acc.x, g_cam_acc.y, g_cam_acc.z};
    float* gp = grad_view_params + (size_t)n * TRB_VIEW_PARAM_STRIDE;
    if (uniform) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float sum = warp_sum(vals[i]);
        if ((threadIdx.x & 31) == 0 && sum != 0.0f) atomicAdd(gp + (i < 3 ? i : 10 + i), sum);
      }
    } else if (live) {
#pragma unroll
      for (int i = 0; i < 6; ++i)
        if (vals[i] != 0.0f) atomicAdd(gp + (i < 3 ? i : 10 + i), vals[i]);
    }
  }
  if (!live) return;

  // ---- pass C: blend-weight gradients -> grad_dists, grad_zbuf
  if (hard) {
    for (int k = 0; k < K; ++k) {
      if (grad_dists) grad_dists[s0 + k] = 0.0f;
      if (grad_zbuf) grad_zbuf[s0 + k] = 0.0f;
      if (k >= nloop) {
        if (grad_bary) { grad_bary[(s0 + k) * 3] = 0.0f; grad_bary[(s0 + k) * 3 + 1] = 0.0f; grad_bary[(s0 + k) * 3 + 2] = 0.0f; }
        if (TEX == TRB_TEX_TEXELS && grad_texels) { grad_texels[(s0 + k) * 3] = 0.0f; grad_texels[(s0 + k) * 3 + 1] = 0.0f; grad_texels[(s0 + k) * 3 + 2] = 0.0f; }
      }
    }
    return;
  }
  const F3 bg = {cfg.background[0], cfg.background[1], cfg.background[2]};
  const F3 rgb = {(acc.x + delta * bg.x) * inv_den, (acc.y + delta * bg.y) * inv_den,
                  (acc.z + delta * bg.z) * inv_den};
  const float g_rgb = g.x * rgb.x + g.y * rgb.y + g.z * rgb.z;
  const float g_delta = ((g.x * bg.x + g.y * bg.y + g.z * bg.z) - g_rgb) * inv_den;
  const bool delta_clamped = !(expf((eps - zmax) / cfg.gamma) > eps);
  float g_zmax = delta_clamped ? 0.0f : -g_delta * delta / cfg.gamma;
  // first sweep: everything except the arg-max routing (needs the complete g_zmax)
  for (int k = 0; k < K; ++k) {
    const long long s = s0 + k;
    if (k >= nk) {
      if (grad_dists) grad_dists[s] = 0.0f;
      if (grad_zbuf) grad_zbuf[s] = 0.0f;
      if (grad_bary) { grad_bary[s * 3] = 0.0f; grad_bary[s * 3 + 1] = 0.0f; grad_bary[s * 3 + 2] = 0.0f; }
      if (TEX == TRB_TEX_TEXELS && grad_texels) { grad_texels[s * 3] = 0.0f; grad_texels[s * 3 + 1] = 0.0f; grad_texels[s * 3 + 2] = 0.0f; }
      continue;
    }
    const float p = sigmoidf(-dists[s] / cfg.sigma);
    const float q = 1.0f - p;
    const float zinv = (vp.zfar - zbuf[s]) / zrange;
    const float E = expf((zinv - zmax) / cfg.gamma);
    const float w = p * E;
    const float gdotc = grad_dists ? grad_dists[s] : 0.0f;  // parked g . colour_k
    const float g_w = (gdotc - g_rgb) * inv_den;
    const float others = zeros == 0 ? prod_nz / q : (zeros == 1 && q == 0.0f ? prod_nz : 0.0f);
    const float g_p = g_w * E + g.w * others;
    const float g_zinv = g_w * w / cfg.gamma;
    g_zmax -= g_zinv;
    if (grad_dists) grad_dists[s] = g_p * p * q * (-1.0f / cfg.sigma);
    if (grad_zbuf) grad_zbuf[s] = -g_zinv / zrange;
  }
  if (kmax >= 0 && grad_zbuf) grad_zbuf[s0 + kmax] += -g_zmax / zrange;
}

// ------------------------------------------------------------------------------------------
// Stand-alone interpolate_face_attributes (used for D != 3 attributes such as UVs).
__global__ void __launch_bounds__(256)
interp_forward_kernel(const long long* __restrict__ p2f, const float* __restrict__ bary,
                      const float* __restrict__ attrs, long long P, int D, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * D) return;
  const long long p = i / D;
  const int d = (int)(i - p * D);
  const long long f = p2f[p];
  float v = 0.0f;
  if (f >= 0) {
    const float* a = attrs + (size_t)f * 3 * D + d;
    v = bary[p * 3] * __ldg(a) + bary[p * 3 + 1] * __ldg(a + D) + bary[p * 3 + 2] * __ldg(a + 2 * D);
  }
  out[i] = v;
}

__global__ void __launch_bounds__(256)
interp_backward_kernel(const long long* __restrict__ p2f, const float* __restrict__ bary,
                       const float* __restrict__ attrs, const float* __restrict__ grad_out, long long P,
                       int D, float* __restrict__ grad_bary, float* __restrict__ grad_attrs) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long f = p2f[p];
  float gb0 = 0.0f, gb1 = 0.0f, gb2 = 0.0f;
  if (f >= 0) {
    const float b0 = bary[p * 3], b1 = bary[p * 3 + 1], b2 = bary[p * 3 + 2];
    const float* a = attrs + (size_t)f * 3 * D;
    float* ga = grad_attrs + (size_t)f * 3 * D;
    for (int d = 0; d < D; ++d) {
      const float g = grad_out[p * D + d];
      gb0 += g * __ldg(a + d); gb1 += g * __ldg(a + D + d); gb2 += g * __ldg(a + 2 * D + d);
      if (g != 0.0f) {
        atomicAdd(ga + d, b0 * g); atomicAdd(ga + D + d, b1 * g); atomicAdd(ga + 2 * D + d, b2 * g);
      }
    }
  }
  grad_bary[p * 3] = gb0; grad_bary[p * 3 + 1] = gb1; grad_bary[p * 3 + 2] = gb2;
}

static int check_cfg(const trb_shade_config* c) {
  if (!c || c->N < 0 || c->H < 1 || c->W < 1 || c->K < 1) return TRB_ERR_BAD_ARG;
  if (c->K > TRB_MAX_FACES_PER_PIXEL) return TRB_ERR_K_TOO_LARGE;
  if (c->shader < 0 || c->shader > 2 || c->light_kind < 0 || c->light_kind > 2 || c->texture_mode < 0 ||
      c->texture_mode > 1)
    return TRB_ERR_BAD_ARG;
  if (!(c->sigma > 0.0f) || !(c->gamma > 0.0f)) return TRB_ERR_BAD_ARG;
  return TRB_OK;
}

}  // namespace trb

using namespace trb;

// Dispatch over (shader, light, texture) template instances.
#define TRB_DISPATCH_LT(SH, MACRO)                                                          \
  do {                                                                                      \
    const int lk = cfg.light_kind, tm = cfg.texture_mode;                                   \
    if (lk == 0 && tm == 0) MACRO(SH, 0, 0); else if (lk == 0) MACRO(SH, 0, 1);             \
    else if (lk == 1 && tm == 0) MACRO(SH, 1, 0); else if (lk == 1) MACRO(SH, 1, 1);        \
    else if (tm == 0) MACRO(SH, 2, 0); else MACRO(SH, 2, 1);                                \
  } while (0)

extern "C" int trb_shade_forward(const trb_shade_config* host_cfg, const trb_view* views,
                                 const float* view_params, const int64_t* pix_to_face,
                                 const float* bary, const float* zbuf, const float* dists,
                                 const int32_t* faces, const float* verts_world,
                                 const float* vert_normals, const float* vert_colors,
                                 const float* texels, float* images, int device,
                                 trb_stream_t stream) {
  const int rc = check_cfg(host_cfg);
  if (rc != TRB_OK) return rc;
  const trb_shade_config cfg = *host_cfg;
  if (cfg.N == 0) return TRB_OK;
  if (!pix_to_face || !dists || !images) return TRB_ERR_BAD_ARG;
  if (cfg.shader != TRB_SHADER_SOFT_SILHOUETTE) {
    if (!views || !view_params || !bary || !zbuf || !faces || !verts_world || !vert_normals)
      return TRB_ERR_BAD_ARG;
    if (cfg.texture_mode == TRB_TEX_VERTEX ? !vert_colors : !texels) return TRB_ERR_BAD_ARG;
  }
  TRB_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long npix = (long long)cfg.N * cfg.H * cfg.W;
  const unsigned blocks = (unsigned)ceil_div64(npix, 256);
  const long long* p2f = (const long long*)pix_to_face;
#define TRB_FWD(SH, L, T)                                                                        \
  shade_forward_kernel<SH, L, T><<<blocks, 256, 0, st>>>(cfg, views, view_params, p2f, bary, zbuf, \
                                                         dists, faces, verts_world, vert_normals,  \
                                                         vert_colors, texels, images)
  if (cfg.shader == TRB_SHADER_SOFT_SILHOUETTE) TRB_FWD(TRB_SHADER_SOFT_SILHOUETTE, 0, 0);
  else if (cfg.shader == TRB_SHADER_HARD_PHONG) TRB_DISPATCH_LT(TRB_SHADER_HARD_PHONG, TRB_FWD);
  else TRB_DISPATCH_LT(TRB_SHADER_SOFT_PHONG, TRB_FWD);
#undef TRB_FWD
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

extern "C" int trb_shade_backward(const trb_shade_config* host_cfg, const trb_view* views,
                                  const float* view_params, const int64_t* pix_to_face,
                                  const float* bary, const float* zbuf, const float* dists,
                                  const int32_t* faces, const float* verts_world,
                                  const float* vert_normals, const float* vert_colors,
                                  const float* texels, const float* grad_images, float* grad_bary,
                                  float* grad_zbuf, float* grad_dists, float* grad_verts_world,
                                  float* grad_vert_normals, float* grad_vert_colors,
                                  float* grad_texels, float* grad_view_params, int device,
                                  trb_stream_t stream) {
  const int rc = check_cfg(host_cfg);
  if (rc != TRB_OK) return rc;
  const trb_shade_config cfg = *host_cfg;
  if (cfg.N == 0) return TRB_OK;
  if (!pix_to_face || !dists || !grad_images) return TRB_ERR_BAD_ARG;
  if (cfg.shader != TRB_SHADER_SOFT_SILHOUETTE) {
    if (!views || !view_params || !bary || !zbuf || !faces || !verts_world || !vert_normals)
      return TRB_ERR_BAD_ARG;
    if (cfg.texture_mode == TRB_TEX_VERTEX ? !vert_colors : !texels) return TRB_ERR_BAD_ARG;
    // the soft blend parks an intermediate in grad_dists
    if (cfg.shader == TRB_SHADER_SOFT_PHONG && grad_zbuf && !grad_dists) return TRB_ERR_BAD_ARG;
  }
  TRB_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long npix = (long long)cfg.N * cfg.H * cfg.W;
  const unsigned blocks = (unsigned)ceil_div64(npix, 256);
  const long long* p2f = (const long long*)pix_to_face;
#define TRB_BWD(SH, L, T)                                                                          \
  shade_backward_kernel<SH, L, T><<<blocks, 256, 0, st>>>(                                          \
      cfg, views, view_params, p2f, bary, zbuf, dists, faces, verts_world, vert_normals, vert_colors, \
      texels, grad_images, grad_bary, grad_zbuf, grad_dists, grad_verts_world, grad_vert_normals,    \
      grad_vert_colors, grad_texels, grad_view_params)
  if (cfg.shader == TRB_SHADER_SOFT_SILHOUETTE) TRB_BWD(TRB_SHADER_SOFT_SILHOUETTE, 0, 0);
  else if (cfg.shader == TRB_SHADER_HARD_PHONG) TRB_DISPATCH_LT(TRB_SHADER_HARD_PHONG, TRB_BWD);
  else TRB_DISPATCH_LT(TRB_SHADER_SOFT_PHONG, TRB_BWD);
#undef TRB_BWD
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

extern "C" int trb_interp_forward(const int64_t* pix_to_face, const float* bary, const float* face_attrs,
                                  int64_t P, int64_t F, int D, float* out, int device,
                                  trb_stream_t stream) {
  if (P < 0 || F < 0 || D < 1) return TRB_ERR_BAD_ARG;
  if (P == 0) return TRB_OK;
  if (!pix_to_face || !bary || !out || (F > 0 && !face_attrs)) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  const unsigned blocks = (unsigned)ceil_div64(P * D, 256);
  interp_forward_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)pix_to_face, bary,
                                                                  face_attrs, P, D, out);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

extern "C" int trb_interp_backward(const int64_t* pix_to_face, const float* bary, const float* face_attrs,
                                   const float* grad_out, int64_t P, int64_t F, int D, float* grad_bary,
                                   float* grad_face_attrs, int device, trb_stream_t stream) {
  if (P < 0 || F < 0 || D < 1) return TRB_ERR_BAD_ARG;
  if (P == 0) return TRB_OK;
  if (!pix_to_face || !bary || !grad_out || !grad_bary || !grad_face_attrs || !face_attrs)
    return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  const unsigned blocks = (unsigned)ceil_div64(P, 256);
  interp_backward_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      (const long long*)pix_to_face, bary, face_attrs, grad_out, P, D, grad_bary, grad_face_attrs);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}
