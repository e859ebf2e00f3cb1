// Per-(pixel, face) arithmetic of the rasteriser, SURVEY.md Appendix A3/A4/A9.
//
// Everything that feeds a coverage / ordering decision is written with the round-to-nearest
// intrinsics (__fmul_rn, __fadd_rn, __fsub_rn, __fdiv_rn): ptxas never contracts those into
// FMAs, so the operator sequence is one IEEE fp32 operation per step -- the same sequence the
// CPU oracle (oracle/trb_oracle.c, built with -ffp-contract=off) executes.  That is what makes
// pix_to_face bit-exact between the two, independent of the nvcc flags of the including TU.
#pragma once
#include "trb_common.cuh"

namespace trb {

constexpr float kEps = 1e-8f;

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
// a / b for a clipped (>= 0) numerator and a positive finite b.  A zero numerator -- one or two of the
// three clipped barycentrics of every sample outside its face -- sends div.rn.f32 down its ~100-instruction
// slow path (ncu: 35% of all instructions of the K=8 fine pass); 0 / b is 0 either way.
__device__ __forceinline__ float div_or_zero(float a, float b) { return a == 0.0f ? 0.0f : __fdiv_rn(a, b); }

// A3: NDC coordinate of pixel-centre i along an axis of S1 pixels (other axis S2).
__device__ __forceinline__ float pix_to_ndc(int i, int S1, int S2) {
  float range = 2.0f;
  if (S1 > S2) range = fdiv(fmul((float)S1, range), (float)S2);
  const float offset = fdiv(range, 2.0f);
  return fadd(-offset, fdiv(fadd(fmul(range, (float)i), offset), (float)S1));
}

// A4.1: edge(p; a, b) = (p.x-a.x)(b.y-a.y) - (p.y-a.y)(b.x-a.x)
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return fsub(fmul(fsub(px, ax), fsub(by, ay)), fmul(fsub(py, ay), fsub(bx, ax)));
}

__device__ __forceinline__ float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// A4.7: squared distance from p to the segment ab.
__device__ __forceinline__ float point_segment_d2(float px, float py, float ax, float ay, float bx,
                                                  float by) {
  const float bax = fsub(bx, ax), bay = fsub(by, ay);
  const float l2 = fadd(fmul(bax, bax), fmul(bay, bay));
  if (l2 <= kEps) {
    const float dx = fsub(px, bx), dy = fsub(py, by);
    return fadd(fmul(dx, dx), fmul(dy, dy));
  }
  float t = fdiv(fadd(fmul(bax, fsub(px, ax)), fmul(bay, fsub(py, ay))), l2);
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  const float qx = fadd(ax, fmul(t, bax)), qy = fadd(ay, fmul(t, bay));
  const float dx = fsub(qx, px), dy = fsub(qy, py);
  return fadd(fmul(dx, dx), fmul(dy, dy));
}

struct FaceXYZ {
  float x0, y0, z0, x1, y1, z1, x2, y2, z2;
};

// Per-face (pixel independent) validity, A4.2 minus the bbox/pixel part.
// `z_cull` = max(0, z_clip_value): a face whose three vertices are all nearer than the clip plane is
// culled, which is what PyTorch3D's clip_faces does to such faces before rasterising (z_cull = 0
// reproduces the rasteriser's own `zmax < 0` test).
__device__ __forceinline__ bool face_is_drawable(const FaceXYZ& v, bool cull_backfaces, float z_cull = 0.0f) {
  const float zmin = min3f(v.z0, v.z1, v.z2), zmax = max3f(v.z0, v.z1, v.z2);
  const float face_area = edge_fn(v.x0, v.y0, v.x1, v.y1, v.x2, v.y2);
  const bool back_face = face_area < 0.0f;
  const bool zero_area = (face_area <= kEps) && (face_area >= -kEps);
  // written so that NaN coordinates make the face undrawable
  if (!(zmin >= kEps)) return false;
  if (zmax < z_cull || zero_area || (cull_backfaces && back_face)) return false;
  return true;
}

struct Sample {
  float z, d, c0, c1, c2;
};

// A4 steps 3-9 for a face already known to be drawable and whose inflated bbox contains p.
// Returns false when the pixel is not a candidate of the face.
template <bool PERSP, bool CLIP>
__device__ __forceinline__ bool eval_pixel_face(const FaceXYZ& v, float px, float py,
                                                float blur_radius, Sample& out) {
  const float area = fadd(edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1), kEps);
  const float w0 = fdiv(edge_fn(px, py, v.x1, v.y1, v.x2, v.y2), area);
  const float w1 = fdiv(edge_fn(px, py, v.x2, v.y2, v.x0, v.y0), area);
  const float w2 = fdiv(edge_fn(px, py, v.x0, v.y0, v.x1, v.y1), area);
  float b0 = w0, b1 = w1, b2 = w2;
  if (PERSP) {
    const float t0 = fmul(fmul(w0, v.z1), v.z2);
    const float t1 = fmul(fmul(w1, v.z0), v.z2);
    const float t2 = fmul(fmul(w2, v.z0), v.z1);
    const float den = fmaxf(fadd(fadd(t0, t1), t2), kEps);
    b0 = fdiv(t0, den); b1 = fdiv(t1, den); b2 = fdiv(t2, den);
  }
  float c0 = b0, c1 = b1, c2 = b2;
  if (CLIP) {
    c0 = fmaxf(b0, 0.0f); c1 = fmaxf(b1, 0.0f); c2 = fmaxf(b2, 0.0f);
    const float s = fmaxf(fadd(fadd(c0, c1), c2), 1e-5f);
    c0 = div_or_zero(c0, s); c1 = div_or_zero(c1, s); c2 = div_or_zero(c2, s);
  }
  const float pz = fadd(fadd(fmul(c0, v.z0), fmul(c1, v.z1)), fmul(c2, v.z2));
  if (pz < 0.0f) return false;
  const bool inside = (b0 > 0.0f) && (b1 > 0.0f) && (b2 > 0.0f);
  const float e01 = point_segment_d2(px, py, v.x0, v.y0, v.x1, v.y1);
  const float e02 = point_segment_d2(px, py, v.x0, v.y0, v.x2, v.y2);
  const float e12 = point_segment_d2(px, py, v.x1, v.y1, v.x2, v.y2);
  const float dist = min3f(e01, e02, e12);
  if (!inside && dist >= blur_radius) return false;
  out.z = pz; out.d = inside ? -dist : dist;
  out.c0 = c0; out.c1 = c1; out.c2 = c2;
  return true;
}

// Same arithmetic as eval_pixel_face, split so that the fused kernel can (a) reuse the three edge
// functions it already evaluated for its early reject and (b) postpone the distance when the
// blur radius is zero (then only strictly-inside samples survive and the distance of the
// winner is all that is ever needed).  `area` = edge(v2; v0, v1) + kEps, exactly as above.
__device__ __forceinline__ bool eval_from_edges(const FaceXYZ& v, float area, float e0, float e1, float e2,
                                                bool persp, bool clip, float& pz, float& c0, float& c1,
                                                float& c2, bool& inside) {
  const float w0 = fdiv(e0, area), w1 = fdiv(e1, area), w2 = fdiv(e2, area);
  float b0 = w0, b1 = w1, b2 = w2;
  if (persp) {
    const float t0 = fmul(fmul(w0, v.z1), v.z2);
    const float t1 = fmul(fmul(w1, v.z0), v.z2);
    const float t2 = fmul(fmul(w2, v.z0), v.z1);
    const float den = fmaxf(fadd(fadd(t0, t1), t2), kEps);
    b0 = fdiv(t0, den); b1 = fdiv(t1, den); b2 = fdiv(t2, den);
  }
  c0 = b0; c1 = b1; c2 = b2;
  if (clip) {
    c0 = fmaxf(b0, 0.0f); c1 = fmaxf(b1, 0.0f); c2 = fmaxf(b2, 0.0f);
    const float s = fmaxf(fadd(fadd(c0, c1), c2), 1e-5f);
    c0 = div_or_zero(c0, s); c1 = div_or_zero(c1, s); c2 = div_or_zero(c2, s);
  }
  pz = fadd(fadd(fmul(c0, v.z0), fmul(c1, v.z1)), fmul(c2, v.z2));
  inside = (b0 > 0.0f) && (b1 > 0.0f) && (b2 > 0.0f);
  return !(pz < 0.0f);
}

__device__ __forceinline__ float triangle_d2(const FaceXYZ& v, float px, float py) {
  const float e01 = point_segment_d2(px, py, v.x0, v.y0, v.x1, v.y1);
  const float e02 = point_segment_d2(px, py, v.x0, v.y0, v.x2, v.y2);
  const float e12 = point_segment_d2(px, py, v.x1, v.y1, v.x2, v.y2);
  return min3f(e01, e02, e12);
}

// Runtime-flag variant of eval_pixel_face (identical operator sequence).
__device__ __forceinline__ bool eval_pixel_face_rt(const FaceXYZ& v, float px, float py, bool persp,
                                                   bool clip, float blur_radius, Sample& out) {
  const float area = fadd(edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1), kEps);
  const float e0 = edge_fn(px, py, v.x1, v.y1, v.x2, v.y2);
  const float e1 = edge_fn(px, py, v.x2, v.y2, v.x0, v.y0);
  const float e2 = edge_fn(px, py, v.x0, v.y0, v.x1, v.y1);
  bool inside;
  if (!eval_from_edges(v, area, e0, e1, e2, persp, clip, out.z, out.c0, out.c1, out.c2, inside)) return false;
  const float dist = triangle_d2(v, px, py);
  if (!inside && dist >= blur_radius) return false;
  out.d = inside ? -dist : dist;
  return true;
}

// (z, face) lexicographic order, A5.
__device__ __forceinline__ bool cand_less(float za, int fa, float zb, int fb) {
  return (za < zb) || (za == zb && fa < fb);
}

// ------------------------------------------------------------------------------------------
// Backward of one sample (A9).  g[9] receives d loss / d (x0,y0,z0,x1,y1,z1,x2,y2,z2).
// The forward quantities are recomputed with the exact forward sequence so that `inside`
// and the clamps take the same branch as in the forward pass; the gradient arithmetic itself
// is ordinary fp32 (FMAs allowed).
__device__ __forceinline__ void edge_bwd(float px, float py, float ax, float ay, float bx, float by,
                                         float g, float& gax, float& gay, float& gbx, float& gby) {
  gax += g * (py - by); gay += g * (bx - px);
  gbx += g * (ay - py); gby += g * (px - ax);
}

__device__ __forceinline__ void point_segment_bwd(float px, float py, float ax, float ay, float bx,
                                                  float by, float g, float& gax, float& gay,
                                                  float& gbx, float& gby) {
  const float bax = fsub(bx, ax), bay = fsub(by, ay);
  const float l2 = fadd(fmul(bax, bax), fmul(bay, bay));
  if (l2 <= kEps) {
    gbx += g * 2.0f * (bx - px); gby += g * 2.0f * (by - py);
    return;
  }
  float t = fdiv(fadd(fmul(bax, fsub(px, ax)), fmul(bay, fsub(py, ay))), l2);
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  const float dx = (ax + t * bax) - px, dy = (ay + t * bay) - py;
  const float ga = g * (1.0f - t) * 2.0f, gb = g * t * 2.0f;
  gax += ga * dx; gay += ga * dy;
  gbx += gb * dx; gby += gb * dy;
}

__device__ __forceinline__ void sample_backward_rt(const FaceXYZ& v, float px, float py, bool PERSP,
                                                   bool CLIP, float gz, float gb0, float gb1, float gb2,
                                                   float gd, float g[9]) {
  const float area = fadd(edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1), kEps);
  const float e0 = edge_fn(px, py, v.x1, v.y1, v.x2, v.y2);
  const float e1 = edge_fn(px, py, v.x2, v.y2, v.x0, v.y0);
  const float e2 = edge_fn(px, py, v.x0, v.y0, v.x1, v.y1);
  const float w0 = fdiv(e0, area), w1 = fdiv(e1, area), w2 = fdiv(e2, area);
  float b0 = w0, b1 = w1, b2 = w2, den = 1.0f, t0 = 0.f, t1 = 0.f, t2 = 0.f, tsum = 0.f;
  if (PERSP) {
    t0 = fmul(fmul(w0, v.z1), v.z2);
    t1 = fmul(fmul(w1, v.z0), v.z2);
    t2 = fmul(fmul(w2, v.z0), v.z1);
    tsum = fadd(fadd(t0, t1), t2);
    den = fmaxf(tsum, kEps);
    b0 = fdiv(t0, den); b1 = fdiv(t1, den); b2 = fdiv(t2, den);
  }
  float c0 = b0, c1 = b1, c2 = b2, m0 = b0, m1 = b1, m2 = b2, s = 1.0f, ssum = 0.f;
  if (CLIP) {
    m0 = fmaxf(b0, 0.0f); m1 = fmaxf(b1, 0.0f); m2 = fmaxf(b2, 0.0f);
    ssum = fadd(fadd(m0, m1), m2);
    s = fmaxf(ssum, 1e-5f);
    c0 = div_or_zero(m0, s); c1 = div_or_zero(m1, s); c2 = div_or_zero(m2, s);
  }
  const bool inside = (b0 > 0.0f) && (b1 > 0.0f) && (b2 > 0.0f);
#pragma unroll
  for (int i = 0; i < 9; ++i) g[i] = 0.0f;
  // zbuf = c . z
  g[2] += gz * c0; g[5] += gz * c1; g[8] += gz * c2;
  const float gc0 = gb0 + gz * v.z0, gc1 = gb1 + gz * v.z1, gc2 = gb2 + gz * v.z2;
  float gbb0 = gc0, gbb1 = gc1, gbb2 = gc2;
  if (CLIP) {
    const float inv_s = 1.0f / s;
    float gs = -(gc0 * m0 + gc1 * m1 + gc2 * m2) * inv_s * inv_s;
    if (!(ssum > 1e-5f)) gs = 0.0f;
    gbb0 = b0 > 0.0f ? gc0 * inv_s + gs : 0.0f;
    gbb1 = b1 > 0.0f ? gc1 * inv_s + gs : 0.0f;
    gbb2 = b2 > 0.0f ? gc2 * inv_s + gs : 0.0f;
  }
  float gw0 = gbb0, gw1 = gbb1, gw2 = gbb2;
  if (PERSP) {
    const float inv_den = 1.0f / den;
    float gden = -(gbb0 * t0 + gbb1 * t1 + gbb2 * t2) * inv_den * inv_den;
    if (!(tsum > kEps)) gden = 0.0f;
    const float gt0 = gbb0 * inv_den + gden, gt1 = gbb1 * inv_den + gden, gt2 = gbb2 * inv_den + gden;
    gw0 = gt0 * v.z1 * v.z2; gw1 = gt1 * v.z0 * v.z2; gw2 = gt2 * v.z0 * v.z1;
    g[2] += gt1 * w1 * v.z2 + gt2 * w2 * v.z1;
    g[5] += gt0 * w0 * v.z2 + gt2 * w2 * v.z0;
    g[8] += gt0 * w0 * v.z1 + gt1 * w1 * v.z0;
  }
  const float inv_area = 1.0f / area;
  const float ge0 = gw0 * inv_area, ge1 = gw1 * inv_area, ge2 = gw2 * inv_area;
  const float garea = -(gw0 * e0 + gw1 * e1 + gw2 * e2) * inv_area * inv_area;
  edge_bwd(px, py, v.x1, v.y1, v.x2, v.y2, ge0, g[3], g[4], g[6], g[7]);
  edge_bwd(px, py, v.x2, v.y2, v.x0, v.y0, ge1, g[6], g[7], g[0], g[1]);
  edge_bwd(px, py, v.x0, v.y0, v.x1, v.y1, ge2, g[0], g[1], g[3], g[4]);
  g[6] += garea * (v.y1 - v.y0); g[7] += garea * (v.x0 - v.x1);
  g[0] += garea * (v.y2 - v.y1); g[1] += garea * (v.x1 - v.x2);
  g[3] += garea * (v.y0 - v.y2); g[4] += garea * (v.x2 - v.x0);
  if (gd != 0.0f) {
    const float e01 = point_segment_d2(px, py, v.x0, v.y0, v.x1, v.y1);
    const float e02 = point_segment_d2(px, py, v.x0, v.y0, v.x2, v.y2);
    const float e12 = point_segment_d2(px, py, v.x1, v.y1, v.x2, v.y2);
    const float gsd = inside ? -gd : gd;
    if (e01 <= e02 && e01 <= e12)
      point_segment_bwd(px, py, v.x0, v.y0, v.x1, v.y1, gsd, g[0], g[1], g[3], g[4]);
    else if (e02 <= e01 && e02 <= e12)
      point_segment_bwd(px, py, v.x0, v.y0, v.x2, v.y2, gsd, g[0], g[1], g[6], g[7]);
    else
      point_segment_bwd(px, py, v.x1, v.y1, v.x2, v.y2, gsd, g[3], g[4], g[6], g[7]);
  }
}

template <bool PERSP, bool CLIP>
__device__ __forceinline__ void sample_backward(const FaceXYZ& v, float px, float py, float gz,
                                                float gb0, float gb1, float gb2, float gd,
                                                float g[9]) {
  sample_backward_rt(v, px, py, PERSP, CLIP, gz, gb0, gb1, gb2, gd, g);
}

// ------------------------------------------------------------------------------------------
// Warp-aggregated scatter: lanes of a warp that hold contributions for the same `key`
// (a face id; < 0 = lane has nothing) are summed with shuffles and issued as ONE atomicAdd per value.
// When the warp holds many distinct keys (small-triangle regime) the per-lane atomics hit distinct
// addresses anyway, so they are issued directly.  The grouping (one match.any) can be shared by
// several scatters that use the same key.
struct WarpGroups {
  unsigned leaders;   // ballot of group leaders that hold a valid key
  bool any;           // some lane has a key
  bool aggregate;     // full warp and few groups: reduce with shuffles
  // runs of CONSECUTIVE lanes with the same key (neighbouring covered pixels of one face): a segmented suffix sum
  // over each run lets its first lane issue one reduction for the whole run
  bool runs;          // full warp and at least one run longer than one lane
  bool run_head;      // this lane starts a run of a valid key
  int run_end;        // last lane of this lane's run
  int run_steps;      // ceil(log2(longest run))
};

__device__ __forceinline__ WarpGroups warp_groups(int key, bool allow_runs = false) {
  WarpGroups wg;
  const unsigned active = __activemask();
  const unsigned have = __ballot_sync(active, key >= 0);
  wg.any = have != 0;
  wg.leaders = 0; wg.aggregate = false; wg.runs = false; wg.run_head = false; wg.run_end = 0; wg.run_steps = 0;
  if (!wg.any) return wg;
  const unsigned peers = __match_any_sync(active, key);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  wg.leaders = __ballot_sync(active, (lane == leader) && key >= 0);
  // shuffle-reduce only when it really merges lanes: few groups AND several lanes per group (a layer that only
  // two or three lanes of the warp still have is cheaper as direct reductions than as 45 shuffles per group)
  wg.aggregate = (active == 0xffffffffu) && (__popc(wg.leaders) <= 4) && (__popc(have) >= 4 * __popc(wg.leaders));
  if (allow_runs && !wg.aggregate && active == 0xffffffffu) {
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
    const unsigned later = (lane == 31) ? 0u : (heads >> (lane + 1));
    wg.run_end = later ? lane + __ffs(later) - 1 : 31;
    wg.run_head = ((heads >> lane) & 1u) && key >= 0;
    // worth it only when the runs remove a good part of the warp's reductions (each costs ~3 L2 sector-ops per
    // scatter; the segmented sum costs ~20 shuffles per step): at least 8 lanes merged away
    const unsigned valid_heads = __ballot_sync(0xffffffffu, wg.run_head);
    if (__popc(have) - __popc(valid_heads) >= 8) {
      // longest run (only runs of valid keys matter, but an upper bound is fine)
      int len = wg.run_head ? wg.run_end - lane + 1 : 1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
      wg.runs = true;
      wg.run_steps = 32 - __clz(len - 1);  // ceil(log2(len)), len >= 2
    }
  }
  return wg;
}

template <int NV>
__device__ __forceinline__ void warp_groups_add(const WarpGroups& wg, int key, const float (&val)[NV],
                                                float* const (&dst)[NV]) {
  if (!wg.any) return;
  if (wg.aggregate) {
    const int lane = threadIdx.x & 31;
    unsigned todo = wg.leaders;
    while (todo) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1;
      const int k = __shfl_sync(0xffffffffu, key, l);
      const bool mine = (key == k);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float s = mine ? val[i] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == l && s != 0.0f) atomicAdd(dst[i], s);
      }
    }
  } else if (key >= 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (val[i] != 0.0f) atomicAdd(dst[i], val[i]);
  }
}

// Same, for three xyz triples going to rows i0, i1, i2 of a float4-strided accumulator: one
// REDG.ADD.F32x4 per vertex instead of three scalar reductions (the fp32 reduction rate of the L2 --
// measured ~60 G sector-ops/s on B200 -- is what bounds the backward scatter).
__device__ __forceinline__ void warp_groups_add_xyz3(const WarpGroups& wg, int key, const float (&val)[9],
                                                     float4* base, int i0, int i1, int i2) {
  if (!wg.any) return;
  if (wg.aggregate) {
    const int lane = threadIdx.x & 31;
    unsigned todo = wg.leaders;
    while (todo) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1;
      const int k = __shfl_sync(0xffffffffu, key, l);
      const bool mine = (key == k);
      float s[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        float t = mine ? val[i] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        s[i] = t;
      }
      if (lane == l) {
        atomicAdd(base + i0, make_float4(s[0], s[1], s[2], 0.0f));
        atomicAdd(base + i1, make_float4(s[3], s[4], s[5], 0.0f));
        atomicAdd(base + i2, make_float4(s[6], s[7], s[8], 0.0f));
      }
    }
  } else if (wg.runs) {
    // segmented suffix sum over runs of equal keys in consecutive lanes; the first lane of a run issues
    const int lane = threadIdx.x & 31;
    float s[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) s[i] = key >= 0 ? val[i] : 0.0f;
    for (int step = 0, d = 1; step < wg.run_steps; ++step, d <<= 1) {
      const bool take = lane + d <= wg.run_end;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const float t = __shfl_down_sync(0xffffffffu, s[i], d);
        if (take) s[i] += t;
      }
    }
    if (wg.run_head) {
      atomicAdd(base + i0, make_float4(s[0], s[1], s[2], 0.0f));
      atomicAdd(base + i1, make_float4(s[3], s[4], s[5], 0.0f));
      atomicAdd(base + i2, make_float4(s[6], s[7], s[8], 0.0f));
    }
  } else if (key >= 0) {
    atomicAdd(base + i0, make_float4(val[0], val[1], val[2], 0.0f));
    atomicAdd(base + i1, make_float4(val[3], val[4], val[5], 0.0f));
    atomicAdd(base + i2, make_float4(val[6], val[7], val[8], 0.0f));
  }
}

template <int NV>
__device__ __forceinline__ void warp_aggregated_add(int key, const float (&val)[NV],
                                                    float* const (&dst)[NV]) {
  const WarpGroups wg = warp_groups(key);
  warp_groups_add<NV>(wg, key, val, dst);
}

// ------------------------------------------------------------------------------------------
// Fast-math variants for backward passes.  Gradients are compared with fp64 autograd at 1e-3
// relative L2, so nothing here has to reproduce the forward's IEEE sequence; the one decision
// that matters -- whether the sample is inside its face -- is read off the sign of the saved
// signed distance instead of being recomputed.
__device__ __forceinline__ float pix_to_ndc_fast(int i, int S1, int S2) {
  const float range = (S1 > S2) ? 2.0f * (float)S1 / (float)S2 : 2.0f;
  return -0.5f * range + (range * (float)i + 0.5f * range) / (float)S1;
}

__device__ __forceinline__ float seg_d2_fast(float px, float py, float ax, float ay, float bx, float by,
                                             float& t_out, bool& degenerate) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = bax * bax + bay * bay;
  degenerate = l2 <= kEps;
  if (degenerate) {
    t_out = 1.0f;
    const float dx = px - bx, dy = py - by;
    return dx * dx + dy * dy;
  }
  float t = __fdividef(bax * (px - ax) + bay * (py - ay), l2);
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  t_out = t;
  const float dx = ax + t * bax - px, dy = ay + t * bay - py;
  return dx * dx + dy * dy;
}

__device__ __forceinline__ void sample_backward_fast(const FaceXYZ& v, float px, float py, bool persp,
                                                     bool clip, bool inside, float gz, float gb0, float gb1,
                                                     float gb2, float gd, float g[9]) {
  const float area = edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1) + kEps;
  const float inv_area = __frcp_rn(area);
  const float e0 = (px - v.x1) * (v.y2 - v.y1) - (py - v.y1) * (v.x2 - v.x1);
  const float e1 = (px - v.x2) * (v.y0 - v.y2) - (py - v.y2) * (v.x0 - v.x2);
  const float e2 = (px - v.x0) * (v.y1 - v.y0) - (py - v.y0) * (v.x1 - v.x0);
  const float w0 = e0 * inv_area, w1 = e1 * inv_area, w2 = e2 * inv_area;
  float b0 = w0, b1 = w1, b2 = w2, inv_den = 1.0f, t0 = 0.f, t1 = 0.f, t2 = 0.f, tsum = 1.0f;
  if (persp) {
    t0 = w0 * v.z1 * v.z2; t1 = w1 * v.z0 * v.z2; t2 = w2 * v.z0 * v.z1;
    tsum = t0 + t1 + t2;
    inv_den = __frcp_rn(fmaxf(tsum, kEps));
    b0 = t0 * inv_den; b1 = t1 * inv_den; b2 = t2 * inv_den;
  }
  float c0 = b0, c1 = b1, c2 = b2, m0 = b0, m1 = b1, m2 = b2, inv_s = 1.0f, ssum = 1.0f;
  if (clip) {
    m0 = fmaxf(b0, 0.0f); m1 = fmaxf(b1, 0.0f); m2 = fmaxf(b2, 0.0f);
    ssum = m0 + m1 + m2;
    inv_s = __frcp_rn(fmaxf(ssum, 1e-5f));
    c0 = m0 * inv_s; c1 = m1 * inv_s; c2 = m2 * inv_s;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) g[i] = 0.0f;
  g[2] = gz * c0; g[5] = gz * c1; g[8] = gz * c2;
  const float gc0 = gb0 + gz * v.z0, gc1 = gb1 + gz * v.z1, gc2 = gb2 + gz * v.z2;
  float gbb0 = gc0, gbb1 = gc1, gbb2 = gc2;
  if (clip) {
    float gs = -(gc0 * m0 + gc1 * m1 + gc2 * m2) * inv_s * inv_s;
    if (!(ssum > 1e-5f)) gs = 0.0f;
    gbb0 = b0 > 0.0f ? gc0 * inv_s + gs : 0.0f;
    gbb1 = b1 > 0.0f ? gc1 * inv_s + gs : 0.0f;
    gbb2 = b2 > 0.0f ? gc2 * inv_s + gs : 0.0f;
  }
  float gw0 = gbb0, gw1 = gbb1, gw2 = gbb2;
  if (persp) {
    float gden = -(gbb0 * t0 + gbb1 * t1 + gbb2 * t2) * inv_den * inv_den;
    if (!(tsum > kEps)) gden = 0.0f;
    const float gt0 = gbb0 * inv_den + gden, gt1 = gbb1 * inv_den + gden, gt2 = gbb2 * inv_den + gden;
    gw0 = gt0 * v.z1 * v.z2; gw1 = gt1 * v.z0 * v.z2; gw2 = gt2 * v.z0 * v.z1;
    g[2] += gt1 * w1 * v.z2 + gt2 * w2 * v.z1;
    g[5] += gt0 * w0 * v.z2 + gt2 * w2 * v.z0;
    g[8] += gt0 * w0 * v.z1 + gt1 * w1 * v.z0;
  }
  const float ge0 = gw0 * inv_area, ge1 = gw1 * inv_area, ge2 = gw2 * inv_area;
  const float garea = -(gw0 * e0 + gw1 * e1 + gw2 * e2) * inv_area * inv_area;
  edge_bwd(px, py, v.x1, v.y1, v.x2, v.y2, ge0, g[3], g[4], g[6], g[7]);
  edge_bwd(px, py, v.x2, v.y2, v.x0, v.y0, ge1, g[6], g[7], g[0], g[1]);
  edge_bwd(px, py, v.x0, v.y0, v.x1, v.y1, ge2, g[0], g[1], g[3], g[4]);
  g[6] += garea * (v.y1 - v.y0); g[7] += garea * (v.x0 - v.x1);
  g[0] += garea * (v.y2 - v.y1); g[1] += garea * (v.x1 - v.x2);
  g[3] += garea * (v.y0 - v.y2); g[4] += garea * (v.x2 - v.x0);
  if (gd != 0.0f) {
    float t01, t02, t12;
    bool d01, d02, d12;
    const float e01 = seg_d2_fast(px, py, v.x0, v.y0, v.x1, v.y1, t01, d01);
    const float e02 = seg_d2_fast(px, py, v.x0, v.y0, v.x2, v.y2, t02, d02);
    const float e12 = seg_d2_fast(px, py, v.x1, v.y1, v.x2, v.y2, t12, d12);
    const float gsd = inside ? -gd : gd;
    // arg-min edge (a, b): d/da = 2(1-t)(proj-p), d/db = 2t(proj-p); degenerate: d/db = 2(b-p)
    float ax, ay, bx, by, t; bool deg; int ia, ib;
    if (e01 <= e02 && e01 <= e12) { ax = v.x0; ay = v.y0; bx = v.x1; by = v.y1; t = t01; deg = d01; ia = 0; ib = 3; }
    else if (e02 <= e01 && e02 <= e12) { ax = v.x0; ay = v.y0; bx = v.x2; by = v.y2; t = t02; deg = d02; ia = 0; ib = 6; }
    else { ax = v.x1; ay = v.y1; bx = v.x2; by = v.y2; t = t12; deg = d12; ia = 3; ib = 6; }
    const float qx = deg ? bx : ax + t * (bx - ax), qy = deg ? by : ay + t * (by - ay);
    const float dx = qx - px, dy = qy - py;
    const float ga = deg ? 0.0f : gsd * (1.0f - t) * 2.0f, gb = gsd * t * 2.0f;
    // (ia, ib) are compile-time constants per branch after inlining; written generically here
    if (ia == 0) { g[0] += ga * dx; g[1] += ga * dy; } else { g[3] += ga * dx; g[4] += ga * dy; }
    if (ib == 3) { g[3] += gb * dx; g[4] += gb * dy; } else { g[6] += gb * dx; g[7] += gb * dy; }
  }
}

}  // namespace trb
