// Per-item device bodies of the small stages around the fine pass: camera transform, camera centre,
// vertex normals, tile binning -- forward and backward.  The stand-alone entry points (transform.cu,
// raster.cu) wrap one body per kernel; the fused render pipeline (render_stages.cu) packs several
// bodies into one launch by block role, which is what takes a render from 10 + 7 graph nodes to 5 + 2.
// Semantics: SURVEY.md A1/A2 (transform), A6 (normals, camera centre), 2c K1-K2 (binning).
#pragma once
#include "raster_internal.cuh"

namespace trb {

// View-space depth of a world-space point: third column of R, third entry of T (A1).  One definition, so that the
// "is any vertex behind the near plane" question (clip.cu) sees the depths the transform writes.
__device__ __forceinline__ float view_depth(float X, float Y, float Z, const float* __restrict__ r,
                                            const float* __restrict__ t) {
  return X * __ldg(r + 2) + Y * __ldg(r + 5) + Z * __ldg(r + 8) + __ldg(t + 2);
}

// ---- world -> view -> NDC of one (view, vertex) -------------------------------------------------
__device__ __forceinline__ void transform_vertex(const float* __restrict__ verts, const float* __restrict__ R,
                                                 const float* __restrict__ T, const float* __restrict__ proj,
                                                 const trb_view& vd, int n, int lv, int perspective,
                                                 float* __restrict__ out) {
  const float* x = verts + 3 * (size_t)(vd.world_vert_start + lv);
  const float* r = R + 9 * (size_t)n;
  const float* t = T + 3 * (size_t)n;
  const float* p = proj + 4 * (size_t)n;
  const float X = __ldg(x), Y = __ldg(x + 1), Z = __ldg(x + 2);
  const float xv = X * __ldg(r + 0) + Y * __ldg(r + 3) + Z * __ldg(r + 6) + __ldg(t + 0);
  const float yv = X * __ldg(r + 1) + Y * __ldg(r + 4) + Z * __ldg(r + 7) + __ldg(t + 1);
  const float zv = view_depth(X, Y, Z, r, t);
  const float den = perspective ? zv : 1.0f;
  float* o = out + 3 * (size_t)(vd.ndc_vert_start + lv);
  o[0] = __ldg(p + 0) * xv / den + __ldg(p + 2);
  o[1] = __ldg(p + 1) * yv / den + __ldg(p + 3);
  o[2] = zv;
}

// Backward of transform_vertex for one (view, vertex) lane; must be called by all 32 lanes of a warp
// (`live` = the lane holds a vertex): vertex gradients are atomically accumulated, the per-view R / T /
// projection gradients are reduced over the warp first (one atomic per warp and value).
__device__ __forceinline__ void transform_vertex_backward(
    const float* __restrict__ verts, const float* __restrict__ R, const float* __restrict__ T,
    const float* __restrict__ proj, const trb_view& vd, int n, int lv, bool live, int perspective,
    const float* __restrict__ grad_ndc, int gstride, float* __restrict__ grad_verts, float* __restrict__ grad_R,
    float* __restrict__ grad_T, float* __restrict__ grad_proj) {
  const float* r = R + 9 * (size_t)n;
  const float* t = T + 3 * (size_t)n;
  const float* p = proj + 4 * (size_t)n;
  float vals[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) vals[i] = 0.0f;
  if (live) {
    const float* x = verts + 3 * (size_t)(vd.world_vert_start + lv);
    const float X = __ldg(x), Y = __ldg(x + 1), Z = __ldg(x + 2);
    const float xv = X * __ldg(r + 0) + Y * __ldg(r + 3) + Z * __ldg(r + 6) + __ldg(t + 0);
    const float yv = X * __ldg(r + 1) + Y * __ldg(r + 4) + Z * __ldg(r + 7) + __ldg(t + 1);
    const float zv = X * __ldg(r + 2) + Y * __ldg(r + 5) + Z * __ldg(r + 8) + __ldg(t + 2);
    const float* g = grad_ndc + (size_t)gstride * (size_t)(vd.ndc_vert_start + lv);
    const float gx = g[0], gy = g[1], gz = g[2];
    const float fx = __ldg(p + 0), fy = __ldg(p + 1);
    float gxv, gyv, gzv = gz, gfx, gfy;
    if (perspective) {
      const float iz = 1.0f / zv;
      gxv = gx * fx * iz; gyv = gy * fy * iz;
      gzv -= (gx * fx * xv + gy * fy * yv) * iz * iz;
      gfx = gx * xv * iz; gfy = gy * yv * iz;
    } else {
      gxv = gx * fx; gyv = gy * fy;
      gfx = gx * xv; gfy = gy * yv;
    }
    // (a vertex no covered sample touched in this view -- about half of a closed mesh -- has nothing to add)
    if (grad_verts && (gx != 0.0f || gy != 0.0f || gz != 0.0f)) {
      float* gv = grad_verts + 3 * (size_t)(vd.world_vert_start + lv);
      atomicAdd(gv + 0, __ldg(r + 0) * gxv + __ldg(r + 1) * gyv + __ldg(r + 2) * gzv);
      atomicAdd(gv + 1, __ldg(r + 3) * gxv + __ldg(r + 4) * gyv + __ldg(r + 5) * gzv);
      atomicAdd(gv + 2, __ldg(r + 6) * gxv + __ldg(r + 7) * gyv + __ldg(r + 8) * gzv);
    }
    vals[0] = X * gxv; vals[1] = X * gyv; vals[2] = X * gzv;
    vals[3] = Y * gxv; vals[4] = Y * gyv; vals[5] = Y * gzv;
    vals[6] = Z * gxv; vals[7] = Z * gyv; vals[8] = Z * gzv;
    vals[9] = gxv; vals[10] = gyv; vals[11] = gzv;
    vals[12] = gfx; vals[13] = gfy; vals[14] = gx; vals[15] = gy;
  }
  if (!grad_R && !grad_T && !grad_proj) return;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float s = warp_sum(vals[i]);
    if (lane == 0 && s != 0.0f) {
      if (i < 9) { if (grad_R) atomicAdd(grad_R + 9 * (size_t)n + i, s); }
      else if (i < 12) { if (grad_T) atomicAdd(grad_T + 3 * (size_t)n + (i - 9), s); }
      else if (grad_proj) atomicAdd(grad_proj + 4 * (size_t)n + (i - 12), s);
    }
  }
}

// ---- camera centre C = -T * inv(R)  (row vectors) ---------------------------------------
__device__ __forceinline__ void inv3(const float* r, float a[9]) {
  const float c00 = r[4] * r[8] - r[5] * r[7], c01 = r[5] * r[6] - r[3] * r[8], c02 = r[3] * r[7] - r[4] * r[6];
  const float det = r[0] * c00 + r[1] * c01 + r[2] * c02;
  const float id = 1.0f / det;
  a[0] = c00 * id; a[1] = (r[2] * r[7] - r[1] * r[8]) * id; a[2] = (r[1] * r[5] - r[2] * r[4]) * id;
  a[3] = c01 * id; a[4] = (r[0] * r[8] - r[2] * r[6]) * id; a[5] = (r[2] * r[3] - r[0] * r[5]) * id;
  a[6] = c02 * id; a[7] = (r[1] * r[6] - r[0] * r[7]) * id; a[8] = (r[0] * r[4] - r[1] * r[3]) * id;
}

__device__ __forceinline__ void camera_center_one(const float* __restrict__ R, const float* __restrict__ T,
                                                  float* __restrict__ vp, int n) {
  float a[9];
  inv3(R + 9 * (size_t)n, a);
  const float* t = T + 3 * (size_t)n;
  float* o = vp + (size_t)n * TRB_VIEW_PARAM_STRIDE + 13;
  o[0] = -(t[0] * a[0] + t[1] * a[3] + t[2] * a[6]);
  o[1] = -(t[0] * a[1] + t[1] * a[4] + t[2] * a[7]);
  o[2] = -(t[0] * a[2] + t[1] * a[5] + t[2] * a[8]);
}

// dC = -dT A - C dR A   =>   gT_i = -sum_k gC_k A_ik ;  gR_ij = -C_i * sum_k A_jk gC_k
// (atomic: the transform backward accumulates into the same grad_R / grad_T concurrently)
__device__ __forceinline__ void camera_center_backward_one(const float* __restrict__ R, const float* __restrict__ vp,
                                                           const float* __restrict__ g_vp, float* __restrict__ gR,
                                                           float* __restrict__ gT, int n) {
  float a[9];
  inv3(R + 9 * (size_t)n, a);
  const float* c = vp + (size_t)n * TRB_VIEW_PARAM_STRIDE + 13;
  const float* g = g_vp + (size_t)n * TRB_VIEW_PARAM_STRIDE + 13;
  float ag[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) ag[j] = a[3 * j] * g[0] + a[3 * j + 1] * g[1] + a[3 * j + 2] * g[2];
  if (gT) {
#pragma unroll
    for (int i = 0; i < 3; ++i) atomicAdd(gT + 3 * (size_t)n + i, -ag[i]);
  }
  if (gR) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) atomicAdd(gR + 9 * (size_t)n + 3 * i + j, -c[i] * ag[j]);
  }
}

// ---- area-weighted vertex normals -------------------------------------------------------------------
__device__ __forceinline__ void face_normal_scatter_one(const float* __restrict__ verts, const int* __restrict__ faces,
                                                        long long f, float* __restrict__ raw) {
  const int i0 = __ldg(faces + 3 * f), i1 = __ldg(faces + 3 * f + 1), i2 = __ldg(faces + 3 * f + 2);
  const float* p0 = verts + 3 * (size_t)i0; const float* p1 = verts + 3 * (size_t)i1;
  const float* p2 = verts + 3 * (size_t)i2;
  const float ax = p2[0] - p1[0], ay = p2[1] - p1[1], az = p2[2] - p1[2];
  const float bx = p0[0] - p1[0], by = p0[1] - p1[1], bz = p0[2] - p1[2];
  const float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
  const int ids[3] = {i0, i1, i2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float* d = raw + 3 * (size_t)ids[k];
    atomicAdd(d, nx); atomicAdd(d + 1, ny); atomicAdd(d + 2, nz);
  }
}

__device__ __forceinline__ void normalize_row_one(const float* __restrict__ raw, long long v, float* __restrict__ out) {
  const float x = raw[3 * v], y = raw[3 * v + 1], z = raw[3 * v + 2];
  const float inv = 1.0f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-6f);
  out[3 * v] = x * inv; out[3 * v + 1] = y * inv; out[3 * v + 2] = z * inv;
}

// gradient of raw / max(|raw|, 1e-6) w.r.t. raw, given the gradient (gx, gy, gz) of the unit normal
__device__ __forceinline__ void normalize_row_backward(float x, float y, float z, float gx, float gy, float gz,
                                                       float& ox, float& oy, float& oz) {
  const float len = sqrtf(x * x + y * y + z * z);
  if (len > 1e-6f) {
    const float inv = 1.0f / len;
    const float ux = x * inv, uy = y * inv, uz = z * inv;
    const float d = ux * gx + uy * gy + uz * gz;
    ox = (gx - ux * d) * inv; oy = (gy - uy * d) * inv; oz = (gz - uz * d) * inv;
  } else {
    ox = gx * 1e6f; oy = gy * 1e6f; oz = gz * 1e6f;
  }
}

// Backward of the face-normal scatter for face f.  (gx, gy, gz) = sum over the three vertices of the
// gradient of their RAW normal.
__device__ __forceinline__ void face_normal_backward_apply(const float* __restrict__ verts, int i0, int i1, int i2,
                                                           float gx, float gy, float gz, float* __restrict__ gverts) {
  const float* p0 = verts + 3 * (size_t)i0; const float* p1 = verts + 3 * (size_t)i1;
  const float* p2 = verts + 3 * (size_t)i2;
  const float ax = p2[0] - p1[0], ay = p2[1] - p1[1], az = p2[2] - p1[2];
  const float bx = p0[0] - p1[0], by = p0[1] - p1[1], bz = p0[2] - p1[2];
  // n = a x b  =>  dL/da = b x g,  dL/db = g x a
  const float gax = by * gz - bz * gy, gay = bz * gx - bx * gz, gaz = bx * gy - by * gx;
  const float gbx = gy * az - gz * ay, gby = gz * ax - gx * az, gbz = gx * ay - gy * ax;
  float* d0 = gverts + 3 * (size_t)i0; float* d1 = gverts + 3 * (size_t)i1; float* d2 = gverts + 3 * (size_t)i2;
  atomicAdd(d2, gax); atomicAdd(d2 + 1, gay); atomicAdd(d2 + 2, gaz);
  atomicAdd(d0, gbx); atomicAdd(d0 + 1, gby); atomicAdd(d0 + 2, gbz);
  atomicAdd(d1, -gax - gbx); atomicAdd(d1 + 1, -gay - gby); atomicAdd(d1 + 2, -gaz - gbz);
}

// ---- tile binning -----------------------------------------------------------------------------------
// One lane per (view, face); must be called by all 32 lanes of a warp (`lf` may be out of range).
// Count (FILL=false) or write (FILL=true) the face into every tile its blur-inflated bounding box can
// touch.  Neighbouring faces of a mesh mostly land in the same tiles, so the lanes that address the same
// tile in the same step are grouped with match.any and issue ONE atomic per group (the same-address
// atomics of the naive version serialised in the L2: 2.2 ms for 4 x 1M faces).  A pair is (face, bits of
// the face's min vertex depth); the K > 1 fine pass orders its tile lists by that depth.
template <bool FILL>
__device__ __forceinline__ void bin_face(const float* __restrict__ verts, const int* __restrict__ faces,
                                         const trb_view& vd, int n, int lf, int H, int W, const TileGrid& tg,
                                         float sqrt_blur, bool cull, int* __restrict__ tile_count,
                                         int* __restrict__ tile_fill, const int* __restrict__ tile_offset,
                                         int2* __restrict__ pairs, float z_cull) {
  const int lane = threadIdx.x & 31;
  int tx0 = 0, ty0 = 0, nx = 0, ny = 0;
  float zmin = 0.0f;
  if (lf < vd.face_count) {
    const FaceXYZ v = load_face(verts, faces, vd, lf);
    if (face_is_drawable(v, cull, z_cull)) {
      const float xmin = min3f(v.x0, v.x1, v.x2) - sqrt_blur, xmax = max3f(v.x0, v.x1, v.x2) + sqrt_blur;
      const float ymin = min3f(v.y0, v.y1, v.y2) - sqrt_blur, ymax = max3f(v.y0, v.y1, v.y2) + sqrt_blur;
      int px0, px1, py0, py1;
      pixel_range(xmin, xmax, W, H, px0, px1);
      pixel_range(ymin, ymax, H, W, py0, py1);
      if (px0 <= px1 && py0 <= py1) {
        tx0 = px0 >> tg.ltx; ty0 = py0 >> tg.lty;
        nx = (px1 >> tg.ltx) - tx0 + 1; ny = (py1 >> tg.lty) - ty0 + 1;
        zmin = min3f(v.z0, v.z1, v.z2);
      }
    }
  }
  const int cnt = nx * ny;
  const int steps = __reduce_max_sync(0xffffffffu, cnt);  // warp-uniform trip count
  const int tbase = n * tg.tiles_x * tg.tiles_y;
  // Neighbouring faces of a mesh cover nearly the same tiles, but lane by lane they reach a given tile at different
  // steps of their own row-major walk, so grouping lanes per step (below) merges few of them.  When the UNION of the
  // warp's tile boxes is small the warp walks that union instead: one ballot and ONE atomic per tile for all lanes
  // that cover it (1M-face sphere with a 15-pixel blur band: ~16 tiles per face, ~25 in the union of 32 faces).
  {
    const int ux0 = __reduce_min_sync(0xffffffffu, cnt > 0 ? tx0 : (1 << 30));
    const int ux1 = __reduce_max_sync(0xffffffffu, cnt > 0 ? tx0 + nx - 1 : -1);
    const int uy0 = __reduce_min_sync(0xffffffffu, cnt > 0 ? ty0 : (1 << 30));
    const int uy1 = __reduce_max_sync(0xffffffffu, cnt > 0 ? ty0 + ny - 1 : -1);
    if (ux1 < ux0) return;   // no lane has a tile
    const int uarea = (ux1 - ux0 + 1) * (uy1 - uy0 + 1);
    if (uarea <= 2 * steps + 8) {
      for (int ty = uy0; ty <= uy1; ++ty)
        for (int tx = ux0; tx <= ux1; ++tx) {
          const bool in = cnt > 0 && tx >= tx0 && tx < tx0 + nx && ty >= ty0 && ty < ty0 + ny;
          const unsigned b = __ballot_sync(0xffffffffu, in);
          if (b == 0) continue;
          const int t = tbase + ty * tg.tiles_x + tx;
          const int leader = __ffs(b) - 1, npeers = __popc(b);
          if (!FILL) {
            if (lane == leader) atomicAdd(tile_count + t, npeers);
          } else {
            const int off = tile_offset[t];
            int base = 0;
            if (lane == leader && off >= 0) base = atomicAdd(tile_fill + t, npeers);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (in && off >= 0)
              pairs[(size_t)off + base + __popc(b & ((1u << lane) - 1u))] = make_int2(lf, __float_as_int(zmin));
          }
        }
      return;
    }
  }
  int ix = 0, iy = 0;
  for (int i = 0; i < steps; ++i) {
    const bool have = i < cnt;
    // lanes without a tile in this step get distinct negative keys: singleton groups, skipped
    const int t = have ? tbase + (ty0 + iy) * tg.tiles_x + tx0 + ix : -1 - lane;
    const unsigned peers = __match_any_sync(0xffffffffu, t);
    if (have) {
      const int leader = __ffs(peers) - 1;
      const int npeers = __popc(peers);
      if (!FILL) {
        if (lane == leader) atomicAdd(tile_count + t, npeers);
      } else {
        const int off = tile_offset[t];
        int base = 0;
        if (lane == leader && off >= 0) base = atomicAdd(tile_fill + t, npeers);
        base = __shfl_sync(peers, base, leader);
        if (off >= 0)
          pairs[(size_t)off + base + __popc(peers & ((1u << lane) - 1u))] = make_int2(lf, __float_as_int(zmin));
      }
      if (++ix == nx) { ix = 0; ++iy; }
    }
  }
}

// Hands every non-empty tile a contiguous slice of `pairs` (order between tiles is irrelevant) and
// appends it to the compact list of non-empty tiles.  One lane per tile; all 32 lanes of a warp.
__device__ __forceinline__ void alloc_tile(const int* __restrict__ tile_count, int* __restrict__ tile_offset,
                                           int ntiles, int* __restrict__ header, long long pair_capacity,
                                           int* __restrict__ busy_list, int t) {
  const int c = t < ntiles ? tile_count[t] : 0;
  const int lane = threadIdx.x & 31;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (lane == 31 && warp_total > 0) {
    // header[0..1] is a 64-bit cursor so that the needed total is exact even past capacity
    base = (int)min((unsigned long long)0x7fffffff,
                    atomicAdd((unsigned long long*)header, (unsigned long long)warp_total));
  }
  base = __shfl_sync(0xffffffffu, base, 31);
  if (t < ntiles) {
    const long long off = (long long)base + (incl - c);
    const bool fits = off + c <= pair_capacity;
    tile_offset[t] = (c == 0) ? 0 : (fits ? (int)off : -1);
    if (c > 0 && !fits) atomicAdd(header + 2, 1);
  }
  const unsigned busy = __ballot_sync(0xffffffffu, c > 0);
  if (busy) {
    int bbase = 0;
    if (lane == 0) bbase = atomicAdd(header + 4, __popc(busy));
    bbase = __shfl_sync(0xffffffffu, bbase, 0);
    if (c > 0) busy_list[bbase + __popc(busy & ((1u << lane) - 1u))] = t;
  }
}

// ---- programmatic dependent launch --------------------------------------------------------------------
// Every kernel of the fused pipeline starts with pdl_wait(): launched with the programmatic-serialisation
// attribute its CTAs may be scheduled while the previous kernel drains, and this is where they wait for
// that kernel's memory to be visible.  Without the attribute it is a no-op.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace trb
