// Pieces of the rasteriser shared between raster.cu (stand-alone entry points) and render.cu (fused
// render pipeline): tile grid / workspace layout, the face fetch, and the binning launcher.
#pragma once
#include "raster_math.cuh"

namespace trb {

struct TileGrid {
  int tiles_x, tiles_y, ltx, lty;
};

__host__ inline TileGrid make_tile_grid(int H, int W, int K) {
  TileGrid g;
  if (K <= 24) { g.ltx = 4; g.lty = 4; } else { g.ltx = 3; g.lty = 3; }
  g.tiles_x = (W + (1 << g.ltx) - 1) >> g.ltx;
  g.tiles_y = (H + (1 << g.lty) - 1) >> g.lty;
  return g;
}

struct WsLayout {
  size_t header, count, offset, fill, busy, pairs, total;
};

__host__ inline WsLayout make_ws_layout(int N, const TileGrid& g, int64_t pair_capacity) {
  WsLayout w;
  const size_t ntiles = (size_t)N * g.tiles_x * g.tiles_y;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  w.header = 0;
  w.count = align(64);
  w.fill = w.count + align(ntiles * 4);
  w.offset = w.fill + align(ntiles * 4);
  w.busy = w.offset + align(ntiles * 4);   // compact list of non-empty tiles (header[4] = its length)
  w.pairs = w.busy + align(ntiles * 4);
  w.total = w.pairs + align((size_t)pair_capacity * 8);  // int2 (face, bits of the min vertex z)
  return w;
}

__device__ __forceinline__ FaceXYZ load_face(const float* __restrict__ verts,
                                             const int* __restrict__ faces, const trb_view& vd,
                                             int local_face) {
  const int r = vd.face_start + local_face;
  int i0, i1, i2;
  if (faces != nullptr) {
    i0 = __ldg(faces + 3 * (size_t)r) + vd.vert_delta;
    i1 = __ldg(faces + 3 * (size_t)r + 1) + vd.vert_delta;
    i2 = __ldg(faces + 3 * (size_t)r + 2) + vd.vert_delta;
  } else {
    i0 = 3 * r; i1 = i0 + 1; i2 = i0 + 2;
  }
  FaceXYZ v;
  const float* p0 = verts + 3 * (size_t)i0;
  const float* p1 = verts + 3 * (size_t)i1;
  const float* p2 = verts + 3 * (size_t)i2;
  v.x0 = __ldg(p0); v.y0 = __ldg(p0 + 1); v.z0 = __ldg(p0 + 2);
  v.x1 = __ldg(p1); v.y1 = __ldg(p1 + 1); v.z1 = __ldg(p1 + 2);
  v.x2 = __ldg(p2); v.y2 = __ldg(p2 + 1); v.z2 = __ldg(p2 + 2);
  return v;
}

// Conservative range of pixel indices (in output order) whose centre can lie in [lo, hi].
// Pixel-centre i' = S-1-i has NDC coordinate -off + (range*i' + off)/S  (A3), i.e. fractional index
// (x + off) * S / range - 0.5.  The 0.02 px of slack is ~10x the fp32 error of that expression for
// images up to 4096 px; the exact per-pixel bbox test of A4.2 is still applied afterwards.
__device__ __forceinline__ void pixel_range(float lo, float hi, int S1, int S2, int& p_lo, int& p_hi) {
  float range = 2.0f;
  if (S1 > S2) range = (float)S1 * 2.0f / (float)S2;
  const float off = 0.5f * range;
  const float scale = (float)S1 / range;
  float a = (lo + off) * scale - 0.5f;
  float b = (hi + off) * scale - 0.5f;
  a = fminf(fmaxf(a, -2.0f), (float)S1 + 1.0f);
  b = fminf(fmaxf(b, -2.0f), (float)S1 + 1.0f);
  int i_lo = (int)ceilf(a - 0.02f), i_hi = (int)floorf(b + 0.02f);
  i_lo = max(i_lo, 0); i_hi = min(i_hi, S1 - 1);
  p_lo = S1 - 1 - i_hi; p_hi = S1 - 1 - i_lo;
}

// Runs count -> allocate -> fill on `stream`; the workspace must have been laid out with
// make_ws_layout and is zeroed here.  Defined in raster.cu.
int run_binning(const float* verts_ndc, const int* faces, const trb_view* views, int N, int max_face_count,
                int H, int W, const TileGrid& tg, const WsLayout& ws, void* workspace, float sqrt_blur, bool cull,
                long long pair_capacity, cudaStream_t st, float z_cull = 0.0f);

__global__ void write_stats_kernel(const int* __restrict__ header, long long pair_capacity,
                                   int* __restrict__ stats);

}  // namespace trb
