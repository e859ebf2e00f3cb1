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
  size_t header, count, offset, fill, pairs, total;
};

__host__ inline WsLayout make_ws_layout(int N, const TileGrid& g, int64_t pair_capacity) {
  WsLayout w;
  const size_t ntiles = (size_t)N * g.tiles_x * g.tiles_y;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  w.header = 0;
  w.count = align(64);
  w.fill = w.count + align(ntiles * 4);
  w.offset = w.fill + align(ntiles * 4);
  w.pairs = w.offset + align(ntiles * 4);
  w.total = w.pairs + align((size_t)pair_capacity * 4);
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

// Runs count -> allocate -> fill on `stream`; the workspace must have been laid out with
// make_ws_layout and is zeroed here.  Defined in raster.cu.
int run_binning(const float* verts_ndc, const int* faces, const trb_view* views, int N, int max_face_count,
                int H, int W, const TileGrid& tg, const WsLayout& ws, void* workspace, float sqrt_blur, bool cull,
                long long pair_capacity, cudaStream_t st);

__global__ void write_stats_kernel(const int* __restrict__ header, long long pair_capacity,
                                   int* __restrict__ stats);

}  // namespace trb
