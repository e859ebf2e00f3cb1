// Stand-alone rasteriser entry points for sm_100a: per-tile face binning (count -> allocate -> fill), the fine
// pass (the fused kernels of render.cu / render_kn.cu with the shading compiled out) and the backward scatter.
// Replaces pytorch3d._C.rasterize_meshes / rasterize_meshes_backward (reference call sites: torch_renderer.py:113,
// camera_pose_optimizer.py:244, batch_rendering_test.py:274); semantics: SURVEY.md Appendix A3-A5, A9.
//
// Differences from the upstream design, all deliberate:
//  * bin lists are compact (a global cursor hands every tile exactly `count` slots) instead of
//    N*BH*BW*M fixed slots, so the fine pass never scans sentinels; entries carry the face's min depth;
//  * a tile whose list does not fit the workspace is rasterised by scanning the whole mesh --
//    faces are never dropped;
//  * the per-pixel queue keeps only (z, face) -- 8 B per entry in shared memory -- and barycentrics /
//    distances are recomputed for the K survivors;
//  * ties are broken by (z, face index), so the result does not depend on list order.
#include "render_internal.cuh"
#include "stages.cuh"

namespace trb {

// ------------------------------------------------------------------------------------------
// Binning (bodies in stages.cuh): count -> allocate -> fill.
template <bool FILL>
__global__ void __launch_bounds__(256)
bin_faces_kernel(const float* __restrict__ verts, const int* __restrict__ faces,
                 const trb_view* __restrict__ views, int H, int W, TileGrid tg, float sqrt_blur,
                 bool cull, int* __restrict__ tile_count, int* __restrict__ tile_fill,
                 const int* __restrict__ tile_offset, int2* __restrict__ pairs, float z_cull) {
  const int n = blockIdx.y;
  const trb_view vd = views[n];
  bin_face<FILL>(verts, faces, vd, n, blockIdx.x * blockDim.x + threadIdx.x, H, W, tg, sqrt_blur, cull, tile_count,
                 tile_fill, tile_offset, pairs, z_cull);
}

__global__ void __launch_bounds__(256)
alloc_tiles_kernel(const int* __restrict__ tile_count, int* __restrict__ tile_offset, int ntiles,
                   int* __restrict__ header, long long pair_capacity, int* __restrict__ busy_list) {
  alloc_tile(tile_count, tile_offset, ntiles, header, pair_capacity, busy_list, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void write_stats_kernel(const int* __restrict__ header, long long pair_capacity,
                                   int* __restrict__ stats) {
  const unsigned long long need = *(const unsigned long long*)header;
  stats[0] = (int)min(need, (unsigned long long)0x7fffffff);
  stats[1] = header[2];
  stats[2] = (int)min(pair_capacity, (long long)0x7fffffff);
  stats[3] = 0;
}

// ------------------------------------------------------------------------------------------
// Backward: one thread per pixel, k in the inner loop so that the 32 lanes of a warp hold
// horizontally adjacent pixels of the same layer -- they mostly hit the same face, and
// warp_aggregated_add turns 32 x 9 atomics into 9.
template <bool PERSP, bool CLIP>
__global__ void __launch_bounds__(256)
raster_backward_kernel(const float* __restrict__ verts, const int* __restrict__ faces,
                       const trb_view* __restrict__ views, int N, int H, int W, int K,
                       const long long* __restrict__ p2f, const float* __restrict__ grad_z,
                       const float* __restrict__ grad_bary, const float* __restrict__ grad_d,
                       float* __restrict__ grad_verts) {
  const long long npix = (long long)N * H * W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = pix < npix;
  int n = 0, xi = 0, yi = 0;
  if (live) {
    n = (int)(pix / ((long long)H * W));
    const int rem = (int)(pix - (long long)n * H * W);
    yi = rem / W; xi = rem - yi * W;
  }
  const trb_view vd = views[n];
  const float px = pix_to_ndc(W - 1 - xi, W, H);
  const float py = pix_to_ndc(H - 1 - yi, H, W);
  for (int k = 0; k < K; ++k) {
    const long long f = live ? p2f[pix * K + k] : -1;
    if (!__any_sync(0xffffffffu, f >= 0)) break;  // layers are sorted: nothing further either
    float g[9];
    int i0 = 0, i1 = 0, i2 = 0;
    int key = -1;
#pragma unroll
    for (int i = 0; i < 9; ++i) g[i] = 0.0f;
    if (f >= 0) {
      const int lf = (int)(f - vd.p2f_base);
      const int r = vd.face_start + lf;
      if (faces != nullptr) {
        i0 = __ldg(faces + 3 * (size_t)r) + vd.vert_delta;
        i1 = __ldg(faces + 3 * (size_t)r + 1) + vd.vert_delta;
        i2 = __ldg(faces + 3 * (size_t)r + 2) + vd.vert_delta;
      } else {
        i0 = 3 * r; i1 = i0 + 1; i2 = i0 + 2;
      }
      const FaceXYZ v = load_face(verts, faces, vd, lf);
      const long long s = pix * K + k;
      sample_backward<PERSP, CLIP>(v, px, py, grad_z ? grad_z[s] : 0.0f,
                                   grad_bary ? grad_bary[s * 3] : 0.0f,
                                   grad_bary ? grad_bary[s * 3 + 1] : 0.0f,
                                   grad_bary ? grad_bary[s * 3 + 2] : 0.0f,
                                   grad_d ? grad_d[s] : 0.0f, g);
      key = (int)f;
    }
    float* const dst[9] = {grad_verts + 3 * (size_t)i0, grad_verts + 3 * (size_t)i0 + 1,
                           grad_verts + 3 * (size_t)i0 + 2, grad_verts + 3 * (size_t)i1,
                           grad_verts + 3 * (size_t)i1 + 1, grad_verts + 3 * (size_t)i1 + 2,
                           grad_verts + 3 * (size_t)i2, grad_verts + 3 * (size_t)i2 + 1,
                           grad_verts + 3 * (size_t)i2 + 2};
    warp_aggregated_add<9>(key, g, dst);
  }
}

int run_binning(const float* verts_ndc, const int* faces, const trb_view* views, int N, int max_face_count,
                int H, int W, const TileGrid& tg, const WsLayout& ws, void* workspace, float sqrt_blur, bool cull,
                long long pair_capacity, cudaStream_t st, float z_cull) {
  unsigned char* wsb = (unsigned char*)workspace;
  int* header = (int*)(wsb + ws.header);
  int* tile_count = (int*)(wsb + ws.count);
  int* tile_fill = (int*)(wsb + ws.fill);
  int* tile_offset = (int*)(wsb + ws.offset);
  int2* pairs = (int2*)(wsb + ws.pairs);
  const int ntiles = N * tg.tiles_x * tg.tiles_y;
  // header, tile_count and tile_fill are contiguous: one memset
  TRB_CUDA_TRY(cudaMemsetAsync(wsb, 0, ws.offset, st));
  if (max_face_count > 0) {
    dim3 bgrid(ceil_div(max_face_count, 256), N);
    bin_faces_kernel<false><<<bgrid, 256, 0, st>>>(verts_ndc, faces, views, H, W, tg, sqrt_blur, cull,
                                                   tile_count, tile_fill, tile_offset, pairs, z_cull);
    TRB_LAUNCH_CHECK();
    alloc_tiles_kernel<<<ceil_div(ntiles, 256), 256, 0, st>>>(tile_count, tile_offset, ntiles, header,
                                                              pair_capacity, (int*)(wsb + ws.busy));
    TRB_LAUNCH_CHECK();
    bin_faces_kernel<true><<<bgrid, 256, 0, st>>>(verts_ndc, faces, views, H, W, tg, sqrt_blur, cull,
                                                  tile_count, tile_fill, tile_offset, pairs, z_cull);
    TRB_LAUNCH_CHECK();
  }
  return TRB_OK;
}

}  // namespace trb

using namespace trb;

extern "C" int trb_raster_workspace_bytes(int N, int H, int W, int K, int64_t pair_capacity,
                                          size_t* bytes) {
  if (!bytes || N < 0 || H < 1 || W < 1 || pair_capacity < 0) return TRB_ERR_BAD_ARG;
  if (K < 1) return TRB_ERR_BAD_ARG;
  if (K > TRB_MAX_FACES_PER_PIXEL) return TRB_ERR_K_TOO_LARGE;
  const TileGrid tg = make_tile_grid(H, W, K);
  *bytes = make_ws_layout(N, tg, pair_capacity).total;
  return TRB_OK;
}

extern "C" int trb_raster_forward(const float* verts_ndc, const int32_t* faces, const trb_view* views,
                                  int N, int max_face_count, int H, int W, int K, float blur_radius,
                                  uint32_t flags, int64_t pair_capacity, void* workspace,
                                  size_t workspace_bytes, int64_t* pix_to_face, float* zbuf,
                                  float* bary, float* dists, int32_t* stats, int device,
                                  trb_stream_t stream) {
  if (N < 0 || H < 1 || W < 1 || K < 1 || max_face_count < 0 || pair_capacity < 0 ||
      !(blur_radius >= 0.0f))
    return TRB_ERR_BAD_ARG;
  if (K > TRB_MAX_FACES_PER_PIXEL) return TRB_ERR_K_TOO_LARGE;
  if (N == 0) return TRB_OK;
  if (N > 65535) return TRB_ERR_BAD_ARG;
  if (!views || !pix_to_face || !zbuf || !bary || !dists || !workspace) return TRB_ERR_BAD_ARG;
  if (max_face_count > 0 && !verts_ndc) return TRB_ERR_BAD_ARG;
  const TileGrid tg = make_tile_grid(H, W, K);
  const WsLayout ws = make_ws_layout(N, tg, pair_capacity);
  if (workspace_bytes < ws.total) return TRB_ERR_WORKSPACE;
  TRB_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* wsb = (unsigned char*)workspace;
  const float sqrt_blur = sqrtf(blur_radius);
  {
    const int brc = run_binning(verts_ndc, faces, views, N, max_face_count, H, W, tg, ws, workspace, sqrt_blur,
                                flags & TRB_CULL_BACKFACES, (long long)pair_capacity, st);
    if (brc != TRB_OK) return brc;
  }
  // the fused fine kernels with the shading epilogue compiled out (TRB_SHADER_NONE)
  FineArgs a = {};
  a.verts_ndc = verts_ndc; a.faces = faces; a.views = views;
  a.H = H; a.W = W; a.K = K; a.blur_radius = blur_radius; a.sqrt_blur = sqrt_blur; a.z_cull = 0.0f;
  a.flags = flags; a.tg = tg;
  a.tile_count = (const int*)(wsb + ws.count); a.tile_offset = (const int*)(wsb + ws.offset);
  a.pairs = (const int2*)(wsb + ws.pairs);
  a.ws_header = (const int*)(wsb + ws.header); a.busy_tiles = (const int*)(wsb + ws.busy);
  a.p2f = (long long*)pix_to_face; a.zbuf = zbuf; a.bary = bary; a.dists = dists;
  a.sigma = 1.0f; a.gamma = 1.0f;
  a.uv = {nullptr, nullptr, nullptr, 0, 0};
  const int rc = launch_render_fine(TRB_SHADER_NONE, 0, N, st, a);
  if (rc != TRB_OK) return rc;
  if (stats) {
    write_stats_kernel<<<1, 1, 0, st>>>((const int*)(wsb + ws.header), (long long)pair_capacity, stats);
    TRB_LAUNCH_CHECK();
  }
  return TRB_OK;
}

extern "C" int trb_raster_backward(const float* verts_ndc, const int32_t* faces, const trb_view* views,
                                   int N, int H, int W, int K, uint32_t flags,
                                   const int64_t* pix_to_face, const float* grad_zbuf,
                                   const float* grad_bary, const float* grad_dists,
                                   float* grad_verts_ndc, int device, trb_stream_t stream) {
  if (N < 0 || H < 1 || W < 1 || K < 1) return TRB_ERR_BAD_ARG;
  if (K > TRB_MAX_FACES_PER_PIXEL) return TRB_ERR_K_TOO_LARGE;
  if (N == 0) return TRB_OK;
  if (!verts_ndc || !views || !pix_to_face || !grad_verts_ndc) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long npix = (long long)N * H * W;
  const unsigned blocks = (unsigned)ceil_div64(npix, 256);
  const bool persp = flags & TRB_PERSPECTIVE_CORRECT, clip = flags & TRB_CLIP_BARYCENTRIC;
  const long long* p2f = (const long long*)pix_to_face;
#define TRB_BWD(P, C)                                                                            \
  raster_backward_kernel<P, C><<<blocks, 256, 0, st>>>(verts_ndc, faces, views, N, H, W, K, p2f,   \
                                                       grad_zbuf, grad_bary, grad_dists,          \
                                                       grad_verts_ndc)
  if (persp && clip) TRB_BWD(true, true);
  else if (persp) TRB_BWD(true, false);
  else if (clip) TRB_BWD(false, true);
  else TRB_BWD(false, false);
#undef TRB_BWD
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}
