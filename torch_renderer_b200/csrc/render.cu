// Fused render pipeline for sm_100a: everything `MeshRenderer.forward` and `loss.backward()` do for
// one batch of views, in one C-ABI call each way (reference call sites: renderer.py:100-101,
// torch_renderer.py:113-120,158, camera_pose_optimizer.py:244-250,303, mesh_deformer.py:197,221).
//
// Forward: four small stage kernels packed by block role (render_stages.cu: prep -> count -> alloc -> fill) and
// ONE fine kernel that rasterises and, in its epilogue, interpolates attributes, samples the texture, lights and
// blends -- this file for faces_per_pixel == 1, render_kn.cu otherwise.  HBM traffic of the big kernel is the
// compulsory 28*K + 16 bytes per pixel of output (Fragments + RGBA) plus L2-resident mesh reads.
//
// Backward: ONE kernel that walks the compact list of covered pixels the forward left (`hit_pixels`), re-reads
// their Fragments + the image gradient, runs the lighting model and the blend backward once per sample and chains
// straight into the rasteriser backward (no grad_bary / grad_zbuf / grad_dists tensors exist), scattering with
// warp-aggregated float4 reductions; then one post kernel (render_stages.cu: NDC->world, camera centre, vertex
// normals, accumulator unpacking).  Every launch is a programmatic dependent launch; there are no memset nodes.
#include "render_internal.cuh"
#include "stages.cuh"

namespace trb {

// ---- fused fine pass, faces_per_pixel == 1 ---------------------------------------------------------
// The meshes this path sees most (cow at 512^2: ~3.5 px per face, ~200 faces per 16x16 tile) make a
// per-pixel walk over the tile's face list 95% wasted bbox rejects.  Here the walk is FACE-parallel:
// each thread takes one face of the list and visits only the pixels of that face's (tile-clipped)
// bounding box, publishing candidates with a shared-memory atomicMin on a 64-bit key
// (z bits << 32 | face) -- which is exactly the (z, face index) order of A5 because z >= 0.
// Faces whose clipped bbox is large go to a second, pixel-parallel pass (one thread per pixel walking
// the few big faces), so neither regime degenerates.  Tiles with an empty list only stream out the
// background.

// (k * kInvWidth[w]) >> 16 == k / w for k < 256, w in 1..16
__constant__ int kInvWidth[17] = {0, 65537, 32769, 21846, 16385, 13108, 10923, 9363, 8193,
                                  7282, 6554, 5958, 5462, 5042, 4682, 4370, 4097};

__device__ __forceinline__ unsigned long long pack_key(float z, int f) {
  const unsigned zb = (z == 0.0f) ? 0u : __float_as_uint(z);  // -0.0f must not sort last
  return ((unsigned long long)zb << 32) | (unsigned)f;
}

#ifndef TRB_K1_STRIP
#define TRB_K1_STRIP 8
#endif
// Diagnostic (TRB_KN_STATS builds): clock64 cycles of thread 0 spent in the phases of a busy K=1 tile, summed over
// tiles: [0] tiles, [1] header + staging (to the first barrier), [2] prefix sums, [3] (face, pixel) items,
// [4] covered-pixel compaction incl. the list atomic, [5] finishing the covered pixels (exact sample + shading)
#ifdef TRB_KN_STATS
__device__ unsigned long long g_k1_phase[16];
#define K1_T(i) do { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&g_k1_phase[i], (unsigned long long)(now_ - k1_t_)); k1_t_ = now_; } } while (0)
#else
#define K1_T(i) ((void)0)
#endif
constexpr int kStrip = TRB_K1_STRIP;  // tiles per CTA strip
constexpr int kK1ItemTable = 4096;    // (face, pixel) items of a staging chunk with a face lookup entry (else: search)

// Background of one tile whose face list is empty: a pure streaming store of -1 Fragments and
// the background colour.
template <int SHADER>
__device__ __forceinline__ void fill_empty_tile(const FineArgs& a, int n, int tbx, int tby) {
  const int tid = threadIdx.x;
  const int xi = tbx * 16 + (tid & 15), yi = tby * 16 + (tid >> 4);
  if (xi >= a.W || yi >= a.H) return;
  const size_t pix = ((size_t)n * a.H + yi) * a.W + xi;
  if (!a.sparse) {
    st_cs(a.p2f + pix, -1ll);
    st_cs(a.zbuf + pix, -1.0f);
    st_cs(a.dists + pix, -1.0f);
    st_cs(a.bary + pix * 3 + 0, -1.0f);
    st_cs(a.bary + pix * 3 + 1, -1.0f);
    st_cs(a.bary + pix * 3 + 2, -1.0f);
  }
  if (SHADER == TRB_SHADER_NONE) return;
  const float4 bgv = (SHADER == TRB_SHADER_SOFT_SILHOUETTE) ? make_float4(1.0f, 1.0f, 1.0f, 0.0f)
                                                            : make_float4(a.bg0, a.bg1, a.bg2, 0.0f);
  st_cs(reinterpret_cast<float4*>(a.images) + pix, bgv);
}

// The same for a tile that lies entirely inside an image whose rows keep 16-byte alignment (W % 4 == 0):
// 16-byte stores only -- 2.75 per thread instead of 7, and every store instruction covers whole sectors
// (the stride-3 barycentric stores above touch 12 sectors for 4 sectors' worth of data).
template <int SHADER>
__device__ __forceinline__ void fill_empty_tile_v4(const FineArgs& a, int n, int tbx, int tby) {
  const int tid = threadIdx.x;
  const size_t pix0 = ((size_t)n * a.H + tby * 16) * a.W + tbx * 16;  // top-left pixel of the tile
  const float4 m4 = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
  const size_t W = a.W;
  if (!a.sparse) {
    if (tid < 64) {          // zbuf: 16 rows x 4
      st_cs(reinterpret_cast<float4*>(a.zbuf + pix0 + (tid >> 2) * W) + (tid & 3), m4);
    } else if (tid < 128) {  // dists
      const int i = tid - 64;
      st_cs(reinterpret_cast<float4*>(a.dists + pix0 + (i >> 2) * W) + (i & 3), m4);
    } else {                 // pix_to_face: 16 rows x 8 (two int64 per store)
      const int i = tid - 128;
      __stcs(reinterpret_cast<longlong2*>(a.p2f + pix0 + (i >> 3) * W) + (i & 7), make_longlong2(-1ll, -1ll));
    }
    if (tid < 192) {         // barycentrics: 16 rows x 12
      const int r = tid / 12, c = tid - r * 12;
      st_cs(reinterpret_cast<float4*>(a.bary + 3 * (pix0 + r * W)) + c, m4);
    }
  }
  if (SHADER == TRB_SHADER_NONE) return;
  const float4 bgv = (SHADER == TRB_SHADER_SOFT_SILHOUETTE) ? make_float4(1.0f, 1.0f, 1.0f, 0.0f)
                                                            : make_float4(a.bg0, a.bg1, a.bg2, 0.0f);
  st_cs(reinterpret_cast<float4*>(a.images) + pix0 + (tid >> 4) * W + (tid & 15), bgv);
}

// A whole strip of kStrip empty tiles inside the image: row-contiguous 16-byte stores, every warp
// instruction writes 512 consecutive bytes.
template <int SHADER>
__device__ __forceinline__ void fill_empty_strip_v4(const FineArgs& a, int n, int tbx0, int tby) {
  const int tid = threadIdx.x;
  const size_t W = a.W;
  const size_t pix0 = ((size_t)n * a.H + tby * 16) * W + tbx0 * 16;
  const float4 m4 = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
  constexpr int ZW = kStrip * 4, PW = kStrip * 8, BW = kStrip * 12, IW = kStrip * 16;  // 16-byte words per row
  static_assert(kStrip % 4 == 0, "the strip fill deals whole rounds of 256 stores");
  if (!a.sparse) {
#pragma unroll
    for (int j = 0; j < kStrip / 4; ++j) {       // zbuf, dists: 16 rows x ZW
      const int i = tid + 256 * j;
      st_cs(reinterpret_cast<float4*>(a.zbuf + pix0 + (i / ZW) * W) + (i % ZW), m4);
      st_cs(reinterpret_cast<float4*>(a.dists + pix0 + (i / ZW) * W) + (i % ZW), m4);
    }
#pragma unroll
    for (int j = 0; j < kStrip / 2; ++j) {       // pix_to_face: 16 rows x PW
      const int i = tid + 256 * j;
      __stcs(reinterpret_cast<longlong2*>(a.p2f + pix0 + (i / PW) * W) + (i % PW), make_longlong2(-1ll, -1ll));
    }
#pragma unroll
    for (int j = 0; j < kStrip * 3 / 4; ++j) {   // barycentrics: 16 rows x BW
      const int i = tid + 256 * j;
      const int r = i / BW, c = i - r * BW;
      st_cs(reinterpret_cast<float4*>(a.bary + 3 * (pix0 + r * W)) + c, m4);
    }
  }
  if (SHADER == TRB_SHADER_NONE) return;
  const float4 bgv = (SHADER == TRB_SHADER_SOFT_SILHOUETTE) ? make_float4(1.0f, 1.0f, 1.0f, 0.0f)
                                                            : make_float4(a.bg0, a.bg1, a.bg2, 0.0f);
#pragma unroll
  for (int j = 0; j < kStrip; ++j) {           // RGBA: 16 rows x IW
    const int i = tid + 256 * j;
    st_cs(reinterpret_cast<float4*>(a.images) + pix0 + (i / IW) * W + (i % IW), bgv);
  }
}

template <int SHADER, int LIGHT>
__device__ __forceinline__ void raster_tile_k1(const FineArgs& a, int t);

// Every CTA (a) owns a strip of kStrip horizontally adjacent tiles, fetches their list lengths with
// one round trip and streams out the empty ones (94% of the tiles for the cow at 512^2), and (b)
// rasterises its share of the compact list of non-empty tiles the binning pass left in the workspace.
// The non-empty tiles are dealt out evenly over the whole grid, so at any moment a fraction of the
// resident CTAs is doing latency-bound raster work while the rest keeps the HBM write stream busy
// (one tile per CTA bound the kernel by CTA turnover x the latency of the list-length load;
// rasterising the strip's own tiles serialised the 4-8 neighbouring busy tiles of an object in one CTA).
#ifndef TRB_K1_CTAS
#define TRB_K1_CTAS 4   // 64 registers: 4 x 256 threads per SM (same-box A/B: 3 -> 4 CTAs: fine 0.160 -> 0.147 ms)
#endif

// STRIP = kStrip for batches (the streaming fill wants long rows per CTA), 1 for grids too small to fill the GPU
// that way: one 512^2 view is 1,024 tiles = 128 strips of 8 for 148 SMs, and a 256^2 view 32 -- its ~150 busy tiles
// were rasterised five deep by 32 CTAs (C1: 0.042 ms of fine kernel for 0.4 M pixels).
template <int SHADER, int LIGHT, int STRIP>
__global__ void __launch_bounds__(256, TRB_K1_CTAS)
render_fine_k1_kernel(const FineArgs a) {
  pdl_wait();
  const int n = blockIdx.z, tby = blockIdx.y, tbx0 = blockIdx.x * STRIP;
  const int* counts = a.tile_count + (size_t)(n * a.tg.tiles_y + tby) * a.tg.tiles_x;
  int cnt[STRIP];
#pragma unroll
  for (int s = 0; s < STRIP; ++s) cnt[s] = (tbx0 + s < a.tg.tiles_x) ? __ldg(counts + tbx0 + s) : -1;
  const int nbusy = __ldg(a.ws_header + 4);
  const bool rows_inside = (tby + 1) * 16 <= a.H && (a.W & 3) == 0;
  bool all_empty = true;
#pragma unroll
  for (int s = 0; s < STRIP; ++s) all_empty = all_empty && cnt[s] == 0;
  if (STRIP == kStrip && all_empty && rows_inside && (tbx0 + STRIP) * 16 <= a.W) {
    fill_empty_strip_v4<SHADER>(a, n, tbx0, tby);
  } else {
#pragma unroll
    for (int s = 0; s < STRIP; ++s)
      if (cnt[s] == 0) {
        if (rows_inside && (tbx0 + s + 1) * 16 <= a.W) fill_empty_tile_v4<SHADER>(a, n, tbx0 + s, tby);
        else fill_empty_tile<SHADER>(a, n, tbx0 + s, tby);
      }
  }
  // deal the non-empty tiles out evenly over the grid (32-bit arithmetic only)
  const unsigned ncta = gridDim.x * gridDim.y * gridDim.z;
  const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  unsigned lo, hi;
  if ((unsigned)nbusy <= ncta) {
    const unsigned step = ncta / max(nbusy, 1);          // every step-th CTA takes one tile
    const unsigned q = cta / step;
    const bool mine = (cta - q * step == 0) && (q < (unsigned)nbusy);
    lo = q; hi = mine ? q + 1 : q;
  } else {
    const unsigned per = (unsigned)nbusy / ncta, rem = (unsigned)nbusy - per * ncta;
    lo = cta * per + min(cta, rem);
    hi = lo + per + (cta < rem ? 1u : 0u);
  }
  for (unsigned i = lo; i < hi; ++i) raster_tile_k1<SHADER, LIGHT>(a, __ldg(a.busy_tiles + i));
}

// Rasterises one non-empty tile.  The unit of parallel work is a (face, pixel-of-its-clipped-bbox)
// pair: the pairs of up to 256 staged faces are numbered with a block-wide prefix sum and dealt out in
// equal contiguous runs to the 256 threads, so a tile with 200 three-pixel faces and a tile with
// three 200-pixel faces cost the same.  Candidates are published with a shared-memory atomicMin on a
// 64-bit key (z bits << 32 | face), which is exactly the (z, face index) order of A5 because z >= 0.
template <int SHADER, int LIGHT>
__device__ __forceinline__ void raster_tile_k1(const FineArgs& a, int t) {
  constexpr int TX = 16, TY = 16, NT = 256;
  const int tid = threadIdx.x;
  const int H = a.H, W = a.W;
  const int tiles_per_view = a.tg.tiles_x * a.tg.tiles_y;
  const int n = t / tiles_per_view;
  const int trem = t - n * tiles_per_view;
  const int tby = trem / a.tg.tiles_x, tbx = trem - tby * a.tg.tiles_x;
  const int lx = tid & (TX - 1), ly = tid >> 4;
  const int tile_x0 = tbx * TX, tile_y0 = tby * TY;
  const int xi = tile_x0 + lx, yi = tile_y0 + ly;
  const bool live = (xi < W) && (yi < H);
  const size_t pix = ((size_t)n * H + yi) * W + xi;

  __shared__ unsigned long long s_key[NT];
  __shared__ float s_px[TX], s_py[TY];
  __shared__ float4 s_bb[NT];
  __shared__ float4 s_va[NT];
  __shared__ float4 s_vb[NT];
  __shared__ float2 s_vc[NT];
  __shared__ int s_id[NT];
  __shared__ int s_rng[NT];        // c0 - tile_x0 | (r0 - tile_y0) << 4 | (bbox width - 1) << 8
  __shared__ int s_start[NT + 1];  // exclusive prefix sum of bbox pixel counts
  __shared__ int s_wtot[NT / 32];
#ifndef TRB_K1_NESTED_ITEMS
  __shared__ unsigned char s_item_face[kK1ItemTable];  // staged face of every (face, pixel) item of the chunk
  __shared__ unsigned s_zlo_bits[NT];                  // bits of every staged face's depth lower bound
#endif

  const trb_view vd = a.views[n];
  const bool persp = a.flags & TRB_PERSPECTIVE_CORRECT, clip = a.flags & TRB_CLIP_BARYCENTRIC;
  const bool cull = a.flags & TRB_CULL_BACKFACES;
  const bool hard_edges = a.blur_radius == 0.0f;
  const float blur = a.blur_radius;
  int nlist = a.tile_count[t];
  const int off = a.tile_offset[t];
  const bool overflow = off < 0;
  if (overflow) nlist = vd.face_count;

#ifdef TRB_KN_STATS
  long long k1_t_ = clock64();
  if (tid == 0) atomicAdd(&g_k1_phase[0], 1ull);
#endif
  __syncthreads();  // a CTA may rasterise several tiles: the previous one's epilogue still reads these arrays
  s_key[tid] = ~0ull;
  if (tid < TX) s_px[tid] = pix_to_ndc(W - 1 - (tile_x0 + tid), W, H);
  else if (tid < TX + TY) s_py[tid - TX] = pix_to_ndc(H - 1 - (tile_y0 + tid - TX), H, W);
  // last pixel column / row of the tile that exists in the image
  const int x_hi = min(tile_x0 + TX, W) - 1, y_hi = min(tile_y0 + TY, H) - 1;
  const int lane = tid & 31, warp = tid >> 5;

  for (int base = 0; base < nlist; base += NT) {
    // ---- stage up to NT faces and the size of their clipped bounding boxes
    const int j = base + tid;
    int npx = 0;
    if (j < nlist) {
      const int lf = overflow ? j : a.pairs[off + j].x;
      const FaceXYZ v = load_face(a.verts_ndc, a.faces, vd, lf);
      const bool ok = overflow ? face_is_drawable(v, cull, a.z_cull) : true;
      if (ok) {
        float4 bb;
        bb.x = fsub(min3f(v.x0, v.x1, v.x2), a.sqrt_blur); bb.y = fadd(max3f(v.x0, v.x1, v.x2), a.sqrt_blur);
        bb.z = fsub(min3f(v.y0, v.y1, v.y2), a.sqrt_blur); bb.w = fadd(max3f(v.y0, v.y1, v.y2), a.sqrt_blur);
        int c0, c1, r0, r1;
        pixel_range(bb.x, bb.y, W, H, c0, c1);
        pixel_range(bb.z, bb.w, H, W, r0, r1);
        c0 = max(c0, tile_x0); c1 = min(c1, x_hi); r0 = max(r0, tile_y0); r1 = min(r1, y_hi);
        if (c1 >= c0 && r1 >= r0) {
          npx = (c1 - c0 + 1) * (r1 - r0 + 1);
          s_bb[tid] = bb;
          s_va[tid] = make_float4(v.x0, v.y0, v.z0, v.x1);
          s_vb[tid] = make_float4(v.y1, v.z1, v.x2, v.y2);
          s_vc[tid] = make_float2(v.z2, fadd(edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1), kEps));
          s_id[tid] = lf;
          // box origin, width - 1 and the reciprocal-multiply constant of the width (k / w for k < 256)
          s_rng[tid] = (c0 - tile_x0) | ((r0 - tile_y0) << 4) | ((c1 - c0) << 8) | (kInvWidth[c1 - c0 + 1] << 12);
#ifndef TRB_K1_NESTED_ITEMS
          {
            // early depth reject bound of this face (see the item loop), computed once per face
            const float area_s = fadd(edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1), kEps);
            float zlo = 0.0f;
            if (hard_edges) {
              const float zmin = min3f(v.z0, v.z1, v.z2);
              if (persp) zlo = zmin >= 1e-3f ? zmin * 0.99999f : 0.0f;
              else if (clip) zlo = zmin * 0.99999f;
              else zlo = zmin * 0.99999f * (area_s > 0.0f ? fmaxf(0.0f, (area_s - 2e-8f) / area_s) : 1.0f);
            }
            s_zlo_bits[tid] = __float_as_uint(zlo);
          }
#endif
        }
      }
    }
    // ---- block-wide exclusive prefix sum of npx
    int incl = npx;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();  // also publishes s_key / s_px / s_py on the first pass and the staging arrays
    K1_T(1);
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
      const int c = s_wtot[w];
      if (w < warp) wbase += c;
      total += c;
    }
    s_start[tid] = wbase + incl - npx;
    if (tid == 0) s_start[NT] = total;
    __syncthreads();
    K1_T(2);
#ifdef TRB_KN_STATS
    if (tid == 0) { atomicAdd(&g_k1_phase[6], (unsigned long long)total); atomicAdd(&g_k1_phase[7], (unsigned long long)min(NT, nlist - base)); }
#endif

#ifdef TRB_KN_STATS
    __shared__ unsigned long long s_itm_max, s_itm_sum;
    if (tid == 0) { s_itm_max = 0; s_itm_sum = 0; }
    __syncthreads();
    const long long itm_t0 = clock64();
#endif
#ifndef TRB_K1_NESTED_ITEMS
    // ---- (face, pixel) items, FLAT: item j of the chunk goes to thread j % 256 in round j / 256, so the lanes of a
    // warp hold consecutive items -- mostly of the same face (its staged record is a shared-memory broadcast) -- and
    // all run the same straight-line code: the face of an item comes from a lookup table the staging threads fill
    // (one entry per item), not from per-thread nested face / pixel loops whose trip counts differ lane by lane
    // (the nested version issued ~1,700 warp instructions per tile for ~4 items per thread).
    const int m = min(NT, nlist - base);  // staged faces
    const bool use_table = total <= kK1ItemTable;
    if (use_table) {
      const int my0 = s_start[tid], my1 = s_start[tid + 1];   // empty for tid >= m (npx == 0)
      for (int j = my0; j < my1; ++j) s_item_face[j] = (unsigned char)tid;
    }
    __syncthreads();
    for (int item = tid; item < total; item += NT) {
      int fj;
      if (use_table) {
        fj = s_item_face[item];
      } else {
        int lo = 0, hi = m - 1;  // last staged face whose run starts at or before `item`
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (s_start[mid] <= item) lo = mid; else hi = mid - 1;
        }
        fj = lo;
      }
      const int k = item - s_start[fj];
      const float4 bb = s_bb[fj], va = s_va[fj], vb = s_vb[fj];
      const float2 vc = s_vc[fj];
      const int rng = s_rng[fj], lf = s_id[fj];
      FaceXYZ v;
      v.x0 = va.x; v.y0 = va.y; v.z0 = va.z; v.x1 = va.w;
      v.y1 = vb.x; v.z1 = vb.y; v.x2 = vb.z; v.y2 = vb.w; v.z2 = vc.x;
      const float area = vc.y;
      const int c0 = rng & 15, r0 = (rng >> 4) & 15, bw = ((rng >> 8) & 15) + 1;
      const int dr = (k * (rng >> 12)) >> 16;
      const int lr = r0 + dr, lc = c0 + (k - dr * bw);
      const float qx = s_px[lc], qy = s_py[lr];
      if ((qx > bb.y) || (qx < bb.x) || (qy > bb.w) || (qy < bb.z)) continue;
      // Early depth reject (blur 0 only, where every candidate is strictly inside its face): the interpolated depth
      // is then a convex combination of the vertex depths up to a few ulp -- times area_raw / (area_raw + kEps)
      // without perspective correction or clipping, where the weights are not renormalised -- so a pixel whose
      // current front-most candidate is nearer than that bound cannot be won by this face.  Bound 0 disables it.
      if (reinterpret_cast<const unsigned*>(s_key)[2 * (lr * TX + lc) + 1] < s_zlo_bits[fj]) continue;
      const float e0 = edge_fn(qx, qy, v.x1, v.y1, v.x2, v.y2);
      const float e1 = edge_fn(qx, qy, v.x2, v.y2, v.x0, v.y0);
      const float e2 = edge_fn(qx, qy, v.x0, v.y0, v.x1, v.y1);
      if (hard_edges) {
        // blur 0: w_i = e_i / area keeps the sign of e_i * area exactly => exact reject
        if (area > 0.0f ? (e0 <= 0.0f || e1 <= 0.0f || e2 <= 0.0f)
                        : (area < 0.0f && (e0 >= 0.0f || e1 >= 0.0f || e2 >= 0.0f)))
          continue;
      }
      float pz, b0, b1, b2;
      bool inside;
      if (!eval_from_edges(v, area, e0, e1, e2, persp, clip, pz, b0, b1, b2, inside)) continue;
      if (!inside) {
        if (hard_edges) continue;
        if (triangle_d2(v, qx, qy) >= blur) continue;
      }
      atomicMin(&s_key[lr * TX + lc], pack_key(pz, lf));
    }
#else
    // ---- deal the (face, pixel) pairs out in equal contiguous runs
    const int m = min(NT, nlist - base);  // staged faces
    const int ipt = (total + NT - 1) / NT;
    int item = tid * ipt;
    const int item_end = min(item + ipt, total);
    if (item < item_end) {
      int lo = 0, hi = m - 1;  // last staged face whose run starts at or before `item`
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_start[mid] <= item) lo = mid; else hi = mid - 1;
      }
      int fj = lo;
      int k = item - s_start[fj];
      for (;;) {
        const int fend = s_start[fj + 1] - s_start[fj];
        const float4 bb = s_bb[fj], va = s_va[fj], vb = s_vb[fj];
        const float2 vc = s_vc[fj];
        const int rng = s_rng[fj], lf = s_id[fj];
        FaceXYZ v;
        v.x0 = va.x; v.y0 = va.y; v.z0 = va.z; v.x1 = va.w;
        v.y1 = vb.x; v.z1 = vb.y; v.x2 = vb.z; v.y2 = vb.w; v.z2 = vc.x;
        const float area = vc.y;
        // Early depth reject (blur 0 only, where every candidate is strictly inside its face): the
        // interpolated depth is then a convex combination of the vertex depths up to a few ulp -- times
        // area_raw / (area_raw + kEps) without perspective correction or clipping, where the weights
        // are not renormalised -- so a pixel whose current front-most candidate is nearer than that
        // bound cannot be won by this face and the six IEEE divisions are skipped.  zlo = 0 disables it.
        float zlo = 0.0f;
        if (hard_edges) {
          const float zmin = min3f(v.z0, v.z1, v.z2);
          if (persp) zlo = zmin >= 1e-3f ? zmin * 0.99999f : 0.0f;
          else if (clip) zlo = zmin * 0.99999f;
          else zlo = zmin * 0.99999f * (area > 0.0f ? fmaxf(0.0f, (area - 2e-8f) / area) : 1.0f);
        }
        const unsigned zlo_bits = __float_as_uint(zlo);
        const int c0 = rng & 15, r0 = (rng >> 4) & 15, bw = ((rng >> 8) & 15) + 1;
        const int inv_bw = kInvWidth[bw];
        const int kend = min(fend, k + (item_end - item));
        item += kend - k;
        for (; k < kend; ++k) {
          const int dr = (k * inv_bw) >> 16;
          const int lr = r0 + dr, lc = c0 + (k - dr * bw);
          const float qx = s_px[lc], qy = s_py[lr];
          if ((qx > bb.y) || (qx < bb.x) || (qy > bb.w) || (qy < bb.z)) continue;
          if (reinterpret_cast<const unsigned*>(s_key)[2 * (lr * TX + lc) + 1] < zlo_bits) continue;
          const float e0 = edge_fn(qx, qy, v.x1, v.y1, v.x2, v.y2);
          const float e1 = edge_fn(qx, qy, v.x2, v.y2, v.x0, v.y0);
          const float e2 = edge_fn(qx, qy, v.x0, v.y0, v.x1, v.y1);
          if (hard_edges) {
            // blur 0: w_i = e_i / area keeps the sign of e_i * area exactly => exact reject
            if (area > 0.0f ? (e0 <= 0.0f || e1 <= 0.0f || e2 <= 0.0f)
                            : (area < 0.0f && (e0 >= 0.0f || e1 >= 0.0f || e2 >= 0.0f)))
              continue;
          }
          float pz, b0, b1, b2;
          bool inside;
          if (!eval_from_edges(v, area, e0, e1, e2, persp, clip, pz, b0, b1, b2, inside)) continue;
          if (!inside) {
            if (hard_edges) continue;
            if (triangle_d2(v, qx, qy) >= blur) continue;
          }
          atomicMin(&s_key[lr * TX + lc], pack_key(pz, lf));
        }
        if (item >= item_end) break;
        // next staged face with a non-empty box
        do { ++fj; } while (fj < m - 1 && s_start[fj + 1] == s_start[fj]);
        k = 0;
      }
    }
#endif
#ifdef TRB_KN_STATS
    {
      const unsigned long long dt = (unsigned long long)(clock64() - itm_t0);
      atomicMax(&s_itm_max, dt); atomicAdd(&s_itm_sum, dt);
    }
#endif
    __syncthreads();  // staging arrays are rewritten by the next chunk
#ifdef TRB_KN_STATS
    if (tid == 0) { atomicAdd(&g_k1_phase[8], s_itm_max); atomicAdd(&g_k1_phase[9], s_itm_sum / NT); }
#endif
    K1_T(3);
  }
  // ---- epilogue.  Only the covered pixels (a third of a busy tile of the cow batch) need the division-heavy
  // sample evaluation and the shading: they are compacted over the CTA, so that ceil(covered / 32) warps run that
  // code instead of all eight; every other pixel just streams out its background.
  __shared__ int s_hitlist[NT];
  __shared__ int s_hcnt[NT / 32];
  __shared__ int s_hbase, s_nhit;
  const bool hit = live && (s_key[tid] != ~0ull);
  const unsigned ball = __ballot_sync(0xffffffffu, hit);
  if (lane == 0) s_hcnt[warp] = __popc(ball);
  __syncthreads();
  if (tid == 0) {
    int tot = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) { const int c = s_hcnt[w]; s_hcnt[w] = tot; tot += c; }
    s_nhit = tot;
    s_hbase = (tot > 0 && a.hit_pixels != nullptr) ? atomicAdd(a.hit_pixels, tot) : 0;
  }
  __syncthreads();
  if (hit) {
    const int r = s_hcnt[warp] + __popc(ball & ((1u << lane) - 1u));
    s_hitlist[r] = tid;
    // list of covered pixels for the backward (row-major inside the tile: neighbouring entries, neighbouring pixels)
    if (a.hit_pixels != nullptr) a.hit_pixels[1 + s_hbase + r] = (int)pix;
  } else if (live) {
    if (!a.sparse) {
      st_cs(a.p2f + pix, -1ll);
      st_cs(a.zbuf + pix, -1.0f);
      st_cs(a.dists + pix, -1.0f);
      st_cs(a.bary + pix * 3 + 0, -1.0f);
      st_cs(a.bary + pix * 3 + 1, -1.0f);
      st_cs(a.bary + pix * 3 + 2, -1.0f);
    }
    if (SHADER != TRB_SHADER_NONE) {
      const float4 bgv = (SHADER == TRB_SHADER_SOFT_SILHOUETTE) ? make_float4(1.0f, 1.0f, 1.0f, 0.0f)
                                                                : make_float4(a.bg0, a.bg1, a.bg2, 0.0f);
      st_cs(reinterpret_cast<float4*>(a.images) + pix, bgv);
    }
  }
  __syncthreads();
  K1_T(4);
  if (tid >= s_nhit) return;
  const int hp = s_hitlist[tid];           // pixel of the tile this thread finishes
  const int hx = hp & (TX - 1), hy = hp >> 4;
  const float px = s_px[hx], py = s_py[hy];
  const size_t hpix = ((size_t)n * H + tile_y0 + hy) * W + tile_x0 + hx;
  const int best_f = (int)(unsigned)(s_key[hp] & 0xffffffffull);
  Sample s = {-1.0f, -1.0f, -1.0f, -1.0f, -1.0f};
  {
    const FaceXYZ v = load_face(a.verts_ndc, a.faces, vd, best_f);
    eval_pixel_face_rt(v, px, py, persp, clip, blur, s);  // same operator sequence => same z as the key
  }
  st_cs(a.p2f + hpix, (long long)vd.p2f_base + best_f);
  st_cs(a.zbuf + hpix, s.z);
  st_cs(a.dists + hpix, s.d);
  st_cs(a.bary + hpix * 3 + 0, s.c0);
  st_cs(a.bary + hpix * 3 + 1, s.c1);
  st_cs(a.bary + hpix * 3 + 2, s.c2);
  if (SHADER == TRB_SHADER_NONE) return;
  float4 out;
  if (SHADER == TRB_SHADER_SOFT_SILHOUETTE) {
    // 1 - (1 - p) is not bit-identical to p; keep the product form of sigmoid_alpha_blend
    out = make_float4(1.0f, 1.0f, 1.0f, 1.0f - (1.0f - sigmoidf(-s.d / a.sigma)));
  } else {
    const ViewParams vp = load_view_params(a.view_params, n);
    const ShadeIn sin = {a.verts_world, a.normals, a.colors, a.faces, a.uv};
    const F3 c = shade_sample<LIGHT>(sin, vd, vp, best_f, s.c0, s.c1, s.c2);
    if (SHADER == TRB_SHADER_HARD_PHONG) {
      out = make_float4(c.x, c.y, c.z, 1.0f);
    } else {
      const float eps = 1e-10f;
      const float zrange = vp.zfar - vp.znear;
      const float zinv = (vp.zfar - s.z) / zrange;
      const float zmax = fmaxf(zinv, eps);
      const float prob = sigmoidf(-s.d / a.sigma);
      const float w = prob * expf((zinv - zmax) / a.gamma);
      const float delta = fmaxf(expf((eps - zmax) / a.gamma), eps);
      const float inv = 1.0f / (w + delta);
      out = make_float4((w * c.x + delta * a.bg0) * inv, (w * c.y + delta * a.bg1) * inv,
                        (w * c.z + delta * a.bg2) * inv, 1.0f - (1.0f - prob));
    }
  }
  st_cs(reinterpret_cast<float4*>(a.images) + hpix, out);
  if (a.alpha_sum != nullptr) {
    // per-view sum of the alpha channel (a coverage metric / silhouette-area loss without re-reading the image):
    // lanes [0, wcnt) of this warp got here
    const int wcnt = min(32, s_nhit - warp * 32);
    const unsigned m = wcnt >= 32 ? 0xffffffffu : ((1u << wcnt) - 1u);
    float asum = out.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float t = __shfl_down_sync(m, asum, o);
      if (lane + o < wcnt) asum += t;
    }
    if (lane == 0) atomicAdd(a.alpha_sum + n, asum);
  }
  K1_T(5);
}

// ---- fused backward ----------------------------------------------------------------------------
struct BwdArgs {
  const float* verts_ndc; const int* faces; const trb_view* views;
  int H, W, K; unsigned flags; const int* hit_pixels; const int* hit_counts;
  const long long* p2f; const float* zbuf; const float* bary; const float* dists;
  const float* view_params; const float* verts_world; const float* normals; const float* colors;
  const float* g_images; const float* g_zbuf; const float* g_bary; const float* g_dists;
  float4* g_verts_ndc; float4* g_verts_world; float4* g_normals; float4* g_colors;  // xyz_ accumulators
  float* g_view_params; int want_light_grad, want_cam_grad;
  float sigma, gamma, bg0, bg1, bg2;
  UvTex uv; float* g_tex_map;  // TexturesUV: texture lookup instead of vertex colours; gradient of the map
};

// The K > 1 backward stalls 32 % of its samples on instruction fetch (82 KB of SASS, ncu r02h): its four scatters
// share ONE out-of-line copy of the warp-aggregated reduction instead of four inlined ones.
#ifndef TRB_BWD_SCATTER_INLINE
__device__ __noinline__ void scatter_xyz3_outlined(unsigned leaders, bool any, bool aggregate, int key, float v0,
                                                   float v1, float v2, float v3, float v4, float v5, float v6,
                                                   float v7, float v8, float4* base, int i0, int i1, int i2) {
  WarpGroups wg;
  wg.leaders = leaders; wg.any = any; wg.aggregate = aggregate;
  wg.runs = false; wg.run_head = false; wg.run_end = 0; wg.run_steps = 0;
  const float v[9] = {v0, v1, v2, v3, v4, v5, v6, v7, v8};
  warp_groups_add_xyz3(wg, key, v, base, i0, i1, i2);
}
#endif

template <int SHADER, int LIGHT>
__device__ __forceinline__ void render_backward_pixel_k1(const BwdArgs& a, bool live, int pixi);
template <bool K1, int SHADER, int LIGHT>
__device__ __forceinline__ void render_backward_pixel(const BwdArgs& a, bool live, int pixi, int nk, float4* s_park);

// Diagnostic counters of the K > 1 backward scatter (TRB_KN_STATS builds only; trb_debug_bw_stats):
// [0] warp x layer rounds of the rasteriser-backward scatter, [1] lanes with a sample, [2] distinct faces
#ifdef TRB_KN_STATS
__device__ unsigned long long g_bw_stats[8];
#endif

#ifndef TRB_BWD_CTAS
#define TRB_BWD_CTAS 4
#endif
// K > 1 backward: capped at 128 registers (4 CTAs = 16 warps per SM instead of 3 at 167 registers; a few spills).
// Same-box A/B on the 1M-face sphere: 4.02 -> 3.04 ms (5 CTAs / 96 registers: 3.38 ms).
#ifndef TRB_BWD_KN_CTAS
#define TRB_BWD_KN_CTAS 4
#endif
template <bool K1, int SHADER, int LIGHT>
__global__ void __launch_bounds__(128, K1 ? TRB_BWD_CTAS : TRB_BWD_KN_CTAS)
render_backward_kernel(const BwdArgs a) {
  pdl_wait();
  extern __shared__ float4 s_park[];  // K>1 Phong: (g_bary from shading, g . colour_k) per [k][tid]
  const int count = a.hit_pixels[0];
  const int count_up = (count + 31) & ~31;
  const int stride = gridDim.x * blockDim.x;
  // the next iteration's list entry is requested before this one's pixel is processed (one of the four dependent
  // round trips of an iteration: list entry -> Fragments -> face -> vertices)
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int pixi = i < count ? a.hit_pixels[1 + i] : 0;
  int nk = (!K1 && i < count) ? a.hit_counts[i] : 0;
  for (; i < count_up; i += stride) {
    const bool live = i < count;
    const int inext = i + stride;
    const int pixi_next = inext < count ? a.hit_pixels[1 + inext] : 0;
    const int nk_next = (!K1 && inext < count) ? a.hit_counts[inext] : 0;
    if (K1) render_backward_pixel_k1<SHADER, LIGHT>(a, live, pixi);
    else render_backward_pixel<false, SHADER, LIGHT>(a, live, pixi, nk, s_park);
    pixi = pixi_next; nk = nk_next;
  }
}

// faces_per_pixel == 1: straight-line code, every quantity computed once, fast-math divisions and
// transcendentals, all pixel-indexed loads issued up front, one match.any shared by the four scatters.
template <int SHADER, int LIGHT>
__device__ __forceinline__ void render_backward_pixel_k1(const BwdArgs& a, bool live, int pixi) {
  constexpr bool PHONG = (SHADER == TRB_SHADER_SOFT_PHONG || SHADER == TRB_SHADER_HARD_PHONG);
  constexpr bool SOFT = (SHADER == TRB_SHADER_SOFT_PHONG);
  constexpr bool SIL = (SHADER == TRB_SHADER_SOFT_SILHOUETTE);
  constexpr bool LIT = PHONG && LIGHT != TRB_LIGHT_AMBIENT;
  const int H = a.H, W = a.W, HW = H * W;
  const int n = pixi / HW;
  const int prem = pixi - n * HW;
  const int yi = prem / W, xi = prem - yi * W;
  const size_t pix = (size_t)pixi;
  const bool persp = a.flags & TRB_PERSPECTIVE_CORRECT, clip = a.flags & TRB_CLIP_BARYCENTRIC;

  // ---- every pixel-indexed load first (independent of each other: one DRAM round trip)
  const long long f = live ? a.p2f[pix] : -1;
  const bool on = f >= 0;
  float4 g = make_float4(0, 0, 0, 0);
  float z = 0.0f, d = 0.0f, b0 = 0.0f, b1 = 0.0f, b2 = 0.0f;
  float gz = 0.0f, gd = 0.0f, gb0 = 0.0f, gb1 = 0.0f, gb2 = 0.0f;
  if (on) {
    if (SHADER != TRB_SHADER_NONE) g = __ldg(reinterpret_cast<const float4*>(a.g_images) + pix);
    z = a.zbuf[pix]; d = a.dists[pix];
    if (PHONG) { b0 = a.bary[pix * 3]; b1 = a.bary[pix * 3 + 1]; b2 = a.bary[pix * 3 + 2]; }
    if (a.g_zbuf) gz = a.g_zbuf[pix];
    if (a.g_dists) gd = a.g_dists[pix];
    if (a.g_bary) { gb0 = a.g_bary[pix * 3]; gb1 = a.g_bary[pix * 3 + 1]; gb2 = a.g_bary[pix * 3 + 2]; }
  }
  const trb_view vd = a.views[n];
  const int lf = on ? (int)(f - vd.p2f_base) : 0;
  const size_t r = (size_t)(vd.face_start + lf);
  int w0i = 0, w1i = 0, w2i = 0;  // world-space vertex rows
  if (on && a.faces != nullptr) { w0i = __ldg(a.faces + 3 * r); w1i = __ldg(a.faces + 3 * r + 1); w2i = __ldg(a.faces + 3 * r + 2); }
  const int key = on ? (int)f : -1;
  // K = 1: the reductions (12 per covered pixel) are what bounds this kernel, so runs of neighbouring pixels on
  // one face are summed in the warp first (cow batch: 3.5 M -> 1.9 M L2 sector-ops, 0.059 -> 0.047 ms)
  const WarpGroups wg = warp_groups(key, true);

  // ---- blend (K = 1) and lighting
  if (SOFT || SIL || PHONG) {
    float p = 0.0f, q = 1.0f;
    if (SOFT || SIL) { p = __frcp_rn(1.0f + __expf(d / a.sigma)); q = 1.0f - p; }  // sigmoid(-d/sigma)
    float g_p = g.w;  // alpha = p for a single layer
    if (PHONG) {
      ViewParams vp = load_view_params(a.view_params, n);
      F3 C0 = {0, 0, 0}, C1 = C0, C2 = C0, X0 = C0, X1 = C0, X2 = C0, N0 = C0, N1 = C0, N2 = C0;
      const bool use_uv = a.uv.map != nullptr;
      F2 t0 = {0, 0}, t1 = t0, t2 = t0;
      if (on) {
        if (use_uv) {
          t0 = ld2(a.uv.verts_uvs, __ldg(a.uv.faces_uvs + 3 * r)); t1 = ld2(a.uv.verts_uvs, __ldg(a.uv.faces_uvs + 3 * r + 1));
          t2 = ld2(a.uv.verts_uvs, __ldg(a.uv.faces_uvs + 3 * r + 2));
        } else {
          C0 = ld3(a.colors, w0i); C1 = ld3(a.colors, w1i); C2 = ld3(a.colors, w2i);
        }
        if (LIT) {
          X0 = ld3(a.verts_world, w0i); X1 = ld3(a.verts_world, w1i); X2 = ld3(a.verts_world, w2i);
          N0 = ld3(a.normals, w0i); N1 = ld3(a.normals, w1i); N2 = ld3(a.normals, w2i);
        }
      }
      F3 tex = interp3(b0, b1, b2, C0, C1, C2);
      UvTap tap = {};
      F3 dtu = {0, 0, 0}, dtv = {0, 0, 0};
      if (use_uv && on) {
        tap = uv_tap(a.uv, b0 * t0.x + b1 * t1.x + b2 * t2.x, b0 * t0.y + b1 * t1.y + b2 * t2.y);
        tex = uv_sample(a.uv, tap, &dtu, &dtv);
      }
      const F3 P = interp3(b0, b1, b2, X0, X1, X2);
      const F3 nr = interp3(b0, b1, b2, N0, N1, N2);
      Lit lit;
      const F3 c = phong_color<LIGHT, true>(vp, P, nr, tex, lit);
      float wn = 1.0f;
      if (SOFT) {
        const float eps = 1e-10f;
        const float inv_zrange = __frcp_rn(vp.zfar - vp.znear);
        const float inv_gamma = __frcp_rn(a.gamma);
        const float zinv = (vp.zfar - z) * inv_zrange;
        const float zmax = fmaxf(zinv, eps);
        const float E = __expf((zinv - zmax) * inv_gamma);
        const float w = p * E;
        const float dexp = __expf((eps - zmax) * inv_gamma);
        const float delta = fmaxf(dexp, eps);
        const float inv_den = __frcp_rn(w + delta);
        const F3 rgb = {(w * c.x + delta * a.bg0) * inv_den, (w * c.y + delta * a.bg1) * inv_den,
                        (w * c.z + delta * a.bg2) * inv_den};
        const float g_rgb = g.x * rgb.x + g.y * rgb.y + g.z * rgb.z;
        const float g_w = ((g.x * c.x + g.y * c.y + g.z * c.z) - g_rgb) * inv_den;
        const float g_delta = ((g.x * a.bg0 + g.y * a.bg1 + g.z * a.bg2) - g_rgb) * inv_den;
        g_p += g_w * E;
        // d/dz_inv: directly through w and (when z_inv is the softmax max) through z_max
        const float g_zinv = g_w * w * inv_gamma;
        float g_zmax = (dexp > eps ? -g_delta * delta * inv_gamma : 0.0f) - g_zinv;
        const float g_zinv_total = g_zinv + (zinv > eps ? g_zmax : 0.0f);
        gz += -g_zinv_total * inv_zrange;
        wn = w * inv_den;
      }
      const F3 gc = {g.x * wn, g.y * wn, g.z * wn};
      F3 gT, gP, gN, g_lv, g_cam;
      phong_color_bwd<LIGHT, true>(vp, tex, lit, gc, gT, gP, gN, g_lv, g_cam);
      if (use_uv) {
        const float gu = dot3(gT, dtu), gvv = dot3(gT, dtv);
        gb0 += gu * t0.x + gvv * t0.y; gb1 += gu * t1.x + gvv * t1.y; gb2 += gu * t2.x + gvv * t2.y;
        if (a.g_tex_map && on) uv_scatter(a.g_tex_map, tap, gT);
      } else {
        gb0 += dot3(gT, C0); gb1 += dot3(gT, C1); gb2 += dot3(gT, C2);
      }
      if (LIT) {
        gb0 += dot3(gP, X0) + dot3(gN, N0); gb1 += dot3(gP, X1) + dot3(gN, N1); gb2 += dot3(gP, X2) + dot3(gN, N2);
      }
      auto scatter = [&](float4* base, F3 gv) {
        const float v[9] = {b0 * gv.x, b0 * gv.y, b0 * gv.z, b1 * gv.x, b1 * gv.y, b1 * gv.z,
                            b2 * gv.x, b2 * gv.y, b2 * gv.z};
        warp_groups_add_xyz3(wg, key, v, base, w0i, w1i, w2i);
      };
      if (a.g_colors && !use_uv) scatter(a.g_colors, gT);
      if (LIT) {
        if (a.g_verts_world) scatter(a.g_verts_world, gP);
        if (a.g_normals) scatter(a.g_normals, gN);
        if (a.g_view_params) {
          float vals[6] = {g_lv.x, g_lv.y, g_lv.z, g_cam.x, g_cam.y, g_cam.z};
          const int n0 = __shfl_sync(0xffffffffu, n, 0);
          const bool uniform = __all_sync(0xffffffffu, !on || n == n0);
          float* gp = a.g_view_params + (size_t)(uniform ? n0 : n) * TRB_VIEW_PARAM_STRIDE;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            if (i < 3 ? !a.want_light_grad : !a.want_cam_grad) continue;  // uniform
            if (uniform) {
              const float sum = warp_sum(on ? vals[i] : 0.0f);
              if ((threadIdx.x & 31) == 0 && sum != 0.0f) atomicAdd(gp + (i < 3 ? i : 10 + i), sum);
            } else if (on && vals[i] != 0.0f) {
              atomicAdd(gp + (i < 3 ? i : 10 + i), vals[i]);
            }
          }
        }
      }
    }
    if (SOFT || SIL) gd += g_p * p * q * (-1.0f / a.sigma);
  }

  // ---- rasteriser backward
  if (!a.g_verts_ndc) return;
  float gv[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) gv[i] = 0.0f;
  int i0 = 0, i1 = 0, i2 = 0;
  if (on) {
    if (a.faces != nullptr) { i0 = w0i + vd.vert_delta; i1 = w1i + vd.vert_delta; i2 = w2i + vd.vert_delta; }
    else { i0 = 3 * (int)r; i1 = i0 + 1; i2 = i0 + 2; }
    FaceXYZ v;
    const float* p0 = a.verts_ndc + 3 * (size_t)i0; const float* p1 = a.verts_ndc + 3 * (size_t)i1;
    const float* p2 = a.verts_ndc + 3 * (size_t)i2;
    v.x0 = __ldg(p0); v.y0 = __ldg(p0 + 1); v.z0 = __ldg(p0 + 2);
    v.x1 = __ldg(p1); v.y1 = __ldg(p1 + 1); v.z1 = __ldg(p1 + 2);
    v.x2 = __ldg(p2); v.y2 = __ldg(p2 + 1); v.z2 = __ldg(p2 + 2);
    const float px = pix_to_ndc_fast(W - 1 - xi, W, H), py = pix_to_ndc_fast(H - 1 - yi, H, W);
    // the saved distance is negative (or -0) exactly when the forward found the sample inside
    sample_backward_fast(v, px, py, persp, clip, signbit(d), gz, gb0, gb1, gb2, gd, gv);
  }
  warp_groups_add_xyz3(wg, key, gv, a.g_verts_ndc, i0, i1, i2);
}

template <bool K1, int SHADER, int LIGHT>
__device__ __forceinline__ void render_backward_pixel(const BwdArgs& a, bool live, int pixi, int nk, float4* s_park) {
  const int NT = blockDim.x;
  const int tid = threadIdx.x;
  const int H = a.H, W = a.W;
  const int K = K1 ? 1 : a.K;
  const int n = pixi / (H * W);
  const int prem = pixi - n * (H * W);
  const int yi = prem / W, xi = prem - yi * W;
  const trb_view vd = a.views[n];
  const float px = pix_to_ndc(W - 1 - xi, W, H);
  const float py = pix_to_ndc(H - 1 - yi, H, W);
  const bool persp = a.flags & TRB_PERSPECTIVE_CORRECT, clip = a.flags & TRB_CLIP_BARYCENTRIC;
  const size_t pix = (size_t)pixi;
  const size_t s0 = pix * K;
  constexpr bool PHONG = (SHADER == TRB_SHADER_SOFT_PHONG || SHADER == TRB_SHADER_HARD_PHONG);
  constexpr bool SOFT = (SHADER == TRB_SHADER_SOFT_PHONG);
  constexpr bool SIL = (SHADER == TRB_SHADER_SOFT_SILHOUETTE);

  // nk: layers of this pixel, written next to the covered-pixel list by the forward (no scan for the first -1)
  float4 g = make_float4(0, 0, 0, 0);
  if (SHADER != TRB_SHADER_NONE && nk > 0) g = __ldg(reinterpret_cast<const float4*>(a.g_images) + pix);

  ViewParams vp;
  if (PHONG) vp = load_view_params(a.view_params, n);
  const float eps = 1e-10f;
  const float zrange = PHONG ? vp.zfar - vp.znear : 1.0f;
  // fast-math throughout: gradients are held to 1e-3 against fp64; the forward image came from the precise kernels
  const float inv_zrange = __frcp_rn(zrange), inv_sigma = __frcp_rn(a.sigma), inv_gamma = __frcp_rn(a.gamma);

  // ---- pass A: blend bookkeeping (no colours yet)
  float zmax = eps; int kmax = -1;
  float prod_nz = 1.0f; int zeros = 0;
  float wsum = 0.0f, delta = 0.0f, den = 1.0f;
  if (SOFT || SIL) {
    for (int k = 0; k < nk; ++k) {
      const float q = 1.0f - sigmoid_fast(-a.dists[s0 + k] * inv_sigma);
      if (q == 0.0f) ++zeros; else prod_nz *= q;
      if (SOFT) {
        const float zinv = (vp.zfar - a.zbuf[s0 + k]) * inv_zrange;
        if (zinv > zmax) { zmax = zinv; kmax = k; }
      }
    }
  }
  if (SOFT) {
    for (int k = 0; k < nk; ++k) {
      const float zinv = (vp.zfar - a.zbuf[s0 + k]) * inv_zrange;
      wsum += sigmoid_fast(-a.dists[s0 + k] * inv_sigma) * __expf((zinv - zmax) * inv_gamma);
    }
    delta = fmaxf(__expf((eps - zmax) * inv_gamma), eps);
    den = wsum + delta;
  }
  const float inv_den = __frcp_rn(den);

  // ---- pass B: lighting model forward + backward, vertex-attribute scatters
  const int nshade = (SHADER == TRB_SHADER_HARD_PHONG) ? min(nk, 1) : (PHONG ? nk : 0);
  F3 acc = {0, 0, 0};
  F3 g_lv_acc = {0, 0, 0}, g_cam_acc = {0, 0, 0};
  float4 park1 = make_float4(0, 0, 0, 0);
  if (PHONG) {
    const int nloop = __reduce_max_sync(0xffffffffu, nshade);
    for (int k = 0; k < nloop; ++k) {
      const bool on = k < nshade;
      int key = -1, i0 = 0, i1 = 0, i2 = 0;
      float b0 = 0, b1 = 0, b2 = 0;
      F3 gP = {0, 0, 0}, gN = {0, 0, 0}, gT = {0, 0, 0};
      if (on) {
        const size_t s = s0 + k;
        const long long f = a.p2f[s];
        key = (int)f;
        const size_t r = (size_t)(vd.face_start + (int)(f - vd.p2f_base));
        i0 = __ldg(a.faces + 3 * r); i1 = __ldg(a.faces + 3 * r + 1); i2 = __ldg(a.faces + 3 * r + 2);
        b0 = a.bary[s * 3]; b1 = a.bary[s * 3 + 1]; b2 = a.bary[s * 3 + 2];
        const bool use_uv = a.uv.map != nullptr;
        F3 C0 = {0, 0, 0}, C1 = C0, C2 = C0, tex, dtu = C0, dtv = C0;
        F2 t0 = {0, 0}, t1 = t0, t2 = t0;
        UvTap tap = {};
        if (use_uv) {
          t0 = ld2(a.uv.verts_uvs, __ldg(a.uv.faces_uvs + 3 * r)); t1 = ld2(a.uv.verts_uvs, __ldg(a.uv.faces_uvs + 3 * r + 1));
          t2 = ld2(a.uv.verts_uvs, __ldg(a.uv.faces_uvs + 3 * r + 2));
          tap = uv_tap(a.uv, b0 * t0.x + b1 * t1.x + b2 * t2.x, b0 * t0.y + b1 * t1.y + b2 * t2.y);
          tex = uv_sample(a.uv, tap, &dtu, &dtv);
        } else {
          C0 = ld3(a.colors, i0); C1 = ld3(a.colors, i1); C2 = ld3(a.colors, i2);
          tex = interp3(b0, b1, b2, C0, C1, C2);
        }
        F3 X0 = {0, 0, 0}, X1 = X0, X2 = X0, N0 = X0, N1 = X0, N2 = X0, P = X0, nr = X0;
        if (LIGHT != TRB_LIGHT_AMBIENT) {
          X0 = ld3(a.verts_world, i0); X1 = ld3(a.verts_world, i1); X2 = ld3(a.verts_world, i2);
          N0 = ld3(a.normals, i0); N1 = ld3(a.normals, i1); N2 = ld3(a.normals, i2);
          P = interp3(b0, b1, b2, X0, X1, X2);
          nr = interp3(b0, b1, b2, N0, N1, N2);
        }
        Lit lit;
        const F3 c = phong_color<LIGHT, true>(vp, P, nr, tex, lit);
        float wn = 1.0f;
        float gdotc = 0.0f;
        if (SOFT) {
          const float zinv = (vp.zfar - a.zbuf[s]) * inv_zrange;
          const float w = sigmoid_fast(-a.dists[s] * inv_sigma) * __expf((zinv - zmax) * inv_gamma);
          acc.x += w * c.x; acc.y += w * c.y; acc.z += w * c.z;
          gdotc = g.x * c.x + g.y * c.y + g.z * c.z;
          wn = w * inv_den;
        }
        const F3 gc = {g.x * wn, g.y * wn, g.z * wn};
        F3 g_lv, g_cam;
        phong_color_bwd<LIGHT, true>(vp, tex, lit, gc, gT, gP, gN, g_lv, g_cam);
        g_lv_acc.x += g_lv.x; g_lv_acc.y += g_lv.y; g_lv_acc.z += g_lv.z;
        g_cam_acc.x += g_cam.x; g_cam_acc.y += g_cam.y; g_cam_acc.z += g_cam.z;
        float gb0, gb1, gb2;
        if (use_uv) {
          const float gu = dot3(gT, dtu), gvv = dot3(gT, dtv);
          gb0 = gu * t0.x + gvv * t0.y; gb1 = gu * t1.x + gvv * t1.y; gb2 = gu * t2.x + gvv * t2.y;
          if (a.g_tex_map) uv_scatter(a.g_tex_map, tap, gT);
        } else {
          gb0 = dot3(gT, C0); gb1 = dot3(gT, C1); gb2 = dot3(gT, C2);
        }
        if (LIGHT != TRB_LIGHT_AMBIENT) {
          gb0 += dot3(gP, X0) + dot3(gN, N0); gb1 += dot3(gP, X1) + dot3(gN, N1); gb2 += dot3(gP, X2) + dot3(gN, N2);
        }
        const float4 pk = make_float4(gb0, gb1, gb2, gdotc);
        if (K1) park1 = pk; else s_park[k * NT + tid] = pk;
      }
      {
        const WarpGroups wg = warp_groups(key);  // one grouping for the three scatters of this layer
        auto scatter = [&](float4* base, F3 gv) {
#ifndef TRB_BWD_SCATTER_INLINE
          scatter_xyz3_outlined(wg.leaders, wg.any, wg.aggregate, key, b0 * gv.x, b0 * gv.y, b0 * gv.z, b1 * gv.x,
                                b1 * gv.y, b1 * gv.z, b2 * gv.x, b2 * gv.y, b2 * gv.z, base, i0, i1, i2);
#else
          const float v[9] = {b0 * gv.x, b0 * gv.y, b0 * gv.z, b1 * gv.x, b1 * gv.y, b1 * gv.z,
                              b2 * gv.x, b2 * gv.y, b2 * gv.z};
          warp_groups_add_xyz3(wg, key, v, base, i0, i1, i2);
#endif
        };
        if (a.g_colors && a.uv.map == nullptr) scatter(a.g_colors, gT);
        if (LIGHT != TRB_LIGHT_AMBIENT) {
          if (a.g_verts_world) scatter(a.g_verts_world, gP);
          if (a.g_normals) scatter(a.g_normals, gN);
        }
      }
    }
    if (LIGHT != TRB_LIGHT_AMBIENT && a.g_view_params) {
      // one atomic per warp and component while the warp stays inside one view
      float vals[6] = {g_lv_acc.x, g_lv_acc.y, g_lv_acc.z, g_cam_acc.x, g_cam_acc.y, g_cam_acc.z};
      float* gp = a.g_view_params + (size_t)n * TRB_VIEW_PARAM_STRIDE;
      const int n0 = __shfl_sync(0xffffffffu, n, 0);
      if (__all_sync(0xffffffffu, !live || n == n0)) {
        float* gp0 = a.g_view_params + (size_t)n0 * TRB_VIEW_PARAM_STRIDE;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float sum = warp_sum(live ? vals[i] : 0.0f);
          if ((tid & 31) == 0 && sum != 0.0f) atomicAdd(gp0 + (i < 3 ? i : 10 + i), sum);
        }
      } else if (live) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
          if (vals[i] != 0.0f) atomicAdd(gp + (i < 3 ? i : 10 + i), vals[i]);
      }
    }
  }

  // ---- pass C: blend backward -> (g_z, g_bary, g_dist) per sample -> rasteriser backward
  if (!a.g_verts_ndc) return;
  float g_rgb = 0.0f, g_zmax = 0.0f;
  if (SOFT) {
    const F3 rgb = {(acc.x + delta * a.bg0) * inv_den, (acc.y + delta * a.bg1) * inv_den,
                    (acc.z + delta * a.bg2) * inv_den};
    g_rgb = g.x * rgb.x + g.y * rgb.y + g.z * rgb.z;
    const float g_delta = ((g.x * a.bg0 + g.y * a.bg1 + g.z * a.bg2) - g_rgb) * inv_den;
    const bool delta_clamped = !(__expf((eps - zmax) * inv_gamma) > eps);
    g_zmax = delta_clamped ? 0.0f : -g_delta * delta * inv_gamma;
  }
  const int nloop = __reduce_max_sync(0xffffffffu, nk);
  for (int k = 0; k <= nloop; ++k) {
    // iteration k == nloop routes the softmax-max gradient to its arg-max layer (soft Phong only)
    const bool tail = (k == nloop);
    if (tail && !SOFT) break;
    const int kk = tail ? kmax : k;
    const bool on = tail ? (kmax >= 0 && g_zmax != 0.0f) : (k < nk);
    int key = -1, i0 = 0, i1 = 0, i2 = 0;
    float gv[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) gv[i] = 0.0f;
    if (on) {
      const size_t s = s0 + kk;
      const long long f = a.p2f[s];
      float gz = 0.0f, gd = 0.0f, gb0 = 0.0f, gb1 = 0.0f, gb2 = 0.0f;
      if (tail) {
        gz = -g_zmax * inv_zrange;
      } else {
        if (SOFT || SIL) {
          const float p = sigmoid_fast(-a.dists[s] * inv_sigma);
          const float q = 1.0f - p;
          const float others = zeros == 0 ? prod_nz * __frcp_rn(q) : (zeros == 1 && q == 0.0f ? prod_nz : 0.0f);
          float g_p = g.w * others;
          if (SOFT) {
            const float zinv = (vp.zfar - a.zbuf[s]) * inv_zrange;
            const float E = __expf((zinv - zmax) * inv_gamma);
            const float4 pk = K1 ? park1 : s_park[k * NT + tid];
            const float g_w = (pk.w - g_rgb) * inv_den;
            g_p += g_w * E;
            const float g_zinv = g_w * (p * E) * inv_gamma;
            g_zmax -= g_zinv;
            gz = -g_zinv * inv_zrange;
          }
          gd = g_p * p * q * (-inv_sigma);
        }
        if (PHONG && k < nshade) {
          const float4 pk = K1 ? park1 : s_park[k * NT + tid];
          gb0 = pk.x; gb1 = pk.y; gb2 = pk.z;
        }
        if (a.g_zbuf) gz += a.g_zbuf[s];
        if (a.g_dists) gd += a.g_dists[s];
        if (a.g_bary) { gb0 += a.g_bary[s * 3]; gb1 += a.g_bary[s * 3 + 1]; gb2 += a.g_bary[s * 3 + 2]; }
      }
      const int lf = (int)(f - vd.p2f_base);
      const size_t r = (size_t)(vd.face_start + lf);
      if (a.faces != nullptr) {
        i0 = __ldg(a.faces + 3 * r) + vd.vert_delta;
        i1 = __ldg(a.faces + 3 * r + 1) + vd.vert_delta;
        i2 = __ldg(a.faces + 3 * r + 2) + vd.vert_delta;
      } else {
        i0 = 3 * (int)r; i1 = i0 + 1; i2 = i0 + 2;
      }
      const FaceXYZ v = load_face(a.verts_ndc, a.faces, vd, lf);
      // the saved distance is negative (or -0) exactly when the forward found the sample inside
      sample_backward_fast(v, px, py, persp, clip, signbit(a.dists[s]), gz, gb0, gb1, gb2, gd, gv);
      key = (int)f;
    }
    {
      // K > 1 with Phong shading is bound by instruction issue / fetch, not by the reductions: no run merging there
      const WarpGroups wg = warp_groups(key, !PHONG);
#ifndef TRB_BWD_SCATTER_INLINE
      if (PHONG) {
        scatter_xyz3_outlined(wg.leaders, wg.any, wg.aggregate, key, gv[0], gv[1], gv[2], gv[3], gv[4], gv[5], gv[6],
                              gv[7], gv[8], a.g_verts_ndc, i0, i1, i2);
        continue;
      }
#endif
#ifdef TRB_KN_STATS
      const unsigned bw_have = __ballot_sync(__activemask(), key >= 0);
      if ((tid & 31) == 0 && wg.any) {
        atomicAdd(&g_bw_stats[0], 1ull);
        atomicAdd(&g_bw_stats[1], (unsigned long long)__popc(bw_have));
        atomicAdd(&g_bw_stats[2], (unsigned long long)__popc(wg.leaders));
        atomicAdd(&g_bw_stats[3], wg.aggregate ? 1ull : 0ull);
      }
#endif
      warp_groups_add_xyz3(wg, key, gv, a.g_verts_ndc, i0, i1, i2);
    }
  }
}

// ---- soft silhouette, faces_per_pixel > 1: FOUR warps per group of 32 covered pixels -------------------------
// camera_pose_optimizer.py:116-121 (K = 50): one thread per pixel walked up to 50 layers one after the other, each
// step a chain of dependent loads (face -> vertices) -- 41 k covered pixels are 1,300 warps of long serial loops
// (C3: 0.125 ms at 8 % occupancy).  Here a 128-thread CTA takes 32 consecutive covered pixels; warp w handles
// layers w, w+4, ... of all 32, so a warp still sees NEIGHBOURING PIXELS AT THE SAME LAYER (runs of one face merge
// into one reduction, as in the one-warp version) while the serial chain is a quarter as long and four times as
// many warps are in flight.  The only cross-layer quantity of sigmoid_alpha_blend, prod_k (1 - p_k), goes through
// shared memory once.
__global__ void __launch_bounds__(128, 6)
silhouette_backward_kn_kernel(const BwdArgs a) {
  pdl_wait();
  __shared__ float s_prod[4][32];
  __shared__ int s_zero[4][32];
  const int count = a.hit_pixels[0];
  const int H = a.H, W = a.W, K = a.K, HW = H * W;
  const bool persp = a.flags & TRB_PERSPECTIVE_CORRECT, clip = a.flags & TRB_CLIP_BARYCENTRIC;
  const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5;
  const float inv_sigma = 1.0f / a.sigma;
  const int groups = (count + 31) >> 5;
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int hi = grp * 32 + lane;
    const bool live = hi < count;
    const int pixi = live ? a.hit_pixels[1 + hi] : 0;
    const int nk = live ? a.hit_counts[hi] : 0;
    const int n = pixi / HW;
    const int prem = pixi - n * HW;
    const int yi = prem / W, xi = prem - yi * W;
    const trb_view vd = a.views[n];
    const float px = pix_to_ndc_fast(W - 1 - xi, W, H), py = pix_to_ndc_fast(H - 1 - yi, H, W);
    const size_t s0 = (size_t)pixi * K;
    // ---- pass A: this warp's factors of the product
    int zeros = 0;
    float prod_nz = 1.0f;
    for (int k = sub; k < nk; k += 4) {
      const float q = 1.0f - sigmoid_fast(-a.dists[s0 + k] * inv_sigma);
      if (q == 0.0f) ++zeros; else prod_nz *= q;
    }
    s_prod[sub][lane] = prod_nz; s_zero[sub][lane] = zeros;
    __syncthreads();
    prod_nz = s_prod[0][lane] * s_prod[1][lane] * s_prod[2][lane] * s_prod[3][lane];
    zeros = s_zero[0][lane] + s_zero[1][lane] + s_zero[2][lane] + s_zero[3][lane];
    __syncthreads();   // the next group overwrites s_prod / s_zero
    const float gw = live ? __ldg(a.g_images + 4 * (size_t)pixi + 3) : 0.0f;
    // ---- pass C: per layer blend backward -> rasteriser backward -> scatter
    const int mine = nk > sub ? (nk - sub + 3) >> 2 : 0;
    const int nloop = __reduce_max_sync(0xffffffffu, mine);
    for (int r = 0; r < nloop; ++r) {
      const bool on = r < mine;
      int key = -1, i0 = 0, i1 = 0, i2 = 0;
      float gv[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) gv[i] = 0.0f;
      if (on) {
        const size_t s = s0 + sub + 4 * r;
        const long long f = a.p2f[s];
        const float d = a.dists[s];
        const float p = sigmoid_fast(-d * inv_sigma), q = 1.0f - p;
        const float others = zeros == 0 ? prod_nz * __frcp_rn(q) : (zeros == 1 && q == 0.0f ? prod_nz : 0.0f);
        float gd = gw * others * p * q * (-inv_sigma);
        float gz = 0.0f, gb0 = 0.0f, gb1 = 0.0f, gb2 = 0.0f;
        if (a.g_zbuf) gz = a.g_zbuf[s];
        if (a.g_dists) gd += a.g_dists[s];
        if (a.g_bary) { gb0 = a.g_bary[s * 3]; gb1 = a.g_bary[s * 3 + 1]; gb2 = a.g_bary[s * 3 + 2]; }
        const int lf = (int)(f - vd.p2f_base);
        const size_t row = (size_t)(vd.face_start + lf);
        if (a.faces != nullptr) {
          i0 = __ldg(a.faces + 3 * row) + vd.vert_delta;
          i1 = __ldg(a.faces + 3 * row + 1) + vd.vert_delta;
          i2 = __ldg(a.faces + 3 * row + 2) + vd.vert_delta;
        } else {
          i0 = 3 * (int)row; i1 = i0 + 1; i2 = i0 + 2;
        }
        const FaceXYZ v = load_face(a.verts_ndc, a.faces, vd, lf);
        sample_backward_fast(v, px, py, persp, clip, signbit(d), gz, gb0, gb1, gb2, gd, gv);
        key = (int)f;
      }
      const WarpGroups wg = warp_groups(key, true);
      warp_groups_add_xyz3(wg, key, gv, a.g_verts_ndc, i0, i1, i2);
    }
  }
}

// ---- host side ----------------------------------------------------------------------------------
static int check_render_cfg(const trb_render_config* c) {
  if (!c) return TRB_ERR_BAD_ARG;
  const trb_shade_config& s = c->shade;
  if (s.N < 0 || s.H < 1 || s.W < 1 || s.K < 1) return TRB_ERR_BAD_ARG;
  if (s.K > TRB_MAX_FACES_PER_PIXEL) return TRB_ERR_K_TOO_LARGE;
  if (s.N > 65535 || (int64_t)s.N * s.H * s.W >= 2147483647ll) return TRB_ERR_BAD_ARG;
  if (s.shader < -1 || s.shader > 2 || s.light_kind < 0 || s.light_kind > 2) return TRB_ERR_BAD_ARG;
  if (s.shader != TRB_SHADER_NONE && (!(s.sigma > 0.0f) || !(s.gamma > 0.0f))) return TRB_ERR_BAD_ARG;
  if (s.shader >= 0 && s.shader != TRB_SHADER_SOFT_SILHOUETTE && s.texture_mode != TRB_TEX_VERTEX &&
      s.texture_mode != TRB_TEX_UV)
    return TRB_ERR_BAD_ARG;
  if (!(c->blur_radius >= 0.0f) || !(c->z_clip_value == c->z_clip_value) || c->max_face_count < 0 || c->max_vert_count < 0 || c->pair_capacity < 0 ||
      c->num_world_verts < 0 || c->num_faces < 0 || c->num_ndc_verts < 0)
    return TRB_ERR_BAD_ARG;
  return TRB_OK;
}

static inline bool is_phong(int shader) {
  return shader == TRB_SHADER_SOFT_PHONG || shader == TRB_SHADER_HARD_PHONG;
}

static int launch_render_fine_k1(int shader, int light, int N, cudaStream_t st, const FineArgs& a) {
  // strips of kStrip tiles once that still leaves two full waves of CTAs (148 SMs x TRB_K1_CTAS resident)
  const long long tiles = (long long)a.tg.tiles_x * a.tg.tiles_y * N;
  const bool strips = tiles >= 2ll * kStrip * kNumSMs * TRB_K1_CTAS;
  const dim3 grid(strips ? ceil_div(a.tg.tiles_x, kStrip) : a.tg.tiles_x, a.tg.tiles_y, N);
#define TRB_RF1(SH, L)                                                                                   \
  do {                                                                                                   \
    if (strips) TRB_CUDA_TRY(launch_pdl(render_fine_k1_kernel<SH, L, kStrip>, grid, dim3(256), 0, st, a)); \
    else TRB_CUDA_TRY(launch_pdl(render_fine_k1_kernel<SH, L, 1>, grid, dim3(256), 0, st, a));           \
  } while (0)
  if (shader == TRB_SHADER_NONE) TRB_RF1(TRB_SHADER_NONE, 0);
  else if (shader == TRB_SHADER_SOFT_SILHOUETTE) TRB_RF1(TRB_SHADER_SOFT_SILHOUETTE, 0);
  else if (shader == TRB_SHADER_SOFT_PHONG) {
    if (light == 0) TRB_RF1(TRB_SHADER_SOFT_PHONG, 0); else if (light == 1) TRB_RF1(TRB_SHADER_SOFT_PHONG, 1);
    else TRB_RF1(TRB_SHADER_SOFT_PHONG, 2);
  } else {
    if (light == 0) TRB_RF1(TRB_SHADER_HARD_PHONG, 0); else if (light == 1) TRB_RF1(TRB_SHADER_HARD_PHONG, 1);
    else TRB_RF1(TRB_SHADER_HARD_PHONG, 2);
  }
#undef TRB_RF1
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

template <bool K1>
static int launch_render_backward(int shader, int light, dim3 grid, int nt, size_t dyn, cudaStream_t st,
                                  const BwdArgs& a) {
#define TRB_RB(SH, L)                                                                             \
  do {                                                                                            \
    auto kern = render_backward_kernel<K1, SH, L>;                                                \
    if (dyn > 0)                                                                                  \
      TRB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
    TRB_CUDA_TRY(launch_pdl(kern, grid, dim3(nt), dyn, st, a));                                   \
  } while (0)
  if (shader == TRB_SHADER_NONE) TRB_RB(TRB_SHADER_NONE, 0);
  else if (shader == TRB_SHADER_SOFT_SILHOUETTE) TRB_RB(TRB_SHADER_SOFT_SILHOUETTE, 0);
  else if (shader == TRB_SHADER_SOFT_PHONG) {
    if (light == 0) TRB_RB(TRB_SHADER_SOFT_PHONG, 0); else if (light == 1) TRB_RB(TRB_SHADER_SOFT_PHONG, 1);
    else TRB_RB(TRB_SHADER_SOFT_PHONG, 2);
  } else {
    if (light == 0) TRB_RB(TRB_SHADER_HARD_PHONG, 0); else if (light == 1) TRB_RB(TRB_SHADER_HARD_PHONG, 1);
    else TRB_RB(TRB_SHADER_HARD_PHONG, 2);
  }
#undef TRB_RB
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

int launch_render_fine(int shader, int light, int N, cudaStream_t st, const FineArgs& a) {
  if (a.K == 1)
    return launch_render_fine_k1(shader, light, N, st, a);
  return launch_render_fine_kn(shader, light, N, st, a);
}

static cudaEvent_t g_dbg_events[4] = {nullptr, nullptr, nullptr, nullptr};

}  // namespace trb

using namespace trb;

extern "C" int trb_debug_set_events(void* fwd_start, void* fwd_stop, void* bwd_start, void* bwd_stop) {
  g_dbg_events[0] = (cudaEvent_t)fwd_start; g_dbg_events[1] = (cudaEvent_t)fwd_stop;
  g_dbg_events[2] = (cudaEvent_t)bwd_start; g_dbg_events[3] = (cudaEvent_t)bwd_stop;
  return TRB_OK;
}

extern "C" int trb_render_sizes(const trb_render_config* cfg, size_t* workspace_bytes, int64_t* num_tiles,
                                int64_t* backward_scratch_floats) {
  const int rc = check_render_cfg(cfg);
  if (rc != TRB_OK) return rc;
  const trb_shade_config& s = cfg->shade;
  const TileGrid tg = make_tile_grid(s.H, s.W, s.K);
  if (workspace_bytes) *workspace_bytes = make_ws_layout(s.N, tg, cfg->pair_capacity).total;
  // hit_pixels = [count, pixel ids ... (N*H*W slots), layer counts ... (N*H*W slots, faces_per_pixel > 1 only)]
  // count, pixel ids, (K > 1) layer counts, and N floats: the per-view sums of the alpha channel
  if (num_tiles) *num_tiles = (int64_t)s.N * s.H * s.W * (s.K > 1 ? 2 : 1) + 1 + s.N;
  // backward scratch: float4 accumulators for grad NDC verts [num_ndc_verts], world verts, colours and
  // normals [V each], then grad of the raw normals [V,3]
  if (backward_scratch_floats) *backward_scratch_floats = 4 * cfg->num_ndc_verts + 15 * cfg->num_world_verts;
  return TRB_OK;
}

extern "C" int trb_render_forward(const trb_render_config* cfg, const trb_view* views,
                                  const float* verts_world, const int32_t* faces, const float* vert_colors,
                                  const float* R, const float* T, const float* proj, float* view_params,
                                  float* verts_ndc, float* normals_raw, float* normals, int64_t* pix_to_face,
                                  float* zbuf, float* bary, float* dists, float* images, int32_t* tile_hit,
                                  void* workspace, size_t workspace_bytes, int32_t* stats,
                                  const trb_uv_texture* uv, const trb_render_extras* extras, int device,
                                  trb_stream_t stream) {
  int rc = check_render_cfg(cfg);
  if (rc != TRB_OK) return rc;
  const trb_shade_config& sc = cfg->shade;
  const int N = sc.N, H = sc.H, W = sc.W, K = sc.K;
  if (N == 0) return TRB_OK;
  if (!views || !R || !T || !proj || !verts_ndc || !pix_to_face || !zbuf || !bary || !dists || !tile_hit ||
      !workspace)
    return TRB_ERR_BAD_ARG;
  if (cfg->max_face_count > 0 && (!verts_world || !faces)) return TRB_ERR_BAD_ARG;
  if (sc.shader != TRB_SHADER_NONE && !images) return TRB_ERR_BAD_ARG;
  const bool use_uv = is_phong(sc.shader) && sc.texture_mode == TRB_TEX_UV;
  if (use_uv && (!uv || !uv->map || !uv->verts_uvs || !uv->faces_uvs || uv->map_h < 1 || uv->map_w < 1))
    return TRB_ERR_BAD_ARG;
  if (is_phong(sc.shader) && (!view_params || (!vert_colors && !use_uv) || !normals_raw || !normals))
    return TRB_ERR_BAD_ARG;
  const TileGrid tg = make_tile_grid(H, W, K);
  const WsLayout ws = make_ws_layout(N, tg, cfg->pair_capacity);
  if (workspace_bytes < ws.total) return TRB_ERR_WORKSPACE;
  TRB_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;

  const bool lit = is_phong(sc.shader) && sc.light_kind != TRB_LIGHT_AMBIENT;
  rc = run_forward_stages(cfg, views, verts_world, faces, R, T, proj, view_params, verts_ndc, normals_raw, normals,
                          tile_hit, workspace, tg, ws, lit, st, extras);
  if (rc != TRB_OK) return rc;
  const float sqrt_blur = sqrtf(cfg->blur_radius);
  const float z_cull = fmaxf(cfg->z_clip_value, 0.0f);

  unsigned char* wsb = (unsigned char*)workspace;
  FineArgs a;
  a.verts_ndc = verts_ndc; a.faces = faces; a.views = views;
  a.H = H; a.W = W; a.K = K; a.blur_radius = cfg->blur_radius; a.sqrt_blur = sqrt_blur; a.z_cull = z_cull;
  a.flags = cfg->raster_flags; a.tg = tg;
  a.tile_count = (const int*)(wsb + ws.count); a.tile_offset = (const int*)(wsb + ws.offset);
  a.pairs = (const int2*)(wsb + ws.pairs);
  a.ws_header = (const int*)(wsb + ws.header); a.busy_tiles = (const int*)(wsb + ws.busy);
  a.p2f = (long long*)pix_to_face; a.zbuf = zbuf; a.bary = bary; a.dists = dists; a.images = images;
  a.hit_pixels = tile_hit;
  a.hit_counts = K > 1 ? tile_hit + 1 + (size_t)N * H * W : nullptr;
  a.alpha_sum = sc.shader != TRB_SHADER_NONE
                    ? reinterpret_cast<float*>(tile_hit + 1 + (size_t)N * H * W * (K > 1 ? 2 : 1)) : nullptr;
  a.view_params = view_params; a.verts_world = verts_world; a.normals = normals; a.colors = vert_colors;
  a.sigma = sc.sigma; a.gamma = sc.gamma; a.bg0 = sc.background[0]; a.bg1 = sc.background[1];
  a.bg2 = sc.background[2];
  a.uv = {nullptr, nullptr, nullptr, 0, 0};
  if (use_uv) a.uv = {uv->map, uv->verts_uvs, uv->faces_uvs, uv->map_h, uv->map_w};
  // sparse Fragments need the covered-pixel list (it is what tells the backward which samples exist)
  a.sparse = (cfg->sparse_fragments && sc.shader != TRB_SHADER_NONE) ? 1 : 0;
  if (g_dbg_events[0]) TRB_CUDA_TRY(cudaEventRecord(g_dbg_events[0], st));
  rc = launch_render_fine(sc.shader, sc.light_kind, N, st, a);
  if (rc != TRB_OK) return rc;
  if (g_dbg_events[1]) TRB_CUDA_TRY(cudaEventRecord(g_dbg_events[1], st));
  if (stats) {
    write_stats_kernel<<<1, 1, 0, st>>>((const int*)(wsb + ws.header), (long long)cfg->pair_capacity, stats);
    TRB_LAUNCH_CHECK();
  }
  return TRB_OK;
}

static int render_backward_impl(const trb_render_config* cfg, const trb_view* views,
                                const float* verts_world, const int32_t* faces, const float* vert_colors,
                                const float* R, const float* T, const float* proj, const float* view_params,
                                const float* verts_ndc, const float* normals_raw, const float* normals,
                                const int64_t* pix_to_face, const float* zbuf, const float* bary,
                                const float* dists, const int32_t* tile_hit, const float* grad_images,
                                const float* grad_zbuf, const float* grad_bary, const float* grad_dists,
                                float* grad_verts_world, float* grad_vert_colors, float* grad_R,
                                float* grad_T, float* grad_proj, float* grad_view_params, float* scratch,
                                const trb_uv_texture* uv, int device, trb_stream_t stream,
                                const trb_peer_sum* peer) {
  int rc = check_render_cfg(cfg);
  if (rc != TRB_OK) return rc;
  const trb_shade_config& sc = cfg->shade;
  const int N = sc.N, H = sc.H, W = sc.W, K = sc.K;
  // (the fused all-reduce is a collective: a rank with nothing to render cannot skip it)
  if (N == 0 || cfg->max_face_count == 0) return peer ? TRB_ERR_BAD_ARG : TRB_OK;
  if (!views || !verts_world || !faces || !R || !T || !proj || !verts_ndc || !pix_to_face || !zbuf || !bary ||
      !dists || !tile_hit || !scratch)
    return TRB_ERR_BAD_ARG;
  if (sc.shader != TRB_SHADER_NONE && !grad_images) return TRB_ERR_BAD_ARG;
  const bool phong = is_phong(sc.shader);
  const bool lit = phong && sc.light_kind != TRB_LIGHT_AMBIENT;
  const bool use_uv = phong && sc.texture_mode == TRB_TEX_UV;
  if (use_uv && (!uv || !uv->map || !uv->verts_uvs || !uv->faces_uvs || uv->map_h < 1 || uv->map_w < 1))
    return TRB_ERR_BAD_ARG;
  if (phong && (!view_params || (!vert_colors && !use_uv))) return TRB_ERR_BAD_ARG;
  if (lit && (!normals_raw || !normals)) return TRB_ERR_BAD_ARG;
  // multi-GPU: the shared gradients (vertices, then vertex colours) leave for the peers from inside the tail kernel
  ArPush push;
  if (peer) {
    float* segs[2]; int64_t cnts[2]; int ns = 0;
    if (grad_verts_world) { segs[ns] = grad_verts_world; cnts[ns++] = 3 * cfg->num_world_verts; }
    if (grad_vert_colors && phong && !use_uv) { segs[ns] = grad_vert_colors; cnts[ns++] = 3 * cfg->num_world_verts; }
    if (ns == 0 || !peer->epochs || !peer->error_flag || !peer->done_counter) return TRB_ERR_BAD_ARG;
    long long total = 0;
    rc = ar_fill_tables(segs, cnts, ns, peer->host_peer_inbox, peer->capacity_floats, peer->rank, peer->world,
                        push.seg, push.peers, total);
    if (rc != TRB_OK) return rc;
    push.capacity = peer->capacity_floats; push.rank = peer->rank; push.world = peer->world;
    push.epochs = peer->epochs; push.error = peer->error_flag; push.done = peer->done_counter;
  }
  TRB_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_ndc = (size_t)cfg->num_ndc_verts, V = (size_t)cfg->num_world_verts;
  if (!cfg->scratch_is_zeroed) TRB_CUDA_TRY(cudaMemsetAsync(scratch, 0, (4 * n_ndc + 15 * V) * sizeof(float), st));
  float4* g_ndc4 = reinterpret_cast<float4*>(scratch);
  float4* g_world4 = g_ndc4 + n_ndc;
  float4* g_col4 = g_world4 + V;
  float4* g_norm4 = g_col4 + V;
  const bool geom = grad_verts_world || grad_R || grad_T || grad_proj;
  // the camera centre (when derived from R, T) feeds grad_R / grad_T through grad_view_params
  float* g_vp = grad_view_params;
  const bool cam_chain = lit && cfg->camera_center_from_rt && (grad_R || grad_T);
  if (cam_chain && !g_vp) return TRB_ERR_BAD_ARG;  // caller provides the f32[N,20] buffer

  BwdArgs a;
  a.verts_ndc = verts_ndc; a.faces = faces; a.views = views; a.H = H; a.W = W; a.K = K;
  a.flags = cfg->raster_flags; a.hit_pixels = tile_hit;
  a.hit_counts = K > 1 ? tile_hit + 1 + (size_t)N * H * W : nullptr;
  a.p2f = (const long long*)pix_to_face; a.zbuf = zbuf; a.bary = bary; a.dists = dists;
  a.view_params = view_params; a.verts_world = verts_world; a.normals = normals; a.colors = vert_colors;
  a.g_images = grad_images; a.g_zbuf = grad_zbuf; a.g_bary = grad_bary; a.g_dists = grad_dists;
  a.g_verts_ndc = geom ? g_ndc4 : nullptr;
  a.g_verts_world = (lit && grad_verts_world) ? g_world4 : nullptr;
  a.g_normals = (lit && grad_verts_world) ? g_norm4 : nullptr;
  a.g_colors = (phong && grad_vert_colors && !use_uv) ? g_col4 : nullptr;
  a.uv = {nullptr, nullptr, nullptr, 0, 0};
  a.g_tex_map = nullptr;
  if (use_uv) { a.uv = {uv->map, uv->verts_uvs, uv->faces_uvs, uv->map_h, uv->map_w}; a.g_tex_map = uv->grad_map; }
  a.g_view_params = lit ? g_vp : nullptr;
  a.want_light_grad = (lit && g_vp && cfg->want_light_grad) ? 1 : 0;
  a.want_cam_grad = (lit && g_vp && (cam_chain || cfg->want_light_grad)) ? 1 : 0;
  a.sigma = sc.sigma; a.gamma = sc.gamma; a.bg0 = sc.background[0]; a.bg1 = sc.background[1];
  a.bg2 = sc.background[2];
  // block size: the K>1 Phong path parks 16 B per (layer, thread) in shared memory
  const int nt = (K == 1 || !phong) ? 128 : (K <= 24 ? 128 : (K <= 100 ? 64 : 32));
  const size_t dyn = (K > 1 && phong) ? (size_t)K * nt * 16 : 0;
  const dim3 grid(kNumSMs * (K == 1 ? TRB_BWD_CTAS : 512 / nt));
  if (g_dbg_events[2]) TRB_CUDA_TRY(cudaEventRecord(g_dbg_events[2], st));
  if (K == 1) rc = launch_render_backward<true>(sc.shader, sc.light_kind, grid, nt, 0, st, a);
  else if (sc.shader == TRB_SHADER_SOFT_SILHOUETTE && K >= 4 && a.g_verts_ndc != nullptr) {
    TRB_CUDA_TRY(launch_pdl(silhouette_backward_kn_kernel, dim3(kNumSMs * 6), dim3(128), 0, st, a));
    TRB_LAUNCH_CHECK();
  }
  else rc = launch_render_backward<false>(sc.shader, sc.light_kind, grid, nt, dyn, st, a);
  if (rc != TRB_OK) return rc;
  if (g_dbg_events[3]) TRB_CUDA_TRY(cudaEventRecord(g_dbg_events[3], st));

  const bool normals_chain = lit && grad_verts_world;
  rc = run_backward_post(cfg, views, verts_world, faces, R, T, proj, view_params, g_vp, normals_raw, g_ndc4,
                         a.g_verts_world ? g_world4 : nullptr, a.g_colors ? g_col4 : nullptr,
                         normals_chain ? g_norm4 : nullptr, grad_verts_world, grad_vert_colors, grad_R, grad_T,
                         grad_proj, geom, cam_chain, normals_chain, st, peer ? &push : nullptr);
  if (rc != TRB_OK) return rc;
  return TRB_OK;
}

extern "C" int trb_render_backward(const trb_render_config* cfg, const trb_view* views,
                                   const float* verts_world, const int32_t* faces, const float* vert_colors,
                                   const float* R, const float* T, const float* proj, const float* view_params,
                                   const float* verts_ndc, const float* normals_raw, const float* normals,
                                   const int64_t* pix_to_face, const float* zbuf, const float* bary,
                                   const float* dists, const int32_t* tile_hit, const float* grad_images,
                                   const float* grad_zbuf, const float* grad_bary, const float* grad_dists,
                                   float* grad_verts_world, float* grad_vert_colors, float* grad_R,
                                   float* grad_T, float* grad_proj, float* grad_view_params, float* scratch,
                                   const trb_uv_texture* uv, int device, trb_stream_t stream) {
  return render_backward_impl(cfg, views, verts_world, faces, vert_colors, R, T, proj, view_params, verts_ndc,
                              normals_raw, normals, pix_to_face, zbuf, bary, dists, tile_hit, grad_images, grad_zbuf,
                              grad_bary, grad_dists, grad_verts_world, grad_vert_colors, grad_R, grad_T, grad_proj,
                              grad_view_params, scratch, uv, device, stream, nullptr);
}

extern "C" int trb_render_backward_allreduce(const trb_render_config* cfg, const trb_view* views,
                                             const float* verts_world, const int32_t* faces,
                                             const float* vert_colors, const float* R, const float* T,
                                             const float* proj, const float* view_params, const float* verts_ndc,
                                             const float* normals_raw, const float* normals,
                                             const int64_t* pix_to_face, const float* zbuf, const float* bary,
                                             const float* dists, const int32_t* tile_hit, const float* grad_images,
                                             const float* grad_zbuf, const float* grad_bary, const float* grad_dists,
                                             float* grad_verts_world, float* grad_vert_colors, float* grad_R,
                                             float* grad_T, float* grad_proj, float* grad_view_params, float* scratch,
                                             const trb_uv_texture* uv, const trb_peer_sum* host_peer, int device,
                                             trb_stream_t stream) {
  if (!host_peer) return TRB_ERR_BAD_ARG;
  return render_backward_impl(cfg, views, verts_world, faces, vert_colors, R, T, proj, view_params, verts_ndc,
                              normals_raw, normals, pix_to_face, zbuf, bary, dists, tile_hit, grad_images, grad_zbuf,
                              grad_bary, grad_dists, grad_verts_world, grad_vert_colors, grad_R, grad_T, grad_proj,
                              grad_view_params, scratch, uv, device, stream, host_peer);
}

#ifdef TRB_KN_STATS
extern "C" int trb_debug_k1_phases(unsigned long long* host_out) {
  unsigned long long zero[16] = {0};
  if (cudaMemcpyFromSymbol(host_out, trb::g_k1_phase, sizeof(zero)) != cudaSuccess) return TRB_ERR_CUDA;
  if (cudaMemcpyToSymbol(trb::g_k1_phase, zero, sizeof(zero)) != cudaSuccess) return TRB_ERR_CUDA;
  return TRB_OK;
}
// Copies the backward scatter counters to `host_out[8]` and clears them (diagnostic builds only; synchronises).
extern "C" int trb_debug_bw_stats(unsigned long long* host_out) {
  unsigned long long zero[8] = {0};
  if (cudaMemcpyFromSymbol(host_out, trb::g_bw_stats, sizeof(zero)) != cudaSuccess) return TRB_ERR_CUDA;
  if (cudaMemcpyToSymbol(trb::g_bw_stats, zero, sizeof(zero)) != cudaSuccess) return TRB_ERR_CUDA;
  return TRB_OK;
}
#endif
