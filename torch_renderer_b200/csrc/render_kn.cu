// Fused fine pass for faces_per_pixel > 1 (sm_100a): per-pixel top-K rasterisation of one tile per CTA
// with the shading + blending epilogue (reference call sites: camera_pose_optimizer.py:116-121 --
// SoftSilhouette K=50 --, mesh_deformer.py:135-145, BASELINE config 5 -- 1M faces, K=8; semantics SURVEY
// A3-A8).  What differs from a plain "every pixel walks the tile's face list":
//
//  * DEPTH-ORDERED LISTS.  The binning pass stores (face, min vertex depth) pairs.  A tile whose list is
//    longer than one staging chunk is bucket-sorted by that depth in shared memory (256 buckets between the
//    list's min and max, histogram + scan + scatter with shared-memory atomics; exact per-bucket minima).
//    Faces are then staged front to back, every pixel's top-K fills with near faces first, the existing
//    exact early reject (a face whose depth lower bound is behind the pixel's K-th layer cannot enter) starts
//    to fire after a few faces, and the CTA stops as soon as every pixel's K-th layer is nearer than the
//    lower bound of everything not yet staged.  On the 1M-face sphere at 1024^2 (4,400 candidates per
//    16x16 tile, 1,500 inside each pixel's blur disc) this walks ~2 chunks instead of ~17.
//  * COALESCED FRAGMENT STORES.  Layer k of pixel p lives at (p*K + k): a thread-per-pixel store touches
//    one 4-byte word in each of 32 different sectors (ncu: 468 MB written + 150 MB read-modify-write for
//    367 MB of output, every warp stalled behind the LSU).  Here the epilogue parks KG layers of the whole
//    tile in shared memory (bank-conflict-free slot rotation) and the CTA writes them out as contiguous
//    runs: K*TX consecutive words per tile row.
//  * Tiles with an empty list stream out the -1 background with the same coalesced pattern.
#include "render_internal.cuh"
#include "stages.cuh"

namespace trb {

// 256 uniform depth buckets over a tile list's key range.  (A two-level, piecewise-linear map that gave the
// nearest ~1,500 entries 384 buckets of their own was measured on the 1M-face sphere and changed nothing --
// 40.65 M insertions either way: that mesh's per-vertex radial noise makes every face span a depth range far wider
// than the spacing of its neighbours' keys, so the order of min-vertex depths is not the order of sample depths.)
constexpr int kBuckets = 256;
// bit 31 of a top-K list entry's face id: the sample lies inside its face (decided during the walk)
constexpr int kFaceMask = 0x7fffffff;
constexpr int kInsideBit = (int)0x80000000u;

// Diagnostic counters of the K > 1 walk (build with TRB_EXTRA_NVCC_FLAGS=-DTRB_KN_STATS; read with
// trb_debug_kn_stats).  Not compiled into the product library.
#ifdef TRB_KN_STATS
__device__ unsigned long long g_kn_stats[16];
__device__ unsigned long long g_kn_phase[8];   // thread 0: [0] busy tiles, cycles in [1] list ordering, [2] staging, [3] walk, [4] epilogue
#define KN_PH(i) do { if (tid == 0) { const long long n_ = clock64(); atomicAdd(&g_kn_phase[i], (unsigned long long)(n_ - kn_pt)); kn_pt = n_; } } while (0)
#define KN_STAT(i, v) (kn_st[i] += (v))
#else
#define KN_STAT(i, v) ((void)0)
#define KN_PH(i) ((void)0)
#endif

// Block-wide maximum of a small non-negative int; contains __syncthreads (every thread must call it).
template <int NT>
__device__ __forceinline__ int block_max_sync(int v) {
  __shared__ int s_max[NT / 32];
  const int w = __reduce_max_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = w;
  __syncthreads();
  int m = 0;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) m = max(m, s_max[i]);
  return m;
}

template <int LT>
struct KnCfg {
  static constexpr int TX = 1 << LT, NT = TX * TX;   // NT: pixels of the tile
  // 8x8 tiles (K > 24, e.g. the K = 50 soft silhouette of camera_pose_optimizer.py:116-121) put TWO threads on
  // every pixel: thread s of a pixel walks every second staged face into its own top-K list and the epilogue merges
  // the two lists on the fly.  The busiest tile of the teapot at 512^2 (296 faces) kept one thread walking for
  // 467 k cycles = 243 us of a 318 us kernel: that kernel is as long as its longest serial chain, not its work.
#ifndef TRB_KN_SPLIT8
#define TRB_KN_SPLIT8 2
#endif
  static constexpr int SPLIT = (LT == 4) ? 1 : TRB_KN_SPLIT8;
  static constexpr int NTH = NT * SPLIT;              // threads of the CTA
// 16x16 tiles park 4 layers per output pass: 28.7 KB instead of 57 KB of shared memory, which is what lets THREE
// CTAs (24 warps) share an SM instead of two (same-box A/B on the 1M-face sphere: fine kernel 6.04 -> 4.71 ms)
#ifndef TRB_KN_LOGKG16
#define TRB_KN_LOGKG16 2
#endif
// 8x8 tiles (K > 24) park 4 layers per pass as well: at K = 50 the per-pixel lists already take 25.6 KB of a
// 64-thread CTA; 28.7 KB of parking on top left 3 CTAs = 6 warps per SM for a kernel that stalls 38 % on
// fixed-latency dependencies
#ifndef TRB_KN_LOGKG8
#define TRB_KN_LOGKG8 2
#endif
  static constexpr int LOGKG = (LT == 4) ? TRB_KN_LOGKG16 : TRB_KN_LOGKG8;  // layers parked per output pass
  static constexpr int KG = 1 << LOGKG;
  static constexpr int ROT = 5 - LOGKG;             // slot rotation: (kk + (p >> ROT)) & (KG - 1)
  static constexpr int STAGE_BYTES = NTH * 64;      // bb, va, vb (float4) + vc (float2) + zlo + id, one face per thread
  static constexpr int OUT_BYTES = NT * KG * 28;    // p2f (8) + zbuf (4) + dists (4) + bary (12)
  // list entries ordered per super-chunk (face ids, 4 B each)
  __host__ __device__ static constexpr int cap(int K) {
    return (LT == 4) ? (K <= 12 ? 8192 : 4096) : 2048;
  }
  __host__ __device__ static constexpr int union_bytes(int K) {
    return (STAGE_BYTES + cap(K) * 4 > OUT_BYTES) ? STAGE_BYTES + cap(K) * 4 : OUT_BYTES;
  }
};

// -1 background of one tile, written as contiguous runs (cols*K words per tile row), row by row so that no
// per-element division is needed.  16-byte stores whenever every run starts 16-byte aligned and is a whole number
// of them: full-width tiles (cols == TX: TX*K is a multiple of 4 for TX = 8, 16) of an image whose rows keep the
// alignment ((W*K) % 4 == 0) -- for K = 50 at 512^2 the scalar version was 35% of all instructions of the kernel.
template <int LT, int SHADER>
__device__ __forceinline__ void fill_tile_kn(const FineArgs& a, int n, int x0, int y0) {
  constexpr int TX = 1 << LT, NP = TX * TX, NT = KnCfg<LT>::NTH;   // NT: threads (the store loops' stride)
  const int tid = threadIdx.x;
  const int K = a.K, H = a.H, W = a.W;
  const int rows = min(TX, H - y0), cols = min(TX, W - x0);
  const int run = cols * K;
  const size_t row_stride = (size_t)W * K;
  const size_t base0 = ((size_t)(n * H + y0) * W + x0) * K;
  if (a.sparse) {
    // sparse Fragments: background samples stay unwritten
  } else if (cols == TX && (((long long)W * K) & 3) == 0) {
    const int run4 = run >> 2, run2 = run >> 1;
    const float4 m4 = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
    const longlong2 l2 = make_longlong2(-1ll, -1ll);
    for (int ly = 0; ly < rows; ++ly) {
      const size_t g = base0 + ly * row_stride;
      float4* z4 = reinterpret_cast<float4*>(a.zbuf + g);
      float4* d4 = reinterpret_cast<float4*>(a.dists + g);
      longlong2* p2 = reinterpret_cast<longlong2*>(a.p2f + g);
      float4* b4 = reinterpret_cast<float4*>(a.bary + 3 * g);
      for (int i = tid; i < run4; i += NT) { st_cs(z4 + i, m4); st_cs(d4 + i, m4); }
      for (int i = tid; i < run2; i += NT) __stcs(p2 + i, l2);
      for (int i = tid; i < 3 * run4; i += NT) st_cs(b4 + i, m4);
    }
  } else {
    for (int ly = 0; ly < rows; ++ly) {
      const size_t g = base0 + ly * row_stride;
      for (int i = tid; i < run; i += NT) {
        st_cs(a.p2f + g + i, -1ll);
        st_cs(a.zbuf + g + i, -1.0f);
        st_cs(a.dists + g + i, -1.0f);
      }
      for (int i = tid; i < 3 * run; i += NT) st_cs(a.bary + 3 * g + i, -1.0f);
    }
  }
  if (SHADER == TRB_SHADER_NONE) return;
  const int lx = tid & (TX - 1), ly = tid >> LT;
  if (tid < NP && lx < cols && ly < rows) {
    const float4 bgv = (SHADER == TRB_SHADER_SOFT_SILHOUETTE) ? make_float4(1.0f, 1.0f, 1.0f, 0.0f)
                                                              : make_float4(a.bg0, a.bg1, a.bg2, 0.0f);
    st_cs(reinterpret_cast<float4*>(a.images) + ((size_t)(n * H + y0 + ly) * W + x0 + lx), bgv);
  }
}

template <int LT, int SHADER, int LIGHT>
#ifndef TRB_KN_CTAS16
#define TRB_KN_CTAS16 3
#endif
__global__ void __launch_bounds__(KnCfg<LT>::NTH, LT == 4 ? TRB_KN_CTAS16 : (KnCfg<LT>::SPLIT == 2 ? 3 : 8))
render_fine_kn_kernel(const FineArgs a) {
  using C = KnCfg<LT>;
  // NP: pixels of the tile; NT: threads of the CTA (= NP * SPLIT); thread tid works on pixel tid % NP
  constexpr int TX = C::TX, NP = C::NT, NT = C::NTH, SPLIT = C::SPLIT, KG = C::KG, LOGKG = C::LOGKG, ROT = C::ROT;
  pdl_wait();
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const int K = a.K;
  const int CAP = C::cap(K);
  float* kz = reinterpret_cast<float*>(s_dyn);   // [K][NT] depth of the pixel's k-th nearest candidate
  int* kf = reinterpret_cast<int*>(s_dyn) + (size_t)K * NT;  // [K][NT] its face
  unsigned char* s_un = s_dyn + (size_t)K * NT * 8;
  // (a) list walk: staged faces + the depth-ordered list
  float4* s_bb = reinterpret_cast<float4*>(s_un);  // xmin, xmax, ymin, ymax (blur inflated; empty if undrawable)
  float4* s_va = s_bb + NT;                        // x0 y0 z0 x1
  float4* s_vb = s_va + NT;                        // y1 z1 x2 y2
  float2* s_vc = reinterpret_cast<float2*>(s_vb + NT);  // z2, area (= edge(v2;v0,v1) + kEps)
  float* s_zlo = reinterpret_cast<float*>(s_vc + NT);   // lower bound of the depth this face can produce
  int* s_id = reinterpret_cast<int*>(s_zlo + NT);
  int* ord_id = reinterpret_cast<int*>(s_un + C::STAGE_BYTES);  // [CAP] face ids in depth-bucket order
  // (b) epilogue: KG layers of the tile parked for the coalesced write-out
  long long* o_p2f = reinterpret_cast<long long*>(s_un);  // [NP][KG]
  float* o_z = reinterpret_cast<float*>(o_p2f + NP * KG);
  float* o_d = o_z + NP * KG;
  float* o_b = o_d + NP * KG;                              // [NP][KG][3]
  __shared__ int s_hist[kBuckets];
  __shared__ unsigned s_bmin[kBuckets];
  __shared__ float s_bound[kBuckets];  // min depth key over this bucket and every later one
  __shared__ float s_red[2 * (NT / 32)];
  __shared__ unsigned char s_chunk_bucket[8192 / NT + 1];  // bucket of the first entry of every staging chunk

  const int n = blockIdx.z;
  const trb_view vd = a.views[n];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = a.H, W = a.W;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TX;
  const int p = tid & (NP - 1);      // pixel of the tile
  const int slice = tid / NP;        // which share of the staged faces this thread walks (0 when SPLIT == 1)
  const int xi = x0 + (p & (TX - 1));
  const int yi = y0 + (p >> LT);
  const bool live = (xi < W) && (yi < H);

  const int t = (n * a.tg.tiles_y + blockIdx.y) * a.tg.tiles_x + blockIdx.x;
  int nlist = a.tile_count[t];
  if (nlist == 0) {
    fill_tile_kn<LT, SHADER>(a, n, x0, y0);
    return;
  }
  const int off = a.tile_offset[t];
  const bool overflow = off < 0;
  if (overflow) nlist = vd.face_count;

  const float px = pix_to_ndc(W - 1 - xi, W, H);
  const float py = pix_to_ndc(H - 1 - yi, H, W);
  const bool persp = a.flags & TRB_PERSPECTIVE_CORRECT, clip = a.flags & TRB_CLIP_BARYCENTRIC;
  const bool cull = a.flags & TRB_CULL_BACKFACES;
  const bool hard_edges = a.blur_radius == 0.0f;
  const float blur = a.blur_radius;
  // The depth lower bound of a face is a monotone function of its min vertex depth in these modes (see
  // the staging code below), which is what lets the CTA stop once the remaining keys are behind every
  // pixel's K-th layer.
  const bool can_bound = clip || (hard_edges && persp);

#ifdef TRB_KN_STATS
  unsigned kn_st[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (tid == 0) { KN_STAT(0, 1); KN_STAT(1, nlist); }
  const long long kn_t0 = clock64();
  long long kn_pt = kn_t0;
  if (tid == 0) { atomicMax(&g_kn_stats[14], (unsigned long long)nlist); atomicAdd(&g_kn_phase[0], 1ull); }
#endif
  int cnt = 0;
  float kth = 3.0e38f;  // depth of the pixel's K-th layer once the list is full (register copy of kz[K-1])
  for (int sbase = 0; sbase < nlist; sbase += CAP) {
    const int m = min(CAP, nlist - sbase);
    const bool ordered = !overflow && can_bound && m > NT;
    if (ordered) {
      float key_lo, key_scale;
      const int2* lst = a.pairs + (size_t)off + sbase;
      // ---- 1. range of the depth keys
      float lo = 3.0e38f, hi = -3.0e38f;
      for (int i = tid; i < m; i += NT) {
        const float z = __int_as_float(__ldg(&lst[i].y));
        lo = fminf(lo, z); hi = fmaxf(hi, z);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
      if (lane == 0) { s_red[2 * warp] = lo; s_red[2 * warp + 1] = hi; }
      for (int i = tid; i < kBuckets; i += NT) { s_hist[i] = 0; s_bmin[i] = 0x7f800000u; }
      __syncthreads();
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) { lo = fminf(lo, s_red[2 * w]); hi = fmaxf(hi, s_red[2 * w + 1]); }
      key_lo = lo;
      key_scale = (hi > lo) ? (float)kBuckets / (hi - lo) : 0.0f;
      // ---- 2. histogram + exact minimum of every bucket (keys are positive: uint order == float order)
      for (int i = tid; i < m; i += NT) {
        const float z = __int_as_float(__ldg(&lst[i].y));
        const int b = min(kBuckets - 1, (int)((z - key_lo) * key_scale));
        atomicAdd(&s_hist[b], 1);
        atomicMin(&s_bmin[b], __float_as_uint(z));
      }
      __syncthreads();
      // ---- 3. bucket starts (exclusive scan) and suffix minima, by the first warp
      if (warp == 0) {
        constexpr int PER = kBuckets / 32;
        int c[PER], sum = 0;
        unsigned mn = 0x7f800000u;
#pragma unroll
        for (int q = 0; q < PER; ++q) { c[q] = s_hist[lane * PER + q]; sum += c[q]; mn = min(mn, s_bmin[lane * PER + q]); }
        int incl = sum;
        unsigned smn = mn;  // min over this lane's buckets and all later lanes'
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += u;
          const unsigned d = __shfl_down_sync(0xffffffffu, smn, o);
          if (lane + o < 32) smn = min(smn, d);
        }
        int run = incl - sum;
#pragma unroll
        for (int q = 0; q < PER; ++q) { s_hist[lane * PER + q] = run; run += c[q]; }
        unsigned after = __shfl_down_sync(0xffffffffu, smn, 1);  // min over all later lanes
        if (lane == 31) after = 0x7f800000u;
#pragma unroll
        for (int q = PER - 1; q >= 0; --q) {
          after = min(after, s_bmin[lane * PER + q]);
          s_bound[lane * PER + q] = __uint_as_float(after);
        }
      }
      __syncthreads();
      // ---- 4. scatter into bucket order
      for (int i = tid; i < m; i += NT) {
        const int2 e = __ldg(&lst[i]);
        const float z = __int_as_float(e.y);
        const int b = min(kBuckets - 1, (int)((z - key_lo) * key_scale));
        const int pos = atomicAdd(&s_hist[b], 1);
        ord_id[pos] = e.x;
        if ((pos & (NT - 1)) == 0) s_chunk_bucket[pos / NT] = (unsigned char)b;
      }
      __syncthreads();
    }

    KN_PH(1);
    bool stop = false;
    for (int base = 0; base < m; base += NT) {
      // ---- stage up to NT faces of the list
      const int j = base + tid;
      float4 bb = make_float4(3.0e38f, -3.0e38f, 3.0e38f, -3.0e38f);
      if (j < m) {
        const int lf = ordered ? ord_id[j] : (overflow ? sbase + j : __ldg(&a.pairs[(size_t)off + sbase + j].x));
        const FaceXYZ v = load_face(a.verts_ndc, a.faces, vd, lf);
        const bool ok = overflow ? face_is_drawable(v, cull, a.z_cull) : true;
        if (ok) {
          bb.x = fsub(min3f(v.x0, v.x1, v.x2), a.sqrt_blur); bb.y = fadd(max3f(v.x0, v.x1, v.x2), a.sqrt_blur);
          bb.z = fsub(min3f(v.y0, v.y1, v.y2), a.sqrt_blur); bb.w = fadd(max3f(v.y0, v.y1, v.y2), a.sqrt_blur);
        }
        const float area_e = fadd(edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1), kEps);
        s_va[tid] = make_float4(v.x0, v.y0, v.z0, v.x1);
        s_vb[tid] = make_float4(v.y1, v.z1, v.x2, v.y2);
        s_vc[tid] = make_float2(v.z2, area_e);
        s_id[tid] = lf;
        // Lower bound of the depth any pixel can get from this face (0 = no bound).  With barycentric
        // clipping the stored weights are a convex combination, so z >= min vertex z up to rounding;
        // with blur 0 every candidate is strictly inside and the same holds (times area_raw/(area_raw+kEps)
        // when the weights are not renormalised).  A full top-K list whose K-th depth is already below
        // this bound cannot change.
        const float zmin = min3f(v.z0, v.z1, v.z2);
        float zlo = 0.0f;
        if (clip) zlo = (!persp || zmin >= 1e-3f) ? zmin * 0.99999f : 0.0f;
        else if (hard_edges) {
          if (persp) zlo = zmin >= 1e-3f ? zmin * 0.99999f : 0.0f;
          else zlo = zmin * 0.99999f * (area_e > 0.0f ? fmaxf(0.0f, (area_e - 2e-8f) / area_e) : 1.0f);
        }
        s_zlo[tid] = zlo;
      }
      s_bb[tid] = bb;
      __syncthreads();
      const int mm = min(NT, m - base);
      if (tid == 0) KN_STAT(2, mm);
      KN_PH(2);
      if (live) {
        for (int q = slice; q < mm; q += SPLIT) {
          KN_STAT(3, 1);
          // cheapest test first: once the K-th layer has settled it rejects ~90% of a depth-ordered list
          if (s_zlo[q] > kth) continue;
          KN_STAT(4, 1);
          const float4 b = s_bb[q];
          if ((px > b.y) || (px < b.x) || (py > b.w) || (py < b.z)) continue;
          if (!hard_edges) {
            // the face lies inside its un-inflated box: a pixel farther than the blur radius from that box
            // (the rounded corners of the inflated one, 21% of it) fails the distance test by a wide margin
            const float dx = fmaxf(fmaxf(b.x + a.sqrt_blur - px, px - (b.y - a.sqrt_blur)), 0.0f);
            const float dy = fmaxf(fmaxf(b.z + a.sqrt_blur - py, py - (b.w - a.sqrt_blur)), 0.0f);
            if (dx * dx + dy * dy > blur * 1.001f) continue;
          }
          KN_STAT(5, 1);
          const float4 va = s_va[q], vb = s_vb[q];
          const float2 vc = s_vc[q];
          FaceXYZ v;
          v.x0 = va.x; v.y0 = va.y; v.z0 = va.z; v.x1 = va.w;
          v.y1 = vb.x; v.z1 = vb.y; v.x2 = vb.z; v.y2 = vb.w; v.z2 = vc.x;
          const float area = vc.y;
          const float e0 = edge_fn(px, py, v.x1, v.y1, v.x2, v.y2);
          const float e1 = edge_fn(px, py, v.x2, v.y2, v.x0, v.y0);
          const float e2 = edge_fn(px, py, v.x0, v.y0, v.x1, v.y1);
          if (hard_edges) {
            // blur 0: only strictly-inside samples survive, and w_i = e_i / area keeps the sign of
            // e_i * area exactly, so a non-positive edge is an exact (not approximate) reject.
            if (area > 0.0f ? (e0 <= 0.0f || e1 <= 0.0f || e2 <= 0.0f)
                            : (area < 0.0f && (e0 >= 0.0f || e1 >= 0.0f || e2 >= 0.0f)))
              continue;
          }
          const int f = s_id[q];
          float pz;
          bool inside;
          bool have_pz = false;
          // s_zlo > 0 <=> the convex-combination argument of the staging code holds for this face: the
          // perspective / clip renormalisations cannot hit their epsilon clamps, so the area and the
          // perspective denominator cancel out of the depth.
          if (s_zlo[q] > 0.0f && (clip || persp)) {
            // Clipped barycentrics are zero for every vertex whose edge function has the opposite sign of
            // the area, so the depth is a convex combination of the remaining vertices only -- one or two
            // of them for a sample outside its face, i.e. nearly every candidate inside a wide blur band.
            const bool up = area > 0.0f;
            const bool p0 = up ? e0 > 0.0f : e0 < 0.0f;
            const bool p1 = up ? e1 > 0.0f : e1 < 0.0f;
            const bool p2 = up ? e2 > 0.0f : e2 < 0.0f;
            const int npos = (int)p0 + (int)p1 + (int)p2;
            if (clip && npos == 1) {
              // One positive weight w_i >= w_0 + w_1 + w_2 ~ 1: the clip + renormalise step yields
              // exactly (1, 0, 0) -- b_i / max(b_i + 0 + 0, 1e-5) with b_i >= 1 -- and the depth is exactly
              // z_i (1 * z_i + 0 + 0): no division is needed to know it.
              pz = p0 ? v.z0 : (p1 ? v.z1 : v.z2);
              inside = false;
              have_pz = true;
              KN_STAT(6, 1);
            } else if (cnt == K && (clip || npos == 3)) {
              // Approximate depth first.  With a_i = e_i for the positive weights (0 for the clipped ones)
              //   z = z0 z1 z2 (a0+a1+a2) / (a0 z1 z2 + a1 z0 z2 + a2 z0 z1)   (perspective-correct)
              //   z = (a0 z0 + a1 z1 + a2 z2) / (a0 + a1 + a2)                 (otherwise)
              // is what the exact IEEE sequence (9 divisions) computes up to ~2e-6 relative: every term has
              // the same sign, nothing cancels.  A candidate that loses by more than 1e-4 relative is out.
              const float a0 = p0 ? e0 : 0.0f, a1 = p1 ? e1 : 0.0f, a2 = p2 ? e2 : 0.0f;
              float num, den;
              if (persp) {
                const float z12 = v.z1 * v.z2, z02 = v.z0 * v.z2, z01 = v.z0 * v.z1;
                num = v.z0 * z12 * (a0 + a1 + a2);
                den = a0 * z12 + a1 * z02 + a2 * z01;
              } else {
                num = a0 * v.z0 + a1 * v.z1 + a2 * v.z2;
                den = a0 + a1 + a2;
              }
              if (__fdividef(num, den) * 0.9999f > kth) continue;
              KN_STAT(7, 1);
            }
          }
          if (!have_pz) {
            KN_STAT(8, 1);
            float c0, c1, c2;
            if (!eval_from_edges(v, area, e0, e1, e2, persp, clip, pz, c0, c1, c2, inside)) continue;
          }
          // the depth is known before the (three more divisions of the) distance: losers leave here
          if (cnt == K && !cand_less(pz, f, kth, kf[(K - 1) * NT + tid] & kFaceMask)) continue;
          KN_STAT(9, 1);
          if (!inside) {
            if (hard_edges) continue;
            if (triangle_d2(v, px, py) >= blur) continue;
          }
          KN_STAT(10, 1);
          int pos = cnt < K ? cnt : K - 1;
          while (pos > 0 && cand_less(pz, f, kz[(pos - 1) * NT + tid], kf[(pos - 1) * NT + tid] & kFaceMask)) {
            kz[pos * NT + tid] = kz[(pos - 1) * NT + tid];
            kf[pos * NT + tid] = kf[(pos - 1) * NT + tid];
            --pos;
            KN_STAT(11, 1);
          }
          kz[pos * NT + tid] = pz; kf[pos * NT + tid] = inside ? (f | kInsideBit) : f;
          if (cnt < K) ++cnt;
          if (cnt == K) kth = kz[(K - 1) * NT + tid];
        }
      }
      if (ordered && base + NT < m) {
        // everything not yet staged has a depth key >= bound; stop when that is behind every K-th layer
        const float bound = s_bound[s_chunk_bucket[base / NT + 1]];
        const float zl = (persp && bound < 1e-3f) ? 0.0f : bound * 0.99999f;
        const bool done = !live || zl > kth;
        if (__syncthreads_and(done)) { stop = true; if (tid == 0) KN_STAT(12, 1); KN_PH(3); break; }
      } else {
        __syncthreads();
      }
      KN_PH(3);
    }
    (void)stop;
  }

#ifdef TRB_KN_STATS
  if (live && cnt < K) KN_STAT(13, 1);
  atomicMax(&g_kn_stats[15], (unsigned long long)(clock64() - kn_t0));   // longest list walk of any thread (cycles)
#pragma unroll
  for (int i = 0; i < 14; ++i) {
    unsigned v = kn_st[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v) atomicAdd(&g_kn_stats[i], (unsigned long long)v);
  }
#endif
  // ---- SPLIT == 2: the two threads of a pixel hold two sorted lists A (slice 0) and B (slice 1); the pixel's
  // layers are their merge, produced on the fly by both threads (same sequence), each evaluating every second layer
  int cntA = cnt, cntB = 0, total = cnt;
  const int colA = (SPLIT == 2) ? p : tid, colB = p + NP;
  if (SPLIT == 2) {
    __shared__ int s_cnt2[NT];
    s_cnt2[tid] = cnt;
    __syncthreads();
    cntA = s_cnt2[colA]; cntB = s_cnt2[colB];
    total = min(K, cntA + cntB);
  }
  const bool hit = live && slice == 0 && total > 0;
  const size_t pix = ((size_t)n * H + yi) * W + xi;
  append_hit_pixels<NT>(a.hit_pixels, hit, (int)pix, a.hit_counts, total);
  // Sparse Fragments (the caller only wants the image): covered pixels write layers [0, total) and one -1
  // terminator layer when total < K; nothing else is written, and the layer loop ends at the tile's deepest pixel.
  // s_wr[p] = number of layers pixel p writes (K for every pixel of the image in the dense layout).
  __shared__ int s_wr[NP];
  const bool sparse = a.sparse != 0;
  const int n_write = (!live || slice != 0) ? 0 : (sparse ? (total > 0 ? min(total + 1, K) : 0) : K);
  if (slice == 0) s_wr[p] = n_write;
  const bool dist_only_pre = sparse && SHADER == TRB_SHADER_SOFT_SILHOUETTE;   // see dist_only below
  const int k_end = dist_only_pre ? 0 : (sparse ? block_max_sync<NT>(n_write) : K);   // s_wr: published by the barrier after parking

  ViewParams vp;
  const ShadeIn sin = {a.verts_world, a.normals, a.colors, a.faces, a.uv};
  if (SHADER == TRB_SHADER_SOFT_PHONG || SHADER == TRB_SHADER_HARD_PHONG) vp = load_view_params(a.view_params, n);
  const float eps = 1e-10f;
  // layers are sorted front to back, so the softmax's max z_inv belongs to layer 0
  float alpha = 1.0f, wsum = 0.0f, zmax = eps, zrange = 1.0f;
  F3 acc = {0, 0, 0};
  F3 hard_c = {a.bg0, a.bg1, a.bg2};
  if (SHADER == TRB_SHADER_SOFT_PHONG) {
    zrange = vp.zfar - vp.znear;
    if (total > 0) {
      float z_first = cntA > 0 ? kz[colA] : 3.0e38f;
      if (SPLIT == 2 && cntB > 0 && (cntA == 0 || cand_less(kz[colB], kf[colB] & kFaceMask, kz[colA], kf[colA] & kFaceMask)))
        z_first = kz[colB];
      zmax = fmaxf(eps, (vp.zfar - z_first) / zrange);
    }
  }
  const int rot = p >> ROT;
  // sparse Fragments of a soft-silhouette render: zbuf and barycentrics are never read again (sigmoid_alpha_blend and
  // its backward use the distances only), so the epilogue skips the nine IEEE divisions per layer that produce them
  const bool dist_only = sparse && SHADER == TRB_SHADER_SOFT_SILHOUETTE;
  int ia = 0, ib = 0;   // merge cursors
  if (dist_only) {
    // No parking, no barriers: the only things written are 12 bytes per sample (face + signed distance) of COVERED
    // pixels -- a few MB -- so every thread stores its own layers straight to global memory.  (The coalesced,
    // parked write-out below exists for the dense layout's 28*K bytes per pixel.)  The backward takes the number of
    // layers from the covered-pixel list, so no terminator layer is written either.
    if (live) {
      const size_t g0 = (((size_t)n * H + yi) * W + xi) * K;
      for (int k = 0; k < total; ++k) {
        bool take_a = true;
        if (SPLIT == 2)
          take_a = ib >= cntB || (ia < cntA && cand_less(kz[ia * NT + colA], kf[ia * NT + colA] & kFaceMask,
                                                         kz[ib * NT + colB], kf[ib * NT + colB] & kFaceMask));
        const int f_raw = take_a ? kf[ia * NT + colA] : kf[ib * NT + colB];
        if (take_a) ++ia; else ++ib;
        if (SPLIT == 2 && (k & 1) != slice) continue;
        const int f = f_raw & kFaceMask;
        const FaceXYZ v = load_face(a.verts_ndc, a.faces, vd, f);
        const float d2 = triangle_d2(v, px, py);
        const float d = (f_raw & kInsideBit) ? -d2 : d2;
        alpha *= 1.0f - sigmoidf(-d / a.sigma);
        st_cs(a.p2f + g0 + k, (long long)vd.p2f_base + f);
        st_cs(a.dists + g0 + k, d);
      }
    }
  } else
  for (int k0 = 0; k0 < k_end; k0 += KG) {
    const int kg = min(KG, k_end - k0);
    for (int kk = 0; kk < kg; ++kk) {
      const int k = k0 + kk;
      const bool mine = (SPLIT == 1) || ((k & 1) == slice);
      Sample s = {-1.0f, -1.0f, -1.0f, -1.0f, -1.0f};
      long long pf = -1;
      if (k < total) {
        bool take_a = true;
        if (SPLIT == 2)
          take_a = ib >= cntB || (ia < cntA && cand_less(kz[ia * NT + colA], kf[ia * NT + colA] & kFaceMask,
                                                         kz[ib * NT + colB], kf[ib * NT + colB] & kFaceMask));
        const int f_raw = take_a ? kf[ia * NT + colA] : kf[ib * NT + colB];
        const int f = f_raw & kFaceMask;
        if (take_a) ++ia; else ++ib;
        if (mine) {
          const FaceXYZ v = load_face(a.verts_ndc, a.faces, vd, f);
          if (dist_only) {
            // image-only soft silhouette: the blend and the backward read the signed distance and nothing else;
            // `inside` was decided when the candidate entered the list (same arithmetic as eval_pixel_face_rt)
            const float d2 = triangle_d2(v, px, py);
            s.d = (f_raw & kInsideBit) ? -d2 : d2;
          } else {
            eval_pixel_face_rt(v, px, py, persp, clip, blur, s);
          }
          pf = (long long)vd.p2f_base + f;
          if (SHADER == TRB_SHADER_SOFT_SILHOUETTE) {
            alpha *= 1.0f - sigmoidf(-s.d / a.sigma);
          } else if (SHADER == TRB_SHADER_HARD_PHONG) {
            if (k == 0) hard_c = shade_sample<LIGHT>(sin, vd, vp, f, s.c0, s.c1, s.c2);
          } else if (SHADER == TRB_SHADER_SOFT_PHONG) {
            const float prob = sigmoidf(-s.d / a.sigma);
            alpha *= 1.0f - prob;
            const float zinv = (vp.zfar - s.z) / zrange;
            const float w = prob * expf((zinv - zmax) / a.gamma);
            const F3 c = shade_sample<LIGHT>(sin, vd, vp, f, s.c0, s.c1, s.c2);
            wsum += w;
            acc.x += w * c.x; acc.y += w * c.y; acc.z += w * c.z;
          }
        }
      }
      if (mine) {
        const int slot = p * KG + ((kk + rot) & (KG - 1));
        o_p2f[slot] = pf; o_z[slot] = s.z; o_d[slot] = s.d;
        o_b[3 * slot] = s.c0; o_b[3 * slot + 1] = s.c1; o_b[3 * slot + 2] = s.c2;
      }
    }
    __syncthreads();
    // ---- the CTA writes the parked layers out as contiguous runs
    for (int idx = tid; idx < NP * KG; idx += NT) {
      const int pp = idx >> LOGKG;
      const int kk = ((idx & (KG - 1)) - (pp >> ROT)) & (KG - 1);
      const int gx = x0 + (pp & (TX - 1)), gy = y0 + (pp >> LT);
      if (kk < kg && k0 + kk < s_wr[pp]) {
        const size_t g = ((size_t)(n * H + gy) * W + gx) * K + k0 + kk;
        st_cs(a.p2f + g, o_p2f[idx]);
        if (!dist_only) st_cs(a.zbuf + g, o_z[idx]);
        st_cs(a.dists + g, o_d[idx]);
      }
    }
    for (int idx = tid; idx < (dist_only ? 0 : NP * KG * 3); idx += NT) {
      const int e = idx / 3, c = idx - 3 * e;
      const int pp = e >> LOGKG;
      const int kk = ((e & (KG - 1)) - (pp >> ROT)) & (KG - 1);
      const int gx = x0 + (pp & (TX - 1)), gy = y0 + (pp >> LT);
      if (kk < kg && k0 + kk < s_wr[pp]) {
        const size_t g = ((size_t)(n * H + gy) * W + gx) * K + k0 + kk;
        st_cs(a.bary + 3 * g + c, o_b[idx]);
      }
    }
    __syncthreads();
  }
  KN_PH(4);
  if (SHADER == TRB_SHADER_NONE) return;
  if (SPLIT == 2 && SHADER != TRB_SHADER_HARD_PHONG) {
    // slice 1 hands its share of the blend sums to slice 0 (the parking area is free again)
    float* s_part = reinterpret_cast<float*>(s_un);   // [NP][5]
    if (slice == 1) {
      s_part[5 * p] = alpha; s_part[5 * p + 1] = wsum;
      s_part[5 * p + 2] = acc.x; s_part[5 * p + 3] = acc.y; s_part[5 * p + 4] = acc.z;
    }
    __syncthreads();
    if (slice == 0) {
      alpha *= s_part[5 * p]; wsum += s_part[5 * p + 1];
      acc.x += s_part[5 * p + 2]; acc.y += s_part[5 * p + 3]; acc.z += s_part[5 * p + 4];
    }
  }
  const int cnt_total = total;
  const bool writer = live && slice == 0;
  float4 out = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (writer) {
    if (SHADER == TRB_SHADER_SOFT_SILHOUETTE) {
      out = make_float4(1.0f, 1.0f, 1.0f, 1.0f - alpha);
    } else if (SHADER == TRB_SHADER_HARD_PHONG) {
      out = make_float4(hard_c.x, hard_c.y, hard_c.z, cnt_total > 0 ? 1.0f : 0.0f);
    } else {
      const float delta = fmaxf(expf((eps - zmax) / a.gamma), eps);
      const float inv = 1.0f / (wsum + delta);
      out = make_float4((acc.x + delta * a.bg0) * inv, (acc.y + delta * a.bg1) * inv,
                        (acc.z + delta * a.bg2) * inv, 1.0f - alpha);
    }
    st_cs(reinterpret_cast<float4*>(a.images) + pix, out);
  }
  if (a.alpha_sum != nullptr) {
    // per-view sum of the alpha channel (every thread of the CTA is here)
    float asum = writer ? out.w : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
    if (lane == 0 && asum != 0.0f) atomicAdd(a.alpha_sum + n, asum);
  }
}

template <int LT>
static int launch_kn(int shader, int light, dim3 grid, cudaStream_t st, const FineArgs& a) {
  using C = KnCfg<LT>;
  const size_t dyn = (size_t)a.K * C::NTH * 8 + C::union_bytes(a.K);
#define TRB_RKN(SH, L)                                                                              \
  do {                                                                                              \
    auto kern = render_fine_kn_kernel<LT, SH, L>;                                                   \
    TRB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
    TRB_CUDA_TRY(launch_pdl(kern, grid, dim3(C::NTH), dyn, st, a));                                 \
  } while (0)
  if (shader == TRB_SHADER_NONE) TRB_RKN(TRB_SHADER_NONE, 0);
  else if (shader == TRB_SHADER_SOFT_SILHOUETTE) TRB_RKN(TRB_SHADER_SOFT_SILHOUETTE, 0);
  else if (shader == TRB_SHADER_SOFT_PHONG) {
    if (light == 0) TRB_RKN(TRB_SHADER_SOFT_PHONG, 0); else if (light == 1) TRB_RKN(TRB_SHADER_SOFT_PHONG, 1);
    else TRB_RKN(TRB_SHADER_SOFT_PHONG, 2);
  } else {
    if (light == 0) TRB_RKN(TRB_SHADER_HARD_PHONG, 0); else if (light == 1) TRB_RKN(TRB_SHADER_HARD_PHONG, 1);
    else TRB_RKN(TRB_SHADER_HARD_PHONG, 2);
  }
#undef TRB_RKN
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

int launch_render_fine_kn(int shader, int light, int N, cudaStream_t st, const FineArgs& a) {
  const dim3 grid(a.tg.tiles_x, a.tg.tiles_y, N);
  if (a.tg.ltx == 4) return launch_kn<4>(shader, light, grid, st, a);
  return launch_kn<3>(shader, light, grid, st, a);
}

}  // namespace trb

#ifdef TRB_KN_STATS
// Copies the 16 walk counters to `host_out` and clears them (diagnostic builds only; synchronises).
extern "C" int trb_debug_kn_phases(unsigned long long* host_out) {
  unsigned long long zero[8] = {0};
  if (cudaMemcpyFromSymbol(host_out, trb::g_kn_phase, sizeof(zero)) != cudaSuccess) return TRB_ERR_CUDA;
  if (cudaMemcpyToSymbol(trb::g_kn_phase, zero, sizeof(zero)) != cudaSuccess) return TRB_ERR_CUDA;
  return TRB_OK;
}
extern "C" int trb_debug_kn_stats(unsigned long long* host_out) {
  unsigned long long zero[16] = {0};
  if (cudaMemcpyFromSymbol(host_out, trb::g_kn_stats, sizeof(zero)) != cudaSuccess) return TRB_ERR_CUDA;
  if (cudaMemcpyToSymbol(trb::g_kn_stats, zero, sizeof(zero)) != cudaSuccess) return TRB_ERR_CUDA;
  return TRB_OK;
}
#endif
