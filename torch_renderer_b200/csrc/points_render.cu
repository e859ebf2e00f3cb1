// Point-cloud rendering for sm_100a: per-pixel top-K point rasteriser (forward, backward) and the two compositors
// PointsRenderer uses (alpha compositing, normalised weighted sum; forward, backward).  Replaces
// pytorch3d._C.rasterize_points(_backward), _C.accum_alphacomposite(_backward), _C.accum_weightedsumnorm(_backward)
// behind the reference's AlphaPointRender / NormPointRender (torch_renderer.py:163-208) -- SURVEY 8f rank 4, last item.
// Semantics: the point-rasteriser section of oracle/trb_oracle.c and oracle/points_render_ref.py.
//
// Rasteriser: one CTA per 16x16 pixel tile.  The view's points stream through the CTA 256 at a time; a point whose
// disc can reach the tile is compacted into shared memory (x, y, z, r^2, index), then every thread tests its pixel
// against the compacted points and keeps a (z, index)-sorted top-K in local memory.  No workspace, no bin capacity,
// no dependence on point order: ties are ordered by (z, point index) like the oracle.  Coverage decisions use the
// round-to-nearest intrinsics (one IEEE operation per operator), so idx is bit-exact against the CPU oracle.
//
// Channels-last layouts throughout: idx i32 / zbuf / dists / alphas [N,H,W,K], features [P,C], images [N,H,W,C].
#include "raster_internal.cuh"
#include "raster_math.cuh"
#include "trb_internal.cuh"

namespace trb {

constexpr int kPtsTile = 16;
constexpr int kPtsThreads = kPtsTile * kPtsTile;

__global__ void __launch_bounds__(kPtsThreads)
points_raster_kernel(const float* __restrict__ points, const float* __restrict__ radius,
                     const trb_view* __restrict__ views, int H, int W, int K, int tiles_x, int tiles_y,
                     int* __restrict__ idx, float* __restrict__ zbuf, float* __restrict__ dists) {
  __shared__ float s_x[kPtsThreads], s_y[kPtsThreads], s_z[kPtsThreads], s_r2[kPtsThreads];
  __shared__ int s_id[kPtsThreads];
  __shared__ int s_n;
  const int tid = threadIdx.x;
  const int n = blockIdx.y;
  const int tile = blockIdx.x;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const trb_view vd = views[n];
  const int xi = tx * kPtsTile + (tid & (kPtsTile - 1));
  const int yi = ty * kPtsTile + (tid >> 4);
  const bool live = xi < W && yi < H;
  const float xf = pix_to_ndc(W - 1 - min(xi, W - 1), W, H);
  const float yf = pix_to_ndc(H - 1 - min(yi, H - 1), H, W);
  // NDC extent of the tile's pixel centres (x and y decrease with the pixel index)
  const int x_last = min(tx * kPtsTile + kPtsTile - 1, W - 1), y_last = min(ty * kPtsTile + kPtsTile - 1, H - 1);
  const float tx_hi = pix_to_ndc(W - 1 - tx * kPtsTile, W, H), tx_lo = pix_to_ndc(W - 1 - x_last, W, H);
  const float ty_hi = pix_to_ndc(H - 1 - ty * kPtsTile, H, W), ty_lo = pix_to_ndc(H - 1 - y_last, H, W);

  float qz[TRB_MAX_FACES_PER_PIXEL];
  float qd[TRB_MAX_FACES_PER_PIXEL];
  int qi[TRB_MAX_FACES_PER_PIXEL];
  int qn = 0;

  for (int base = 0; base < vd.face_count; base += kPtsThreads) {
    if (tid == 0) s_n = 0;
    __syncthreads();
    const int lp = base + tid;
    if (lp < vd.face_count) {
      const size_t row = (size_t)(vd.face_start + lp);
      const float px = __ldg(points + 3 * row), py = __ldg(points + 3 * row + 1), pz = __ldg(points + 3 * row + 2);
      const float r = __ldg(radius + row);
      // conservative reach test (a hair of slack; the exact test follows per pixel); NaNs fail it
      const float slack = r * 1.0001f + 1e-6f;
      if (pz >= 0.0f && px >= tx_lo - slack && px <= tx_hi + slack && py >= ty_lo - slack && py <= ty_hi + slack) {
        const int pos = atomicAdd(&s_n, 1);
        s_x[pos] = px; s_y[pos] = py; s_z[pos] = pz; s_r2[pos] = fmul(r, r);
        s_id[pos] = vd.p2f_base + lp;
      }
    }
    __syncthreads();
    const int m = s_n;
    if (live) {
      for (int j = 0; j < m; ++j) {
        const float dx = fsub(xf, s_x[j]), dy = fsub(yf, s_y[j]);
        const float d2 = fadd(fmul(dx, dx), fmul(dy, dy));
        if (!(d2 < s_r2[j])) continue;
        const float z = s_z[j];
        const int id = s_id[j];
        if (qn == K && !cand_less(z, id, qz[K - 1], qi[K - 1])) continue;
        int pos = qn < K ? qn : K - 1;
        while (pos > 0 && cand_less(z, id, qz[pos - 1], qi[pos - 1])) {
          qz[pos] = qz[pos - 1]; qd[pos] = qd[pos - 1]; qi[pos] = qi[pos - 1];
          --pos;
        }
        qz[pos] = z; qd[pos] = d2; qi[pos] = id;
        if (qn < K) ++qn;
      }
    }
    __syncthreads();
  }
  if (!live) return;
  const size_t o = (((size_t)n * H + yi) * W + xi) * K;
  for (int k = 0; k < K; ++k) {
    const bool hit = k < qn;
    idx[o + k] = hit ? qi[k] : -1;
    zbuf[o + k] = hit ? qz[k] : -1.0f;
    dists[o + k] = hit ? qd[k] : -1.0f;
  }
}

// ---- binned rasteriser: count -> allocate -> fill per-tile point lists, then one CTA per tile walks ITS list ------
// Points are their own bounding discs.  The kernel above streams the whole cloud through every tile (4 x 10^8 point
// tests for 100 k points x 4 views at 512^2, of which ~10^6 matter); with per-tile lists a tile sees only the points
// whose disc can reach it.  Lists live in caller-provided workspace ([counts | fill cursors | offsets] per tile, one
// global cursor, the entries); a tile whose list does not fit the entry capacity falls back to the whole-cloud scan
// (points are never dropped).  List order is whatever the atomics gave: the per-pixel top-K is ordered by
// (z, point index), so the result does not depend on it.
struct PtsWs {
  int* count; int* fill; int* offset; int* cursor; int* entries; long long capacity;
};

__device__ __forceinline__ bool point_tile_range(float px, float py, float pz, float r, int H, int W, int tiles_x,
                                                 int tiles_y, int& tx0, int& tx1, int& ty0, int& ty1) {
  if (!(pz >= 0.0f) || !(r >= 0.0f)) return false;   // behind the camera / NaN: never drawn
  const float slack = r * 1.0001f + 1e-6f;
  int c0, c1, r0, r1;
  pixel_range(px - slack, px + slack, W, H, c0, c1);
  pixel_range(py - slack, py + slack, H, W, r0, r1);
  if (c1 < c0 || r1 < r0) return false;
  tx0 = c0 / kPtsTile; tx1 = min(c1 / kPtsTile, tiles_x - 1);
  ty0 = r0 / kPtsTile; ty1 = min(r1 / kPtsTile, tiles_y - 1);
  return true;
}

template <bool FILL>
__global__ void __launch_bounds__(256)
points_bin_kernel(const float* __restrict__ points, const float* __restrict__ radius,
                  const trb_view* __restrict__ views, int H, int W, int tiles_x, int tiles_y, int max_points,
                  PtsWs ws) {
  const int n = blockIdx.y;
  const trb_view vd = views[n];
  const int lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= vd.face_count || lp >= max_points) return;
  const size_t row = (size_t)(vd.face_start + lp);
  int tx0, tx1, ty0, ty1;
  if (!point_tile_range(__ldg(points + 3 * row), __ldg(points + 3 * row + 1), __ldg(points + 3 * row + 2),
                        __ldg(radius + row), H, W, tiles_x, tiles_y, tx0, tx1, ty0, ty1))
    return;
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx) {
      const int t = (n * tiles_y + ty) * tiles_x + tx;
      if (!FILL) {
        atomicAdd(&ws.count[t], 1);
      } else {
        const int off = ws.offset[t];
        if (off >= 0) ws.entries[(size_t)off + atomicAdd(&ws.fill[t], 1)] = lp;
      }
    }
}

__global__ void __launch_bounds__(256) points_alloc_kernel(int ntiles, PtsWs ws) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = t < ntiles ? ws.count[t] : 0;
  // warp-aggregated slice reservation
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += u;
  }
  const int wtot = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if ((threadIdx.x & 31) == 31 && wtot > 0) base = atomicAdd(ws.cursor, wtot);
  base = __shfl_sync(0xffffffffu, base, 31);
  if (t < ntiles) {
    const long long start = (long long)base + incl - c;
    ws.offset[t] = (c > 0 && start + c <= ws.capacity) ? (int)start : -1;
  }
}

// One CTA per tile; per-pixel (z, index)-sorted top-K in SHARED memory ([K][256] columns, conflict free) when it
// fits, in local memory otherwise (K up to 150).
template <bool SMEM_Q>
__global__ void __launch_bounds__(kPtsThreads)
points_raster_binned_kernel(const float* __restrict__ points, const float* __restrict__ radius,
                            const trb_view* __restrict__ views, int H, int W, int K, int tiles_x, int tiles_y,
                            PtsWs ws, int* __restrict__ idx, float* __restrict__ zbuf, float* __restrict__ dists) {
  __shared__ float s_x[kPtsThreads], s_y[kPtsThreads], s_z[kPtsThreads], s_r2[kPtsThreads];
  __shared__ int s_id[kPtsThreads];
  extern __shared__ float s_q[];   // SMEM_Q: z [K][256], d [K][256], id [K][256]
  const int tid = threadIdx.x;
  const int n = blockIdx.y;
  const int tile = blockIdx.x;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const trb_view vd = views[n];
  const int xi = tx * kPtsTile + (tid & (kPtsTile - 1));
  const int yi = ty * kPtsTile + (tid >> 4);
  const bool live = xi < W && yi < H;
  const float xf = pix_to_ndc(W - 1 - min(xi, W - 1), W, H);
  const float yf = pix_to_ndc(H - 1 - min(yi, H - 1), H, W);
  const int t = (n * tiles_y + ty) * tiles_x + tx;
  const int off = ws.offset[t];
  const bool overflow = ws.count[t] > 0 && off < 0;
  const int nlist = overflow ? vd.face_count : ws.count[t];

  float lz[SMEM_Q ? 1 : TRB_MAX_FACES_PER_PIXEL];
  float ld[SMEM_Q ? 1 : TRB_MAX_FACES_PER_PIXEL];
  int li[SMEM_Q ? 1 : TRB_MAX_FACES_PER_PIXEL];
  float* qz = SMEM_Q ? s_q + tid : lz;
  float* qd = SMEM_Q ? s_q + (size_t)K * kPtsThreads + tid : ld;
  int* qi = SMEM_Q ? reinterpret_cast<int*>(s_q + 2 * (size_t)K * kPtsThreads) + tid : li;
  constexpr int QS = SMEM_Q ? kPtsThreads : 1;   // stride between a pixel's consecutive layers
  int qn = 0;
  float kth = 3.0e38f;

  // NDC y extent of the two pixel rows this warp owns (y decreases with the row index)
  const int lane = tid & 31, warp = tid >> 5;
  const int row0 = ty * kPtsTile + 2 * warp;
  const float wy_hi = pix_to_ndc(H - 1 - min(row0, H - 1), H, W), wy_lo = pix_to_ndc(H - 1 - min(row0 + 1, H - 1), H, W);
  __shared__ float s_rad[kPtsThreads];
  __shared__ unsigned char s_wlist[kPtsThreads / 32][kPtsThreads];   // per warp: staged points that reach its rows

  for (int base = 0; base < nlist; base += kPtsThreads) {
    __syncthreads();   // the previous chunk's readers are done
    const int j = base + tid;
    float r2 = -1.0f, rad = -1.0f;   // empty slot: fails every d2 < r2 test
    if (j < nlist) {
      const int lp = overflow ? j : ws.entries[(size_t)off + j];
      const size_t row = (size_t)(vd.face_start + lp);
      const float pz = __ldg(points + 3 * row + 2);
      const float r = __ldg(radius + row);
      s_x[tid] = __ldg(points + 3 * row); s_y[tid] = __ldg(points + 3 * row + 1); s_z[tid] = pz;
      s_id[tid] = vd.p2f_base + lp;
      if (pz >= 0.0f && r >= 0.0f) { r2 = fmul(r, r); rad = r * 1.0001f + 1e-6f; }
    }
    s_r2[tid] = r2; s_rad[tid] = rad;
    __syncthreads();
    const int m = min(kPtsThreads, nlist - base);
    // Sub-tile culling: a disc of the reference's radius (0.003 NDC = 0.8 px at 512^2) reaches one or two of a tile's
    // sixteen rows, so each warp first compacts the staged points that can reach ITS two rows (one ballot per 32
    // points) and its pixels test only those -- 280 staged points per tile became ~50 per warp.
    int wn = 0;
    for (int q0 = 0; q0 < m; q0 += 32) {
      const int q = q0 + lane;
      bool ov = false;
      if (q < m) {
        const float rq = s_rad[q], yq = s_y[q];
        ov = rq >= 0.0f && yq + rq >= wy_lo && yq - rq <= wy_hi;
      }
      const unsigned b = __ballot_sync(0xffffffffu, ov);
      if (ov) s_wlist[warp][wn + __popc(b & ((1u << lane) - 1u))] = (unsigned char)q;
      wn += __popc(b);
    }
    __syncwarp();
    if (live) {
      for (int jj = 0; jj < wn; ++jj) {
        const int q = s_wlist[warp][jj];
        const float dx = fsub(xf, s_x[q]), dy = fsub(yf, s_y[q]);
        const float d2 = fadd(fmul(dx, dx), fmul(dy, dy));
        if (!(d2 < s_r2[q])) continue;
        const float z = s_z[q];
        if (z > kth) continue;
        const int id = s_id[q];
        if (qn == K && !cand_less(z, id, qz[(K - 1) * QS], qi[(K - 1) * QS])) continue;
        int pos = qn < K ? qn : K - 1;
        while (pos > 0 && cand_less(z, id, qz[(pos - 1) * QS], qi[(pos - 1) * QS])) {
          qz[pos * QS] = qz[(pos - 1) * QS]; qd[pos * QS] = qd[(pos - 1) * QS]; qi[pos * QS] = qi[(pos - 1) * QS];
          --pos;
        }
        qz[pos * QS] = z; qd[pos * QS] = d2; qi[pos * QS] = id;
        if (qn < K) ++qn;
        if (qn == K) kth = qz[(K - 1) * QS];
      }
    }
  }
  if (SMEM_Q) {
    // Coalesced write-out: layer k of pixel p lives at (p*K + k), so a thread-per-pixel store touches one word in
    // each of 32 sectors; here the CTA walks the tile's rows as contiguous runs of 16*K words instead.
    __shared__ int s_qn[kPtsThreads];
    s_qn[tid] = qn;
    __syncthreads();
    const int rows = min(kPtsTile, H - ty * kPtsTile), cols = min(kPtsTile, W - tx * kPtsTile);
    const int run = cols * K;
    const float* gz = s_q;
    const float* gd = s_q + (size_t)K * kPtsThreads;
    const int* gi = reinterpret_cast<const int*>(s_q + 2 * (size_t)K * kPtsThreads);
    for (int r = 0; r < rows; ++r) {
      const size_t o = (((size_t)n * H + ty * kPtsTile + r) * W + tx * kPtsTile) * K;
      for (int i = tid; i < run; i += kPtsThreads) {
        const int c = i / K, k = i - c * K;
        const int pp = r * kPtsTile + c;
        const bool hit = k < s_qn[pp];
        idx[o + i] = hit ? gi[k * kPtsThreads + pp] : -1;
        zbuf[o + i] = hit ? gz[k * kPtsThreads + pp] : -1.0f;
        dists[o + i] = hit ? gd[k * kPtsThreads + pp] : -1.0f;
      }
    }
    return;
  }
  if (!live) return;
  const size_t o = (((size_t)n * H + yi) * W + xi) * K;
  for (int k = 0; k < K; ++k) {
    const bool hit = k < qn;
    idx[o + k] = hit ? qi[k * QS] : -1;
    zbuf[o + k] = hit ? qz[k * QS] : -1.0f;
    dists[o + k] = hit ? qd[k * QS] : -1.0f;
  }
}

// One thread per sample: d dist^2 / d (px, py) = 2 (p - pixel), d zbuf / d pz = 1.
__global__ void __launch_bounds__(256)
points_raster_backward_kernel(const float* __restrict__ points, const int* __restrict__ idx,
                              const float* __restrict__ grad_zbuf, const float* __restrict__ grad_dists, long long total,
                              int H, int W, int K, float* __restrict__ grad_points) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  const int p = idx[s];
  if (p < 0) return;
  const long long pix = s / K;
  const int xi = (int)(pix % W), yi = (int)((pix / W) % H);
  const float xf = pix_to_ndc(W - 1 - xi, W, H), yf = pix_to_ndc(H - 1 - yi, H, W);
  const float gd = grad_dists ? grad_dists[s] : 0.0f, gz = grad_zbuf ? grad_zbuf[s] : 0.0f;
  const float px = __ldg(points + 3 * (size_t)p), py = __ldg(points + 3 * (size_t)p + 1);
  if (gd != 0.0f) {
    atomicAdd(grad_points + 3 * (size_t)p, 2.0f * gd * (px - xf));
    atomicAdd(grad_points + 3 * (size_t)p + 1, 2.0f * gd * (py - yf));
  }
  if (gz != 0.0f) atomicAdd(grad_points + 3 * (size_t)p + 2, gz);
}

// ---- compositors --------------------------------------------------------------------------------------------
// mode 0: alpha compositing   out_c = sum_k f[idx_k, c] a_k prod_{j<k} (1 - a_j)
// mode 1: normalised weights  out_c = sum_k f[idx_k, c] a_k / max(sum_k a_k, 1e-4)
// Empty slots (idx < 0) are skipped; a pixel whose first slot is empty takes `background` when given.
constexpr float kNormEps = 1e-4f;

__global__ void __launch_bounds__(256)
points_composite_kernel(int mode, const int* __restrict__ idx, const float* __restrict__ alphas,
                        const float* __restrict__ features, long long npix, int K, int C,
                        const float* __restrict__ background, float* __restrict__ images) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npix * C) return;
  const long long pix = t / C;
  const int c = (int)(t - pix * C);
  const int* pi = idx + pix * K;
  const float* pa = alphas + pix * K;
  if (background != nullptr && pi[0] < 0) { images[t] = background[c]; return; }
  float out = 0.0f;
  if (mode == 0) {
    float cum = 1.0f;
    for (int k = 0; k < K; ++k) {
      const int p = pi[k];
      if (p < 0) continue;
      const float a = pa[k];
      out += __ldg(features + (size_t)p * C + c) * cum * a;
      cum *= 1.0f - a;
    }
  } else {
    float total = 0.0f;
    for (int k = 0; k < K; ++k)
      if (pi[k] >= 0) total += pa[k];
    total = fmaxf(total, kNormEps);
    for (int k = 0; k < K; ++k) {
      const int p = pi[k];
      if (p < 0) continue;
      out += __ldg(features + (size_t)p * C + c) * pa[k] / total;
    }
  }
  images[t] = out;
}

// One thread per pixel (all channels): grad_alphas is written (0 for empty slots), grad_features accumulated.
// Alpha mode, division-free: with cum_k = prod_{j<k}(1 - a_j) and R_k = sum_{m>k} f_m a_m prod_{k<j<m}(1 - a_j)
// (R_{k-1} = f_k a_k + (1 - a_k) R_k),  d out / d a_k = cum_k (f_k - R_k).
__global__ void __launch_bounds__(128)
points_composite_backward_kernel(int mode, const int* __restrict__ idx, const float* __restrict__ alphas,
                                 const float* __restrict__ features, const float* __restrict__ grad_images,
                                 long long npix, int K, int C, int has_background, float* __restrict__ grad_alphas,
                                 float* __restrict__ grad_features) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int* pi = idx + pix * K;
  const float* pa = alphas + pix * K;
  const float* g = grad_images + pix * C;
  float ga[TRB_MAX_FACES_PER_PIXEL];
  float cumv[TRB_MAX_FACES_PER_PIXEL];
  for (int k = 0; k < K; ++k) ga[k] = 0.0f;
  const bool masked = has_background && pi[0] < 0;   // the pixel shows the background: no gradient
  if (!masked) {
    if (mode == 0) {
      float cum = 1.0f;
      for (int k = 0; k < K; ++k) {
        cumv[k] = cum;
        if (pi[k] >= 0) cum *= 1.0f - pa[k];
      }
      for (int c = 0; c < C; ++c) {
        const float gc = g[c];
        if (gc == 0.0f) continue;
        float R = 0.0f;
        for (int k = K - 1; k >= 0; --k) {
          const int p = pi[k];
          if (p < 0) continue;
          const float a = pa[k];
          const float f = __ldg(features + (size_t)p * C + c);
          ga[k] += gc * cumv[k] * (f - R);
          if (grad_features) atomicAdd(grad_features + (size_t)p * C + c, gc * cumv[k] * a);
          R = f * a + (1.0f - a) * R;
        }
      }
    } else {
      float total = 0.0f;
      for (int k = 0; k < K; ++k)
        if (pi[k] >= 0) total += pa[k];
      const bool clamped = total < kNormEps;
      const float t = clamped ? kNormEps : total;
      for (int c = 0; c < C; ++c) {
        const float gc = g[c];
        if (gc == 0.0f) continue;
        float out = 0.0f;
        for (int k = 0; k < K; ++k)
          if (pi[k] >= 0) out += __ldg(features + (size_t)pi[k] * C + c) * pa[k];
        out /= t;
        for (int k = 0; k < K; ++k) {
          const int p = pi[k];
          if (p < 0) continue;
          const float f = __ldg(features + (size_t)p * C + c);
          ga[k] += gc * (f - (clamped ? 0.0f : out)) / t;
          if (grad_features) atomicAdd(grad_features + (size_t)p * C + c, gc * pa[k] / t);
        }
      }
    }
  }
  if (grad_alphas)
    for (int k = 0; k < K; ++k) grad_alphas[pix * K + k] = ga[k];
}

}  // namespace trb

using namespace trb;

extern "C" int trb_points_raster_forward(const float* points_ndc, const float* radius, const trb_view* views, int N,
                                         int H, int W, int K, int32_t* idx, float* zbuf, float* dists, int device,
                                         trb_stream_t stream) {
  if (N < 0 || H < 1 || W < 1 || K < 1) return TRB_ERR_BAD_ARG;
  if (K > TRB_MAX_FACES_PER_PIXEL) return TRB_ERR_K_TOO_LARGE;
  if (N == 0) return TRB_OK;
  if (N > 65535 || !views || !idx || !zbuf || !dists) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  const int tiles_x = ceil_div(W, kPtsTile), tiles_y = ceil_div(H, kPtsTile);
  points_raster_kernel<<<dim3(tiles_x * tiles_y, N), kPtsThreads, 0, (cudaStream_t)stream>>>(
      points_ndc, radius, views, H, W, K, tiles_x, tiles_y, idx, zbuf, dists);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

/* workspace of trb_points_raster_forward_binned: tile counters, fill cursors, offsets, one global cursor, entries */
static PtsWs points_ws(void* workspace, int N, int H, int W, int64_t entry_capacity) {
  const size_t ntiles = (size_t)N * ceil_div(W, kPtsTile) * ceil_div(H, kPtsTile);
  PtsWs ws;
  int* base = (int*)workspace;
  ws.cursor = base; ws.count = base + 4; ws.fill = ws.count + ntiles; ws.offset = ws.fill + ntiles;
  ws.entries = ws.offset + ntiles; ws.capacity = entry_capacity;
  return ws;
}

extern "C" int trb_points_raster_workspace_bytes(int N, int H, int W, int64_t entry_capacity, size_t* bytes) {
  if (N < 0 || H < 1 || W < 1 || entry_capacity < 0 || !bytes) return TRB_ERR_BAD_ARG;
  const size_t ntiles = (size_t)N * ceil_div(W, kPtsTile) * ceil_div(H, kPtsTile);
  *bytes = (4 + 3 * ntiles + (size_t)entry_capacity) * sizeof(int);
  return TRB_OK;
}

extern "C" int trb_points_raster_forward_binned(const float* points_ndc, const float* radius, const trb_view* views,
                                                int N, int max_points, int H, int W, int K, int64_t entry_capacity,
                                                void* workspace, size_t workspace_bytes, int32_t* idx, float* zbuf,
                                                float* dists, int device, trb_stream_t stream) {
  if (N < 0 || H < 1 || W < 1 || K < 1 || max_points < 0 || entry_capacity < 0) return TRB_ERR_BAD_ARG;
  if (K > TRB_MAX_FACES_PER_PIXEL) return TRB_ERR_K_TOO_LARGE;
  if (N == 0) return TRB_OK;
  if (N > 65535 || !views || !idx || !zbuf || !dists || !workspace) return TRB_ERR_BAD_ARG;
  if (entry_capacity > 0x7fffffffll) return TRB_ERR_BAD_ARG;
  size_t need = 0;
  trb_points_raster_workspace_bytes(N, H, W, entry_capacity, &need);
  if (workspace_bytes < need) return TRB_ERR_WORKSPACE;
  TRB_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_x = ceil_div(W, kPtsTile), tiles_y = ceil_div(H, kPtsTile);
  const int ntiles = N * tiles_x * tiles_y;
  const PtsWs ws = points_ws(workspace, N, H, W, entry_capacity);
  TRB_CUDA_TRY(cudaMemsetAsync(workspace, 0, (4 + 2 * (size_t)ntiles) * sizeof(int), st));   // cursor, counts, fills
  if (max_points > 0) {
    const dim3 gp(ceil_div(max_points, 256), N);
    points_bin_kernel<false><<<gp, 256, 0, st>>>(points_ndc, radius, views, H, W, tiles_x, tiles_y, max_points, ws);
    points_alloc_kernel<<<ceil_div(ntiles, 256), 256, 0, st>>>(ntiles, ws);
    points_bin_kernel<true><<<gp, 256, 0, st>>>(points_ndc, radius, views, H, W, tiles_x, tiles_y, max_points, ws);
  } else {
    points_alloc_kernel<<<ceil_div(ntiles, 256), 256, 0, st>>>(ntiles, ws);
  }
  const size_t dyn = (size_t)K * kPtsThreads * 12;
  const dim3 grid(tiles_x * tiles_y, N);
  if (dyn <= 96 * 1024) {
    auto kern = points_raster_binned_kernel<true>;
    TRB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<grid, kPtsThreads, dyn, st>>>(points_ndc, radius, views, H, W, K, tiles_x, tiles_y, ws, idx, zbuf, dists);
  } else {
    points_raster_binned_kernel<false><<<grid, kPtsThreads, 0, st>>>(points_ndc, radius, views, H, W, K, tiles_x,
                                                                      tiles_y, ws, idx, zbuf, dists);
  }
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

extern "C" int trb_points_raster_backward(const float* points_ndc, const int32_t* idx, const float* grad_zbuf,
                                          const float* grad_dists, int N, int H, int W, int K, float* grad_points,
                                          int device, trb_stream_t stream) {
  if (N < 0 || H < 1 || W < 1 || K < 1) return TRB_ERR_BAD_ARG;
  if (N == 0) return TRB_OK;
  if (!points_ndc || !idx || !grad_points) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  const long long total = (long long)N * H * W * K;
  points_raster_backward_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
      points_ndc, idx, grad_zbuf, grad_dists, total, H, W, K, grad_points);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

extern "C" int trb_points_composite_forward(int mode, const int32_t* idx, const float* alphas, const float* features,
                                            int64_t num_pixels, int K, int C, const float* background,
                                            float* images, int device, trb_stream_t stream) {
  if ((mode != 0 && mode != 1) || num_pixels < 0 || K < 1 || C < 1) return TRB_ERR_BAD_ARG;
  if (K > TRB_MAX_FACES_PER_PIXEL) return TRB_ERR_K_TOO_LARGE;
  if (num_pixels == 0) return TRB_OK;
  if (!idx || !alphas || !features || !images) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  points_composite_kernel<<<(unsigned)ceil_div64(num_pixels * C, 256), 256, 0, (cudaStream_t)stream>>>(
      mode, idx, alphas, features, (long long)num_pixels, K, C, background, images);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

extern "C" int trb_points_composite_backward(int mode, const int32_t* idx, const float* alphas, const float* features,
                                             const float* grad_images, int64_t num_pixels, int K, int C,
                                             int has_background, float* grad_alphas, float* grad_features, int device,
                                             trb_stream_t stream) {
  if ((mode != 0 && mode != 1) || num_pixels < 0 || K < 1 || C < 1) return TRB_ERR_BAD_ARG;
  if (K > TRB_MAX_FACES_PER_PIXEL) return TRB_ERR_K_TOO_LARGE;
  if (num_pixels == 0) return TRB_OK;
  if (!idx || !alphas || !features || !grad_images) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  points_composite_backward_kernel<<<(unsigned)ceil_div64(num_pixels, 128), 128, 0, (cudaStream_t)stream>>>(
      mode, idx, alphas, features, grad_images, (long long)num_pixels, K, C, has_background, grad_alphas,
      grad_features);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}
