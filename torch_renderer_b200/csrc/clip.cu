// Near-plane clipping support for the stand-alone rasteriser: the "one of two neighbours per pixel" rule of
// PyTorch3D's rasterize_meshes (clipped_faces_neighbor_idx), SURVEY.md 8f rank 3.
//
// clip_faces (torch_renderer_b200/clip.py) turns a face with one vertex behind the near plane into two triangles
// t1, t2 = t1 + 1.  Upstream's per-pixel queue keeps at most one of them: when t2 arrives and t1 is in the queue,
// t2 replaces t1 iff its unsigned distance is strictly smaller, else t2 is dropped; when t1 is not in the queue
// (never a candidate, or pushed out by K nearer faces) t2 is an ordinary face.  The outcome depends on the order
// faces arrive in (ascending index) and can differ from any order-free rule, so it is reproduced literally --
// but only where it can matter: a pixel is affected iff BOTH halves of some pair pass the candidate test there
// (otherwise the queue never finds a neighbour and upstream degenerates to the plain top-K that
// trb_raster_forward has already written).  Those pixels lie in a thin band around the shared diagonal (none at
// all, up to rounding, when blur_radius == 0).
//
// One CTA per pair: every thread tests one pixel of the pair's common (inflated) bounding box against both halves;
// the flagged pixels are compacted into shared memory and re-rasterised, one warp per pixel: 32 lanes test 32
// faces of the view at a time, in face order, and lane 0 replays upstream's queue in shared memory.  Two pairs
// flagging the same pixel write identical values.
#include "stages.cuh"
#include "trb_internal.cuh"

namespace trb {

constexpr int kClipThreads = 128;
constexpr int kClipWarps = kClipThreads / 32;

// A4 steps 2-9 for one (pixel, face): per-face validity, inflated bounding box, then the sample itself.
__device__ __forceinline__ bool clip_candidate(const FaceXYZ& v, float px, float py, float blur, float sqrt_blur,
                                               bool persp, bool clip, bool cull, Sample& s) {
  if (!face_is_drawable(v, cull, 0.0f)) return false;
  const float xmin = fsub(min3f(v.x0, v.x1, v.x2), sqrt_blur), xmax = fadd(max3f(v.x0, v.x1, v.x2), sqrt_blur);
  const float ymin = fsub(min3f(v.y0, v.y1, v.y2), sqrt_blur), ymax = fadd(max3f(v.y0, v.y1, v.y2), sqrt_blur);
  if ((px > xmax) || (px < xmin) || (py > ymax) || (py < ymin)) return false;
  return eval_pixel_face_rt(v, px, py, persp, clip, blur, s);
}

struct ClipQueue {
  float z[TRB_MAX_FACES_PER_PIXEL];
  float d[TRB_MAX_FACES_PER_PIXEL];
  int f[TRB_MAX_FACES_PER_PIXEL];
};

__global__ void __launch_bounds__(kClipThreads)
clip_resequence_kernel(const float* __restrict__ face_verts, const trb_view* __restrict__ views,
                       const int* __restrict__ pair_face, const int* __restrict__ pair_view,
                       const int* __restrict__ neighbor, int H, int W, int K, float blur, float sqrt_blur,
                       unsigned flags, long long* __restrict__ p2f, float* __restrict__ zbuf,
                       float* __restrict__ bary, float* __restrict__ dists, int* __restrict__ counters) {
  __shared__ ClipQueue s_q[kClipWarps];
  __shared__ int s_flagged[kClipThreads];
  __shared__ int s_nflag;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool persp = flags & TRB_PERSPECTIVE_CORRECT, clip = flags & TRB_CLIP_BARYCENTRIC;
  const bool cull = flags & TRB_CULL_BACKFACES;
  const int n = pair_view[blockIdx.x];
  const trb_view vd = views[n];
  const int lf1 = pair_face[blockIdx.x] - vd.p2f_base;   // local index of t1; t2 = t1 + 1
  const FaceXYZ a = load_face(face_verts, nullptr, vd, lf1);
  const FaceXYZ b = load_face(face_verts, nullptr, vd, lf1 + 1);
  // common part of the two inflated boxes, as a conservative pixel rectangle
  const float xlo = fmaxf(min3f(a.x0, a.x1, a.x2), min3f(b.x0, b.x1, b.x2)) - sqrt_blur;
  const float xhi = fminf(max3f(a.x0, a.x1, a.x2), max3f(b.x0, b.x1, b.x2)) + sqrt_blur;
  const float ylo = fmaxf(min3f(a.y0, a.y1, a.y2), min3f(b.y0, b.y1, b.y2)) - sqrt_blur;
  const float yhi = fminf(max3f(a.y0, a.y1, a.y2), max3f(b.y0, b.y1, b.y2)) + sqrt_blur;
  int c0 = 0, c1 = -1, r0 = 0, r1 = -1;
  if (xlo <= xhi && ylo <= yhi) {   // false for NaN as well
    pixel_range(xlo, xhi, W, H, c0, c1);
    pixel_range(ylo, yhi, H, W, r0, r1);
  }
  const int rw = c1 - c0 + 1, rh = r1 - r0 + 1;
  if (rw <= 0 || rh <= 0) return;
  const long long npix = (long long)rw * rh;
  ClipQueue& q = s_q[warp];
  for (long long base = 0; base < npix; base += kClipThreads) {
    if (tid == 0) s_nflag = 0;
    __syncthreads();
    const long long i = base + tid;
    if (i < npix) {
      const int yi = r0 + (int)(i / rw), xi = c0 + (int)(i % rw);
      const float px = pix_to_ndc(W - 1 - xi, W, H), py = pix_to_ndc(H - 1 - yi, H, W);
      Sample s;
      if (clip_candidate(a, px, py, blur, sqrt_blur, persp, clip, cull, s) &&
          clip_candidate(b, px, py, blur, sqrt_blur, persp, clip, cull, s))
        s_flagged[atomicAdd(&s_nflag, 1)] = yi * W + xi;
    }
    __syncthreads();
    const int nflag = s_nflag;
    if (tid == 0 && nflag > 0 && counters != nullptr) atomicAdd(counters, nflag);
    for (int e = warp; e < nflag; e += kClipWarps) {
      const int pix = s_flagged[e];
      const int yi = pix / W, xi = pix - yi * W;
      const float px = pix_to_ndc(W - 1 - xi, W, H), py = pix_to_ndc(H - 1 - yi, H, W);
      int qn = 0;   // warp-uniform (broadcast from lane 0 after every update)
      for (int fb = 0; fb < vd.face_count; fb += 32) {
        const int lf = fb + lane;
        Sample s;
        bool pass = false;
        if (lf < vd.face_count)
          pass = clip_candidate(load_face(face_verts, nullptr, vd, lf), px, py, blur, sqrt_blur, persp, clip, cull, s);
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {   // ascending face order
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float cz = __shfl_sync(0xffffffffu, s.z, src);
          const float cd = __shfl_sync(0xffffffffu, s.d, src);
          if (lane == 0) {
            const int f = vd.p2f_base + fb + src;
            const int nb = neighbor[vd.face_start + fb + src];
            bool drop = false;
            if (nb >= 0) {
              int at = -1;
              for (int t = 0; t < qn; ++t)
                if (q.f[t] == nb) { at = t; break; }
              if (at >= 0) {
                if (!(fabsf(cd) < fabsf(q.d[at]))) drop = true;
                else {
                  for (int t = at; t + 1 < qn; ++t) { q.z[t] = q.z[t + 1]; q.d[t] = q.d[t + 1]; q.f[t] = q.f[t + 1]; }
                  --qn;
                }
              }
            }
            if (!drop && !(qn == K && !cand_less(cz, f, q.z[K - 1], q.f[K - 1]))) {
              int pos = qn < K ? qn : K - 1;
              while (pos > 0 && cand_less(cz, f, q.z[pos - 1], q.f[pos - 1])) {
                q.z[pos] = q.z[pos - 1]; q.d[pos] = q.d[pos - 1]; q.f[pos] = q.f[pos - 1];
                --pos;
              }
              q.z[pos] = cz; q.d[pos] = cd; q.f[pos] = f;
              if (qn < K) ++qn;
            }
          }
          qn = __shfl_sync(0xffffffffu, qn, 0);
        }
      }
      __syncwarp();
      // write the K layers of this pixel; barycentrics are re-evaluated for the survivors
      const long long obase = ((long long)n * H * W + pix) * K;
      for (int k = lane; k < K; k += 32) {
        long long of = -1;
        Sample s;
        s.z = -1.0f; s.d = -1.0f; s.c0 = -1.0f; s.c1 = -1.0f; s.c2 = -1.0f;
        if (k < qn) {
          of = q.f[k];
          const FaceXYZ v = load_face(face_verts, nullptr, vd, q.f[k] - vd.p2f_base);
          eval_pixel_face_rt(v, px, py, persp, clip, blur, s);
        }
        p2f[obase + k] = of;
        zbuf[obase + k] = s.z;
        dists[obase + k] = s.d;
        bary[(obase + k) * 3] = s.c0; bary[(obase + k) * 3 + 1] = s.c1; bary[(obase + k) * 3 + 2] = s.c2;
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

// Does any (view, vertex) of the batch lie behind the plane z_view = z_plane?  `flag` is a persistent int32 the
// caller never resets: a hit raises it to this call's `epoch` (epochs grow), so flag == epoch answers the question.
__global__ void __launch_bounds__(256)
any_vertex_behind_kernel(const float* __restrict__ verts, const float* __restrict__ R, const float* __restrict__ T,
                         const trb_view* __restrict__ views, float z_plane, int epoch, int* __restrict__ flag) {
  const int n = blockIdx.y;
  const trb_view vd = views[n];
  const int lv = blockIdx.x * blockDim.x + threadIdx.x;
  bool behind = false;
  if (lv < vd.vert_count) {
    const float* x = verts + 3 * (size_t)(vd.world_vert_start + lv);
    behind = view_depth(__ldg(x), __ldg(x + 1), __ldg(x + 2), R + 9 * (size_t)n, T + 3 * (size_t)n) < z_plane;
  }
  if (__any_sync(0xffffffffu, behind) && (threadIdx.x & 31) == 0) atomicMax(flag, epoch);
}

}  // namespace trb

using namespace trb;

extern "C" int trb_any_vertex_behind(const float* verts_world, const float* R, const float* T, const trb_view* views,
                                     int N, int max_vert_count, float z_plane, int32_t epoch, int32_t* flag,
                                     int32_t* host_flag, void* event, int device, trb_stream_t stream) {
  if (N < 0 || max_vert_count < 0 || !(z_plane == z_plane) || !flag) return TRB_ERR_BAD_ARG;
  if (N > 65535) return TRB_ERR_BAD_ARG;
  if (N > 0 && max_vert_count > 0 && (!verts_world || !R || !T || !views)) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  if (N > 0 && max_vert_count > 0) {
    any_vertex_behind_kernel<<<dim3(ceil_div(max_vert_count, 256), N), 256, 0, st>>>(verts_world, R, T, views, z_plane,
                                                                                      epoch, flag);
    TRB_LAUNCH_CHECK();
  }
  if (host_flag) TRB_CUDA_TRY(cudaMemcpyAsync(host_flag, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (event) TRB_CUDA_TRY(cudaEventRecord((cudaEvent_t)event, st));
  return TRB_OK;
}

extern "C" int trb_clip_resequence(const float* face_verts, const trb_view* views, const int32_t* pair_face,
                                   const int32_t* pair_view, int64_t num_pairs, const int32_t* neighbor, int N,
                                   int H, int W, int K, float blur_radius, uint32_t flags, int64_t* pix_to_face,
                                   float* zbuf, float* bary, float* dists, int32_t* counters, int device,
                                   trb_stream_t stream) {
  if (N < 0 || H < 1 || W < 1 || K < 1 || num_pairs < 0 || !(blur_radius >= 0.0f)) return TRB_ERR_BAD_ARG;
  if (K > TRB_MAX_FACES_PER_PIXEL) return TRB_ERR_K_TOO_LARGE;
  if (N == 0 || num_pairs == 0) return TRB_OK;
  if (num_pairs > 0x7fffffff) return TRB_ERR_BAD_ARG;
  if (!face_verts || !views || !pair_face || !pair_view || !neighbor || !pix_to_face || !zbuf || !bary || !dists)
    return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  clip_resequence_kernel<<<(unsigned)num_pairs, kClipThreads, 0, st>>>(
      face_verts, views, pair_face, pair_view, neighbor, H, W, K, blur_radius, sqrtf(blur_radius), flags,
      (long long*)pix_to_face, zbuf, bary, dists, counters);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}
