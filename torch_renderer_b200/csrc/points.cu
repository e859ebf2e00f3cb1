// Nearest neighbour between two batched point sets and its backward: the arithmetic of
// pytorch3d.loss.chamfer_distance (knn_points with K = 1) that the reference's deformation loops call right
// after the render (mesh_deformer.py:307-311, deform_mesh_from_pcd.py:168-172: 1,000 points sampled from each
// mesh).  One thread per query point; the other set streams through shared memory 256 points at a time.
// Ties keep the lowest index (strict <), like a sequential scan.
#include "trb_common.cuh"

namespace trb {

__global__ void __launch_bounds__(256)
nn_forward_kernel(const float* __restrict__ x, const float* __restrict__ y, int P1, int P2,
                  float* __restrict__ dist, int* __restrict__ idx) {
  __shared__ float sy[256 * 3];
  const int n = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float* xb = x + (size_t)n * P1 * 3;
  const float* yb = y + (size_t)n * P2 * 3;
  float px = 0.0f, py = 0.0f, pz = 0.0f;
  if (i < P1) { px = xb[3 * i]; py = xb[3 * i + 1]; pz = xb[3 * i + 2]; }
  float best = 3.0e38f;
  int best_j = 0;
  for (int base = 0; base < P2; base += 256) {
    const int m = min(256, P2 - base);
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * m; t += 256) sy[t] = yb[3 * (size_t)base + t];
    __syncthreads();
    if (i < P1) {
      for (int j = 0; j < m; ++j) {
        const float dx = px - sy[3 * j], dy = py - sy[3 * j + 1], dz = pz - sy[3 * j + 2];
        const float d = dx * dx + dy * dy + dz * dz;
        if (d < best) { best = d; best_j = base + j; }
      }
    }
  }
  if (i < P1) {
    dist[(size_t)n * P1 + i] = P2 > 0 ? best : 0.0f;
    idx[(size_t)n * P1 + i] = best_j;
  }
}

// d dist_i / d x_i = 2 (x_i - y_j*),  d dist_i / d y_j* = -2 (x_i - y_j*)
__global__ void __launch_bounds__(256)
nn_backward_kernel(const float* __restrict__ x, const float* __restrict__ y, const int* __restrict__ idx,
                   const float* __restrict__ g_dist, int P1, int P2, float* __restrict__ g_x,
                   float* __restrict__ g_y) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= P1 || P2 == 0) return;
  const size_t xi = ((size_t)n * P1 + i) * 3;
  const int j = idx[(size_t)n * P1 + i];
  const size_t yj = ((size_t)n * P2 + j) * 3;
  const float g = 2.0f * g_dist[(size_t)n * P1 + i];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float d = g * (x[xi + c] - y[yj + c]);
    if (g_x) atomicAdd(g_x + xi + c, d);
    if (g_y) atomicAdd(g_y + yj + c, -d);
  }
}

}  // namespace trb

using namespace trb;

extern "C" int trb_nn_forward(const float* x, const float* y, int N, int P1, int P2, float* dist, int32_t* idx,
                              int device, trb_stream_t stream) {
  if (N < 0 || P1 < 0 || P2 < 0 || N > 65535) return TRB_ERR_BAD_ARG;
  if (N == 0 || P1 == 0) return TRB_OK;
  if (!x || !dist || !idx || (P2 > 0 && !y)) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  nn_forward_kernel<<<dim3(ceil_div(P1, 256), N), 256, 0, (cudaStream_t)stream>>>(x, y, P1, P2, dist, idx);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

extern "C" int trb_nn_backward(const float* x, const float* y, const int32_t* idx, const float* grad_dist, int N,
                               int P1, int P2, float* grad_x, float* grad_y, int device, trb_stream_t stream) {
  if (N < 0 || P1 < 0 || P2 < 0 || N > 65535) return TRB_ERR_BAD_ARG;
  if (N == 0 || P1 == 0 || P2 == 0) return TRB_OK;
  if (!x || !y || !idx || !grad_dist) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  nn_backward_kernel<<<dim3(ceil_div(P1, 256), N), 256, 0, (cudaStream_t)stream>>>(x, y, idx, grad_dist, P1, P2,
                                                                                 grad_x, grad_y);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}
