// World -> view -> NDC vertex transform and area-weighted vertex normals, forward and backward.
// Replaces MeshRasterizer.transform (Transform3d composition + bmm + divide) and
// Meshes.verts_normals_packed (A6); reference call sites: every `rasterizer(meshes, R=, T=)`
// (torch_renderer.py:113, camera_pose_optimizer.py:244) and every Phong shader call.
//
// One thread per (view, vertex): the N*V*3 NDC array is the only thing written; the (N*F,3,3)
// face_verts gather PyTorch3D materialises never exists.
#include "stages.cuh"

namespace trb {

__global__ void __launch_bounds__(256)
transform_forward_kernel(const float* __restrict__ verts, const float* __restrict__ R,
                         const float* __restrict__ T, const float* __restrict__ proj,
                         const trb_view* __restrict__ views, int perspective,
                         float* __restrict__ out) {
  const int n = blockIdx.y;
  const trb_view vd = views[n];
  const int lv = blockIdx.x * blockDim.x + threadIdx.x;
  if (lv >= vd.vert_count) return;
  transform_vertex(verts, R, T, proj, vd, n, lv, perspective, out);
}

__global__ void __launch_bounds__(256)
transform_backward_kernel(const float* __restrict__ verts, const float* __restrict__ R,
                          const float* __restrict__ T, const float* __restrict__ proj,
                          const trb_view* __restrict__ views, int perspective,
                          const float* __restrict__ grad_ndc, int gstride, float* __restrict__ grad_verts,
                          float* __restrict__ grad_R, float* __restrict__ grad_T,
                          float* __restrict__ grad_proj) {
  const int n = blockIdx.y;
  const trb_view vd = views[n];
  const int lv = blockIdx.x * blockDim.x + threadIdx.x;
  transform_vertex_backward(verts, R, T, proj, vd, n, lv, lv < vd.vert_count, perspective, grad_ndc, gstride,
                            grad_verts, grad_R, grad_T, grad_proj);
}

// ---- vertex normals ------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
face_normal_scatter_kernel(const float* __restrict__ verts, const int* __restrict__ faces, long long F,
                           float* __restrict__ raw) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  face_normal_scatter_one(verts, faces, f, raw);
}

__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* __restrict__ raw, long long V, float* __restrict__ out) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  normalize_row_one(raw, v, out);
}

__global__ void __launch_bounds__(256)
normalize_rows_backward_kernel(const float* __restrict__ raw, const float* __restrict__ gout, long long V,
                               float* __restrict__ graw) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  normalize_row_backward(raw[3 * v], raw[3 * v + 1], raw[3 * v + 2], gout[3 * v], gout[3 * v + 1], gout[3 * v + 2],
                         graw[3 * v], graw[3 * v + 1], graw[3 * v + 2]);
}

__global__ void __launch_bounds__(256)
face_normal_backward_kernel(const float* __restrict__ verts, const int* __restrict__ faces, long long F,
                            const float* __restrict__ graw, float* __restrict__ gverts) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int i0 = __ldg(faces + 3 * f), i1 = __ldg(faces + 3 * f + 1), i2 = __ldg(faces + 3 * f + 2);
  const float gx = graw[3 * (size_t)i0] + graw[3 * (size_t)i1] + graw[3 * (size_t)i2];
  const float gy = graw[3 * (size_t)i0 + 1] + graw[3 * (size_t)i1 + 1] + graw[3 * (size_t)i2 + 1];
  const float gz = graw[3 * (size_t)i0 + 2] + graw[3 * (size_t)i1 + 2] + graw[3 * (size_t)i2 + 2];
  face_normal_backward_apply(verts, i0, i1, i2, gx, gy, gz, gverts);
}

}  // namespace trb

#include "trb_internal.cuh"

using namespace trb;

extern "C" int trb_transform_forward(const float* verts_world, const float* R, const float* T,
                                     const float* proj, const trb_view* views, int N, int max_vert_count,
                                     int perspective, float* verts_ndc, int device, trb_stream_t stream) {
  if (N < 0 || max_vert_count < 0) return TRB_ERR_BAD_ARG;
  if (N == 0 || max_vert_count == 0) return TRB_OK;
  if (N > 65535) return TRB_ERR_BAD_ARG;
  if (!verts_world || !R || !T || !proj || !views || !verts_ndc) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  dim3 grid(ceil_div(max_vert_count, 256), N);
  transform_forward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(verts_world, R, T, proj, views,
                                                                   perspective, verts_ndc);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

extern "C" int trb_transform_backward(const float* verts_world, const float* R, const float* T,
                                      const float* proj, const trb_view* views, int N, int max_vert_count,
                                      int perspective, const float* grad_verts_ndc, float* grad_verts_world,
                                      float* grad_R, float* grad_T, float* grad_proj, int device,
                                      trb_stream_t stream) {
  return trb::transform_backward_strided(verts_world, R, T, proj, views, N, max_vert_count, perspective,
                                         grad_verts_ndc, 3, grad_verts_world, grad_R, grad_T, grad_proj, device,
                                         stream);
}

int trb::transform_backward_strided(const float* verts_world, const float* R, const float* T, const float* proj,
                                    const trb_view* views, int N, int max_vert_count, int perspective,
                                    const float* grad_verts_ndc, int grad_stride, float* grad_verts_world,
                                    float* grad_R, float* grad_T, float* grad_proj, int device,
                                    trb_stream_t stream) {
  if (N < 0 || max_vert_count < 0) return TRB_ERR_BAD_ARG;
  if (N == 0 || max_vert_count == 0) return TRB_OK;
  if (N > 65535) return TRB_ERR_BAD_ARG;
  if (!verts_world || !R || !T || !proj || !views || !grad_verts_ndc) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  dim3 grid(ceil_div(max_vert_count, 256), N);
  transform_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      verts_world, R, T, proj, views, perspective, grad_verts_ndc, grad_stride, grad_verts_world, grad_R,
      grad_T, grad_proj);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

extern "C" int trb_vertex_normals_forward(const float* verts, const int32_t* faces, int64_t V, int64_t F,
                                          float* raw, float* normals, int device, trb_stream_t stream) {
  if (V < 0 || F < 0) return TRB_ERR_BAD_ARG;
  if (V == 0) return TRB_OK;
  if (!verts || !raw || !normals || (F > 0 && !faces)) return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  TRB_CUDA_TRY(cudaMemsetAsync(raw, 0, (size_t)V * 12, st));
  if (F > 0) {
    face_normal_scatter_kernel<<<(unsigned)ceil_div64(F, 256), 256, 0, st>>>(verts, faces, F, raw);
    TRB_LAUNCH_CHECK();
  }
  normalize_rows_kernel<<<(unsigned)ceil_div64(V, 256), 256, 0, st>>>(raw, V, normals);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}

extern "C" int trb_vertex_normals_backward(const float* verts, const int32_t* faces, int64_t V, int64_t F,
                                           const float* raw, const float* grad_normals, float* grad_raw,
                                           float* grad_verts, int device, trb_stream_t stream) {
  if (V < 0 || F < 0) return TRB_ERR_BAD_ARG;
  if (V == 0) return TRB_OK;
  if (!verts || !raw || !grad_normals || !grad_raw || !grad_verts || (F > 0 && !faces))
    return TRB_ERR_BAD_ARG;
  TRB_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  normalize_rows_backward_kernel<<<(unsigned)ceil_div64(V, 256), 256, 0, st>>>(raw, grad_normals, V, grad_raw);
  TRB_LAUNCH_CHECK();
  if (F > 0) {
    face_normal_backward_kernel<<<(unsigned)ceil_div64(F, 256), 256, 0, st>>>(verts, faces, F, grad_raw,
                                                                             grad_verts);
    TRB_LAUNCH_CHECK();
  }
  return TRB_OK;
}
