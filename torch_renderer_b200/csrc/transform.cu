// World -> view -> NDC vertex transform and area-weighted vertex normals, forward and backward.
// Replaces MeshRasterizer.transform (Transform3d composition + bmm + divide) and
// Meshes.verts_normals_packed (A6); reference call sites: every `rasterizer(meshes, R=, T=)`
// (torch_renderer.py:113, camera_pose_optimizer.py:244) and every Phong shader call.
//
// One thread per (view, vertex): the N*V*3 NDC array is the only thing written; the (N*F,3,3)
// face_verts gather PyTorch3D materialises never exists.
#include "trb_common.cuh"

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
  const float* x = verts + 3 * (size_t)(vd.world_vert_start + lv);
  const float* r = R + 9 * (size_t)n;
  const float* t = T + 3 * (size_t)n;
  const float* p = proj + 4 * (size_t)n;
  const float X = __ldg(x), Y = __ldg(x + 1), Z = __ldg(x + 2);
  const float xv = X * __ldg(r + 0) + Y * __ldg(r + 3) + Z * __ldg(r + 6) + __ldg(t + 0);
  const float yv = X * __ldg(r + 1) + Y * __ldg(r + 4) + Z * __ldg(r + 7) + __ldg(t + 1);
  const float zv = X * __ldg(r + 2) + Y * __ldg(r + 5) + Z * __ldg(r + 8) + __ldg(t + 2);
  const float den = perspective ? zv : 1.0f;
  float* o = out + 3 * (size_t)(vd.ndc_vert_start + lv);
  o[0] = __ldg(p + 0) * xv / den + __ldg(p + 2);
  o[1] = __ldg(p + 1) * yv / den + __ldg(p + 3);
  o[2] = zv;
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
  const bool live = lv < vd.vert_count;
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
    if (grad_verts) {
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
  // block reduction: warp shuffles, then one atomic per warp and value
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

// ---- vertex normals ------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
face_normal_scatter_kernel(const float* __restrict__ verts, const int* __restrict__ faces, long long F,
                           float* __restrict__ raw) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
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

__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* __restrict__ raw, long long V, float* __restrict__ out) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float x = raw[3 * v], y = raw[3 * v + 1], z = raw[3 * v + 2];
  const float inv = 1.0f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-6f);
  out[3 * v] = x * inv; out[3 * v + 1] = y * inv; out[3 * v + 2] = z * inv;
}

__global__ void __launch_bounds__(256)
normalize_rows_backward_kernel(const float* __restrict__ raw, const float* __restrict__ gout, long long V,
                               float* __restrict__ graw) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float x = raw[3 * v], y = raw[3 * v + 1], z = raw[3 * v + 2];
  const float gx = gout[3 * v], gy = gout[3 * v + 1], gz = gout[3 * v + 2];
  const float len = sqrtf(x * x + y * y + z * z);
  if (len > 1e-6f) {
    const float inv = 1.0f / len;
    const float ux = x * inv, uy = y * inv, uz = z * inv;
    const float d = ux * gx + uy * gy + uz * gz;
    graw[3 * v] = (gx - ux * d) * inv; graw[3 * v + 1] = (gy - uy * d) * inv; graw[3 * v + 2] = (gz - uz * d) * inv;
  } else {
    graw[3 * v] = gx * 1e6f; graw[3 * v + 1] = gy * 1e6f; graw[3 * v + 2] = gz * 1e6f;
  }
}

__global__ void __launch_bounds__(256)
face_normal_backward_kernel(const float* __restrict__ verts, const int* __restrict__ faces, long long F,
                            const float* __restrict__ graw, float* __restrict__ gverts) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int i0 = __ldg(faces + 3 * f), i1 = __ldg(faces + 3 * f + 1), i2 = __ldg(faces + 3 * f + 2);
  const float* p0 = verts + 3 * (size_t)i0; const float* p1 = verts + 3 * (size_t)i1;
  const float* p2 = verts + 3 * (size_t)i2;
  const float ax = p2[0] - p1[0], ay = p2[1] - p1[1], az = p2[2] - p1[2];
  const float bx = p0[0] - p1[0], by = p0[1] - p1[1], bz = p0[2] - p1[2];
  const float gx = graw[3 * (size_t)i0] + graw[3 * (size_t)i1] + graw[3 * (size_t)i2];
  const float gy = graw[3 * (size_t)i0 + 1] + graw[3 * (size_t)i1 + 1] + graw[3 * (size_t)i2 + 1];
  const float gz = graw[3 * (size_t)i0 + 2] + graw[3 * (size_t)i1 + 2] + graw[3 * (size_t)i2 + 2];
  // n = a x b  =>  dL/da = b x g,  dL/db = g x a
  const float gax = by * gz - bz * gy, gay = bz * gx - bx * gz, gaz = bx * gy - by * gx;
  const float gbx = gy * az - gz * ay, gby = gz * ax - gx * az, gbz = gx * ay - gy * ax;
  float* d0 = gverts + 3 * (size_t)i0; float* d1 = gverts + 3 * (size_t)i1; float* d2 = gverts + 3 * (size_t)i2;
  atomicAdd(d2, gax); atomicAdd(d2 + 1, gay); atomicAdd(d2 + 2, gaz);
  atomicAdd(d0, gbx); atomicAdd(d0 + 1, gby); atomicAdd(d0 + 2, gbz);
  atomicAdd(d1, -gax - gbx); atomicAdd(d1 + 1, -gay - gby); atomicAdd(d1 + 2, -gaz - gbz);
}

}  // namespace trb

#include "trb_internal.cuh"

namespace trb {
int face_normals_backward(const float* verts, const int32_t* faces, int64_t F, const float* grad_raw,
                          float* grad_verts, cudaStream_t st) {
  if (F <= 0) return TRB_OK;
  face_normal_backward_kernel<<<(unsigned)ceil_div64(F, 256), 256, 0, st>>>(verts, faces, F, grad_raw, grad_verts);
  TRB_LAUNCH_CHECK();
  return TRB_OK;
}
}  // namespace trb

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
