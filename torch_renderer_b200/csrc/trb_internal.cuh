// Host-side helpers shared between translation units of libtrb.so (not part of the C ABI).
#pragma once
#include "trb_common.cuh"

namespace trb {

// trb_transform_backward with a configurable row stride of grad_verts_ndc (3, or 4 for the float4
// accumulators of the fused backward).
int transform_backward_strided(const float* verts_world, const float* R, const float* T, const float* proj,
                               const trb_view* views, int N, int max_vert_count, int perspective,
                               const float* grad_verts_ndc, int grad_stride, float* grad_verts_world,
                               float* grad_R, float* grad_T, float* grad_proj, int device, trb_stream_t stream);

// Second half of trb_vertex_normals_backward: grad of the raw (un-normalised) normals -> grad_verts.
int face_normals_backward(const float* verts, const int32_t* faces, int64_t F, const float* grad_raw,
                          float* grad_verts, cudaStream_t st);

}  // namespace trb
