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

// Launch with the programmatic-serialisation attribute (programmatic dependent launch): the kernel may be
// scheduled while its predecessor in the stream drains; it must begin with pdl_wait() (stages.cuh).  Captured
// into a CUDA graph this becomes a programmatic edge.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace trb
