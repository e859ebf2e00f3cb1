// Declarations shared by the fused forward kernels (render.cu: faces_per_pixel == 1; render_kn.cu:
// faces_per_pixel > 1): kernel arguments, the covered-pixel list, the per-sample shading call.
#pragma once
#include "raster_internal.cuh"
#include "shade_math.cuh"
#include "trb_internal.cuh"
#include "allreduce.cuh"

namespace trb {

struct FineArgs {
  const float* verts_ndc; const int* faces; const trb_view* views;
  int H, W, K; float blur_radius, sqrt_blur, z_cull; unsigned flags; TileGrid tg;
  const int* tile_count; const int* tile_offset; const int2* pairs;  // (face, bits of min vertex z)
  long long* p2f; float* zbuf; float* bary; float* dists; float* images; int* hit_pixels;
  int* hit_counts;  // K > 1: layers of each covered pixel, parallel to hit_pixels[1..] (the backward's loop bounds)
  const int* ws_header; const int* busy_tiles;
  const float* view_params; const float* verts_world; const float* normals; const float* colors;
  float sigma, gamma, bg0, bg1, bg2;
  UvTex uv;  // uv.map != nullptr: TexturesUV instead of per-vertex colours
  int sparse;  // trb_render_config::sparse_fragments: Fragments only for covered pixels (+ a -1 terminator layer)
  float* alpha_sum;  // f32[N] behind the covered-pixel list: per-view sum of images[..., 3] (zeroed by prep_kernel)
};

// Appends the linear indices of the pixels of this CTA that got at least one face to the global
// list the backward pass walks (hit_pixels[0] = count, [1..] = pixel ids, row-major inside a tile so
// that neighbouring lanes of the backward still see neighbouring pixels).  Must be reached by every
// thread of the CTA.
template <int NT>
__device__ __forceinline__ void append_hit_pixels(int* hit_pixels, bool hit, int pix, int* hit_counts = nullptr,
                                                  int cnt = 0) {
  __shared__ int s_wcnt[NT / 32];
  __shared__ int s_base;
  if (hit_pixels == nullptr) return;  // stand-alone rasteriser: no fused backward follows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned ball = __ballot_sync(0xffffffffu, hit);
  if (lane == 0) s_wcnt[warp] = __popc(ball);
  __syncthreads();
  if (tid == 0) {
    int tot = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) { const int c = s_wcnt[w]; s_wcnt[w] = tot; tot += c; }
    s_base = tot > 0 ? atomicAdd(hit_pixels, tot) : 0;
  }
  __syncthreads();
  if (hit) {
    const int at = s_base + s_wcnt[warp] + __popc(ball & ((1u << lane) - 1u));
    hit_pixels[1 + at] = pix;
    if (hit_counts != nullptr) hit_counts[at] = cnt;
  }
}

struct ShadeIn {
  const float* verts_world; const float* normals; const float* colors; const int* faces; UvTex uv;
};

template <int LIGHT>
__device__ __forceinline__ F3 shade_sample(const ShadeIn& in, const trb_view& vd, const ViewParams& vp,
                                           int local_face, float b0, float b1, float b2) {
  const size_t r = (size_t)(vd.face_start + local_face);
  const int i0 = __ldg(in.faces + 3 * r), i1 = __ldg(in.faces + 3 * r + 1), i2 = __ldg(in.faces + 3 * r + 2);
  F3 tex;
  if (in.uv.map != nullptr) {
    const F2 t0 = ld2(in.uv.verts_uvs, __ldg(in.uv.faces_uvs + 3 * r)), t1 = ld2(in.uv.verts_uvs, __ldg(in.uv.faces_uvs + 3 * r + 1)),
             t2 = ld2(in.uv.verts_uvs, __ldg(in.uv.faces_uvs + 3 * r + 2));
    const UvTap k = uv_tap(in.uv, b0 * t0.x + b1 * t1.x + b2 * t2.x, b0 * t0.y + b1 * t1.y + b2 * t2.y);
    tex = uv_sample(in.uv, k, nullptr, nullptr);
  } else {
    tex = interp3(b0, b1, b2, ld3(in.colors, i0), ld3(in.colors, i1), ld3(in.colors, i2));
  }
  if (LIGHT == TRB_LIGHT_AMBIENT) return {vp.amb[0] * tex.x, vp.amb[1] * tex.y, vp.amb[2] * tex.z};
  const F3 P = interp3(b0, b1, b2, ld3(in.verts_world, i0), ld3(in.verts_world, i1), ld3(in.verts_world, i2));
  const F3 nr = interp3(b0, b1, b2, ld3(in.normals, i0), ld3(in.normals, i1), ld3(in.normals, i2));
  Lit lit;
  return phong_color<LIGHT>(vp, P, nr, tex, lit);
}

// The small stages around the fine pass, packed by block role (render_stages.cu).
int run_forward_stages(const trb_render_config* cfg, const trb_view* views, const float* verts_world,
                       const int32_t* faces, const float* R, const float* T, const float* proj,
                       float* view_params, float* verts_ndc, float* normals_raw, float* normals,
                       int32_t* hit_pixels, void* workspace, const TileGrid& tg, const WsLayout& ws, bool lit,
                       cudaStream_t st, const trb_render_extras* extras = nullptr);
int run_backward_post(const trb_render_config* cfg, const trb_view* views, const float* verts_world,
                      const int32_t* faces, const float* R, const float* T, const float* proj,
                      const float* view_params, const float* g_view_params, const float* normals_raw,
                      const float4* g_ndc4, const float4* g_world4, const float4* g_col4, const float4* g_norm4,
                      float* grad_verts, float* grad_colors, float* grad_R, float* grad_T, float* grad_proj,
                      bool geom, bool cam_chain, bool normals_chain, cudaStream_t st, const ArPush* push = nullptr);

// Fine pass of one batch: the K == 1 strip kernel (render.cu) or the K > 1 kernel (render_kn.cu).
int launch_render_fine(int shader, int light, int N, cudaStream_t st, const FineArgs& a);
// faces_per_pixel > 1 (render_kn.cu)
int launch_render_fine_kn(int shader, int light, int N, cudaStream_t st, const FineArgs& a);

}  // namespace trb
