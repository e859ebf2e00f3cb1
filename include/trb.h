/*
 * trb.h -- C ABI of the B200-native mesh-rendering hot path (libtrb.so).
 *
 * This is the drop-in boundary for the four native entry points the reference scripts reach
 * through PyTorch3D's pybind11 module `pytorch3d._C` (un-vendored dependency of
 * YufengJin/torch_renderer; call sites: torch_renderer.py:113,120,158,
 * camera_pose_optimizer.py:244-250, mesh_deformer.py:153,197, batch_rendering_test.py:252,274,
 * myrenderer.py:103-105, renderer.py:100-101), plus the fused shading/blending that PyTorch3D
 * expresses as ~40 ATen ops (renderer/mesh/shading.py, renderer/blending.py, renderer/lighting.py):
 *
 *   _C.rasterize_meshes            -> trb_raster_forward
 *   _C.rasterize_meshes_backward   -> trb_raster_backward
 *   clipped_faces_neighbor_idx     -> trb_clip_resequence (after clip_faces, torch_renderer_b200/clip.py)
 *   _C.interp_face_attrs_forward   -> trb_interp_forward
 *   _C.interp_face_attrs_backward  -> trb_interp_backward
 *   phong_shading + softmax_rgb_blend / sigmoid_alpha_blend / hard_rgb_blend
 *                                  -> trb_shade_forward / trb_shade_backward
 *   MeshRasterizer.transform (Transform3d bmm + divide)
 *                                  -> trb_transform_forward / trb_transform_backward
 *   Meshes.verts_normals_packed    -> trb_vertex_normals_forward / _backward
 *   MeshRenderer.forward + loss.backward() end to end (Fragments written once, read once)
 *                                  -> trb_render_forward / trb_render_backward
 *
 * Contract:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - the caller allocates every input, output and workspace buffer; the library never
 *     allocates or frees device memory, keeps no global state and never synchronises;
 *   - every call takes the CUDA device ordinal and the stream explicitly (autograd's backward
 *     runs on a different host thread than forward);
 *   - returns TRB_OK or a trb_status; trb_last_cuda_error() gives the cudaError_t of the
 *     calling thread's last TRB_ERR_CUDA;
 *   - layouts are PyTorch3D's: row-major contiguous, int64 pix_to_face, -1 background fill.
 */
#ifndef TRB_H_
#define TRB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 5 (round 2): trb_render_backward_allreduce + trb_peer_sum (the all-reduce's push half inside the backward tail);
 * per-view alpha sums at the end of hit_pixels (trb_render_sizes reports the new length); trb_render_extras
 * (trb_render_forward takes one more argument).
 * 4 (round 2): trb_render_config.sparse_fragments (was `reserved`), layer counts appended to the covered-pixel list
 * (trb_render_sizes reports the new length), trb_points_raster_forward_binned / _workspace_bytes,
 * trb_allreduce_set_timing. */
#define TRB_ABI_VERSION 5
#define TRB_MAX_FACES_PER_PIXEL 150

typedef void* trb_stream_t; /* cudaStream_t */

typedef enum trb_status {
  TRB_OK = 0,
  TRB_ERR_BAD_ARG = 1,      /* -> ValueError  */
  TRB_ERR_K_TOO_LARGE = 2,  /* -> ValueError: faces_per_pixel > 150 (PyTorch3D kMaxPointsPerPixel) */
  TRB_ERR_WORKSPACE = 3,    /* -> RuntimeError: workspace smaller than trb_raster_workspace_bytes */
  TRB_ERR_CUDA = 4          /* -> RuntimeError: see trb_last_cuda_error() */
} trb_status;

/* flags for the rasteriser entry points */
#define TRB_PERSPECTIVE_CORRECT 1u
#define TRB_CLIP_BARYCENTRIC 2u
#define TRB_CULL_BACKFACES 4u

/*
 * One record per view (image) of the batch; int32[8], device memory, `N` records.
 * A "view" is one (mesh, camera) pair.  Several views may share one mesh (the
 * `Meshes.extend(N)` pattern of batch_rendering_test.py:326 / mesh_deformer.py:150) without
 * the mesh being replicated in memory.
 */
typedef struct trb_view {
  int32_t face_start;      /* first row of `faces` drawn by this view */
  int32_t face_count;      /* number of rows */
  int32_t vert_delta;      /* faces[r][i] + vert_delta = row of that vertex in `verts_ndc` */
  int32_t p2f_base;        /* pix_to_face value of row face_start (packed face index) */
  int32_t world_vert_start;/* first row of this view's mesh in the world-space vertex arrays */
  int32_t vert_count;      /* number of vertices of that mesh */
  int32_t ndc_vert_start;  /* first row of this view's block in `verts_ndc` */
  int32_t reserved;
} trb_view;

int trb_abi_version(void);
const char* trb_status_string(int status);
int trb_last_cuda_error(void);
/* sizeof(trb_view | trb_shade_config | trb_render_config | trb_uv_texture | trb_peer_sum | trb_render_extras) for
 * which = 0 | 1 | 2 | 3 | 4 | 5 as compiled into the
 * library: bindings compare it with their own layout and refuse a stale build. */
int trb_abi_struct_size(int which);

/* ------------------------------------------------------------------------------------------
 * Camera transform (replaces MeshRasterizer.transform):
 *   X_view = X_world * R + T;  x_ndc = fx*X/Z + px,  y_ndc = fy*Y/Z + py  (perspective)
 *                               x_ndc = fx*X   + px,  y_ndc = fy*Y   + py  (orthographic)
 *   z_ndc  = Z_view.
 * verts_world f32[Vw,3]; R f32[N,3,3]; T f32[N,3]; proj f32[N,4] = (fx,fy,px,py) already in NDC
 * units; verts_ndc f32[sum_n vert_count,3].
 * Backward ACCUMULATES into grad_verts_world / grad_R / grad_T / grad_proj (caller zeroes them;
 * any of the four may be NULL).
 */
int trb_transform_forward(const float* verts_world, const float* R, const float* T,
                          const float* proj, const trb_view* views, int N, int max_vert_count,
                          int perspective, float* verts_ndc, int device, trb_stream_t stream);
int trb_transform_backward(const float* verts_world, const float* R, const float* T,
                           const float* proj, const trb_view* views, int N, int max_vert_count,
                           int perspective, const float* grad_verts_ndc, float* grad_verts_world,
                           float* grad_R, float* grad_T, float* grad_proj, int device,
                           trb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Rasteriser (replaces _C.rasterize_meshes / _C.rasterize_meshes_backward).
 *
 * verts_ndc f32[*,3]: NDC x,y and view-space z of every (view, vertex).
 * faces     i32[F,3] vertex ids, or NULL: then `verts_ndc` is PyTorch3D's face_verts f32[F,3,3]
 *           and face row r owns vertices 3r, 3r+1, 3r+2 (vert_delta is ignored).
 * pair_capacity: number of int32 (tile, face) slots in the workspace; tiles whose list does not
 *           fit are rasterised by scanning the whole mesh, so the result never depends on it.
 * stats     i32[4] (may be NULL): [0] pairs needed, [1] tiles that overflowed, [2] pair_capacity.
 * Outputs   pix_to_face i64[N,H,W,K], zbuf f32[N,H,W,K], bary f32[N,H,W,K,3], dists f32[N,H,W,K];
 *           every element is written (-1 where no face).
 */
int trb_raster_workspace_bytes(int N, int H, int W, int K, int64_t pair_capacity, size_t* bytes);
int trb_raster_forward(const float* verts_ndc, const int32_t* faces, const trb_view* views, int N,
                       int max_face_count, int H, int W, int K, float blur_radius, uint32_t flags,
                       int64_t pair_capacity, void* workspace, size_t workspace_bytes,
                       int64_t* pix_to_face, float* zbuf, float* bary, float* dists,
                       int32_t* stats, int device, trb_stream_t stream);
/* grad_verts_ndc f32, same shape as verts_ndc; ACCUMULATED into (caller zeroes). */
int trb_raster_backward(const float* verts_ndc, const int32_t* faces, const trb_view* views, int N,
                        int H, int W, int K, uint32_t flags, const int64_t* pix_to_face,
                        const float* grad_zbuf, const float* grad_bary, const float* grad_dists,
                        float* grad_verts_ndc, int device, trb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Faces cut by the near plane (replaces the `clipped_faces_neighbor_idx` argument of _C.rasterize_meshes;
 * PyTorch3D renderer/mesh/clip.py + the neighbour rule of rasterize_meshes.cu / rasterize_meshes_cpu.cpp).
 *
 * Call after trb_raster_forward(face_verts, faces = NULL, ...) on the clipped face list, with views whose
 * p2f_base == face_start.  A face cut into a quadrilateral is two consecutive rows t1, t1 + 1 that name each
 * other in `neighbor` i32[F'] (-1 elsewhere); pair_face i32[P] lists every t1, pair_view i32[P] its view.
 * Pixels where both halves of a pair are candidates are re-rasterised with upstream's order-dependent queue
 * (at most one half per pixel) and overwritten in the four outputs; all other pixels are left as they are.
 * counters i32[1] (may be NULL) accumulates the number of re-rasterised pixels.
 */
/* Raises *flag (persistent device i32, never reset by the caller) to `epoch` when any (view, vertex) has
 * view-space depth < z_plane; epochs must grow from call to call.  With host_flag (pinned host i32) the flag is
 * copied there on the stream, and with event (cudaEvent_t) the event is recorded after the copy: the caller waits
 * on the event -- not on the stream -- and reads host_flag >= epoch.  MeshRasterizer asks this for every render
 * with an active near plane (upstream's clip_faces reads two sums on the host at the same place). */
int trb_any_vertex_behind(const float* verts_world, const float* R, const float* T, const trb_view* views, int N,
                          int max_vert_count, float z_plane, int32_t epoch, int32_t* flag, int32_t* host_flag,
                          void* event, int device, trb_stream_t stream);
int trb_clip_resequence(const float* face_verts, const trb_view* views, const int32_t* pair_face,
                        const int32_t* pair_view, int64_t num_pairs, const int32_t* neighbor, int N, int H,
                        int W, int K, float blur_radius, uint32_t flags, int64_t* pix_to_face, float* zbuf,
                        float* bary, float* dists, int32_t* counters, int device, trb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * interpolate_face_attributes (replaces _C.interp_face_attrs_forward/backward).
 * pix_to_face i64[P], bary f32[P,3], face_attrs f32[F,3,D] -> out f32[P,D] (0 where face < 0).
 * Backward writes grad_bary f32[P,3] and ACCUMULATES into grad_face_attrs f32[F,3,D].
 */
int trb_interp_forward(const int64_t* pix_to_face, const float* bary, const float* face_attrs,
                       int64_t P, int64_t F, int D, float* out, int device, trb_stream_t stream);
int trb_interp_backward(const int64_t* pix_to_face, const float* bary, const float* face_attrs,
                        const float* grad_out, int64_t P, int64_t F, int D, float* grad_bary,
                        float* grad_face_attrs, int device, trb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Area-weighted vertex normals (replaces Meshes.verts_normals_packed).
 * verts f32[V,3], faces i32[F,3] -> raw f32[V,3] (un-normalised sum, kept for backward) and
 * normals f32[V,3] = raw / max(|raw|, 1e-6).  Backward ACCUMULATES into grad_verts.
 */
int trb_vertex_normals_forward(const float* verts, const int32_t* faces, int64_t V, int64_t F,
                               float* raw, float* normals, int device, trb_stream_t stream);
int trb_vertex_normals_backward(const float* verts, const int32_t* faces, int64_t V, int64_t F,
                                const float* raw, const float* grad_normals, float* grad_raw,
                                float* grad_verts, int device, trb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused shading + blending.
 */
#define TRB_SHADER_SOFT_PHONG 0      /* phong_shading + softmax_rgb_blend   */
#define TRB_SHADER_HARD_PHONG 1      /* phong_shading + hard_rgb_blend      */
#define TRB_SHADER_SOFT_SILHOUETTE 2 /* sigmoid_alpha_blend, RGB = 1        */
#define TRB_LIGHT_AMBIENT 0
#define TRB_LIGHT_POINT 1
#define TRB_LIGHT_DIRECTIONAL 2
#define TRB_TEX_VERTEX 0 /* per-vertex RGB interpolated in the kernel (TexturesVertex) */
#define TRB_TEX_TEXELS 1 /* caller supplies texels f32[N,H,W,K,3] (trb_shade_* only)      */
#define TRB_TEX_UV 2     /* TexturesUV sampled inside the fused kernels (trb_render_* only) */
#define TRB_VIEW_PARAM_STRIDE 20
/* view_params f32[N,20]: [0:3] light location/direction, [3:6] ambient (material*light),
 * [6:9] diffuse (material*light), [9:12] specular (material*light), [12] shininess,
 * [13:16] camera centre, [16] znear, [17] zfar, [18:20] reserved. */

typedef struct trb_shade_config {
  int32_t N, H, W, K;
  int32_t shader;       /* TRB_SHADER_*  */
  int32_t light_kind;   /* TRB_LIGHT_*   */
  int32_t texture_mode; /* TRB_TEX_*     */
  float sigma, gamma;
  float background[3];
} trb_shade_config;

/* faces i32[F,3] index verts_world / vert_normals / vert_colors (all f32[Vw,3]).
 * images f32[N,H,W,4]. */
int trb_shade_forward(const trb_shade_config* host_cfg, const trb_view* views,
                      const float* view_params, const int64_t* pix_to_face, const float* bary,
                      const float* zbuf, const float* dists, const int32_t* faces,
                      const float* verts_world, const float* vert_normals,
                      const float* vert_colors, const float* texels, float* images, int device,
                      trb_stream_t stream);
/* Writes grad_bary/grad_zbuf/grad_dists (same shapes as the fragments; any may be NULL) and
 * ACCUMULATES into grad_verts_world, grad_vert_normals, grad_vert_colors (f32[Vw,3]),
 * grad_texels is written (f32[N,H,W,K,3]); grad_view_params f32[N,20] accumulates the light
 * vector and camera-centre gradients.  Any output pointer may be NULL. */
int trb_shade_backward(const trb_shade_config* host_cfg, const trb_view* views,
                       const float* view_params, const int64_t* pix_to_face, const float* bary,
                       const float* zbuf, const float* dists, const int32_t* faces,
                       const float* verts_world, const float* vert_normals,
                       const float* vert_colors, const float* texels, const float* grad_images,
                       float* grad_bary, float* grad_zbuf, float* grad_dists,
                       float* grad_verts_world, float* grad_vert_normals, float* grad_vert_colors,
                       float* grad_texels, float* grad_view_params, int device,
                       trb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused render pipeline: everything MeshRenderer.forward / loss.backward() does for one batch of
 * views, in one call each way (vertex normals, camera transform, tile binning, fine rasterisation
 * with the shading + blending epilogue; and the reverse).  Fragments are written once in the
 * forward and read once in the backward; nothing else of size N*H*W*K touches HBM.
 *
 * shade.shader may be TRB_SHADER_NONE: rasterise only (MeshRasterizer.forward).  Textures are
 * per-vertex colours (TexturesVertex, texture_mode TRB_TEX_VERTEX) or one UV map (TRB_TEX_UV, host_uv);
 * anything else goes through trb_shade_forward with texels.
 */
#define TRB_SHADER_NONE (-1)

typedef struct trb_render_config {
  trb_shade_config shade;
  float blur_radius;
  uint32_t raster_flags;         /* TRB_PERSPECTIVE_CORRECT | TRB_CLIP_BARYCENTRIC | TRB_CULL_BACKFACES */
  int32_t perspective;           /* camera model: 1 = divide by view z, 0 = orthographic */
  int32_t max_face_count;
  int32_t max_vert_count;
  int32_t camera_center_from_rt; /* 1: view_params[n][13:16] := -T[n] * inv(R[n]) (and its gradient
                                    flows back into grad_R / grad_T) */
  int32_t want_light_grad;       /* backward: also accumulate d/d(light location|direction) into
                                    grad_view_params[:, 0:3] (camera centre is handled automatically) */
  float z_clip_value;            /* > 0: faces whose three vertices are all nearer than this view depth
                                    are culled (clip_faces' "fully behind the clip plane" case); 0: off */
  int64_t num_world_verts;       /* rows of verts_world / vert_colors */
  int64_t num_faces;             /* rows of faces */
  int64_t num_ndc_verts;         /* rows of verts_ndc = sum_n vert_count */
  int64_t pair_capacity;
  int32_t scratch_is_zeroed;     /* backward: the caller already zeroed `scratch` (e.g. it lives in the same
                                    zero-filled allocation as the gradient outputs): skip the memset */
  int32_t sparse_fragments;      /* forward: 1 = the caller will not look at the Fragments (MeshRenderer returns the
                                    image only; upstream's `renderer(meshes)` never exposes them): pix_to_face / zbuf /
                                    bary / dists are written ONLY for covered pixels -- layers [0, count) plus one -1
                                    terminator layer when count < K -- which is all the fused backward reads (it walks
                                    the covered-pixel list).  Background samples are left unwritten.  0 = PyTorch3D's
                                    dense layout with -1 fill (MeshRasterizer, MeshRendererWithFragments). */
} trb_render_config;

/* workspace_bytes: scratch for the forward; hit_pixels_len: length of hit_pixels (int32: a count
 * followed by the linear ids of the pixels that got at least one face -- the backward visits only
 * those); backward_scratch_floats: length of the f32 scratch the backward needs. */
/* TexturesUV for the fused pipeline (replaces TexturesUV.sample_textures: interp_face_attrs on the UVs +
 * grid_sample(flip(maps), 2uv-1, bilinear, align_corners=True, padding border); reference:
 * every load_objs_as_meshes() of data/cow_mesh/cow.obj, deform_mesh_with_color.py:266-271,329).
 * One map shared by all views of the batch; host struct holding device pointers. */
typedef struct trb_uv_texture {
  const float* map;         /* f32 [map_h, map_w, 3] */
  const float* verts_uvs;   /* f32 [Vt, 2] */
  const int32_t* faces_uvs; /* i32 [F, 3], row r belongs to row r of `faces` */
  float* grad_map;          /* backward only: f32 [map_h, map_w, 3], ACCUMULATED into; may be NULL */
  int32_t map_h, map_w;
} trb_uv_texture;

/* Optional work folded into the forward's FIRST kernel, so that a step has no copy / fill nodes of its own (host
 * struct, device pointers; NULL / 0 switches a member off; host_extras itself may be NULL):
 *   view_params_src  f32[N,20]: copied into `view_params` before anything reads it -- the caller keeps its (cached)
 *                    parameter block untouched although the call writes the camera centres into `view_params`;
 *   zero_buffer      f32[zero_count]: zero-filled -- meant for the ONE allocation that will hold the backward's
 *                    gradient outputs and scratch (then trb_render_backward needs no memset before it). */
typedef struct trb_render_extras {
  const float* view_params_src;
  float* zero_buffer;
  int64_t zero_count;
} trb_render_extras;

int trb_render_sizes(const trb_render_config* host_cfg, size_t* workspace_bytes, int64_t* hit_pixels_len,
                     int64_t* backward_scratch_floats);
/* view_params f32[N,20] is in/out (camera centre filled in when camera_center_from_rt).
 * Outputs: verts_ndc f32[num_ndc_verts,3]; normals_raw, normals f32[num_world_verts,3] (Phong only);
 * Fragments; images f32[N,H,W,4] (NULL when shader is NONE); hit_pixels i32[hit_pixels_len]: [0] = number of
 * covered pixels C, [1 .. 1+C) their linear pixel ids (tile by tile), and for faces_per_pixel > 1
 * [1+N*H*W .. 1+N*H*W+C) the number of layers each of them got -- the list the fused backward walks.  The LAST N
 * words are f32: the sum of the alpha channel images[n,:,:,3] of every view, accumulated by the fine kernel as it
 * writes the pixels (a coverage metric / silhouette-area term that costs no second pass over the image; float atomics:
 * the last bits depend on the order of arrival).  Untouched when shader is NONE. */
int trb_render_forward(const trb_render_config* host_cfg, const trb_view* views,
                       const float* verts_world, const int32_t* faces, const float* vert_colors,
                       const float* R, const float* T, const float* proj, float* view_params,
                       float* verts_ndc, float* normals_raw, float* normals, int64_t* pix_to_face,
                       float* zbuf, float* bary, float* dists, float* images, int32_t* hit_pixels,
                       void* workspace, size_t workspace_bytes, int32_t* stats,
                       const trb_uv_texture* host_uv /* NULL unless shade.texture_mode == TRB_TEX_UV */,
                       const trb_render_extras* host_extras /* may be NULL */, int device, trb_stream_t stream);
/* grad_images may be NULL (shader NONE); grad_zbuf / grad_bary / grad_dists are optional extra
 * upstream gradients on the Fragments.  Every grad_* output is ACCUMULATED into (caller zeroes;
 * any may be NULL); `scratch` is zeroed by the call unless cfg->scratch_is_zeroed. */
int trb_render_backward(const trb_render_config* host_cfg, const trb_view* views,
                        const float* verts_world, const int32_t* faces, const float* vert_colors,
                        const float* R, const float* T, const float* proj, const float* view_params,
                        const float* verts_ndc, const float* normals_raw, const float* normals,
                        const int64_t* pix_to_face, const float* zbuf, const float* bary,
                        const float* dists, const int32_t* hit_pixels, const float* grad_images,
                        const float* grad_zbuf, const float* grad_bary, const float* grad_dists,
                        float* grad_verts_world, float* grad_vert_colors, float* grad_R, float* grad_T,
                        float* grad_proj, float* grad_view_params, float* scratch,
                        const trb_uv_texture* host_uv, int device, trb_stream_t stream);

/* The same backward with the multi-GPU sum of the view-shared gradients fused into its tail (the
 * exchange step of the path): the last blocks of the kernel that finalises grad_verts_world / grad_vert_colors push
 * the finished values, tagged with the call's epoch, straight into every peer's inbox (trb_allreduce_sum_f32's
 * protocol, inbox and epoch counters -- fused and stand-alone calls may alternate), and a short receive kernel sums
 * what arrives, in rank order, back into grad_verts_world / grad_vert_colors: on return (in stream order) they hold
 * the sums over all ranks, bit-identical on every rank.  grad_R / grad_T / grad_proj / grad_view_params / the UV
 * map's gradient stay local.  Segments: grad_verts_world (3 * num_world_verts floats) when non-NULL, then
 * grad_vert_colors (Phong shaders with vertex colours) when non-NULL; together they must fit capacity_floats and
 * every rank must make the same sequence of calls with the same layout.  A rank with an empty batch cannot take
 * part (TRB_ERR_BAD_ARG).  A peer that does not arrive within the spin limit raises *error_flag (device int) and
 * leaves the affected values un-reduced.
 * Host struct; every pointer inside is a DEVICE pointer except host_peer_inbox (host array of `world` device
 * pointers: rank r's inbox as mapped into this process, as for trb_allreduce_sum_f32). */
typedef struct trb_peer_sum {
  void* const* host_peer_inbox;
  int64_t capacity_floats;
  int32_t rank, world;
  uint32_t* epochs;        /* trb_allreduce_grid(capacity_floats) counters, shared with trb_allreduce_sum_f32 */
  int32_t* error_flag;
  uint32_t* done_counter;  /* one zero-initialised word; the receive kernel leaves it zero again */
} trb_peer_sum;

int trb_render_backward_allreduce(const trb_render_config* host_cfg, const trb_view* views,
                                  const float* verts_world, const int32_t* faces, const float* vert_colors,
                                  const float* R, const float* T, const float* proj, const float* view_params,
                                  const float* verts_ndc, const float* normals_raw, const float* normals,
                                  const int64_t* pix_to_face, const float* zbuf, const float* bary,
                                  const float* dists, const int32_t* hit_pixels, const float* grad_images,
                                  const float* grad_zbuf, const float* grad_bary, const float* grad_dists,
                                  float* grad_verts_world, float* grad_vert_colors, float* grad_R, float* grad_T,
                                  float* grad_proj, float* grad_view_params, float* scratch,
                                  const trb_uv_texture* host_uv, const trb_peer_sum* host_peer, int device,
                                  trb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Point-cloud rendering (replaces _C.rasterize_points / _C.rasterize_points_backward and the compositors
 * _C.accum_alphacomposite / _C.accum_weightedsumnorm (+ _backward) behind PointsRasterizer / PointsRenderer /
 * AlphaCompositor / NormWeightedCompositor; reference: torch_renderer.py:163-208 -- SURVEY 8f rank 4, last item).
 *
 * points_ndc f32[P,3]: NDC x, y + view-space z of every point, packed over the batch; radius f32[P] per point (NDC).
 * views: one record per cloud -- face_start / face_count = first point / number of points, p2f_base = packed index
 * of the first point (the value written to idx).  A point covers a pixel iff z >= 0 and squared distance < radius^2;
 * the K nearest in z are kept, ordered by (z, index).
 * idx i32[N,H,W,K], zbuf f32[N,H,W,K], dists f32[N,H,W,K] (squared NDC distance); -1 where no point.
 * Backward ACCUMULATES into grad_points f32[P,3]: x, y from grad_dists, z from grad_zbuf (either may be NULL).
 */
int trb_points_raster_forward(const float* points_ndc, const float* radius, const trb_view* views, int N, int H,
                              int W, int K, int32_t* idx, float* zbuf, float* dists, int device,
                              trb_stream_t stream);
/* The same rasteriser with per-tile point lists (count -> allocate -> fill; points are their own bounding discs):
 * a tile only sees the points whose disc can reach it instead of streaming the whole cloud.  `workspace` (caller
 * allocated, trb_points_raster_workspace_bytes; contents need not be initialised) holds the tile counters and
 * `entry_capacity` list entries; tiles whose list does not fit fall back to the whole-cloud scan, so the result
 * never depends on the capacity.  `max_points` = largest point count of a view.  Same outputs, bit for bit. */
int trb_points_raster_workspace_bytes(int N, int H, int W, int64_t entry_capacity, size_t* bytes);
int trb_points_raster_forward_binned(const float* points_ndc, const float* radius, const trb_view* views, int N,
                                     int max_points, int H, int W, int K, int64_t entry_capacity, void* workspace,
                                     size_t workspace_bytes, int32_t* idx, float* zbuf, float* dists, int device,
                                     trb_stream_t stream);
int trb_points_raster_backward(const float* points_ndc, const int32_t* idx, const float* grad_zbuf,
                               const float* grad_dists, int N, int H, int W, int K, float* grad_points, int device,
                               trb_stream_t stream);
/* Compositors, channels-last: idx i32 / alphas f32 [num_pixels,K], features f32[P,C] -> images f32[num_pixels,C].
 * mode 0 (alpha_composite):   out_c = sum_k f[idx_k,c] a_k prod_{j<k} (1 - a_j)
 * mode 1 (norm_weighted_sum): out_c = sum_k f[idx_k,c] a_k / max(sum_k a_k, 1e-4)
 * Slots with idx < 0 are skipped; with `background` f32[C] (may be NULL) a pixel whose first slot is empty takes it.
 * Backward WRITES grad_alphas f32[num_pixels,K] and ACCUMULATES into grad_features f32[P,C] (either may be NULL). */
#define TRB_COMPOSITE_ALPHA 0
#define TRB_COMPOSITE_NORM_WEIGHTED 1
int trb_points_composite_forward(int mode, const int32_t* idx, const float* alphas, const float* features,
                                 int64_t num_pixels, int K, int C, const float* background, float* images,
                                 int device, trb_stream_t stream);
int trb_points_composite_backward(int mode, const int32_t* idx, const float* alphas, const float* features,
                                  const float* grad_images, int64_t num_pixels, int K, int C, int has_background,
                                  float* grad_alphas, float* grad_features, int device, trb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Nearest neighbour between batched point sets (replaces knn_points(K=1) inside pytorch3d.loss.chamfer_distance;
 * reference: mesh_deformer.py:307-311, deform_mesh_from_pcd.py:168-172 -- SURVEY 8f rank 4, the step right after
 * the render in the deformation loops).  x f32[N,P1,3], y f32[N,P2,3] -> dist f32[N,P1] squared distance to the
 * nearest y (lowest index on ties), idx i32[N,P1].  Backward ACCUMULATES into grad_x / grad_y (either may be NULL). */
int trb_nn_forward(const float* x, const float* y, int N, int P1, int P2, float* dist, int32_t* idx, int device,
                   trb_stream_t stream);
int trb_nn_backward(const float* x, const float* y, const int32_t* idx, const float* grad_dist, int N, int P1,
                    int P2, float* grad_x, float* grad_y, int device, trb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU: one-shot sum all-reduce of the view-shared gradients over peer memory.
 * The reference has no distributed path (every script pins cuda:0); the view batch is sharded over the GPUs of
 * one box and this is the step's only exchange.  host_segments[i] / host_counts[i]: up to 4 local f32 buffers,
 * reduced in place (sum over ranks, rank order).  host_peer_inbox[r] (r < world): this process's mapping of
 * rank r's inbox, 2 * world * capacity_floats 8-byte words, zero-initialised once with all ranks synchronised
 * before the first call (e.g. torch symmetric memory).  epochs u32[trb_allreduce_grid(capacity_floats)]
 * (device, local, zero-initialised once): per-block call counters.  error_flag i32[1] (device) is set when a
 * peer's data did not arrive within the spin limit.  Every rank must call with the same segment layout. */
int trb_allreduce_grid(int64_t capacity_floats);
int trb_allreduce_sum_f32(float* const* host_segments, const int64_t* host_counts, int num_segments,
                          void* const* host_peer_inbox, int64_t capacity_floats, int rank, int world,
                          uint32_t* epochs, int32_t* error_flag, int device, trb_stream_t stream);
/* Diagnostics: `device_u64x3` (device memory, zeroed by the caller) accumulates, for every later
 * trb_allreduce_sum_f32 launch of this process, block 0's push time and wait-and-sum time in %globaltimer
 * nanoseconds and the call count; NULL switches it off. */
int trb_allreduce_set_timing(uint64_t* device_u64x3);

/* Measurement hook (bench.py's roofline leg): when non-NULL, the four cudaEvent_t handles are
 * recorded on the call's stream immediately before / after the dominant kernel of
 * trb_render_forward (the fused fine pass) and of trb_render_backward (the fused backward).
 * Process-global and off by default; pass NULLs to switch it off again. */
int trb_debug_set_events(void* fwd_start, void* fwd_stop, void* bwd_start, void* bwd_stop);

#ifdef __cplusplus
}
#endif
#endif /* TRB_H_ */
