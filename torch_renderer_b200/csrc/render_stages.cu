// The small stages of the fused render pipeline, packed by block role so that a render is 5 kernels
// forward (prep -> count -> alloc -> fill -> fine) and 2 backward (fused backward -> post) with no memset
// nodes in between; every launch carries the programmatic-serialisation attribute, so each kernel's CTAs
// are already resident when its predecessor drains (stages.cuh: pdl_wait).  Under CUDA-graph replay of the
// 64-view cow batch the seventeen small nodes this replaces cost ~60 us of a 277 us step.
#include "render_internal.cuh"
#include "stages.cuh"

namespace trb {

#ifdef TRB_STEP_STAMPS
static __device__ StepStamps g_stamps_stages = {nullptr, nullptr};
int set_step_stamps_stages(unsigned long long* ring, unsigned* step) {
  StepStamps s = {ring, step};
  return cudaMemcpyToSymbol(g_stamps_stages, &s, sizeof(s)) == cudaSuccess ? TRB_OK : TRB_ERR_CUDA;
}
#endif

// ---- forward -----------------------------------------------------------------------------------------
struct PrepArgs {
  const float* verts; const float* R; const float* T; const float* proj; const trb_view* views;
  int perspective, bpv, blocks_transform;
  float* verts_ndc; float* view_params;  // view_params != null: write the camera centres
  int* zero_a; long long n_zero_a;       // binning counters (ints)
  float* zero_b; long long n_zero_b;     // raw vertex normals
  int* zero_c;                           // covered-pixel counter
  float* zero_d; int n_zero_d;           // per-view alpha sums behind the covered-pixel list
  const float* vp_src; float* vp_dst;    // trb_render_extras: the parameter block is copied in here ...
  float* zero_e; long long n_zero_e;     // ... and the backward's gradient + scratch allocation zeroed
};

// role 0: world -> NDC of every (view, vertex) + the camera centre of every view; role 1: zero the
// accumulators the next stages add into.
__global__ void __launch_bounds__(256) prep_kernel(const PrepArgs p) {
  pdl_wait();
  const int b = blockIdx.x;
#ifdef TRB_STEP_STAMPS
  if (b == 0 && threadIdx.x == 0 && g_stamps_stages.ring) {
    const unsigned step = *g_stamps_stages.step + 1u;
    unsigned long long* row = g_stamps_stages.ring + (size_t)(step & 255u) * 8;
    row[0] = stamp_now(); row[1] = ~0ull; row[2] = 0ull; row[3] = ~0ull; row[4] = 0ull;
    __threadfence();
    *g_stamps_stages.step = step;
  }
#endif
  if (b < p.blocks_transform) {
    const int n = b / p.bpv, bx = b - n * p.bpv;
    const trb_view vd = p.views[n];
    const int lv = bx * 256 + threadIdx.x;
    if (lv < vd.vert_count) transform_vertex(p.verts, p.R, p.T, p.proj, vd, n, lv, p.perspective, p.verts_ndc);
    if (bx == 0) {
      const int t = threadIdx.x;
      // (slots 13..15 are the camera centre: thread 0 writes them below when the kernel derives it)
      if (p.vp_src && t < TRB_VIEW_PARAM_STRIDE && !(p.view_params && t >= 13 && t < 16))
        p.vp_dst[(size_t)n * TRB_VIEW_PARAM_STRIDE + t] = p.vp_src[(size_t)n * TRB_VIEW_PARAM_STRIDE + t];
      if (p.view_params && t == 0) camera_center_one(p.R, p.T, p.view_params, n);
    }
    return;
  }
  const long long stride = (long long)(gridDim.x - p.blocks_transform) * 256;
  const long long i0 = (long long)(b - p.blocks_transform) * 256 + threadIdx.x;
  for (long long i = i0; i < p.n_zero_a; i += stride) p.zero_a[i] = 0;
  for (long long i = i0; i < p.n_zero_b; i += stride) p.zero_b[i] = 0.0f;
  if (i0 == 0 && p.zero_c) *p.zero_c = 0;
  if (p.zero_d) for (long long i = i0; i < p.n_zero_d; i += stride) p.zero_d[i] = 0.0f;
  if (p.zero_e) {
    if ((reinterpret_cast<size_t>(p.zero_e) & 15) == 0) {
      const long long n4 = p.n_zero_e >> 2;
      float4* z4 = reinterpret_cast<float4*>(p.zero_e);
      for (long long i = i0; i < n4; i += stride) z4[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      for (long long i = (n4 << 2) + i0; i < p.n_zero_e; i += stride) p.zero_e[i] = 0.0f;
    } else {
      for (long long i = i0; i < p.n_zero_e; i += stride) p.zero_e[i] = 0.0f;
    }
  }
}

struct BinArgs {
  const float* verts_ndc; const int* faces; const trb_view* views;
  int H, W; TileGrid tg; float sqrt_blur; bool cull; float z_cull;
  int* tile_count; int* tile_fill; int* tile_offset; int2* pairs;
  int bpf, blocks_bin;  // blocks per view, N * bpf
};

// role 0: count the faces of every tile; role 1: scatter the face normals to their vertices.
__global__ void __launch_bounds__(256)
count_kernel(const BinArgs a, const float* __restrict__ verts_world, long long F, float* __restrict__ normals_raw) {
  pdl_wait();
  const int b = blockIdx.x;
  if (b < a.blocks_bin) {
    const int n = b / a.bpf, bx = b - n * a.bpf;
    const trb_view vd = a.views[n];
    bin_face<false>(a.verts_ndc, a.faces, vd, n, bx * 256 + threadIdx.x, a.H, a.W, a.tg, a.sqrt_blur, a.cull,
                    a.tile_count, a.tile_fill, a.tile_offset, a.pairs, a.z_cull);
    return;
  }
  const long long f = (long long)(b - a.blocks_bin) * 256 + threadIdx.x;
  if (f < F) face_normal_scatter_one(verts_world, a.faces, f, normals_raw);
}

// role 0: hand every non-empty tile its slice of the pair list; role 1: normalise the vertex normals.
__global__ void __launch_bounds__(256)
alloc_kernel(const int* __restrict__ tile_count, int* __restrict__ tile_offset, int ntiles, int* __restrict__ header,
             long long pair_capacity, int* __restrict__ busy_list, int blocks_alloc,
             const float* __restrict__ normals_raw, long long V, float* __restrict__ normals) {
  pdl_wait();
  const int b = blockIdx.x;
  if (b < blocks_alloc) {
    alloc_tile(tile_count, tile_offset, ntiles, header, pair_capacity, busy_list, b * 256 + threadIdx.x);
    return;
  }
  const long long v = (long long)(b - blocks_alloc) * 256 + threadIdx.x;
  if (v < V) normalize_row_one(normals_raw, v, normals);
}

__global__ void __launch_bounds__(256) fill_kernel(const BinArgs a) {
  pdl_wait();
  const int b = blockIdx.x;
  const int n = b / a.bpf, bx = b - n * a.bpf;
  const trb_view vd = a.views[n];
  bin_face<true>(a.verts_ndc, a.faces, vd, n, bx * 256 + threadIdx.x, a.H, a.W, a.tg, a.sqrt_blur, a.cull,
                 a.tile_count, a.tile_fill, a.tile_offset, a.pairs, a.z_cull);
}

int run_forward_stages(const trb_render_config* cfg, const trb_view* views, const float* verts_world,
                       const int32_t* faces, const float* R, const float* T, const float* proj,
                       float* view_params, float* verts_ndc, float* normals_raw, float* normals,
                       int32_t* hit_pixels, void* workspace, const TileGrid& tg, const WsLayout& ws, bool lit,
                       cudaStream_t st, const trb_render_extras* extras) {
  const trb_shade_config& sc = cfg->shade;
  const int N = sc.N;
  unsigned char* wsb = (unsigned char*)workspace;
  const long long V = cfg->num_world_verts, F = cfg->num_faces;
  const int bpv = max(1, ceil_div(cfg->max_vert_count, 256));
  const int bpf = max(1, ceil_div(cfg->max_face_count, 256));
  const int ntiles = N * tg.tiles_x * tg.tiles_y;

  PrepArgs p;
  p.verts = verts_world; p.R = R; p.T = T; p.proj = proj; p.views = views;
  p.perspective = cfg->perspective; p.bpv = bpv; p.blocks_transform = N * bpv;
  p.verts_ndc = verts_ndc;
  p.view_params = (lit && cfg->camera_center_from_rt) ? view_params : nullptr;
  p.zero_a = (int*)wsb; p.n_zero_a = (long long)(ws.offset / 4);  // header, tile_count, tile_fill
  p.zero_b = lit ? normals_raw : nullptr; p.n_zero_b = lit ? 3 * V : 0;
  p.zero_c = hit_pixels;
  p.zero_d = (hit_pixels && sc.shader != TRB_SHADER_NONE)
                 ? reinterpret_cast<float*>(hit_pixels + 1 + (size_t)N * sc.H * sc.W * (sc.K > 1 ? 2 : 1)) : nullptr;
  p.n_zero_d = N;
  p.vp_src = (extras && view_params) ? extras->view_params_src : nullptr; p.vp_dst = view_params;
  p.zero_e = extras ? extras->zero_buffer : nullptr;
  p.n_zero_e = (extras && extras->zero_buffer) ? extras->zero_count : 0;
  if (p.n_zero_e <= 0) p.zero_e = nullptr;
  const long long zero_words = p.n_zero_a + p.n_zero_b + p.n_zero_e / 4;
  long long bz = ceil_div64(zero_words, 256 * 4);
  if (bz < 1) bz = 1;
  if (bz > 4 * kNumSMs) bz = 4 * kNumSMs;
  const int blocks_zero = (int)bz;
  TRB_CUDA_TRY(launch_pdl(prep_kernel, dim3(p.blocks_transform + blocks_zero), dim3(256), 0, st, p));
  if (cfg->max_face_count == 0) return TRB_OK;

  BinArgs b;
  b.verts_ndc = verts_ndc; b.faces = faces; b.views = views; b.H = sc.H; b.W = sc.W; b.tg = tg;
  b.sqrt_blur = sqrtf(cfg->blur_radius); b.cull = cfg->raster_flags & TRB_CULL_BACKFACES;
  b.z_cull = fmaxf(cfg->z_clip_value, 0.0f);
  b.tile_count = (int*)(wsb + ws.count); b.tile_fill = (int*)(wsb + ws.fill);
  b.tile_offset = (int*)(wsb + ws.offset); b.pairs = (int2*)(wsb + ws.pairs);
  b.bpf = bpf; b.blocks_bin = N * bpf;
  const int blocks_scatter = lit ? (int)ceil_div64(F, 256) : 0;
  TRB_CUDA_TRY(launch_pdl(count_kernel, dim3(b.blocks_bin + blocks_scatter), dim3(256), 0, st, b, verts_world, F,
                          normals_raw));
  const int blocks_alloc = ceil_div(ntiles, 256);
  const int blocks_norm = lit ? (int)ceil_div64(V, 256) : 0;
  TRB_CUDA_TRY(launch_pdl(alloc_kernel, dim3(blocks_alloc + blocks_norm), dim3(256), 0, st,
                          (const int*)b.tile_count, b.tile_offset, ntiles, (int*)(wsb + ws.header),
                          (long long)cfg->pair_capacity, (int*)(wsb + ws.busy), blocks_alloc,
                          (const float*)normals_raw, V, normals));
  TRB_CUDA_TRY(launch_pdl(fill_kernel, dim3(b.blocks_bin), dim3(256), 0, st, b));
  return TRB_OK;
}

// ---- backward ------------------------------------------------------------------------------------------
struct PostArgs {
  const float* verts; const int* faces; const float* R; const float* T; const float* proj; const trb_view* views;
  const float* view_params; const float* g_view_params; const float* normals_raw;
  const float4* g_ndc4; const float4* g_world4; const float4* g_col4; const float4* g_norm4;
  float* grad_verts; float* grad_colors; float* grad_R; float* grad_T; float* grad_proj;
  int N, perspective, bpv; long long V, F;
  int b_transform, b_cam, b_final;  // cumulative block boundaries of the roles
  int b_push;                       // first block of the push role (= number of blocks of roles 0-3); grid size when off
  int has_push;
  ArPush push;
};

// Everything after the fused backward kernel, in one launch (all accumulation is atomic):
//  role 0  NDC -> world: grad of the (view, vertex) NDC positions into grad_verts and the per-view R / T / proj;
//  role 1  camera centre -> R, T;
//  role 2  per vertex: unpack the float4 accumulators of the world-position path and of the colours;
//  role 3  per face: gradient of the unit vertex normals through the normalisation (recomputed per corner
//          instead of a separate per-vertex pass) and through the area-weighted face normal;
//  role 4  (multi-GPU, opt-in: trb_render_backward_allreduce) the LAST blocks of the grid push this rank's finished
//          grad_verts / grad_colors into the peers' inboxes (allreduce.cuh) as soon as every block of roles 0-3 has
//          signed off on a device counter: the gradients leave for the peers from inside the kernel that finishes
//          them, and the NVLink flight overlaps this kernel's drain and the launch of the receive kernel.
//          Blocks are dispatched in index order, so every block the push role waits for is resident or finished
//          when it starts; the wait is bounded all the same (error flag instead of a hang).
__global__ void __launch_bounds__(256) post_backward_kernel(const PostArgs p) {
  pdl_wait();
  const int b = blockIdx.x;
#ifdef TRB_STEP_STAMPS
  struct StampExit {   // every exit path of the block stamps "last block exit"
    unsigned long long* row;
    __device__ ~StampExit() { if (row && threadIdx.x == 0) atomicMax(row + 2, stamp_now()); }
  } stamp_exit = {g_stamps_stages.ring ? stamp_row(g_stamps_stages) : nullptr};
  if (stamp_exit.row && threadIdx.x == 0) atomicMin(stamp_exit.row + 1, stamp_now());
#endif
  if (p.has_push && b >= p.b_push) {
    const unsigned epoch = p.push.epochs[b - p.b_push] + 1u;   // fetched while the counter is still moving
    if (threadIdx.x == 0) {
      unsigned spins = 0;
      while (ld_acquire_gpu_u32(p.push.done) < (unsigned)p.b_push)
        if (++spins > (1u << 24)) { atomicExch(p.push.error, 1); break; }
    }
    __syncthreads();
    ar_push_block(p.push.seg, p.push.peers, p.push.rank, p.push.world, p.push.capacity, epoch, b - p.b_push);
    return;
  }
  if (b < p.b_transform) {
    const int n = b / p.bpv, bx = b - n * p.bpv;
    const trb_view vd = p.views[n];
    const int lv = bx * 256 + threadIdx.x;
    transform_vertex_backward(p.verts, p.R, p.T, p.proj, vd, n, lv, lv < vd.vert_count, p.perspective,
                              reinterpret_cast<const float*>(p.g_ndc4), 4, p.grad_verts, p.grad_R, p.grad_T,
                              p.grad_proj);
  } else if (b < p.b_cam) {
    const int n = (b - p.b_transform) * 256 + threadIdx.x;
    if (n < p.N) camera_center_backward_one(p.R, p.view_params, p.g_view_params, p.grad_R, p.grad_T, n);
  } else if (b < p.b_final) {
    const long long v = (long long)(b - p.b_cam) * 256 + threadIdx.x;
    if (v < p.V) {
      if (p.g_world4 && p.grad_verts) {
        const float4 g = p.g_world4[v];
        atomicAdd(p.grad_verts + 3 * v, g.x); atomicAdd(p.grad_verts + 3 * v + 1, g.y);
        atomicAdd(p.grad_verts + 3 * v + 2, g.z);
      }
      if (p.g_col4 && p.grad_colors) {
        const float4 g = p.g_col4[v];
        p.grad_colors[3 * v] += g.x; p.grad_colors[3 * v + 1] += g.y; p.grad_colors[3 * v + 2] += g.z;
      }
    }
  } else {
    const long long f = (long long)(b - p.b_final) * 256 + threadIdx.x;
    if (f < p.F) {
      const int ids[3] = {__ldg(p.faces + 3 * f), __ldg(p.faces + 3 * f + 1), __ldg(p.faces + 3 * f + 2)};
      float gx = 0.0f, gy = 0.0f, gz = 0.0f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float4 g = p.g_norm4[ids[k]];
        const float* r = p.normals_raw + 3 * (size_t)ids[k];
        float ox, oy, oz;
        normalize_row_backward(r[0], r[1], r[2], g.x, g.y, g.z, ox, oy, oz);
        gx += ox; gy += oy; gz += oz;
      }
      face_normal_backward_apply(p.verts, ids[0], ids[1], ids[2], gx, gy, gz, p.grad_verts);
    }
  }
  if (p.has_push) {
    // this block's share of the gradients is in L2 before the counter moves (fence, then barrier, then one atomic)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(p.push.done, 1u);
  }
}

int run_backward_post(const trb_render_config* cfg, const trb_view* views, const float* verts_world,
                      const int32_t* faces, const float* R, const float* T, const float* proj,
                      const float* view_params, const float* g_view_params, const float* normals_raw,
                      const float4* g_ndc4, const float4* g_world4, const float4* g_col4, const float4* g_norm4,
                      float* grad_verts, float* grad_colors, float* grad_R, float* grad_T, float* grad_proj,
                      bool geom, bool cam_chain, bool normals_chain, cudaStream_t st, const ArPush* push) {
  PostArgs p = PostArgs();
  p.verts = verts_world; p.faces = faces; p.R = R; p.T = T; p.proj = proj; p.views = views;
  p.view_params = view_params; p.g_view_params = g_view_params; p.normals_raw = normals_raw;
  p.g_ndc4 = g_ndc4; p.g_world4 = g_world4; p.g_col4 = g_col4; p.g_norm4 = g_norm4;
  p.grad_verts = grad_verts; p.grad_colors = grad_colors; p.grad_R = grad_R; p.grad_T = grad_T;
  p.grad_proj = grad_proj;
  p.N = cfg->shade.N; p.perspective = cfg->perspective; p.bpv = max(1, ceil_div(cfg->max_vert_count, 256));
  p.V = cfg->num_world_verts; p.F = cfg->num_faces;
  const bool finalize = (g_world4 && grad_verts) || (g_col4 && grad_colors);
  p.b_transform = geom ? p.N * p.bpv : 0;
  p.b_cam = p.b_transform + (cam_chain ? ceil_div(p.N, 256) : 0);
  p.b_final = p.b_cam + (finalize ? (int)ceil_div64(p.V, 256) : 0);
  int total = p.b_final + (normals_chain ? (int)ceil_div64(p.F, 256) : 0);
  p.b_push = total;
  p.has_push = 0;
  if (push) {
    p.has_push = 1;
    p.push = *push;
    total += (int)ceil_div64(push->seg.start[push->seg.count], kArChunk);
  }
  if (total == 0) return TRB_OK;
  TRB_CUDA_TRY(launch_pdl(post_backward_kernel, dim3(total), dim3(256), 0, st, p));
  if (push) return launch_allreduce_receive(*push, st);
  return TRB_OK;
}

}  // namespace trb
