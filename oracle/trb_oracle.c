/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.
 *
 * CPU restatement (plain C, fp32, -ffp-contract=off) of the mesh-rasterisation
 * path the reference scripts reach through PyTorch3D on CPU tensors
 * (reference call sites: torch_renderer.py:113, camera_pose_optimizer.py:244,
 * batch_rendering_test.py:274, mesh_deformer.py:197; PyTorch3D is an
 * un-vendored, un-pinned dependency of the reference, inferred >= 0.6.1).
 *
 * The algorithm restated is PyTorch3D's published naive CPU rasteriser
 * (pytorch3d/csrc/rasterize_meshes/rasterize_meshes_cpu.cpp:
 * RasterizeMeshesNaiveCpu / RasterizeMeshesBackwardCpu, geometry_utils.h,
 * rasterization_utils.h) and interp_face_attrs, as written down in
 * SURVEY.md Appendix A (A3, A4, A5, A9).  PyTorch3D is not installable in the
 * build container, so **parity is unpinned**: no reference-owned golden vector
 * exists for this path.  Hand-checkable scenes in
 * tests/test_oracle.py pin the conventions instead.
 *
 * Every fp32 decision (coverage, blur band, z order) is evaluated with one IEEE
 * operation per source operator, left to right; the CUDA kernels use the same
 * operator sequence with FMA contraction disabled, which is what makes
 * pix_to_face bit-exact between the two.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define K_EPS 1e-8f
#define K_MAX_FACES_PER_PIXEL 150

typedef struct { float z; int64_t f; float d; float b0, b1, b2; } cand_t;

/* pixel centre in NDC; the short image side spans [-1, 1]. */
static float pix_to_ndc(int i, int S1, int S2) {
  float range = 2.0f;
  if (S1 > S2) range = ((float)S1 * range) / (float)S2;
  const float offset = range / 2.0f;
  return -offset + (range * (float)i + offset) / (float)S1;
}

/* SURVEY A4.1 */
static float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

static float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }
static float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

/* squared distance from p to segment ab. */
static float point_segment_d2(float px, float py, float ax, float ay, float bx, float by) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = bax * bax + bay * bay;
  if (l2 <= K_EPS) {
    const float dx = px - bx, dy = py - by;
    return dx * dx + dy * dy;
  }
  float t = (bax * (px - ax) + bay * (py - ay)) / l2;
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  const float qx = ax + t * bax, qy = ay + t * bay;
  const float dx = qx - px, dy = qy - py;
  return dx * dx + dy * dy;
}

/* (z, face) lexicographic order, SURVEY A5. */
static int cand_less(float za, int64_t fa, float zb, int64_t fb) {
  return (za < zb) || (za == zb && fa < fb);
}

/*
 * Evaluate one (pixel, face) pair, SURVEY A4 steps 2-9.
 * Returns 1 and fills *out when the face is a candidate for the pixel.
 */
static int eval_pixel_face(const float* v, float px, float py, float blur_radius,
                           float sqrt_blur, int perspective_correct, int clip_bary,
                           int cull_backfaces, cand_t* out) {
  const float x0 = v[0], y0 = v[1], z0 = v[2];
  const float x1 = v[3], y1 = v[4], z1 = v[5];
  const float x2 = v[6], y2 = v[7], z2 = v[8];
  const float xmin = min3f(x0, x1, x2) - sqrt_blur, xmax = max3f(x0, x1, x2) + sqrt_blur;
  const float ymin = min3f(y0, y1, y2) - sqrt_blur, ymax = max3f(y0, y1, y2) + sqrt_blur;
  const float zmin = min3f(z0, z1, z2), zmax = max3f(z0, z1, z2);
  const int outside_bbox = (px > xmax) || (px < xmin) || (py > ymax) || (py < ymin) || (zmin < K_EPS);
  const float face_area = edge_fn(x0, y0, x1, y1, x2, y2);
  const int back_face = face_area < 0.0f;
  const int zero_area = (face_area <= K_EPS) && (face_area >= -K_EPS);
  if (zmax < 0.0f || (cull_backfaces && back_face) || outside_bbox || zero_area) return 0;

  const float area = edge_fn(x2, y2, x0, y0, x1, y1) + K_EPS;
  const float w0 = edge_fn(px, py, x1, y1, x2, y2) / area;
  const float w1 = edge_fn(px, py, x2, y2, x0, y0) / area;
  const float w2 = edge_fn(px, py, x0, y0, x1, y1) / area;
  float b0 = w0, b1 = w1, b2 = w2;
  if (perspective_correct) {
    const float t0 = w0 * z1 * z2, t1 = w1 * z0 * z2, t2 = w2 * z0 * z1;
    const float den = fmaxf(t0 + t1 + t2, K_EPS);
    b0 = t0 / den; b1 = t1 / den; b2 = t2 / den;
  }
  float c0 = b0, c1 = b1, c2 = b2;
  if (clip_bary) {
    c0 = fmaxf(b0, 0.0f); c1 = fmaxf(b1, 0.0f); c2 = fmaxf(b2, 0.0f);
    const float s = fmaxf(c0 + c1 + c2, 1e-5f);
    c0 = c0 / s; c1 = c1 / s; c2 = c2 / s;
  }
  const float pz = c0 * z0 + c1 * z1 + c2 * z2;
  if (pz < 0.0f) return 0;
  const float e01 = point_segment_d2(px, py, x0, y0, x1, y1);
  const float e02 = point_segment_d2(px, py, x0, y0, x2, y2);
  const float e12 = point_segment_d2(px, py, x1, y1, x2, y2);
  const float dist = min3f(e01, e02, e12);
  const int inside = (b0 > 0.0f) && (b1 > 0.0f) && (b2 > 0.0f);
  if (!inside && dist >= blur_radius) return 0;
  out->z = pz; out->d = inside ? -dist : dist;
  out->b0 = c0; out->b1 = c1; out->b2 = c2;
  return 1;
}

typedef struct {
  const float* face_verts; const int64_t* first; const int64_t* count;
  int N, H, W, K; float blur_radius, sqrt_blur; int persp, clip, cull;
  int64_t* pix_to_face; float* zbuf; float* bary; float* dists;
  const int64_t* neighbor;  /* clipped_faces_neighbor_idx i64[F_total] or NULL */
  atomic_llong next_row;
} fwd_job_t;

static void rasterize_row(const fwd_job_t* j, int64_t row) {
  const int H = j->H, W = j->W, K = j->K;
  const int n = (int)(row / H), yi = (int)(row % H);
  const int64_t f0 = j->first[n], f1 = f0 + j->count[n];
  const float yf = pix_to_ndc(H - 1 - yi, H, W);
  cand_t q[K_MAX_FACES_PER_PIXEL + 1];
  for (int xi = 0; xi < W; ++xi) {
    const float xf = pix_to_ndc(W - 1 - xi, W, H);
    int qn = 0;
    for (int64_t f = f0; f < f1; ++f) {
      cand_t c;
      if (!eval_pixel_face(j->face_verts + f * 9, xf, yf, j->blur_radius, j->sqrt_blur, j->persp,
                           j->clip, j->cull, &c))
        continue;
      c.f = f;
      /*
       * A face cut by the near plane into a quadrilateral is drawn as two triangles (t1, t2 = t1 + 1) that
       * name each other in clipped_faces_neighbor_idx (PyTorch3D renderer/mesh/clip.py).  RasterizeMeshesNaiveCpu
       * keeps at most one of the two per pixel: when the neighbour is in the queue at this moment, the arriving
       * face replaces it iff its unsigned distance is strictly smaller, and is dropped otherwise; when the
       * neighbour is not in the queue (never a candidate, or already pushed out) the face is an ordinary one.
       * The rule depends on the order faces arrive in -- ascending face index -- and is restated as is.
       */
      if (j->neighbor != NULL && j->neighbor[f] >= 0) {
        const int64_t nb = j->neighbor[f];
        int at = -1;
        for (int i = 0; i < qn; ++i)
          if (q[i].f == nb) { at = i; break; }
        if (at >= 0) {
          if (!(fabsf(c.d) < fabsf(q[at].d))) continue;
          for (int i = at; i + 1 < qn; ++i) q[i] = q[i + 1];   /* take the neighbour out ... */
          --qn;                                                /* ... and insert the face below */
        }
      }
      /* sorted insert by (z, f); drop the largest when more than K. */
      if (qn == K && !cand_less(c.z, c.f, q[K - 1].z, q[K - 1].f)) continue;
      int pos = qn < K ? qn : K - 1;
      while (pos > 0 && cand_less(c.z, c.f, q[pos - 1].z, q[pos - 1].f)) {
        q[pos] = q[pos - 1];
        --pos;
      }
      q[pos] = c;
      if (qn < K) ++qn;
    }
    const int64_t base = (row * W + xi) * K;
    for (int k = 0; k < K; ++k) {
      const int hit = k < qn;
      j->pix_to_face[base + k] = hit ? q[k].f : -1;
      j->zbuf[base + k] = hit ? q[k].z : -1.0f;
      j->dists[base + k] = hit ? q[k].d : -1.0f;
      j->bary[(base + k) * 3 + 0] = hit ? q[k].b0 : -1.0f;
      j->bary[(base + k) * 3 + 1] = hit ? q[k].b1 : -1.0f;
      j->bary[(base + k) * 3 + 2] = hit ? q[k].b2 : -1.0f;
    }
  }
}

static void* fwd_worker(void* arg) {
  fwd_job_t* j = (fwd_job_t*)arg;
  const int64_t rows = (int64_t)j->N * j->H;
  for (;;) {
    const int64_t r0 = atomic_fetch_add(&j->next_row, 4);
    if (r0 >= rows) break;
    for (int64_t r = r0; r < r0 + 4 && r < rows; ++r) rasterize_row(j, r);
  }
  return NULL;
}

int trb_oracle_num_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (int)n;
}

/*
 * Naive forward rasteriser (restates RasterizeMeshesNaiveCpu).
 *   face_verts              f32 [F_total,3,3]  NDC x,y + view-space z
 *   mesh_to_face_first_idx  i64 [N]
 *   num_faces_per_mesh      i64 [N]
 * Outputs [N,H,W,K] (+[...,3] for bary), all -1 filled where no face.
 * Image rows are handed out to `num_threads` pthreads (<=0: all online cores).
 */
int trb_oracle_rasterize_forward_clipped(const float* face_verts, const int64_t* mesh_to_face_first_idx,
                                         const int64_t* num_faces_per_mesh,
                                         const int64_t* clipped_faces_neighbor_idx, int N, int H, int W, int K,
                                         float blur_radius, int perspective_correct,
                                         int clip_barycentric_coords, int cull_backfaces,
                                         int64_t* pix_to_face, float* zbuf, float* bary, float* dists,
                                         int num_threads);

int trb_oracle_rasterize_forward(const float* face_verts, const int64_t* mesh_to_face_first_idx,
                                 const int64_t* num_faces_per_mesh, int N, int H, int W, int K,
                                 float blur_radius, int perspective_correct,
                                 int clip_barycentric_coords, int cull_backfaces,
                                 int64_t* pix_to_face, float* zbuf, float* bary, float* dists,
                                 int num_threads) {
  return trb_oracle_rasterize_forward_clipped(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, NULL, N, H,
                                              W, K, blur_radius, perspective_correct, clip_barycentric_coords,
                                              cull_backfaces, pix_to_face, zbuf, bary, dists, num_threads);
}

/* Same, with `clipped_faces_neighbor_idx` i64[F_total] (-1 = none; NULL = no clipped faces): the fifth argument
 * of _C.rasterize_meshes that PyTorch3D's clip_faces produces for faces crossing the near plane. */
int trb_oracle_rasterize_forward_clipped(const float* face_verts, const int64_t* mesh_to_face_first_idx,
                                         const int64_t* num_faces_per_mesh,
                                         const int64_t* clipped_faces_neighbor_idx, int N, int H, int W, int K,
                                         float blur_radius, int perspective_correct,
                                         int clip_barycentric_coords, int cull_backfaces,
                                         int64_t* pix_to_face, float* zbuf, float* bary, float* dists,
                                         int num_threads) {
  if (K < 1 || K > K_MAX_FACES_PER_PIXEL) return 2;
  if (N < 0 || H < 1 || W < 1) return 1;
  if (num_threads <= 0) num_threads = trb_oracle_num_threads();
  if (num_threads > 256) num_threads = 256;
  fwd_job_t job = {face_verts, mesh_to_face_first_idx, num_faces_per_mesh, N, H, W, K,
                   blur_radius, sqrtf(blur_radius), perspective_correct, clip_barycentric_coords,
                   cull_backfaces, pix_to_face, zbuf, bary, dists, clipped_faces_neighbor_idx, 0};
  atomic_init(&job.next_row, 0);
  if (num_threads == 1) { fwd_worker(&job); return 0; }
  pthread_t th[256];
  int started = 0;
  for (int t = 0; t < num_threads; ++t)
    if (pthread_create(&th[started], NULL, fwd_worker, &job) == 0) ++started;
  if (started == 0) fwd_worker(&job);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
  return 0;
}

/* d edge(p,a,b) -> accumulates g * d/d{a,b}; p is a pixel centre (constant). */
static void edge_bwd(float px, float py, float ax, float ay, float bx, float by, float g,
                     float* ga, float* gb) {
  ga[0] += g * (py - by); ga[1] += g * (bx - px);
  gb[0] += g * (ay - py); gb[1] += g * (px - ax);
}

/* gradient of the squared point-segment distance, t held constant. */
static void point_segment_bwd(float px, float py, float ax, float ay, float bx, float by,
                              float g, float* ga, float* gb) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = bax * bax + bay * bay;
  if (l2 <= K_EPS) { /* d2 = |p - b|^2 */
    gb[0] += g * 2.0f * (bx - px); gb[1] += g * 2.0f * (by - py);
    return;
  }
  float t = (bax * (px - ax) + bay * (py - ay)) / l2;
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  const float qx = ax + t * bax, qy = ay + t * bay;
  const float dx = qx - px, dy = qy - py;
  ga[0] += g * (1.0f - t) * 2.0f * dx; ga[1] += g * (1.0f - t) * 2.0f * dy;
  gb[0] += g * t * 2.0f * dx;          gb[1] += g * t * 2.0f * dy;
}

/*
 * Backward rasteriser (restates RasterizeMeshesBackwardCpu).
 * The clip backward is evaluated at the perspective-corrected coordinates
 * (the mathematically consistent choice; see SURVEY A9 last paragraph).
 * Contributions are computed in fp32 and accumulated in fp64 so that the
 * result does not depend on pixel order.
 */
int trb_oracle_rasterize_backward(const float* face_verts, const int64_t* pix_to_face,
                                  const float* grad_zbuf, const float* grad_bary,
                                  const float* grad_dists, int N, int H, int W, int K,
                                  int64_t F_total, int perspective_correct,
                                  int clip_barycentric_coords, float* grad_face_verts) {
  double* acc = (double*)calloc((size_t)F_total * 9, sizeof(double));
  if (!acc) return 3;
  for (int n = 0; n < N; ++n)
    for (int yi = 0; yi < H; ++yi) {
      const float py = pix_to_ndc(H - 1 - yi, H, W);
      for (int xi = 0; xi < W; ++xi) {
        const float px = pix_to_ndc(W - 1 - xi, W, H);
        for (int k = 0; k < K; ++k) {
          const int64_t i = (((int64_t)n * H + yi) * W + xi) * K + k;
          const int64_t f = pix_to_face[i];
          if (f < 0) continue;
          const float* v = face_verts + f * 9;
          const float x0 = v[0], y0 = v[1], z0 = v[2], x1 = v[3], y1 = v[4], z1 = v[5];
          const float x2 = v[6], y2 = v[7], z2 = v[8];
          const float gz = grad_zbuf[i], gd = grad_dists[i];
          const float gb0 = grad_bary[i * 3], gb1 = grad_bary[i * 3 + 1], gb2 = grad_bary[i * 3 + 2];
          /* forward recompute */
          const float area = edge_fn(x2, y2, x0, y0, x1, y1) + K_EPS;
          const float e0 = edge_fn(px, py, x1, y1, x2, y2);
          const float e1 = edge_fn(px, py, x2, y2, x0, y0);
          const float e2 = edge_fn(px, py, x0, y0, x1, y1);
          const float w0 = e0 / area, w1 = e1 / area, w2 = e2 / area;
          float b0 = w0, b1 = w1, b2 = w2, den = 1.0f, t0 = 0, t1 = 0, t2 = 0, tsum = 0;
          if (perspective_correct) {
            t0 = w0 * z1 * z2; t1 = w1 * z0 * z2; t2 = w2 * z0 * z1;
            tsum = t0 + t1 + t2;
            den = fmaxf(tsum, K_EPS);
            b0 = t0 / den; b1 = t1 / den; b2 = t2 / den;
          }
          float c0 = b0, c1 = b1, c2 = b2, m0 = b0, m1 = b1, m2 = b2, s = 1.0f, ssum = 0;
          if (clip_barycentric_coords) {
            m0 = fmaxf(b0, 0.0f); m1 = fmaxf(b1, 0.0f); m2 = fmaxf(b2, 0.0f);
            ssum = m0 + m1 + m2;
            s = fmaxf(ssum, 1e-5f);
            c0 = m0 / s; c1 = m1 / s; c2 = m2 / s;
          }
          const int inside = (b0 > 0.0f) && (b1 > 0.0f) && (b2 > 0.0f);
          float g0[3] = {0, 0, 0}, g1[3] = {0, 0, 0}, g2[3] = {0, 0, 0}; /* x,y,z per vertex */
          /* zbuf = c . z */
          g0[2] += gz * c0; g1[2] += gz * c1; g2[2] += gz * c2;
          float gc0 = gb0 + gz * z0, gc1 = gb1 + gz * z1, gc2 = gb2 + gz * z2;
          /* clip backward */
          float gbb0 = gc0, gbb1 = gc1, gbb2 = gc2;
          if (clip_barycentric_coords) {
            /* c_i = m_i / s ; s = max(sum m, 1e-5) */
            float gs = -(gc0 * m0 + gc1 * m1 + gc2 * m2) / (s * s);
            if (!(ssum > 1e-5f)) gs = 0.0f;
            float gm0 = gc0 / s + gs, gm1 = gc1 / s + gs, gm2 = gc2 / s + gs;
            gbb0 = b0 > 0.0f ? gm0 : 0.0f;
            gbb1 = b1 > 0.0f ? gm1 : 0.0f;
            gbb2 = b2 > 0.0f ? gm2 : 0.0f;
          }
          /* perspective backward */
          float gw0 = gbb0, gw1 = gbb1, gw2 = gbb2;
          if (perspective_correct) {
            float gden = -(gbb0 * t0 + gbb1 * t1 + gbb2 * t2) / (den * den);
            if (!(tsum > K_EPS)) gden = 0.0f;
            const float gt0 = gbb0 / den + gden, gt1 = gbb1 / den + gden, gt2 = gbb2 / den + gden;
            gw0 = gt0 * z1 * z2; gw1 = gt1 * z0 * z2; gw2 = gt2 * z0 * z1;
            g0[2] += gt1 * w1 * z2 + gt2 * w2 * z1;
            g1[2] += gt0 * w0 * z2 + gt2 * w2 * z0;
            g2[2] += gt0 * w0 * z1 + gt1 * w1 * z0;
          }
          /* w_i = e_i / area */
          const float ge0 = gw0 / area, ge1 = gw1 / area, ge2 = gw2 / area;
          const float garea = -(gw0 * e0 + gw1 * e1 + gw2 * e2) / (area * area);
          edge_bwd(px, py, x1, y1, x2, y2, ge0, g1, g2);
          edge_bwd(px, py, x2, y2, x0, y0, ge1, g2, g0);
          edge_bwd(px, py, x0, y0, x1, y1, ge2, g0, g1);
          /* area = edge(v2; v0, v1): here the "point" v2 also carries gradient */
          g2[0] += garea * (y1 - y0); g2[1] += garea * (x0 - x1);
          g0[0] += garea * (y2 - y1); g0[1] += garea * (x1 - x2);
          g1[0] += garea * (y0 - y2); g1[1] += garea * (x2 - x0);
          /* signed squared distance through the arg-min edge */
          const float e01 = point_segment_d2(px, py, x0, y0, x1, y1);
          const float e02 = point_segment_d2(px, py, x0, y0, x2, y2);
          const float e12 = point_segment_d2(px, py, x1, y1, x2, y2);
          const float gsd = inside ? -gd : gd;
          if (e01 <= e02 && e01 <= e12) point_segment_bwd(px, py, x0, y0, x1, y1, gsd, g0, g1);
          else if (e02 <= e01 && e02 <= e12) point_segment_bwd(px, py, x0, y0, x2, y2, gsd, g0, g2);
          else point_segment_bwd(px, py, x1, y1, x2, y2, gsd, g1, g2);
          double* a = acc + f * 9;
          a[0] += g0[0]; a[1] += g0[1]; a[2] += g0[2];
          a[3] += g1[0]; a[4] += g1[1]; a[5] += g1[2];
          a[6] += g2[0]; a[7] += g2[1]; a[8] += g2[2];
        }
      }
    }
  for (int64_t i = 0; i < F_total * 9; ++i) grad_face_verts[i] = (float)acc[i];
  free(acc);
  return 0;
}

/* interp_face_attrs forward: out[p,d] = sum_i bary[p,i] * attrs[f_p,i,d]; 0 where f_p < 0. */
int trb_oracle_interp_forward(const int64_t* pix_to_face, const float* bary, const float* face_attrs,
                              int64_t P, int D, float* out) {
  for (int64_t p = 0; p < P; ++p) {
    const int64_t f = pix_to_face[p];
    for (int d = 0; d < D; ++d) {
      float acc = 0.0f;
      if (f >= 0)
        for (int i = 0; i < 3; ++i) acc += bary[p * 3 + i] * face_attrs[(f * 3 + i) * D + d];
      out[p * D + d] = acc;
    }
  }
  return 0;
}

/* interp_face_attrs backward (fp64 accumulation for grad_face_attrs). */
int trb_oracle_interp_backward(const int64_t* pix_to_face, const float* bary, const float* face_attrs,
                               const float* grad_out, int64_t P, int64_t F, int D, float* grad_bary,
                               float* grad_face_attrs) {
  double* acc = (double*)calloc((size_t)F * 3 * D, sizeof(double));
  if (!acc) return 3;
  for (int64_t p = 0; p < P; ++p) {
    const int64_t f = pix_to_face[p];
    for (int i = 0; i < 3; ++i) {
      float gb = 0.0f;
      if (f >= 0)
        for (int d = 0; d < D; ++d) {
          gb += grad_out[p * D + d] * face_attrs[(f * 3 + i) * D + d];
          acc[(f * 3 + i) * D + d] += (double)(bary[p * 3 + i] * grad_out[p * D + d]);
        }
      grad_bary[p * 3 + i] = gb;
    }
  }
  for (int64_t i = 0; i < F * 3 * D; ++i) grad_face_attrs[i] = (float)acc[i];
  free(acc);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Point clouds: restatement of PyTorch3D's naive CPU point rasteriser (pytorch3d/csrc/rasterize_points/
 * rasterize_points_cpu.cpp: RasterizePointsNaiveCpu / RasterizePointsBackwardCpu), which PointsRasterizer reaches
 * for the reference's AlphaPointRender / NormPointRender (torch_renderer.py:163-208; marked untested there).
 *   points   f32 [P,3]  NDC x, y + view-space z, packed over the batch
 *   first, count i64 [N] range of every cloud;  radius f32 [P] per point (NDC units)
 * A point is a candidate of a pixel iff z >= 0 and (xf - px)^2 + (yf - py)^2 < radius^2; the K nearest in z are kept,
 * ordered by (z, point index).  idx i32 [N,H,W,K] indexes the packed points; zbuf, dists (squared) f32; -1 fill.
 */
int trb_oracle_rasterize_points_forward(const float* points, const int64_t* first, const int64_t* count,
                                        const float* radius, int N, int H, int W, int K, int32_t* idx, float* zbuf,
                                        float* dists) {
  if (K < 1 || K > K_MAX_FACES_PER_PIXEL) return 2;
  if (N < 0 || H < 1 || W < 1) return 1;
  cand_t q[K_MAX_FACES_PER_PIXEL + 1];
  for (int n = 0; n < N; ++n) {
    for (int yi = 0; yi < H; ++yi) {
      const float yf = pix_to_ndc(H - 1 - yi, H, W);
      for (int xi = 0; xi < W; ++xi) {
        const float xf = pix_to_ndc(W - 1 - xi, W, H);
        int qn = 0;
        for (int64_t p = first[n]; p < first[n] + count[n]; ++p) {
          const float px = points[3 * p], py = points[3 * p + 1], pz = points[3 * p + 2];
          if (pz < 0.0f) continue;
          const float dx = xf - px, dy = yf - py;
          const float d2 = dx * dx + dy * dy;
          const float r2 = radius[p] * radius[p];
          if (!(d2 < r2)) continue;
          if (qn == K && !cand_less(pz, p, q[K - 1].z, q[K - 1].f)) continue;
          int pos = qn < K ? qn : K - 1;
          while (pos > 0 && cand_less(pz, p, q[pos - 1].z, q[pos - 1].f)) { q[pos] = q[pos - 1]; --pos; }
          q[pos].z = pz; q[pos].f = p; q[pos].d = d2;
          if (qn < K) ++qn;
        }
        const int64_t base = (((int64_t)n * H + yi) * W + xi) * K;
        for (int k = 0; k < K; ++k) {
          idx[base + k] = k < qn ? (int32_t)q[k].f : -1;
          zbuf[base + k] = k < qn ? q[k].z : -1.0f;
          dists[base + k] = k < qn ? q[k].d : -1.0f;
        }
      }
    }
  }
  return 0;
}

/* grad_points f32 [P,3] (zeroed by the caller): d/dx = 2 (px - xf) g_dist, d/dy likewise, d/dz = g_zbuf. */
int trb_oracle_rasterize_points_backward(const float* points, const int32_t* idx, const float* grad_zbuf,
                                         const float* grad_dists, int N, int H, int W, int K, double* grad_points) {
  for (int n = 0; n < N; ++n)
    for (int yi = 0; yi < H; ++yi) {
      const float yf = pix_to_ndc(H - 1 - yi, H, W);
      for (int xi = 0; xi < W; ++xi) {
        const float xf = pix_to_ndc(W - 1 - xi, W, H);
        const int64_t base = (((int64_t)n * H + yi) * W + xi) * K;
        for (int k = 0; k < K; ++k) {
          const int32_t p = idx[base + k];
          if (p < 0) continue;
          grad_points[3 * (int64_t)p] += 2.0 * (double)grad_dists[base + k] * ((double)points[3 * p] - (double)xf);
          grad_points[3 * (int64_t)p + 1] += 2.0 * (double)grad_dists[base + k] * ((double)points[3 * p + 1] - (double)yf);
          grad_points[3 * (int64_t)p + 2] += (double)grad_zbuf[base + k];
        }
      }
    }
  return 0;
}
