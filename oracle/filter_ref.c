/*
 * filter_ref.c -- CPU ORACLE for the grasp-pose filter.  TEST INFRASTRUCTURE ONLY: nothing in
 * catgrasp_b200/ links or loads this; tests, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs do.
 *
 * Pose logic: restates my_cpp/common.cpp:159 (canonical_to_cam), :185-197 (tf * pose, compose,
 * normalise the first three columns), :199-212 (approach-direction test), :253-299 (lateral
 * offset search with the float step accumulator: only 0, 0.001f, 0.001f+0.001f execute; order
 * (0,+),(1,+),(1,-),(2,+),(2,-); first collision-free wins; none -> zero matrix) in the fp32
 * operation order of the reference build (Eigen fixed-size products, SSE2, no FMA contraction).
 *
 * PINNED: oracle/build_ref.py compiles the reference's common.cpp itself (Eigen 3.2.92 as vendored, -O3, SSE2); its
 * filterGraspPose returns bit-identical survivors to this file (tests/test_mycpp_golden.py).
 *
 * Geometry predicate: the reference calls FCL (BVH mesh vs octomap OcTree,
 * my_cpp/collision_manager.cpp:93-111); FCL and octomap are not in /root/reference nor installed,
 * so that boundary is PARITY UNPINNED.  Here, as in the CUDA kernel, the predicate is the gripper
 * SDF one of meshpy/meshpy/sdf.py: scene points are mapped into the posed gripper's grid
 * ((x - origin)/res, sdf.py:252-264) and the pose collides iff any point has sd < 0, with sd either
 * trilinear (sdf.py:292-343) or nearest-cell over in-bounds cells (sdf.py:377-389).
 *
 * Build: gcc -O2 -mfma -ffp-contract=off -fopenmp -shared -fPIC  (oracle/build_oracle.py).
 * -ffp-contract=off keeps every a*b+c below as two roundings; fmaf() is the only fused operation.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  const float *grid;
  int nx, ny, nz;
  float ox, oy, oz;
  float inv_res;
} sdf_view;

static void mm4(const float *A, const float *B, float *O) {
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      float s = A[r * 4 + 0] * B[0 * 4 + c];
      s = s + A[r * 4 + 1] * B[1 * 4 + c];
      s = s + A[r * 4 + 2] * B[2 * 4 + c];
      s = s + A[r * 4 + 3] * B[3 * 4 + c];
      O[r * 4 + c] = s;
    }
}

static void normalize_col(float *G, int col) {
  const float x = G[0 * 4 + col], y = G[1 * 4 + col], z = G[2 * 4 + col];
  const float n = sqrtf((x * x + y * y) + z * z);
  G[0 * 4 + col] = x / n;
  G[1 * 4 + col] = y / n;
  G[2 * 4 + col] = z / n;
}

static void affine_inverse(const float *A, float *inv) {
  const float a = A[0], b = A[1], c = A[2], d = A[4], e = A[5], f = A[6], g = A[8], h = A[9], i = A[10];
  const float c00 = e * i - f * h;
  const float c01 = f * g - d * i;
  const float c02 = d * h - e * g;
  const float det = (a * c00 + b * c01) + c * c02;
  const float r = 1.0f / det;
  inv[0] = c00 * r;
  inv[1] = (c * h - b * i) * r;
  inv[2] = (b * f - c * e) * r;
  inv[3] = c01 * r;
  inv[4] = (a * i - c * g) * r;
  inv[5] = (c * d - a * f) * r;
  inv[6] = c02 * r;
  inv[7] = (b * g - a * h) * r;
  inv[8] = (a * e - b * d) * r;
  const float tx = A[3], ty = A[7], tz = A[11];
  for (int k = 0; k < 3; k++) inv[9 + k] = -((inv[k * 3 + 0] * tx + inv[k * 3 + 1] * ty) + inv[k * 3 + 2] * tz);
}

static float sdf_trilinear(const sdf_view *s, float gx, float gy, float gz) {
  const float cx = fminf(fmaxf(gx, 0.f), (float)(s->nx - 1));
  const float cy = fminf(fmaxf(gy, 0.f), (float)(s->ny - 1));
  const float cz = fminf(fmaxf(gz, 0.f), (float)(s->nz - 1));
  const float lx = floorf(cx), ly = floorf(cy), lz = floorf(cz);
  const int ix = (int)lx, iy = (int)ly, iz = (int)lz;
  const float wx0 = 1.f - (cx - lx), wx1 = 1.f - ((lx + 1.f) - cx);
  const float wy0 = 1.f - (cy - ly), wy1 = 1.f - ((ly + 1.f) - cy);
  const float wz0 = 1.f - (cz - lz), wz1 = 1.f - ((lz + 1.f) - cz);
  const int hx = (ix + 1) < s->nx, hy = (iy + 1) < s->ny, hz = (iz + 1) < s->nz;
  const size_t sx = (size_t)s->ny * s->nz, sy = (size_t)s->nz;
  const float *p = s->grid + (size_t)ix * sx + (size_t)iy * sy + iz;
  /* corner order of Sdf3D (sdf.py:217-225) */
  const float v0 = p[0];
  const float v1 = hx ? p[sx] : 0.f;
  const float v2 = hy ? p[sy] : 0.f;
  const float v3 = hz ? p[1] : 0.f;
  const float v4 = (hx && hy) ? p[sx + sy] : 0.f;
  const float v5 = (hy && hz) ? p[sy + 1] : 0.f;
  const float v6 = (hx && hz) ? p[sx + 1] : 0.f;
  const float v7 = (hx && hy && hz) ? p[sx + sy + 1] : 0.f;
  float sd = 0.f;
  sd = fmaf((wx0 * wy0) * wz0, v0, sd);
  sd = fmaf((wx1 * wy0) * wz0, v1, sd);
  sd = fmaf((wx0 * wy1) * wz0, v2, sd);
  sd = fmaf((wx0 * wy0) * wz1, v3, sd);
  sd = fmaf((wx1 * wy1) * wz0, v4, sd);
  sd = fmaf((wx0 * wy1) * wz1, v5, sd);
  sd = fmaf((wx1 * wy0) * wz1, v6, sd);
  sd = fmaf((wx1 * wy1) * wz1, v7, sd);
  return sd;
}

static float sdf_nearest(const sdf_view *s, float gx, float gy, float gz, int clamp, int *inb) {
  float rx = rintf(gx), ry = rintf(gy), rz = rintf(gz);
  int ok = (rx >= 0.f) && (rx < (float)s->nx) && (ry >= 0.f) && (ry < (float)s->ny) && (rz >= 0.f) && (rz < (float)s->nz);
  if (!ok) {
    if (!clamp) { *inb = 0; return 0.f; }
    rx = fminf(fmaxf(rx, 0.f), (float)(s->nx - 1));
    ry = fminf(fmaxf(ry, 0.f), (float)(s->ny - 1));
    rz = fminf(fmaxf(rz, 0.f), (float)(s->nz - 1));
  }
  *inb = 1;
  return s->grid[((size_t)(int)rx * s->ny + (int)ry) * s->nz + (int)rz];
}

static void fold_grid(const float *inv, const sdf_view *s, float *out) {
  for (int k = 0; k < 9; k++) out[k] = inv[k] * s->inv_res;
  out[9] = (inv[9] - s->ox) * s->inv_res;
  out[10] = (inv[10] - s->oy) * s->inv_res;
  out[11] = (inv[11] - s->oz) * s->inv_res;
}

/* G = camera frame -> grid coordinates ((x - origin)/res of sdf.py:252-264 folded into the inverse gripper pose) */
static int point_hits(const sdf_view *s, const float *G, int mode, float margin, float x, float y, float z) {
  const float gx = fmaf(G[2], z, fmaf(G[1], y, fmaf(G[0], x, G[9])));
  const float gy = fmaf(G[5], z, fmaf(G[4], y, fmaf(G[3], x, G[10])));
  const float gz = fmaf(G[8], z, fmaf(G[7], y, fmaf(G[6], x, G[11])));
  if (mode == 0) return sdf_trilinear(s, gx, gy, gz) < margin;
  int inb;
  const float sd = sdf_nearest(s, gx, gy, gz, 0, &inb);
  return inb && (sd < margin);
}

static int any_hits(const sdf_view *s, const float *inv, int mode, float margin, const float *pts, int P) {
  for (int p = 0; p < P; p++)
    if (point_hits(s, inv, mode, margin, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2])) return 1;
  return 0;
}

/* The geometry predicate alone, for a given gripper pose in the camera frame (row-major 4x4): used by filter_ref below
 * and by the oracle/_ref build of the reference's common.cpp (oracle/ref_shim/collision_manager_sdf.cpp). */
int gripper_hits_ref(const float *gripper_in_cam, const float *grid, const int *dims, const float *origin, float res,
                     int mode, const float *pts, int P) {
  sdf_view s = {grid, dims[0], dims[1], dims[2], origin[0], origin[1], origin[2], 1.0f / res};
  float inv[12], g[12];
  affine_inverse(gripper_in_cam, inv);
  fold_grid(inv, &s, g);
  return any_hits(&s, g, mode, 0.f, pts, P);
}

/* status: 0 accept, 1 approach-direction reject, 3 collision reject; offset: 0..4 or -1.
 * split != 0 and adjust == 0: 3 = the open gripper hits the object's points (common.cpp:231-238, n_open_gripper_rej),
 * 4 = only the enclosed gripper hits the background (:241-248, n_close_gripper_rej).  With adjust != 0 the reference
 * counts every collision rejection as n_open_gripper_rej (:290-294): always 3. */
void filter_ref_ms(const float *nocs_pose, const float *canonical_to_nocs, const float *gripper_in_grasp,
                int filter_dir, int adjust, int sdf_mode, float sdf_margin, int split, const float *grasp_poses, int G, const float *sym, int S,
                const float *grid_open, const int *dims_open, const float *origin_open, float res_open,
                const float *open_pts, int P1, const float *grid_encl, const int *dims_encl,
                const float *origin_encl, float res_encl, const float *encl_pts, int P2, int nthreads,
                uint8_t *out_status, int8_t *out_offset, float *out_poses) {
  sdf_view so = {grid_open, dims_open[0], dims_open[1], dims_open[2], origin_open[0], origin_open[1], origin_open[2],
                 1.0f / res_open};
  sdf_view se = so;
  if (grid_encl) {
    sdf_view t = {grid_encl, dims_encl[0], dims_encl[1], dims_encl[2], origin_encl[0], origin_encl[1], origin_encl[2],
                  1.0f / res_encl};
    se = t;
  }
  float c2c[16];
  mm4(nocs_pose, canonical_to_nocs, c2c); /* common.cpp:159 */
  const float step1 = 0.001f;
  const float step2 = step1 + 0.001f;
  const long Q = (long)G * S;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic)
  for (long q = 0; q < Q; q++) {
    const int i = (int)(q / S), j = (int)(q % S);
    float tmp[16], g[16];
    mm4(sym + (size_t)j * 16, grasp_poses + (size_t)i * 16, tmp); /* :190 */
    mm4(c2c, tmp, g);                                             /* :191 */
    for (int col = 0; col < 3; col++) normalize_col(g, col);      /* :194-197 */
    float *op = out_poses + (size_t)q * 16;
    if (filter_dir) { /* :199-212 */
      const float x = g[0], y = g[4], z = g[8];
      const float n = sqrtf((x * x + y * y) + z * z);
      const float dot = ((x / n) * 0.f + (y / n) * 0.f) + (z / n) * 1.f;
      if (dot < 0.f) {
        out_status[q] = 1;
        out_offset[q] = -1;
        memset(op, 0, 64);
        continue;
      }
    }
    const int n_off = adjust ? 5 : 1;
    int winner = -1, open_hit = 0;
    float cur[16];
    for (int k = 0; k < n_off; k++) { /* :253-287 */
      const float step = (k == 0) ? 0.f : ((k <= 2) ? step1 : step2);
      const float sign = (k == 0 || (k & 1)) ? 1.f : -1.f;
      float gic[16], inv[12], go[12], ge[12];
      memcpy(cur, g, 64);
      for (int r = 0; r < 3; r++) cur[r * 4 + 3] = cur[r * 4 + 3] + (step * g[r * 4 + 1]) * sign; /* :265 */
      mm4(cur, gripper_in_grasp, gic);                                                            /* :266 */
      affine_inverse(gic, inv);
      fold_grid(inv, &so, go);
      fold_grid(inv, &se, ge);
      int coll = any_hits(&so, go, sdf_mode, sdf_margin, open_pts, P1);
      open_hit = coll;
      if (!coll && P2 > 0) coll = any_hits(&se, ge, sdf_mode, sdf_margin, encl_pts, P2);
      if (!coll) { winner = k; break; }
    }
    out_status[q] = (winner >= 0) ? 0 : ((split && !adjust && !open_hit) ? 4 : 3);
    out_offset[q] = (int8_t)winner;
    if (winner >= 0) memcpy(op, cur, 64); else memset(op, 0, 64); /* :289-293 */
  }
}

void filter_ref_m(const float *nocs_pose, const float *canonical_to_nocs, const float *gripper_in_grasp,
                int filter_dir, int adjust, int sdf_mode, float sdf_margin, const float *grasp_poses, int G, const float *sym, int S,
                const float *grid_open, const int *dims_open, const float *origin_open, float res_open,
                const float *open_pts, int P1, const float *grid_encl, const int *dims_encl,
                const float *origin_encl, float res_encl, const float *encl_pts, int P2, int nthreads,
                uint8_t *out_status, int8_t *out_offset, float *out_poses) {
  filter_ref_ms(nocs_pose, canonical_to_nocs, gripper_in_grasp, filter_dir, adjust, sdf_mode, sdf_margin, 0, grasp_poses, G, sym, S,
                grid_open, dims_open, origin_open, res_open, open_pts, P1, grid_encl, dims_encl, origin_encl, res_encl,
                encl_pts, P2, nthreads, out_status, out_offset, out_poses);
}

/* point-wise lookups for the SDF parity tests: mode 0 trilinear, 1 nearest (clamped) */
void sdf_lookup_ref(const float *grid, const int *dims, const float *gc, int P, int mode, float *out) {
  sdf_view s = {grid, dims[0], dims[1], dims[2], 0.f, 0.f, 0.f, 1.f};
  for (int p = 0; p < P; p++) {
    if (mode == 0) out[p] = sdf_trilinear(&s, gc[3 * p], gc[3 * p + 1], gc[3 * p + 2]);
    else { int inb; out[p] = sdf_nearest(&s, gc[3 * p], gc[3 * p + 1], gc[3 * p + 2], 1, &inb); }
  }
}

/* margin 0: the SDF predicate itself (the form the reference build in oracle/_ref is compared with) */
void filter_ref(const float *nocs_pose, const float *canonical_to_nocs, const float *gripper_in_grasp, int filter_dir,
                int adjust, int sdf_mode, const float *grasp_poses, int G, const float *sym, int S, const float *grid_open,
                const int *dims_open, const float *origin_open, float res_open, const float *open_pts, int P1,
                const float *grid_encl, const int *dims_encl, const float *origin_encl, float res_encl,
                const float *encl_pts, int P2, int nthreads, uint8_t *out_status, int8_t *out_offset, float *out_poses) {
  filter_ref_m(nocs_pose, canonical_to_nocs, gripper_in_grasp, filter_dir, adjust, sdf_mode, 0.f, grasp_poses, G, sym, S,
               grid_open, dims_open, origin_open, res_open, open_pts, P1, grid_encl, dims_encl, origin_encl, res_encl,
               encl_pts, P2, nthreads, out_status, out_offset, out_poses);
}
