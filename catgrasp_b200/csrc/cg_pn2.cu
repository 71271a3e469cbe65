// cg_pn2.cu -- PointNet++ sampling / grouping primitives (pointnet2.py:14-149).
//
// These are HBM/L2-bound index kernels: no tensor cores.  Distances follow the
// reference's floating-point forms exactly where the result feeds a comparison:
//   FPS        : direct form  ((dx*dx + dy*dy) + dz*dz), pointnet2.py:71
//   ball query : expanded form -2*<s,d> + |s|^2 + |d|^2,  pointnet2.py:30-32
#include "cg_common.cuh"

namespace {

__device__ __forceinline__ float sq_direct(float x, float y, float z, float cx, float cy, float cz) {
  const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy), dz = __fsub_rn(z, cz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float sq_expanded(float sx, float sy, float sz, float ss, float dx, float dy, float dz) {
  // dist = -2 * (src . dst); dist += sum(src^2); dist += sum(dst^2)
  const float dot = fmaf(sz, dz, fmaf(sy, dy, __fmul_rn(sx, dx)));
  const float dd = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), ss), dd);
}

// ---------------------------------------------------------------- FPS ------
// One CTA per cloud.  Running min-distances live in shared memory; coordinates
// too when they fit (16 B/point), else they are re-read through L1/L2.
constexpr int FPS_T = 1024;

template <bool XYZ_IN_SMEM>
__global__ void __launch_bounds__(FPS_T, 1) fps_kernel(const float *__restrict__ xyz, int N, int npoint,
                                                       const int32_t *__restrict__ start_idx,
                                                       int32_t *__restrict__ out_idx) {
  extern __shared__ __align__(16) float sm[];
  float *dist = sm;                       // [N]
  float *sx = sm + N;                     // [3N] (only when XYZ_IN_SMEM)
  __shared__ float red_v[32];
  __shared__ int red_i[32];
  __shared__ int far_s;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float *P = xyz + (size_t)b * N * 3;
  for (int i = tid; i < N; i += FPS_T) dist[i] = 1e10f;    // pointnet2.py:65
  if (XYZ_IN_SMEM)
    for (int i = tid; i < 3 * N; i += FPS_T) sx[i] = P[i];
  if (tid == 0) far_s = start_idx ? start_idx[b] : 0;      // :66 (explicit instead of torch.randint)
  __syncthreads();
  const float *Q = XYZ_IN_SMEM ? sx : P;
  for (int it = 0; it < npoint; it++) {
    const int far = far_s;
    if (tid == 0) out_idx[(size_t)b * npoint + it] = far;  // :69
    const float cx = Q[3 * far], cy = Q[3 * far + 1], cz = Q[3 * far + 2];
    float best = -1.f;
    int besti = 0x7fffffff;
    for (int i = tid; i < N; i += FPS_T) {
      const float d = sq_direct(Q[3 * i], Q[3 * i + 1], Q[3 * i + 2], cx, cy, cz);  // :71
      float dm = dist[i];
      if (d < dm) { dm = d; dist[i] = d; }                 // :72-73
      if (dm > best) { best = dm; besti = i; }             // first maximum (lowest index) per thread
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if (lane == 0) { red_v[wid] = best; red_i[wid] = besti; }   // previous round's readers passed its last barrier
    __syncthreads();
    if (wid == 0) {
      best = red_v[lane];
      besti = red_i[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
      }
      if (lane == 0) far_s = besti;                        // :74 torch.max -> first max index
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------- ball query ------
// One warp per centroid; 32 points per step, ballot + prefix popcount keeps the
// reference's "nsample smallest indices" order without a sort.
constexpr int BQ_WARPS = 8;

__global__ void __launch_bounds__(BQ_WARPS * 32) ball_query_kernel(float r2, int nsample,
                                                                    const float *__restrict__ xyz,
                                                                    const float *__restrict__ new_xyz, int B, int N,
                                                                    int S, int32_t *__restrict__ out_idx) {
  const long w = (long)blockIdx.x * BQ_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= (long)B * S) return;
  const int b = (int)(w / S);
  const float *P = xyz + (size_t)b * N * 3;
  const float sx = new_xyz[w * 3], sy = new_xyz[w * 3 + 1], sz = new_xyz[w * 3 + 2];
  const float ss = __fadd_rn(__fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy)), __fmul_rn(sz, sz));
  int32_t *out = out_idx + w * nsample;
  int cnt = 0;
  int first = N;   // an empty ball leaves N everywhere (reference behaviour)
  for (int base = 0; base < N && cnt < nsample; base += 32) {
    const int i = base + lane;
    bool in = false;
    if (i < N) {
      const float d = sq_expanded(sx, sy, sz, ss, P[3 * i], P[3 * i + 1], P[3 * i + 2]);
      in = !(d > r2);                                      // pointnet2.py:93
    }
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (m) {
      if (cnt == 0) first = base + __ffs(m) - 1;
      const int slot = cnt + __popc(m & ((1u << lane) - 1u));
      if (in && slot < nsample) out[slot] = i;
      cnt += __popc(m);
    }
  }
  if (cnt > nsample) cnt = nsample;
  for (int s = cnt + lane; s < nsample; s += 32) out[s] = first;   // :95-97
}

// ------------------------------------------------------ dense helpers ------
__global__ void square_distance_kernel(const float *__restrict__ src, const float *__restrict__ dst, int B, int S,
                                       int N, float *__restrict__ out) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)B * S * N;
  if (t >= total) return;
  const int n = (int)(t % N);
  const long bs = t / N;
  const int b = (int)(bs / S);
  const float *s = src + bs * 3;
  const float *d = dst + ((size_t)b * N + n) * 3;
  const float ss = __fadd_rn(__fadd_rn(__fmul_rn(s[0], s[0]), __fmul_rn(s[1], s[1])), __fmul_rn(s[2], s[2]));
  out[t] = sq_expanded(s[0], s[1], s[2], ss, d[0], d[1], d[2]);
}

__global__ void index_points_kernel(const float *__restrict__ points, const int32_t *__restrict__ idx, int B, int N,
                                    int C, int S, float *__restrict__ out) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)B * S * C;
  if (t >= total) return;
  const int c = (int)(t % C);
  const long bs = t / C;
  const int b = (int)(bs / S);
  const int id = idx[bs];
  out[t] = (id >= 0 && id < N) ? points[((size_t)b * N + id) * C + c] : 0.f;
}

__global__ void group_points_kernel(const float *__restrict__ xyz, const float *__restrict__ points,
                                    const float *__restrict__ new_xyz, const int32_t *__restrict__ idx, int B, int N,
                                    int D, int S, int K, float *__restrict__ out) {
  const int Cc = 3 + D;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)B * S * K * Cc;
  if (t >= total) return;
  const int c = (int)(t % Cc);
  const long bsk = t / Cc;
  const long bs = bsk / K;
  const int b = (int)(bs / S);
  const int id = idx[bsk];
  float v = 0.f;
  if (id >= 0 && id < N) {
    if (c < 3) v = __fsub_rn(xyz[((size_t)b * N + id) * 3 + c], new_xyz[bs * 3 + c]);   // pointnet2.py:119
    else v = points[((size_t)b * N + id) * D + (c - 3)];                                 // :122-123
  }
  out[t] = v;
}

}  // namespace

extern "C" int cg_fps_dev(cg_ctx *ctx, const float *xyz, int B, int N, int npoint, const int32_t *start_idx,
                          int32_t *out_idx) {
  if (!ctx) return CG_EINVAL;
  CG_REQUIRE(ctx, xyz && out_idx && B > 0 && N > 0 && npoint > 0, "fps: bad arguments");
  CG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t full = (size_t)N * 16, dist_only = (size_t)N * 4;
  const size_t cap = 220 * 1024;
  CG_REQUIRE(ctx, dist_only <= cap, "fps: N too large for the shared-memory distance array (max 56320)");
  static bool attr_set[CG_MAX_DEVICES] = {};   // the attribute is per device
  if (!attr_set[ctx->device]) {
    CG_CUDA(ctx, cudaFuncSetAttribute(fps_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
    CG_CUDA(ctx, cudaFuncSetAttribute(fps_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
    attr_set[ctx->device] = true;
  }
  if (full <= cap)
    fps_kernel<true><<<B, FPS_T, full, ctx->stream>>>(xyz, N, npoint, start_idx, out_idx);
  else
    fps_kernel<false><<<B, FPS_T, dist_only, ctx->stream>>>(xyz, N, npoint, start_idx, out_idx);
  CG_LAUNCH_CHECK(ctx);
  return CG_OK;
}

extern "C" int cg_ball_query_dev(cg_ctx *ctx, float radius2, int nsample, const float *xyz, const float *new_xyz,
                                 int B, int N, int S, int32_t *out_idx) {
  if (!ctx) return CG_EINVAL;
  CG_REQUIRE(ctx, xyz && new_xyz && out_idx && B > 0 && N > 0 && S > 0 && nsample > 0, "ball_query: bad arguments");
  const long warps = (long)B * S;
  ball_query_kernel<<<(unsigned)((warps + BQ_WARPS - 1) / BQ_WARPS), BQ_WARPS * 32, 0, ctx->stream>>>(
      radius2, nsample, xyz, new_xyz, B, N, S, out_idx);
  CG_LAUNCH_CHECK(ctx);
  return CG_OK;
}

extern "C" int cg_square_distance_dev(cg_ctx *ctx, const float *src, const float *dst, int B, int S, int N,
                                      float *out) {
  if (!ctx) return CG_EINVAL;
  CG_REQUIRE(ctx, src && dst && out && B > 0 && S > 0 && N > 0, "square_distance: bad arguments");
  const long total = (long)B * S * N;
  square_distance_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(src, dst, B, S, N, out);
  CG_LAUNCH_CHECK(ctx);
  return CG_OK;
}

extern "C" int cg_index_points_dev(cg_ctx *ctx, const float *points, const int32_t *idx, int B, int N, int C, int S,
                                   float *out) {
  if (!ctx) return CG_EINVAL;
  CG_REQUIRE(ctx, points && idx && out && B > 0 && N > 0 && C > 0 && S > 0, "index_points: bad arguments");
  const long total = (long)B * S * C;
  index_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(points, idx, B, N, C, S, out);
  CG_LAUNCH_CHECK(ctx);
  return CG_OK;
}

extern "C" int cg_group_points_dev(cg_ctx *ctx, const float *xyz, const float *points, const float *new_xyz,
                                   const int32_t *idx, int B, int N, int D, int S, int K, float *out) {
  if (!ctx) return CG_EINVAL;
  CG_REQUIRE(ctx, xyz && new_xyz && idx && out && B > 0 && N > 0 && S > 0 && K > 0 && D >= 0, "group_points: bad arguments");
  CG_REQUIRE(ctx, D == 0 || points, "group_points: points required when D > 0");
  const long total = (long)B * S * K * (3 + D);
  group_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(xyz, points, new_xyz, idx, B, N, D, S, K,
                                                                           out);
  CG_LAUNCH_CHECK(ctx);
  return CG_OK;
}
