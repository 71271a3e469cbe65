// cg_trunk_tc.cu -- tcgen05 "trunk" kernel (engine 1): same fused chain as cg_trunk_simt.cu, with the
// 128 -> 1024 shared-MLP layer (91.5 % of the path's FLOPs) on the 5th-gen tensor cores.
//
// Contraction per 128-point tile and 128-channel chunk:   D[ch][pt] = sum_k W3[ch][k] * X3[pt][k]
//   A = W3 chunk  (M = 128 channels, K-major)   B = X3 tile (N = 128 points, K-major)   K = 128
//   D lives in TMEM: lane = channel, column = point  -> the max over points is a per-thread reduction
//   over TMEM columns (tcgen05.ld 32x32b), no cross-lane shuffles, and the N x 1024 activation never
//   leaves the SM.
//
// Precision: scores must stay within 1e-4 of the fp32 reference, which rules out one
// bf16 pass (and leaves single-pass TF32 marginal).  Operands are therefore split x = hi + lo with
// hi = bf16(x), lo = bf16(x - hi) and every product is accumulated as  hi*hi + hi*lo + lo*hi  in the fp32
// TMEM accumulator (kind::f16, three UMMAs per K-step); the dropped lo*lo term is ~2^-16 relative.
//
// Shared-memory operand layout = the canonical UMMA K-major SWIZZLE_128B layout: a K-block of 64 bf16 is
// one 128-byte row per M/N index, rows in 8-row / 1024-byte swizzle atoms, 16-byte chunk index XOR (row & 7).
// W3 is pre-arranged in exactly this image on the host (cg_tc_prepare_w3), so a chunk arrives with plain
// 1-D bulk copies (cp.async.bulk -> UBLKCP) completing on an mbarrier; the X3 tile is written in the same
// layout by the epilogue of the 64 -> 128 layer.
#include <cuda_bf16.h>

#include "cg_trunk_common.cuh"

namespace {
using namespace cg_trunk;

constexpr uint32_t PIECE = 16384;        // [128 rows x 64 bf16] one swizzled K-block
constexpr uint32_t OPND = 4 * PIECE;     // {hi,lo} x {kb0,kb1} = 64 KB: one full K=128 operand tile
constexpr uint32_t X3_OFF = 0;
constexpr uint32_t RING_OFF = OPND;      // two 64 KB stages of W3 chunks
constexpr uint32_t MISC_OFF = 3 * OPND;  // 192 KB
constexpr int NCHUNK = 8;                // 1024 output channels / 128
constexpr uint32_t TMEM_COLS = 256;      // two 128-column fp32 accumulators

// scratch of the SIMT front layers; aliases the W3 ring, which is idle while they run
struct Scratch {
  float in_s[8 * TP];     //  4 KB
  float regA[64 * TP];    // 32 KB
  float regB[64 * TP];    // 32 KB
  float w1s[64 * 64];     // 16 KB
  float w2s[64 * 128];    // 32 KB
};
static_assert(sizeof(Scratch) <= 2 * OPND, "front-layer scratch must fit in the W3 ring");

struct Misc {
  uint32_t gmax_s[1024];
  float w0[6 * 64];
  float bias0[64];
  float bias1[64];
  float bias2[128];
  double pinv[12];
  double mean[6];
  double sden[6];
  float T3[9];
  uint32_t tmem_base;
  unsigned long long full_bar[2];   // W3 chunk landed in ring stage s
  unsigned long long done_bar[2];   // UMMAs of the chunk using stage s / accumulator s have completed
};

constexpr size_t SMEM_BYTES = MISC_OFF + sizeof(Misc) + 1024;  // + slack for manual 1024-byte alignment

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: start address (>>4), LBO = 1 (ignored for swizzled
// K-major), SBO = 1024 B between 8-row groups, version = 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread (thread t <-> lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// byte offset of element (row, k) of a K=128 bf16 operand tile, part 0 = hi, 1 = lo
__host__ __device__ __forceinline__ uint32_t opnd_off(int part, int row, int k) {
  const int kb = k >> 6, kk = k & 63;
  return (uint32_t)part * (2 * PIECE) + (uint32_t)kb * PIECE + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u +
         (uint32_t)(((kk >> 3) ^ (row & 7)) << 4) + (uint32_t)(kk & 7) * 2u;
}

// 64 -> 128 layer (+bias, ReLU) whose output is written as the bf16 hi/lo UMMA operand tile X3[pt][ch]
__device__ __forceinline__ void mlp_layer_to_umma(const float *__restrict__ hin, const float *__restrict__ w,
                                                  const float *__restrict__ bias, unsigned char *__restrict__ x3,
                                                  int tx, int ty) {
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
  const int p0 = ty * 4, p1 = 64 + ty * 4;
  const int c0 = tx * 4, c1 = 64 + tx * 4;
#pragma unroll 4
  for (int k = 0; k < 64; k++) {
    float a[8], b[8];
    *reinterpret_cast<float4 *>(&a[0]) = *reinterpret_cast<const float4 *>(&hin[k * TP + p0]);
    *reinterpret_cast<float4 *>(&a[4]) = *reinterpret_cast<const float4 *>(&hin[k * TP + p1]);
    *reinterpret_cast<float4 *>(&b[0]) = *reinterpret_cast<const float4 *>(&w[k * 128 + c0]);
    *reinterpret_cast<float4 *>(&b[4]) = *reinterpret_cast<const float4 *>(&w[k * 128 + c1]);
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int p = (i < 4) ? (p0 + i) : (p1 + i - 4);
#pragma unroll
    for (int q = 0; q < 2; q++) {   // channel quad: c0.. (K-block 0) / c1.. (K-block 1)
      const int c = q ? c1 : c0;
      unsigned short hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float v = fmaxf(acc[i][q * 4 + j] + bias[c + j], 0.f);
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        hi[j] = __bfloat16_as_ushort(h);
        lo[j] = __bfloat16_as_ushort(l);
      }
      const uint32_t off = opnd_off(0, p, c);
      *reinterpret_cast<uint2 *>(x3 + off) = make_uint2(hi[0] | ((uint32_t)hi[1] << 16), hi[2] | ((uint32_t)hi[3] << 16));
      *reinterpret_cast<uint2 *>(x3 + off + 2 * PIECE) =
          make_uint2(lo[0] | ((uint32_t)lo[1] << 16), lo[2] | ((uint32_t)lo[3] << 16));
    }
  }
}

__global__ void __launch_bounds__(NT, 1) trunk_tc_kernel(const cg_trunk_args a, int tiles_per_cta) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char *x3 = smem + X3_OFF;
  unsigned char *ring = smem + RING_OFF;
  Scratch &F = *reinterpret_cast<Scratch *>(ring);
  Misc &S = *reinterpret_cast<Misc *>(smem + MISC_OFF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.y;
  const int N = a.N;
  const int ntiles = (N + TP - 1) / TP;
  const int tile_begin = blockIdx.x * tiles_per_cta;
  const int tile_end = min(ntiles, tile_begin + tiles_per_cta);
  if (tile_begin >= tile_end) return;

  // ---- one-time setup: constants, mbarriers, TMEM ---------------------------
  for (int i = tid; i < 1024; i += NT) S.gmax_s[i] = 0u;
  for (int i = tid; i < 6 * 64; i += NT) S.w0[i] = a.l0.Wt[i];
  if (tid < 64) {
    S.bias0[tid] = a.l0.b[tid];
    S.bias1[tid] = (a.stage1_mode == 1) ? a.l1.b[tid] : 0.f;
  }
  if (tid < 128) S.bias2[tid] = a.l2.b[tid];
  if (tid < 9) S.T3[tid] = a.T3 ? a.T3[b * 9 + tid] : 0.f;
  if (a.in.x_direct == nullptr) {
    if (tid == 0) pose_inverse(a.in.poses + (size_t)b * 16, S.pinv);
    if (tid < 6) {
      S.mean[tid] = a.in.mean ? a.in.mean[tid] : 0.0;
      S.sden[tid] = a.in.stdv ? (a.in.stdv[tid] + 1e-15) : 1.0;
    }
  }
  if (tid == 0) {
    mbar_init(smem_u32(&S.full_bar[0]), 1);
    mbar_init(smem_u32(&S.full_bar[1]), 1);
    mbar_init(smem_u32(&S.done_bar[0]), 1);
    mbar_init(smem_u32(&S.done_bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_base;
  const uint32_t full_bar[2] = {smem_u32(&S.full_bar[0]), smem_u32(&S.full_bar[1])};
  const uint32_t done_bar[2] = {smem_u32(&S.done_bar[0]), smem_u32(&S.done_bar[1])};
  uint32_t full_ph[2] = {0u, 0u}, done_ph[2] = {0u, 0u};
  const unsigned char *w3img = static_cast<const unsigned char *>(a.l3_tc);
  const uint32_t x3_s = smem_u32(x3), ring_s = smem_u32(ring);

  for (int tile = tile_begin; tile < tile_end; tile++) {
    // ================= front layers (fp32 SIMT), scratch aliases the idle W3 ring =================
    if (a.stage1_mode != 0) {
      const float *src = (a.stage1_mode == 1) ? a.l1.Wt : (a.T64 + (size_t)b * 4096);
      for (int e = tid * 4; e < 4096; e += NT * 4) cp_async16(F.w1s + e, src + e);
    }
    for (int e = tid * 4; e < 8192; e += NT * 4) cp_async16(F.w2s + e, a.l2.Wt + e);
    cp_async_commit();
    if (tid < TP) {
      int n = tile * TP + tid;
      if (n >= N) n = N - 1;
      float v[6];
      if (a.in.x_direct) {
        const float *xr = a.in.x_direct + ((size_t)b * N + n) * 6;
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] = xr[k];
      } else {
        const int id = a.in.ids ? a.in.ids[(size_t)b * N + n] : n;
        const double *px = a.in.cloud_xyz + (size_t)id * 3;
        const double *pn = a.in.cloud_nrm + (size_t)id * 3;
        const double x = px[0], y = px[1], z = px[2];
        const double nx = pn[0], ny = pn[1], nz = pn[2];
        const double *R = S.pinv;
        double w[6];
        w[0] = R[0] * x + R[1] * y + R[2] * z + R[9];
        w[1] = R[3] * x + R[4] * y + R[5] * z + R[10];
        w[2] = R[6] * x + R[7] * y + R[8] * z + R[11];
        w[3] = R[0] * nx + R[1] * ny + R[2] * nz;
        w[4] = R[3] * nx + R[4] * ny + R[5] * nz;
        w[5] = R[6] * nx + R[7] * ny + R[8] * nz;
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] = (float)((w[k] - S.mean[k]) / S.sden[k]);
      }
      if (a.T3) {
        const float x = v[0], y = v[1], z = v[2];
        v[0] = fmaf(z, S.T3[6], fmaf(y, S.T3[3], x * S.T3[0]));
        v[1] = fmaf(z, S.T3[7], fmaf(y, S.T3[4], x * S.T3[1]));
        v[2] = fmaf(z, S.T3[8], fmaf(y, S.T3[5], x * S.T3[2]));
      }
#pragma unroll
      for (int k = 0; k < 6; k++) F.in_s[k * TP + tid] = v[k];
    }
    __syncthreads();
    float *h0 = (a.stage1_mode != 0) ? F.regA : F.regB;
    mlp_layer<6, 64, 4, true, true>(F.in_s, S.w0, S.bias0, h0, tx, ty);
    cp_async_wait<0>();
    __syncthreads();
    if (a.stage1_mode == 1) {
      mlp_layer<64, 64, 4, true, true>(F.regA, F.w1s, S.bias1, F.regB, tx, ty);
      __syncthreads();
    } else if (a.stage1_mode == 2) {
      mlp_layer<64, 64, 4, false, false>(F.regA, F.w1s, S.bias1, F.regB, tx, ty);
      __syncthreads();
    }
    if (a.pf_out) {
      const int p = tid & (TP - 1);
      const int n = tile * TP + p;
      if (n < N) {
        float *dst = a.pf_out + ((size_t)b * N + n) * 64;
        for (int c = (tid >> 7) * 4; c < 64; c += 8) {
          float4 o = make_float4(F.regB[(c + 0) * TP + p], F.regB[(c + 1) * TP + p], F.regB[(c + 2) * TP + p],
                                 F.regB[(c + 3) * TP + p]);
          *reinterpret_cast<float4 *>(dst + c) = o;
        }
      }
    }
    mlp_layer_to_umma(F.regB, F.w2s, S.bias2, x3, tx, ty);
    // generic-proxy writes (X3 tile, scratch) -> visible to / ordered before the async proxy (bulk copies, UMMA)
    fence_proxy_async();
    __syncthreads();

    // ================= 128 -> 1024 on tcgen05, fused bias/ReLU/max epilogue =================
    if (tid == 0) {
      mbar_expect_tx(full_bar[0], OPND);
#pragma unroll
      for (int pc = 0; pc < 4; pc++) bulk_g2s(ring_s + pc * PIECE, w3img + (size_t)pc * PIECE, PIECE, full_bar[0]);
    }
    for (int c = 0; c <= NCHUNK; c++) {
      const int s = c & 1;
      if (c < NCHUNK && tid == 0) {
        mbar_wait(full_bar[s], full_ph[s]);
        full_ph[s] ^= 1u;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)s * 128u;
        const uint32_t a_base = ring_s + (uint32_t)s * OPND;
        uint32_t acc = 0u;
#pragma unroll
        for (int kb = 0; kb < 2; kb++) {
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            const uint32_t koff = (uint32_t)kb * PIECE + (uint32_t)ks * 32u;
            const uint64_t a_hi = umma_desc(a_base + koff), a_lo = umma_desc(a_base + 2 * PIECE + koff);
            const uint64_t b_hi = umma_desc(x3_s + koff), b_lo = umma_desc(x3_s + 2 * PIECE + koff);
            umma_bf16(d_tmem, a_lo, b_hi, acc);   // small terms first
            umma_bf16(d_tmem, a_hi, b_lo, 1u);
            umma_bf16(d_tmem, a_hi, b_hi, 1u);
            acc = 1u;
          }
        }
        umma_commit(done_bar[s]);
      }
      __syncwarp();
      if (c >= 1) {
        const int pb = s ^ 1, pc = c - 1;           // accumulator / chunk whose UMMAs were issued last round
        mbar_wait(done_bar[pb], done_ph[pb]);
        done_ph[pb] ^= 1u;
        tc_fence_after();
        const int q = warp & 3, half = warp >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)pb * 128u + (uint32_t)half * 64u;
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 2; j++) {
          float v[32];
          tmem_ld32(taddr + (uint32_t)j * 32u, v);
#pragma unroll
          for (int i = 0; i < 32; i++) m = fmaxf(m, v[i]);
        }
        const int ch = pc * 128 + q * 32 + lane;
        m += __ldg(&a.l3.b[ch]);
        if (a.relu3) m = fmaxf(m, 0.f);
        atomicMax(&S.gmax_s[ch], cg_f2key(m));
        tc_fence_before();
      }
      __syncthreads();   // accumulator pb drained, ring stage pb free
      if (tid == 0 && c + 1 < NCHUNK) {
        const int ns = s ^ 1;
        mbar_expect_tx(full_bar[ns], OPND);
        const unsigned char *src = w3img + (size_t)(c + 1) * OPND;
#pragma unroll
        for (int pc = 0; pc < 4; pc++)
          bulk_g2s(ring_s + (uint32_t)ns * OPND + pc * PIECE, src + (size_t)pc * PIECE, PIECE, full_bar[ns]);
      }
    }
    // all UMMAs of this tile have completed and been consumed: ring and X3 are free again
  }

  for (int i = tid; i < 1024; i += NT) atomicMax(&a.gmax_keys[(size_t)b * 1024 + i], S.gmax_s[i]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

unsigned short bf16_rne(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t lsb = (x >> 16) & 1u;
  x += 0x7fffu + lsb;
  return (unsigned short)(x >> 16);
}
float bf16_to_f(unsigned short h) {
  uint32_t x = (uint32_t)h << 16;
  float f;
  memcpy(&f, &x, 4);
  return f;
}

}  // namespace

size_t cg_tc_w3_bytes() { return (size_t)NCHUNK * OPND; }

// Wt_host: [128][1024] folded fp32 (k-major rows).  Device image: per 128-channel chunk one 64 KB operand tile
// [hi|lo][kb][128 rows x 128 B swizzled].
int cg_tc_prepare_w3(cg_ctx *ctx, const float *Wt_host, void *dst_dev) {
  std::vector<unsigned char> img(cg_tc_w3_bytes());
  for (int ch = 0; ch < NCHUNK; ch++)
    for (int r = 0; r < 128; r++)
      for (int k = 0; k < 128; k++) {
        const float w = Wt_host[(size_t)k * 1024 + ch * 128 + r];
        const unsigned short hi = bf16_rne(w);
        const unsigned short lo = bf16_rne(w - bf16_to_f(hi));
        unsigned char *base = img.data() + (size_t)ch * OPND;
        memcpy(base + opnd_off(0, r, k), &hi, 2);
        memcpy(base + opnd_off(1, r, k), &lo, 2);
      }
  CG_CUDA(ctx, cudaMemcpyAsync(dst_dev, img.data(), img.size(), cudaMemcpyHostToDevice, ctx->stream));
  CG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // img goes out of scope
  return CG_OK;
}

int cg_trunk_launch_tc(cg_ctx *ctx, const cg_trunk_args &a) {
  CG_REQUIRE(ctx, a.B > 0 && a.N > 0, "trunk: B,N must be positive");
  CG_REQUIRE(ctx, a.B <= 65535, "trunk: B > 65535 must be chunked by the caller");
  CG_REQUIRE(ctx, a.l3_tc != nullptr, "trunk: tensor-core weight image missing");
  static bool attr_set = false;
  if (!attr_set) {
    CG_CUDA(ctx, cudaFuncSetAttribute(trunk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set = true;
  }
  const int ntiles = (a.N + TP - 1) / TP;
  int splits = 1;
  while ((long)a.B * splits < 4L * ctx->num_sms && splits < ntiles) splits *= 2;
  const int tiles_per_cta = (ntiles + splits - 1) / splits;
  dim3 grid((ntiles + tiles_per_cta - 1) / tiles_per_cta, a.B);
  trunk_tc_kernel<<<grid, NT, SMEM_BYTES, ctx->stream>>>(a, tiles_per_cta);
  CG_LAUNCH_CHECK(ctx);
  return CG_OK;
}
