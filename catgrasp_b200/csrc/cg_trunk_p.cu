// cg_trunk_p.cu -- persistent tcgen05 trunk kernel (engine 3): one CTA per SM loops over (candidate, 128-point tile)
// work items, so barriers, TMEM, the resident weight tiles and the W3 stream are set up once per SM instead of once
// per candidate, and the 128->1024 layer (91.5 % of the FLOPs, /root/reference/pointnet2.py:156-161,264-265) runs as
// ONE fp16 x fp16 tensor-core pass with fp32 accumulation.
//
//   layer            UMMA (cta_group::1, kind::f16, fp32 accumulators in TMEM)                       epilogue
//   6 -> 64          fp32 FMA (K = 6 is not a tensor-core shape; thread = point)                      -> X1 | X2 tile
//   64 -> 64 (L1)    D1[pt][ch] = X1[pt][k] . W1[ch][k]   bf16 hi/lo x3   M=128 N=64  K=64           bias/ReLU -> X2
//   64 -> 128 (L2)   D2[pt][ch] = X2[pt][k] . W2[ch][k]   bf16 hi/lo x3   M=128 N=128 K=64           bias/ReLU -> X3
//   128 -> 1024 (L3) D3[pt][ch] = X3[pt][k] (TMEM, fp16) . W3[ch][k] (smem, fp16)   x8 chunks        max over points
//
// Precision: the folded W3 and the post-ReLU X3 are rounded once to fp16 (11-bit mantissa); the products are exact in
// the fp32 accumulator.  CPU emulation of exactly this rounding on the reference model (scripts/emulate_engine3.py):
// max |dprob| 4e-6, max |dlogit| 3e-5 against the fp32 oracle (tolerance 1e-4).  X3 values above the fp16 range are
// clamped to 65504 AND reported through cg_trunk_args::ovf_flag so the host can fall back to engine 1.
//
// Warp roles (608 threads, 1 CTA / SM):
//   warps 0-7   front : thread = (point, channel half): input build + 6->64 FMA layer, L1 / L2 epilogues
//   warps 8-15  max   : L3 epilogue: 16x256b TMEM loads + FMNMX3 / shuffle column max over the tile's 128 points;
//                       warp = (TMEM lane quarter, column half) -- one warp alone reads TMEM at only ~31 B/cycle
//   warp  16    W3 producer: streams the fp16 W3 image through a 3 x 32 KB ring with cp.async.bulk (UBLKCP)
//   warp  17    UMMA issuer (one elected lane)
//   warp  18    aux: per-candidate constants (float64 pose inverse, T3, and the per-candidate 64x64 feature transform
//               of pointnet2.py:257 converted into a UMMA B-operand image), double-buffered one candidate ahead
// All hand-overs are mbarriers; the front layers of tile t+1 run in the shadow of tile t's L3 stream.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "cg_tc_ptx.cuh"
#include "cg_trunk_common.cuh"

namespace {
using namespace cg_trunk;
using namespace cg_ptx;

constexpr int NFRONT = 8;
constexpr int NMAXW = 8;
constexpr int PROD_WARP = NFRONT + NMAXW;   // 16
constexpr int MMA_WARP = PROD_WARP + 1;     // 17: issuer A (even L3 chunks)
constexpr int MMAB_WARP = MMA_WARP + 1;     // 18: issuer B (odd L3 chunks + the L1 / L2 UMMAs of the next tile)
constexpr int AUX_WARP = MMAB_WARP + 1;     // 19
constexpr int NTP = (AUX_WARP + 1) * 32;    // 640 threads
constexpr int NFT = NFRONT * 32;            // 256 front threads
constexpr uint32_t PIECE = 16384;           // [128 rows x 64 x 16-bit] one swizzled K-block
constexpr uint32_t XA_OFF = 0;              // [hi|lo] 32 KB: X1 (L1 input), then X2 (L2 input) of the same tile
constexpr uint32_t W1_OFF = 2 * PIECE;      // 2 slots x [hi 8 KB | lo 8 KB]: shared W1 (slot 0) or per-candidate T64
constexpr uint32_t W2_OFF = 4 * PIECE;      // [hi|lo][128 rows x 128 B] 32 KB
constexpr uint32_t RING_OFF = 6 * PIECE;    // NPAIR x 32 KB: one 128-channel chunk of W3 (fp16) = [kb0 | kb1]
constexpr int NPAIR = 3;
constexpr uint32_t MISC_OFF = (6 + 2 * NPAIR) * PIECE;   // 192 KB
constexpr int NCHUNK = 8;                   // 1024 output channels / 128
constexpr uint32_t TMEM_COLS = 512;         // D3 x2 at 0 / 128; activation blocks XB(it) at 256 + (it & 1) * 128
// offsets inside the operand image built by cg_tc_prepare (cg_trunk_tc.cu)
constexpr uint32_t IMG_W3 = NCHUNK * 4 * PIECE, IMG_W2 = 2 * PIECE, IMG_W1 = PIECE;
constexpr uint32_t IMG_W3H_OFF = IMG_W3 + IMG_W2 + IMG_W1;

struct CandConst {
  double pinv[12];
  float T3[12];
};

struct MiscP {
  float2 sacc[NCHUNK][NMAXW][32];   // running max of the current candidate: a private slot per (chunk, max warp, lane)
  float w0[6 * 64];
  float bias0[64];
  float bias1[64];
  float bias2[128];
  double mean[6];
  double sden[6];
  CandConst cc[2];
  unsigned long long full_bar[NPAIR];   // producer -> MMA : W3 chunk landed in ring slot
  unsigned long long free_bar[NPAIR];   // MMA -> producer : UMMAs reading the slot have completed
  unsigned long long acc_bar[2];        // MMA -> max      : chunk accumulated into D3[buf]
  unsigned long long accfree_bar[2];    // max -> MMA      : D3[buf] copied to registers (one arrival per max warp)
  unsigned long long x1_bar, x2_bar, x3_bar;   // front -> MMA : XA holds X1 / XA holds X2 / X3 written to TMEM
  unsigned long long l1_bar, l2_bar;    // MMA -> front    : D1 / D2 complete
  unsigned long long w_bar;             // resident W2 (+ shared W1) landed
  unsigned long long a_done[2];         // issuer A -> issuer B: A's L3 UMMAs of tile it (parity it & 1) completed
  unsigned long long cc_full[2], cc_free[2];   // aux <-> front : per-candidate constants
  unsigned long long w1_full[2], w1_free[2];   // aux <-> MMA   : per-candidate L1 operand (stage1_mode 2)
  uint32_t tmem_base;
};

#define SBAR(field) (misc_s + (uint32_t)offsetof(MiscP, field))
#define SBARI(field, i) (misc_s + (uint32_t)offsetof(MiscP, field) + (uint32_t)(i) * 8u)

#ifdef CG_EXPERIMENTS
#define CG_EXP(a, bit) (((a).exp_flags & (bit)) != 0)
#else
#define CG_EXP(a, bit) false
#endif

// per-role cycle counters exist only in developer builds: a release build reads no clocks in the hot loops
#ifdef CG_EXPERIMENTS
#define CG_CLK() clock64()
#define CG_DBG(a) ((a).dbg)
#else
#define CG_CLK() 0ll
#define CG_DBG(a) ((unsigned long long *)nullptr)
#endif

// developer timeline: clock64 stamps of one steady-state tile of CTA 0 (slot ids are printed by the host side)
#ifdef CG_EXPERIMENTS
#define CG_TRACE_AT(cond, slot)                                                     \
  do {                                                                              \
    if (a.dbg && blockIdx.x == 0 && (cond)) a.dbg[16 * gridDim.x + (slot)] = (unsigned long long)clock64(); \
  } while (0)
#else
#define CG_TRACE_AT(cond, slot) do { } while (0)
#endif

constexpr size_t SMEM_BYTES_P = MISC_OFF + sizeof(MiscP) + 1024;   // + slack for manual 1024-byte alignment
static_assert(SMEM_BYTES_P <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");

__device__ __forceinline__ void bar_max() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

// split 8 fp32 values into bf16 hi / lo and store them as the two 16-byte chunks of an operand row
__device__ __forceinline__ void store_hilo8(unsigned char *hi_dst, unsigned char *lo_dst, const float *v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    const uint32_t hb = *reinterpret_cast<const uint32_t *>(&hh);
    const float2 lo = fsub2(make_float2(v[2 * j], v[2 * j + 1]), make_float2(__uint_as_float(hb << 16), __uint_as_float(hb & 0xffff0000u)));
    const __nv_bfloat162 ll = __floats2bfloat162_rn(lo.x, lo.y);
    h[j] = hb;
    l[j] = *reinterpret_cast<const uint32_t *>(&ll);
  }
  *reinterpret_cast<uint4 *>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4 *>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

// K = 64 layer: 4 K-steps x (x_lo*w_hi + x_hi*w_lo + x_hi*w_hi);  A = activations (M = 128 points), B = weights
__device__ __forceinline__ void issue_k64(uint32_t d, uint32_t x_s, uint32_t x_part, uint32_t w_s, uint32_t w_part,
                                          uint32_t id) {
  uint32_t acc = 0u;
#pragma unroll
  for (int ks = 0; ks < 4; ks++) {
    const uint32_t koff = (uint32_t)ks * 32u;
    const uint64_t a_hi = umma_desc(x_s + koff), a_lo = umma_desc(x_s + x_part + koff);
    const uint64_t b_hi = umma_desc(w_s + koff), b_lo = umma_desc(w_s + w_part + koff);
    umma_ss(d, a_lo, b_hi, id, acc);
    umma_ss(d, a_hi, b_lo, id, 1u);
    umma_ss(d, a_hi, b_hi, id, 1u);
    acc = 1u;
  }
}

__device__ __forceinline__ uint32_t xb_col(int it) { return 256u + (uint32_t)(it & 1) * 128u; }

__global__ void __launch_bounds__(NTP, 1) trunk_p_kernel(const cg_trunk_args a, int ntiles, int total_tiles) {
  // The operand tiles need 1024-byte alignment (SWIZZLE_128B atoms).  The kernel has no static shared memory, so the
  // dynamic window starts at an aligned offset; using the array directly (instead of re-aligning through an integer
  // cast) keeps every access in the shared address space: LDS / STS / ATOMS instead of generic LD / ST / ATOM.
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char *smem = smem_dyn;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  unsigned char *xa = smem + XA_OFF, *w1 = smem + W1_OFF;
  MiscP &S = *reinterpret_cast<MiscP *>(smem + MISC_OFF);
  // shared-space address of the barrier block, converted once: every barrier operand below is misc_s + a constant
  // (a generic-to-shared conversion per use showed up with 5-7 % of the stall samples)
  uint32_t smem_s = smem_u32(smem);
  asm volatile("" : "+r"(smem_s));   // opaque: otherwise the compiler re-derives it from SR_CgaCtaId (an S2R) before every use
  const uint32_t misc_s = smem_s + MISC_OFF;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N;
  // contiguous range of flattened (candidate, tile) work items of this CTA
  const int g0 = (int)(((long long)blockIdx.x * total_tiles) / gridDim.x);
  const int g1 = (int)(((long long)(blockIdx.x + 1) * total_tiles) / gridDim.x);
  const int T = g1 - g0;
  if (T <= 0) return;
  const int b_first = g0 / ntiles;
  const unsigned char *img = static_cast<const unsigned char *>(a.tc_img);
  const bool has_l1 = a.stage1_mode != 0;
  const bool percand_w1 = a.stage1_mode == 2;

  // ---- one-time setup -------------------------------------------------------------------------------------
  for (int i = tid; i < 6 * 64; i += NTP) S.w0[i] = a.l0.Wt[i];
  if (tid < 64) {
    S.bias0[tid] = a.l0.b[tid];
    S.bias1[tid] = (a.stage1_mode == 1) ? a.l1.b[tid] : 0.f;
  }
  if (tid < 128) S.bias2[tid] = a.l2.b[tid];
  if (tid < 6) {
    S.mean[tid] = a.in.mean ? a.in.mean[tid] : 0.0;
    S.sden[tid] = a.in.stdv ? 1.0 / (a.in.stdv[tid] + 1e-15) : 1.0;   // reciprocal: the hot loop multiplies
  }
  if (tid == 0) {
    for (int i = 0; i < NPAIR; i++) {
      mbar_init(SBARI(full_bar, i), 1);
      mbar_init(SBARI(free_bar, i), 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(SBARI(acc_bar, i), 1);
      mbar_init(SBARI(accfree_bar, i), NMAXW);
      mbar_init(SBARI(cc_full, i), 1);
      mbar_init(SBARI(cc_free, i), NFRONT);
      mbar_init(SBARI(w1_full, i), 1);
      mbar_init(SBARI(w1_free, i), 1);
    }
    mbar_init(SBAR(x1_bar), NFRONT);   // front -> issuer hand-overs: one arrival per front warp (no block barrier first)
    mbar_init(SBAR(x2_bar), NFRONT);
    mbar_init(SBAR(x3_bar), NFRONT);
    mbar_init(SBAR(l1_bar), 1);
    mbar_init(SBAR(l2_bar), 1);
    mbar_init(SBAR(w_bar), 1);
    mbar_init(SBARI(a_done, 0), 1);
    mbar_init(SBARI(a_done, 1), 1);
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc(SBAR(tmem_base), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_base;
  const uint32_t xa_s = smem_s + XA_OFF, w1_s = smem_s + W1_OFF, w2_s = smem_s + W2_OFF;
  const uint32_t ring_s = smem_s + RING_OFF;

  // (candidate, tile) of local work item `it`
  auto locate = [&](int it, int &b, int &tile) {
    const int g = g0 + it;
    b = g / ntiles;
    tile = g - b * ntiles;
  };

  if (warp == PROD_WARP) {
    // ======================= producer: stream W3 (fp16, one 32 KB chunk per ring slot) =======================
    const unsigned char *w3src = img + IMG_W3H_OFF;
    const int total = T * NCHUNK;
    int slot = 0;
    uint32_t ph = 1u;   // a fresh barrier passes a wait on parity 1: the first round does not block
    int chunk = 0;   // every tile starts with chunk 0
    for (int gp = 0; gp < total; gp++) {
      mbar_wait(SBARI(free_bar, slot), ph);
      if (elect_one()) {
        const uint32_t fb = SBARI(full_bar, slot);
        if (CG_EXP(a, 1) && gp >= NPAIR) {
          mbar_arrive(fb);   // timing experiment: no W3 traffic after the first round
        } else {
          mbar_expect_tx(fb, 2 * PIECE);
          const unsigned char *src = w3src + (size_t)chunk * 2 * PIECE;
          bulk_g2s(ring_s + (uint32_t)slot * 2 * PIECE, src, PIECE, fb);
          bulk_g2s(ring_s + (uint32_t)slot * 2 * PIECE + PIECE, src + PIECE, PIECE, fb);
        }
      }
      __syncwarp();
      chunk = (chunk + 1) & (NCHUNK - 1);
      if (++slot == NPAIR) { slot = 0; ph ^= 1u; }
    }
  } else if (warp == AUX_WARP) {
    // ======================= aux: per-candidate constants, one candidate ahead =======================
    const int b_last = (g1 - 1) / ntiles;
    for (int b = b_first; b <= b_last; b++) {
      const int lc = b - b_first, slot = lc & 1;
      const uint32_t ph = (((uint32_t)lc >> 1) & 1u) ^ 1u;   // first use of each slot passes immediately
      mbar_wait(SBARI(cc_free, slot), ph);
      if (lane == 0 && a.in.x_direct == nullptr) pose_inverse(a.in.poses + (size_t)b * 16, S.cc[slot].pinv);
      if (lane < 9) S.cc[slot].T3[lane] = a.T3 ? a.T3[(size_t)b * 9 + lane] : 0.f;
      __syncwarp();
      if (lane == 0) mbar_arrive(SBARI(cc_full, slot));
      if (percand_w1) {
        // per-candidate feature transform as the B operand of L1:  B[j][k] = T64[k][j]   (pointnet2.py:257)
        mbar_wait(SBARI(w1_free, slot), ph);
        const float *Tm = a.T64 + (size_t)b * 4096;
        unsigned char *dst = w1 + (size_t)slot * PIECE;
#pragma unroll 4
        for (int idx = lane; idx < 4096; idx += 32) {
          const int k = idx >> 6, j = idx & 63;
          const float v = __ldg(Tm + idx);
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
          const uint32_t off = row_chunk_off(j, k >> 3) + (uint32_t)(k & 7) * 2u;
          *reinterpret_cast<__nv_bfloat16 *>(dst + off) = h;
          *reinterpret_cast<__nv_bfloat16 *>(dst + 8192 + off) = l;
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the async proxy (UMMA operand reads)
        __syncwarp();
        if (lane == 0) mbar_arrive(SBARI(w1_full, slot));
      }
    }
  } else if (warp == MMA_WARP || warp == MMAB_WARP) {
    // ======================= UMMA issuers A / B: the warps stay converged, one elected lane issues =======================
    // Measured: tcgen05.mma issue is nearly synchronous with the tensor pipe (~75 cycles per
    // 128x128x16 UMMA in the issuing thread) and each chunk costs the issuer another ~400 cycles of waits / fences / commits
    // during which a single issuer leaves the pipe empty.  Two issuers alternate over the chunks (A: even chunks ->
    // accumulator 0, B: odd chunks -> accumulator 1): while one is blocked issuing, the other does its hand-overs.  B also
    // issues the L1 / L2 UMMAs of the next tile at the fixed slots that measured best with one issuer (after chunks 1 and 3;
    // later slots 767 vs 826-850 TFLOP/s, run-time placement by probes 789, a separate issuer for L1 / L2 only 768).
    const bool isB = warp == MMAB_WARP;
    const uint32_t wb = SBAR(w_bar);
    const uint32_t l1b = SBAR(l1_bar), l2b = SBAR(l2_bar);
    uint32_t ph_x1 = 0u, ph_x2 = 0u;
    int b_prev = -1;
    long long t_all = CG_CLK(), t_x3 = 0, t_ring = 0, t_accf = 0, t_x12 = 0, tw;
    // front layers (L1, L2) of local tile `itn` (issuer B only)
    // issue_l1 is called with consecutive local tiles: (candidate, tile) is a cursor, not a division per call
    int il_b, il_tile;
    locate(0, il_b, il_tile);
    auto issue_l1 = [&](int itn) {
      const int b = il_b;
      const bool last_of_cand = (itn == T - 1) || (il_tile == ntiles - 1);
      if (++il_tile == ntiles) { il_tile = 0; il_b++; }
      const int lc = b - b_first, slot = percand_w1 ? (lc & 1) : 0;
      const bool new_cand = b != b_prev;
      b_prev = b;
      if (percand_w1 && new_cand) mbar_wait(SBARI(w1_full, slot), ((uint32_t)lc >> 1) & 1u);
      mbar_wait(SBAR(x1_bar), ph_x1);
      ph_x1 ^= 1u;
      tc_fence_after();
      if (elect_one()) {
        if (!CG_EXP(a, 8))
          issue_k64(tmem_base + xb_col(itn), xa_s, PIECE, w1_s + (uint32_t)slot * PIECE, 8192u, umma_idesc(128, 64));
        umma_commit(l1b);
        if (percand_w1 && last_of_cand) umma_commit(SBARI(w1_free, slot));
      }
      __syncwarp();
    };
    auto issue_l2 = [&](int itn) {
      mbar_wait(SBAR(x2_bar), ph_x2);
      ph_x2 ^= 1u;
      tc_fence_after();
      if (elect_one()) {
        if (!CG_EXP(a, 8)) issue_k64(tmem_base + xb_col(itn), xa_s, PIECE, w2_s, PIECE, umma_idesc(128, 128));
        umma_commit(l2b);
      }
      __syncwarp();
    };
    // D1 / D2 of tile itn land in the activation block XB(itn) = XB(itn - 2), whose X3 issuer A's chunks of tile itn - 2
    // read: they must have completed (B's own chunks are ordered before by issue order)
    auto guard_xb = [&](int itn) {
      if (itn >= 2) mbar_wait(SBARI(a_done, itn & 1), (((uint32_t)itn >> 1) - 1u) & 1u);
    };
    if (isB) {
      if (elect_one()) {
        mbar_expect_tx(wb, IMG_W2 + (a.stage1_mode == 1 ? IMG_W1 : 0u));
        bulk_g2s(w2_s, img + IMG_W3, PIECE, wb);
        bulk_g2s(w2_s + PIECE, img + IMG_W3 + PIECE, PIECE, wb);
        if (a.stage1_mode == 1) bulk_g2s(w1_s, img + IMG_W3 + IMG_W2, IMG_W1, wb);
      }
      __syncwarp();
      mbar_wait(wb, 0u);
      if (has_l1) issue_l1(0);
      issue_l2(0);
    }
    constexpr uint32_t id3 = umma_idesc(128, 128, 0u, 0u);   // f16 x f16 -> f32
    const int buf = isB ? 1 : 0;
    for (int it = 0; it < T; it++) {
      tw = CG_CLK();
      mbar_wait(SBAR(x3_bar), (uint32_t)it & 1u);
      t_x3 += CG_CLK() - tw;
      const bool tr = (it == T / 2) && lane == 0 && !isB;
      CG_TRACE_AT(tr, 0);
      const bool has_next = it + 1 < T;
      const uint32_t x3c = tmem_base + xb_col(it);
      for (int c = buf; c < NCHUNK; c += 2) {
        const uint32_t use = (uint32_t)it * 4u + (uint32_t)(c >> 1);   // earlier uses of this accumulator
        const uint32_t gidx = (uint32_t)it * NCHUNK + (uint32_t)c;     // chunk number in the W3 stream
        const uint32_t rslot = gidx % NPAIR, rph = (gidx / NPAIR) & 1u;
        if (CG_DBG(a)) {   // instrumented run: time the two waits separately
          tw = CG_CLK();
          if (use >= 1u) mbar_wait(SBARI(accfree_bar, buf), (use - 1u) & 1u);
          t_accf += CG_CLK() - tw;
          tw = CG_CLK();
          mbar_wait(SBARI(full_bar, rslot), rph);
          t_ring += CG_CLK() - tw;
        } else if (use >= 1u) {
          mbar_wait2(SBARI(accfree_bar, buf), (use - 1u) & 1u, SBARI(full_bar, rslot), rph);
        } else {
          mbar_wait(SBARI(full_bar, rslot), rph);
        }
        tc_fence_after();
        CG_TRACE_AT(tr, 1 + 2 * c);
        const uint32_t d = tmem_base + (uint32_t)buf * 128u;
        if (elect_one()) {
          const uint32_t w_s = ring_s + rslot * 2 * PIECE;
#pragma unroll
          for (int i = 0; i < (CG_EXP(a, 16) ? 0 : 2); i++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
              // D3[pt][ch] += X3[pt][k] (TMEM, 8 packed columns per K-step, K-block i at column i*64) . W3[ch][k] (smem)
              umma_ts(d, x3c + (uint32_t)i * 64u + (uint32_t)ks * 8u,
                      umma_desc(w_s + (uint32_t)i * PIECE + (uint32_t)ks * 32u), id3, (i | ks) ? 1u : 0u);
            }
          }
          umma_commit(SBARI(free_bar, rslot));
          umma_commit(SBARI(acc_bar, buf));
          if (!isB && c == NCHUNK - 2) umma_commit(SBARI(a_done, it & 1));
        }
        __syncwarp();
        CG_TRACE_AT(tr, 2 + 2 * c);
        // front layers of the NEXT tile run in the shadow of this tile's L3 stream (issuer B)
        if (isB && has_next && c == 1 && has_l1) {
          tw = CG_CLK();
          guard_xb(it + 1);
          issue_l1(it + 1);
          t_x12 += CG_CLK() - tw;
        }
        if (isB && has_next && c == (has_l1 ? 3 : 1)) {
          tw = CG_CLK();
          if (!has_l1) guard_xb(it + 1);
          issue_l2(it + 1);
          t_x12 += CG_CLK() - tw;
        }
      }
    }
    if (CG_DBG(a) && lane == 0) {
      unsigned long long *dd = a.dbg + (size_t)blockIdx.x * 16;
      if (!isB) { dd[0] = CG_CLK() - t_all; dd[1] = t_x3; dd[2] = t_ring; dd[3] = t_accf; dd[5] = T; }
      else dd[4] = t_x12;
    }
  } else if (warp >= NFRONT) {
    // ======================= max warps: L3 epilogue =======================
    // D3[pt][ch]: TMEM lanes = points.  Each warp owns the 32 lanes of its quarter; a 16x256b load hands every
    // thread 4 points x 16 columns, so the column max is 2 FMNMX3/FMNMX per value + a 3-step exchange (14 shuffles
    // for 64 columns); the eight warps (lane quarter x column half) meet in the shared running max of the candidate.
    const int q = warp & 3, hsel = (warp - NFRONT) >> 2;   // TMEM lane quarter, column half of every 128-channel chunk
    const uint32_t lane_lo = (uint32_t)(q * 32) << 16, lane_hi = (uint32_t)(q * 32 + 16) << 16;
    const int mt = tid - NFT;   // 0..255
    int b_cur, tile_cur;
    locate(0, b_cur, tile_cur);
    long long m_all = CG_CLK(), m_wait = 0, mw;
    // Running max of the current candidate: after the lane exchange thread t owns columns 2t, 2t+1 of its warp's 64-column
    // half of every chunk and keeps them in a PRIVATE shared-memory slot (one LDS.64 / STS.64 per chunk, no atomics, no
    // key conversion).  Only when the CTA leaves the candidate do the four lane-quarter warps meet: once per candidate
    // instead of once per tile.
    bool first = true;   // first tile of the candidate inside this CTA: the slot is overwritten, not merged
    for (int it = 0; it < T; it++) {
      const bool leaving = (it == T - 1) || (tile_cur == ntiles - 1);   // last tile of this candidate inside the CTA's range
#pragma unroll 1
      for (int c = 0; c < NCHUNK; c++) {
        const int buf = c & 1;
        const uint32_t use = (uint32_t)it * 4u + (uint32_t)(c >> 1);
        mw = CG_CLK();
        mbar_wait(SBARI(acc_bar, buf), use & 1u);
        m_wait += CG_CLK() - mw;
        tc_fence_after();
        const bool trm = (it == T / 2) && tid == NFT;
        CG_TRACE_AT(trm, 24 + 3 * c);
        const uint32_t col0 = tmem_base + (uint32_t)buf * 128u + (uint32_t)hsel * 64u;
        uint32_t ra[32], rb[32];
        float r0[2];
        if (CG_EXP(a, 2)) {   // timing experiment: no TMEM reads, no reduction
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(SBARI(accfree_bar, buf));
          continue;
        }
        tmem_ld_16x256b_x8(col0 + lane_lo, ra);
        tmem_ld_16x256b_x8(col0 + lane_hi, rb);
        tmem_ld_wait();
        // this warp's part of the accumulator is in registers: hand it back before reducing
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(SBARI(accfree_bar, buf));
        CG_TRACE_AT(trm, 25 + 3 * c);
        colmax64_reduce(ra, rb, lane, r0);
        float2 *slot = &S.sacc[c][warp - NFRONT][lane];
        if (!first) {
          const float2 old = *slot;
          r0[0] = fmaxf(r0[0], old.x);
          r0[1] = fmaxf(r0[1], old.y);
        }
        *slot = make_float2(r0[0], r0[1]);
        CG_TRACE_AT(trm, 26 + 3 * c);
      }
      first = false;
      if (leaving) {
        // fold the candidate's max into the global feature: flush thread (warp w, lane l) takes chunk w, both column
        // halves, columns 2l, 2l+1 -- four channels, each the max over the four lane-quarter warps
        bar_max();
        const int cw = warp - NFRONT;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          float2 m = S.sacc[cw][h * 4][lane];
#pragma unroll
          for (int qq = 1; qq < 4; qq++) {
            const float2 o = S.sacc[cw][h * 4 + qq][lane];
            m.x = fmaxf(m.x, o.x);
            m.y = fmaxf(m.y, o.y);
          }
          const int ch = cw * 128 + h * 64 + 2 * lane;
          const float2 bb = __ldg(reinterpret_cast<const float2 *>(&a.l3.b[ch]));   // bias is constant over points: add after the max
          m.x += bb.x;
          m.y += bb.y;
          if (a.relu3) { m.x = fmaxf(m.x, 0.f); m.y = fmaxf(m.y, 0.f); }
          atomicMax(&a.gmax_keys[(size_t)b_cur * 1024 + ch], cg_f2key(m.x));
          atomicMax(&a.gmax_keys[(size_t)b_cur * 1024 + ch + 1], cg_f2key(m.y));
        }
        bar_max();   // every slot has been read: the next candidate's first tile may overwrite
        first = true;
      }
      if (++tile_cur == ntiles) { tile_cur = 0; b_cur++; }
    }
    if (CG_DBG(a) && tid == NFT) {
      unsigned long long *dd = a.dbg + (size_t)blockIdx.x * 16;
      dd[6] = CG_CLK() - m_all; dd[7] = m_wait;
    }
  } else {
    // ======================= front warps: thread = (point, channel half) =======================
    const int p = tid & 127, half = tid >> 7;
    const int q = warp & 3;                         // TMEM lane quadrant of this warp
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    float vmax = 0.f;   // largest 128->1024 input seen by this thread (post-ReLU, >= 0): reported if beyond the fp16 range
    long long f_all = CG_CLK(), f_l2 = 0, f_l1 = 0, fw;
    // raw input row of this thread's point for the tile being prepared (prefetched one tile ahead so that the
    // dependent global loads ids -> cloud row are off the critical path between two tiles)
    double rx[6];
    float rv[6];
    int id_next = 0;   // scene-point index of this thread's point two tiles ahead (the ids load is one more tile ahead
                       // of the dependent cloud-row loads, so neither ever stalls the front pipeline)
    // (candidate, tile) cursors: the lambdas below are called with consecutive local tile indices, so the flattened index
    // is decomposed once and then advanced (an integer division per call and thread showed up in the front warps' time)
    int pid_b, pid_tile, l0_b, l0_tile;
    locate(0, pid_b, pid_tile);
    l0_b = pid_b; l0_tile = pid_tile;
    auto prefetch_id = [&](int it) {
      if (a.in.x_direct || it >= T) return;
      const int b = pid_b, tile = pid_tile;
      if (++pid_tile == ntiles) { pid_tile = 0; pid_b++; }
      int n = tile * TP + p;
      if (n >= N) n = N - 1;   // duplicate a valid point: cannot change a max
      id_next = a.in.ids ? __ldg(a.in.ids + (size_t)b * N + n) : n;
    };
    // rows of local tile `it` (its id was fetched by prefetch_id(it) earlier); then start the id load of tile it + 1
    auto prefetch = [&](int it) {
      if (a.in.x_direct) {
        int b, tile;
        locate(it, b, tile);
        int n = tile * TP + p;
        if (n >= N) n = N - 1;
        const float *xr = a.in.x_direct + ((size_t)b * N + n) * 6;
#pragma unroll
        for (int k = 0; k < 6; k++) rv[k] = xr[k];
      } else {
        const int id = id_next;
        const double *px = a.in.cloud_xyz + (size_t)id * 3;
        const double *pn = a.in.cloud_nrm + (size_t)id * 3;
        rx[0] = px[0]; rx[1] = px[1]; rx[2] = px[2]; rx[3] = pn[0]; rx[4] = pn[1]; rx[5] = pn[2];
        prefetch_id(it + 1);
      }
    };
    int b_l0 = -1;   // candidate of the previous layer0 call
    // 6 -> 64 (+bias, ReLU) of the prefetched row -> this thread's 32-channel slice of the XA tile
    auto layer0 = [&](int it) {
      const int b = l0_b;
      const bool last_of_cand = (it == T - 1) || (l0_tile == ntiles - 1);
      if (++l0_tile == ntiles) { l0_tile = 0; l0_b++; }
      const int lc = b - b_first, slot = lc & 1;
      if (b != b_l0) mbar_wait(SBARI(cc_full, slot), ((uint32_t)lc >> 1) & 1u);
      b_l0 = b;
      const CandConst &C = S.cc[slot];
      float v[6];
      if (a.in.x_direct) {
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] = rv[k];
      } else {
        const double x = rx[0], y = rx[1], z = rx[2];
        const double nx = rx[3], ny = rx[4], nz = rx[5];
        const double *R = C.pinv;
        double w[6];
        w[0] = R[0] * x + R[1] * y + R[2] * z + R[9];
        w[1] = R[3] * x + R[4] * y + R[5] * z + R[10];
        w[2] = R[6] * x + R[7] * y + R[8] * z + R[11];
        w[3] = R[0] * nx + R[1] * ny + R[2] * nz;
        w[4] = R[3] * nx + R[4] * ny + R[5] * nz;
        w[5] = R[6] * nx + R[7] * ny + R[8] * nz;
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] = (float)((w[k] - S.mean[k]) * S.sden[k]);
      }
      if (a.T3) {  // xyz @ T3 (pointnet2.py:248), normals pass through (:245-250)
        const float x = v[0], y = v[1], z = v[2];
        v[0] = fmaf(z, C.T3[6], fmaf(y, C.T3[3], x * C.T3[0]));
        v[1] = fmaf(z, C.T3[7], fmaf(y, C.T3[4], x * C.T3[1]));
        v[2] = fmaf(z, C.T3[8], fmaf(y, C.T3[5], x * C.T3[2]));
      }
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        const int c0 = half * 32 + cc * 8;
        float o[8];
        float2 o2[4];   // channel pairs: FFMA2 does two of the 6 -> 64 multiply-adds per issue slot
#pragma unroll
        for (int j = 0; j < 4; j++) o2[j] = *reinterpret_cast<const float2 *>(&S.bias0[c0 + 2 * j]);
#pragma unroll
        for (int k = 0; k < 6; k++) {
          const float4 wa = *reinterpret_cast<const float4 *>(&S.w0[k * 64 + c0]);
          const float4 wb = *reinterpret_cast<const float4 *>(&S.w0[k * 64 + c0 + 4]);
          const float2 vv = make_float2(v[k], v[k]);
          o2[0] = ffma2(vv, make_float2(wa.x, wa.y), o2[0]);
          o2[1] = ffma2(vv, make_float2(wa.z, wa.w), o2[1]);
          o2[2] = ffma2(vv, make_float2(wb.x, wb.y), o2[2]);
          o2[3] = ffma2(vv, make_float2(wb.z, wb.w), o2[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          o[2 * j] = fmaxf(o2[j].x, 0.f);
          o[2 * j + 1] = fmaxf(o2[j].y, 0.f);
        }
        const uint32_t off = row_chunk_off(p, c0 >> 3);
        store_hilo8(xa + off, xa + PIECE + off, o);
      }
      fence_proxy_async();   // generic-proxy tile writes -> visible to the async proxy (UMMA operand reads)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive((has_l1 ? SBAR(x1_bar) : SBAR(x2_bar)));
        if (last_of_cand) mbar_arrive(SBARI(cc_free, slot));   // this warp is past its reads of cc[slot]
      }
      CG_TRACE_AT(tid == 0 && it == T / 2 + 1, 52);
    };
    // L1 epilogue of local tile `it`: D1 -> (bias, ReLU | nothing) -> XA as the L2 input
    auto l1_epilogue = [&](int it) {
      const long long fw1 = CG_CLK();
      mbar_wait(SBAR(l1_bar), (uint32_t)it & 1u);
      f_l1 += CG_CLK() - fw1;
      tc_fence_after();
      CG_TRACE_AT(tid == 0 && it == T / 2 + 1, 53);
      if (CG_EXP(a, 4)) {   // timing experiment: front warps skip their math
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(SBAR(x2_bar));
        return;
      }
      float v[32];
      tmem_ld32(tmem_base + lane_sel + xb_col(it) + (uint32_t)half * 32u, v);
      if (a.stage1_mode == 1) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const float2 t = fadd2(make_float2(v[2 * j], v[2 * j + 1]), *reinterpret_cast<const float2 *>(&S.bias1[half * 32 + 2 * j]));
          v[2 * j] = fmaxf(t.x, 0.f);
          v[2 * j + 1] = fmaxf(t.y, 0.f);
        }
      }
      if (a.pf_out) {   // PointNetSeg point feature (pointnet2.py:261)
        int b, tile;
        locate(it, b, tile);
        const int n = tile * TP + p;
        if (n < N) {
          float4 *dstg = reinterpret_cast<float4 *>(a.pf_out + ((size_t)b * N + n) * 64 + half * 32);
#pragma unroll
          for (int j = 0; j < 8; j++) dstg[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        const uint32_t off = row_chunk_off(p, half * 4 + cc);
        store_hilo8(xa + off, xa + PIECE + off, v + cc * 8);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(SBAR(x2_bar));
      CG_TRACE_AT(tid == 0 && it == T / 2 + 1, 54);
    };

    // ---- prologue: front layers of the first tile ----
    prefetch_id(0);
    prefetch(0);
    layer0(0);
    if (T > 1) prefetch(1);
    if (has_l1) l1_epilogue(0);

    for (int it = 0; it < T; it++) {
      const bool has_next = it + 1 < T;
      // A. D2(it) complete; its UMMAs no longer read XA
      fw = CG_CLK();
      mbar_wait(SBAR(l2_bar), (uint32_t)it & 1u);
      f_l2 += CG_CLK() - fw;
      tc_fence_after();
      CG_TRACE_AT(tid == 0 && it == T / 2, 48);
      CG_TRACE_AT(tid == 0 && it == T / 2 + 1, 55);
      if (CG_EXP(a, 4)) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(SBAR(x3_bar));
        if (has_next) {
          int b, tile, bn, tn;
          locate(it + 1, b, tile);
          const int lc = b - b_first, slot = lc & 1;
          if (b != b_l0) mbar_wait(SBARI(cc_full, slot), ((uint32_t)lc >> 1) & 1u);
          b_l0 = b;
          const bool last_of_cand = (it + 1 == T - 1) || (locate(it + 2, bn, tn), bn != b);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive((has_l1 ? SBAR(x1_bar) : SBAR(x2_bar)));
            if (last_of_cand) mbar_arrive(SBARI(cc_free, slot));
          }
          if (has_l1) l1_epilogue(it + 1);
        }
        continue;
      }
      // C. L2 epilogue: D2 -> bias, ReLU -> fp16 pairs; written back IN PLACE as the TMEM A operand of L3
      uint32_t ph[32];
#pragma unroll
      for (int j32 = 0; j32 < 2; j32++) {
        float v[32];
        tmem_ld32(tmem_base + lane_sel + xb_col(it) + (uint32_t)half * 64u + (uint32_t)j32 * 32u, v);
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const float2 t = fadd2(make_float2(v[2 * j], v[2 * j + 1]),
                                 *reinterpret_cast<const float2 *>(&S.bias2[half * 64 + j32 * 32 + 2 * j]));
          vmax = fmax3(vmax, t.x, t.y);   // vmax >= 0: negative sums (ReLU'd to 0 below) cannot raise it
          // F2FP.RELU.SATFINITE: ReLU inside the conversion; values beyond the fp16 range saturate to 65504 instead of
          // becoming inf (t.x -> low half)
          asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(ph[j32 * 16 + j]) : "f"(t.y), "f"(t.x));
        }
      }
      // word j of this thread = channels (half*64 + 2j, +1) of its point.  The packed row goes back into the first 32 of
      // the 64 columns this thread has just read (K-block `half` of X3 lives at column half*64): no other thread reads
      // or writes them, so no block barrier separates the D2 reads from the X3 writes
      tmem_st32(tmem_base + lane_sel + xb_col(it) + (uint32_t)half * 64u, ph);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(SBAR(x3_bar));
      CG_TRACE_AT(tid == 0 && it == T / 2, 49);
      CG_TRACE_AT(tid == 0 && it == T / 2 + 1, 56);
      // B. 6 -> 64 of the NEXT tile (inputs were prefetched a tile ago)
      if (has_next) {
        layer0(it + 1);
        if (it + 2 < T) prefetch(it + 2);   // loads stay in flight across the waits below
        CG_TRACE_AT(tid == 0 && it == T / 2, 50);
      }
      // D. L1 epilogue of the next tile
      if (has_next && has_l1) l1_epilogue(it + 1);
    }
    if (vmax > 65504.f && a.ovf_flag) atomicOr(a.ovf_flag, 1u);
    if (CG_DBG(a) && tid == 0) {
      unsigned long long *dd = a.dbg + (size_t)blockIdx.x * 16;
      dd[8] = CG_CLK() - f_all; dd[9] = f_l2; dd[10] = f_l1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// TMEM fragment-layout self test: writes lane*1000 + column with 32x32b stores, reads it back with the 16x256b
// loads the max epilogue relies on and runs the column-max reduction; out[t*2+k] must equal 31*1000+96 + 2t + k ... (host checks)
__global__ void tmem_layout_selftest_kernel(float *out) {
  __shared__ uint32_t tb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(smem_u32(&tb), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tb;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
  uint32_t r[32];
  for (int j = 0; j < 2; j++) {
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = __float_as_uint((float)((warp * 32 + lane) * 1000 + j * 32 + i));
    tmem_st32(base + lane_sel + (uint32_t)j * 32u, r);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t ra[32], rb[32];
  tmem_ld_16x256b_x8(base + lane_sel, ra);
  tmem_ld_16x256b_x8(base + lane_sel + (16u << 16), rb);
  tmem_ld_wait();
  float o[2];
  colmax64_reduce(ra, rb, lane, o);
  out[(warp * 32 + lane) * 2] = o[0];
  out[(warp * 32 + lane) * 2 + 1] = o[1];
  // raw fragment of thread: first 4 registers of the low-lane load
  for (int i = 0; i < 4; i++) out[256 + (warp * 32 + lane) * 4 + i] = __uint_as_float(ra[i]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 64);
}

}  // namespace

int cg_trunk_launch_p(cg_ctx *ctx, const cg_trunk_args &a) {
  CG_REQUIRE(ctx, a.B > 0 && a.N > 0, "trunk: B,N must be positive");
  CG_REQUIRE(ctx, a.tc_img != nullptr, "trunk: tensor-core weight image missing");
  const int ntiles = (a.N + TP - 1) / TP;
  CG_REQUIRE(ctx, (long long)a.B * ntiles < (1ll << 30), "trunk: B * tiles too large for one launch");
  static bool attr_set[CG_MAX_DEVICES] = {};
  if (!attr_set[ctx->device]) {
    CG_CUDA(ctx, cudaFuncSetAttribute(trunk_p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_P));
    attr_set[ctx->device] = true;
  }
  const int total = a.B * ntiles;
  const int grid = total < ctx->num_sms ? total : ctx->num_sms;
  cg_trunk_args aa = a;
#ifdef CG_EXPERIMENTS   // developer builds only (never read from the environment in a release build)
  static const bool debug = getenv("CG_TRUNK_DEBUG") != nullptr;
  static const int exp_flags = getenv("CG_TRUNK_EXP") ? atoi(getenv("CG_TRUNK_EXP")) : 0;
  aa.exp_flags = exp_flags;
  unsigned long long *d_dbg = nullptr;
  if (debug) {
    CG_CUDA(ctx, cudaMalloc(&d_dbg, (size_t)grid * 128 + 64 * 8));
    CG_CUDA(ctx, cudaMemsetAsync(d_dbg, 0, (size_t)grid * 128 + 64 * 8, ctx->stream));
    aa.dbg = d_dbg;
  }
#endif
  trunk_p_kernel<<<grid, NTP, SMEM_BYTES_P, ctx->stream>>>(aa, ntiles, total);
  CG_LAUNCH_CHECK(ctx);
#ifdef CG_EXPERIMENTS
  if (debug) {
    std::vector<unsigned long long> h((size_t)grid * 16 + 64);
    CG_CUDA(ctx, cudaMemcpyAsync(h.data(), d_dbg, (size_t)grid * 128 + 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_dbg);
    double s[16] = {0};
    for (int i = 0; i < grid; i++)
      for (int k = 0; k < 16; k++) s[k] += (double)h[(size_t)i * 16 + k];
    const double tiles = s[5] > 0 ? s[5] : 1;
    fprintf(stderr, "[trunk_p dbg] exp=%d CTAs=%d tiles=%.0f per tile: MMA total %.0f x3-wait %.0f x1/x2-wait %.0f ring-wait %.0f accfree-wait %.0f | "
            "max total %.0f acc-wait %.0f | front total %.0f l2-wait %.0f l1-wait %.0f cycles\n",
            exp_flags, grid, tiles, s[0] / tiles, s[1] / tiles, s[4] / tiles, s[2] / tiles, s[3] / tiles, s[6] / tiles, s[7] / tiles,
            s[8] / tiles, s[9] / tiles, s[10] / tiles);
    {
      const unsigned long long *tr = h.data() + (size_t)grid * 16;
      const long long t0 = (long long)tr[0];
      static const char *names[64] = {"M x3(it) ready", "M c0 waits done", "M c0 issued", "M c1 waits done", "M c1 issued", "M c2 waits done",
        "M c2 issued", "M c3 waits done", "M c3 issued", "M c4 waits done", "M c4 issued", "M c5 waits done", "M c5 issued",
        "M c6 waits done", "M c6 issued", "M c7 waits done", "M c7 issued", "M L1(it+1) begin", "M L1(it+1) issued",
        "M L2(it+1) begin", "M L2(it+1) issued", "", "", "",
        "X c0 acc ready", "X c0 freed", "X c0 reduced", "X c1 acc ready", "X c1 freed", "X c1 reduced", "X c2 acc ready", "X c2 freed",
        "X c2 reduced", "X c3 acc ready", "X c3 freed", "X c3 reduced", "X c4 acc ready", "X c4 freed", "X c4 reduced", "X c5 acc ready",
        "X c5 freed", "X c5 reduced", "X c6 acc ready", "X c6 freed", "X c6 reduced", "X c7 acc ready", "X c7 freed", "X c7 reduced",
        "F D2(it) ready", "F x3(it) arrived", "F layer0(it+1)+prefetch done", "", "F x1(it+1) arrived", "F D1(it+1) ready",
        "F x2(it+1) arrived", "F D2(it+1) ready", "F x3(it+1) arrived", "", "", "", "", "", "", ""};
      fprintf(stderr, "[trunk_p trace] CTA 0, tile T/2, cycles relative to 'M x3(it) ready':\n");
      std::vector<std::pair<long long, int>> ev;
      for (int i = 0; i < 64; i++)
        if (tr[i] && names[i][0]) ev.push_back({(long long)tr[i] - t0, i});
      std::sort(ev.begin(), ev.end());
      for (auto &e : ev) fprintf(stderr, "[trunk_p trace] %8lld  %s\n", e.first, names[e.second]);
    }
  }
#endif
  return CG_OK;
}

// debug / test entry: exercises the TMEM fragment layout the max epilogue assumes; out_host = 256 + 512 floats
extern "C" int cg_tmem_layout_selftest(cg_ctx *ctx, float *out_host) {
  if (!ctx || !out_host) return CG_EINVAL;
  CG_CUDA(ctx, cudaSetDevice(ctx->device));
  float *d = nullptr;
  CG_CUDA(ctx, cudaMalloc(&d, 768 * sizeof(float)));
  tmem_layout_selftest_kernel<<<1, 128, 0, ctx->stream>>>(d);
  CG_LAUNCH_CHECK(ctx);
  CG_CUDA(ctx, cudaMemcpyAsync(out_host, d, 768 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(d);
  return CG_OK;
}
