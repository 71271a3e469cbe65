// cg_common.cuh -- shared internals of libcatgrasp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <utility>
#include <vector>
#include "../../include/catgrasp_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcatgrasp_b200 is written for sm_100a (B200) only"
#endif

constexpr int CG_MAX_DEVICES = 64;   // per-device one-time kernel attributes are tracked in arrays of this size

struct cg_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  // 0 = fp32 SIMT, 1 = tcgen05 bf16 3-pass, 2 = tcgen05 fp16 2-pass, 3 = persistent tcgen05, single fp16 pass (default)
  int engine = 3;
  cudaEvent_t switch_event = nullptr;   // orders a newly selected stream behind the previous one (shared workspaces)
  uint32_t *ovf_flag = nullptr;   // device word: engine 3 saw a 128->1024 input above the fp16 range (clamped)
  int num_sms = 148;
  // optional event-pair timing of trunk launches (bench roofline)
  bool prof = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  // CG_TRACE=1 (diagnostics): an event after every launch; per-call-site durations are printed to stderr every 3000 launches
  bool trace = false;
  std::vector<std::pair<const char *, cudaEvent_t>> trace_events;
  // grow-only device workspace, carved per call
  void *ws = nullptr;
  size_t ws_bytes = 0;
  // grow-only pinned host staging for the *_host entry points
  void *hs = nullptr;
  size_t hs_bytes = 0;
  // second device arena for *_host entry points' device copies of I/O
  void *io = nullptr;
  size_t io_bytes = 0;
};

#define CG_CUDA(ctx, call)                                                        \
  do {                                                                            \
    cudaError_t _e = (call);                                                      \
    if (_e != cudaSuccess) {                                                      \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e);            \
      return CG_ECUDA;                                                            \
    }                                                                             \
  } while (0)

#define CG_REQUIRE(ctx, cond, msg)                                                \
  do {                                                                            \
    if (!(cond)) {                                                                \
      (ctx)->err = std::string("invalid argument: ") + (msg);                     \
      return CG_EINVAL;                                                           \
    }                                                                             \
  } while (0)

void cg_trace_mark(cg_ctx *ctx, const char *where);
#define CG_STR2(x) #x
#define CG_STR(x) CG_STR2(x)

#define CG_LAUNCH_CHECK(ctx)                                                      \
  do {                                                                            \
    (ctx)->launches++;                                                            \
    if ((ctx)->trace) cg_trace_mark((ctx), __FILE__ ":" CG_STR(__LINE__));        \
    cudaError_t _e = cudaGetLastError();                                          \
    if (_e != cudaSuccess) {                                                      \
      (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(_e);       \
      return CG_ECUDA;                                                            \
    }                                                                             \
  } while (0)

int cg_ws_reserve(cg_ctx *ctx, size_t bytes);
int cg_io_reserve(cg_ctx *ctx, size_t bytes);
int cg_hs_reserve(cg_ctx *ctx, size_t bytes);

// bump allocator over a reserved arena (256-byte aligned pieces)
struct cg_arena {
  char *base;
  size_t off = 0;
  explicit cg_arena(void *b) : base(static_cast<char *>(b)) {}
  template <typename T>
  T *take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T *p = reinterpret_cast<T *>(base + off);
    off += n * sizeof(T);
    return p;
  }
  static size_t pad(size_t bytes) { return (bytes + 255) & ~size_t(255); }
};

// ---- order-preserving float <-> uint key (for atomicMax on floats) --------
__host__ __device__ __forceinline__ uint32_t cg_f2key(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  uint32_t b;
  memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float cg_key2f(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}
