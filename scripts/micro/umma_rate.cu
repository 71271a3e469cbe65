// Micro-benchmark: raw tcgen05.mma issue/execute rate for the shapes the trunk uses (run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate scripts/micro/umma_rate.cu && /tmp/umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

template <int N, bool TS, int FMT>   // FMT 0 = f16, 1 = bf16
__global__ void __launch_bounds__(128, 1) rate_kernel(long long *out, int iters) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char *smem = (unsigned char *)(((uintptr_t)smem_dyn + 1023) & ~uintptr_t(1023));
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  constexpr uint32_t id = (1u << 4) | ((uint32_t)FMT << 7) | ((uint32_t)FMT << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (warp == 1) {
    const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 32768);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
          const uint64_t bd = umma_desc(b_s + ks * 32);
          if (TS) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tb + (uint32_t)(it & 1) * 256u), "r"(tb + 512u - 64u + ks * 8), "l"(bd), "r"(id), "r"(1u) : "memory");
          } else {
            const uint64_t ad = umma_desc(a_s + ks * 32);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tb + (uint32_t)(it & 1) * 256u), "l"(ad), "l"(bd), "r"(id), "r"(1u) : "memory");
          }
        }
      }
      __syncwarp();
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

template <int N, bool TS, int FMT>
void run(const char *name, int nblocks) {
  long long *d;
  cudaMalloc(&d, nblocks * sizeof(long long));
  const int iters = 2048;
  cudaFuncSetAttribute(rate_kernel<N, TS, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  rate_kernel<N, TS, FMT><<<nblocks, 128, 100 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, d, nblocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < nblocks; i++) avg += (double)h[i];
  avg /= nblocks;
  printf("%-28s blocks=%3d  %s  cycles/UMMA = %.1f  (ideal %d)\n", name, nblocks, cudaGetErrorString(e), avg / (iters * 4.0), N / 2);
  cudaFree(d);
}

int main() {
  run<128, false, 1>("SS bf16 M128 N128 K16", 1);
  run<128, false, 1>("SS bf16 M128 N128 K16", 148);
  run<256, false, 1>("SS bf16 M128 N256 K16", 148);
  run<128, true, 1>("TS bf16 M128 N128 K16", 148);
  run<256, true, 1>("TS bf16 M128 N256 K16", 148);
  run<128, true, 0>("TS f16  M128 N128 K16", 148);
  run<64, true, 0>("TS f16  M128 N64  K16", 148);
  return 0;
}
