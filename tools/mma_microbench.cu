// Development microbenchmark: raw tcgen05.mma issue/execute rate on one SM (no TMA, no epilogue).
// Variants: dependent accumulation into one TMEM tile vs alternating accumulators, N = 256 / 128.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, int MODE>
__global__ void __launch_bounds__(64, 1) k_bench(int iters, long long *out) {
  extern __shared__ unsigned char raw[];
  unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm = slot;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 0) {
    uint64_t a = make_desc(smem_u32(smem)), b = make_desc(smem_u32(smem + 16384));
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        uint32_t d = tm;
        if (MODE == 1) d = tm + (uint32_t)((k & 1) * 256);          // alternate accumulators every MMA
        if (MODE == 2) d = tm + (uint32_t)((it & 1) * 256);          // alternate per "tile" of 8
        mma(d, a + (uint64_t)(2 * (k & 3)), b + (uint64_t)(2 * (k & 3)), idesc, MODE == 3 ? 0u : 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int N, int MODE>
void run(const char *name, int grid) {
  long long *d;
  cudaMalloc(&d, 8);
  int iters = 2000;
  size_t smem = 1024 + 16384 + 32768;
  cudaFuncSetAttribute(k_bench<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_bench<N, MODE><<<grid, 64, smem>>>(iters, d);
  k_bench<N, MODE><<<grid, 64, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 8.0);
  printf("%-44s grid=%3d: %.1f cycles per MMA (M=128,N=%d,K=16) -> %.0f%% of 4096 MAC/clk/SM  [%s]\n", name, grid, per, N,
         100.0 * (128.0 * N * 16 / per) / 4096.0, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<256, 0>("dependent chain, one accumulator", grid);
    run<256, 1>("alternate 2 accumulators per MMA", grid);
    run<256, 2>("alternate accumulators per 8 MMAs", grid);
    run<256, 3>("no accumulate (overwrite)", grid);
    run<128, 0>("N=128 dependent chain", grid);
  }
  return 0;
}
