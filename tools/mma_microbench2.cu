// Development microbenchmark #2: does concurrent TMEM reading (epilogue) or bulk-copy traffic into
// shared memory (TMA) slow tcgen05.mma down?  One CTA per SM: warp 0 issues MMAs, warps 2-5 optionally
// spin on tcgen05.ld of the other accumulator stage, warp 1 optionally streams cp.async.bulk into smem.
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
#define TC_LD32(taddr, v)                                                                                       \
  asm volatile(                                                                                                 \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, " \
      "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"            \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),           \
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),           \
        "=r"(v[30]), "=r"(v[31])                                                                               \
      : "r"(taddr))

// flags: bit0 = TMEM readers on, bit1 = bulk copies on, bit2 = readers do ALU work (max tree)
__global__ void __launch_bounds__(320, 1) k_bench(int iters, int flags, const unsigned char *gsrc, long long *out,
                                                  float *sink) {
  extern __shared__ unsigned char raw[];
  unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  unsigned char *scratch = smem + 16384 + 32768;   // bulk-copy target (not an MMA operand): 64 KB
  __shared__ uint64_t bar, cbar[2];
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0;
  if (threadIdx.x == 0) {
    done = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm = slot;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (warp == 0) {
    if (lane == 0) {
      uint64_t a = make_desc(smem_u32(smem)), b = make_desc(smem_u32(smem + 16384));
      long long t0 = clock64();
      if (flags & 8) {   // no MMAs: just let the other warps run for a while
        while (clock64() - t0 < (long long)iters * 1024) {}
      } else if (flags & 16) {   // duty cycle: 8 MMAs, then idle for as long as they take
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int k = 0; k < 8; ++k) mma(tm, a + (uint64_t)(2 * (k & 3)), b + (uint64_t)(2 * (k & 3)), idesc, 1u);
          long long w0 = clock64();
          while (clock64() - w0 < 1024) {}
        }
      } else {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) mma(tm, a + (uint64_t)(2 * (k & 3)), b + (uint64_t)(2 * (k & 3)), idesc, 1u);
      }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
      long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      done = 1;
    }
  } else if (warp == 1) {
    if ((flags & 2) && lane == 0) {
      uint32_t ph[2] = {0, 0};
      int n = 0;
      while (!done) {
        int s = n & 1;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[s])), "r"(32768) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(scratch + s * 32768)), "l"(gsrc + (size_t)((n * 37 + blockIdx.x) % 2048) * 32768), "r"(32768),
                     "r"(smem_u32(&cbar[s])) : "memory");
        if (n >= 1) {  // wait for the previous one: two copies in flight
          int q = s ^ 1;
          asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(smem_u32(&cbar[q])), "r"(ph[q]) : "memory");
          ph[q] ^= 1;
        }
        ++n;
      }
      // drain the last copy
      int q = (n - 1) & 1;
      if (n > 0)
        asm volatile("{\n.reg .pred p;\nW3:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D3;\nbra W3;\nD3:\n}\n" ::"r"(smem_u32(&cbar[q])), "r"(ph[q]) : "memory");
      if (blockIdx.x == 0) out[1] = n;
    }
  } else if (flags & 1) {
    const int quarter = warp & 3;
    uint32_t taddr = tm + ((uint32_t)(quarter * 32) << 16) + 256u;   // the OTHER accumulator stage
    float acc = 0.f;
    long long n = 0;
    while (!done) {
      uint32_t v[32];
      if (flags & 32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(v[8*q+0]),"=r"(v[8*q+1]),"=r"(v[8*q+2]),"=r"(v[8*q+3]),"=r"(v[8*q+4]),"=r"(v[8*q+5]),"=r"(v[8*q+6]),"=r"(v[8*q+7])
                       : "r"(taddr + (uint32_t)((n & 7) * 32 + 8 * q)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
      } else {
      TC_LD32(taddr + (uint32_t)((n & 7) * 32), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      if (flags & 4) {
        float m = __uint_as_float(v[0]);
#pragma unroll
        for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
        acc += m;
      } else {
        acc += __uint_as_float(v[lane]);
      }
      ++n;
    }
    if (acc == 12345.f) sink[0] = acc;
    if (blockIdx.x == 0 && warp == 2 && lane == 0) out[2] = n;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  long long *d;
  float *sink;
  unsigned char *gsrc;
  cudaMalloc(&d, 64);
  cudaMalloc(&sink, 64);
  cudaMalloc(&gsrc, (size_t)2048 * 32768);
  cudaMemset(gsrc, 0, (size_t)2048 * 32768);
  size_t smem = 1024 + 16384 + 32768 + 65536;
  cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char *names[] = {"MMA alone", "MMA + 4 warps tcgen05.ld (other stage)", "MMA + bulk copies into smem",
                         "MMA + tcgen05.ld + bulk copies", "", "MMA + tcgen05.ld + max-tree ALU", "",
                         "MMA + tcgen05.ld + ALU + bulk copies"};
  for (int flags : {1, 33, 9, 41, 17, 49}) {
   for (int threads : {192, 320}) {
    int iters = 4000;
    cudaMemset(d, 0, 64);
    k_bench<<<148, threads, smem>>>(iters, flags, gsrc, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3] = {0, 0, 0};
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    double per = (double)h[0] / (iters * 8.0);
    char nmbuf[128];
    snprintf(nmbuf, sizeof(nmbuf), "%s%s, %d reader warps, %s", (flags & 8) ? "no MMA" : (flags & 16) ? "MMA 50% duty" : "MMA 100%", "", (threads - 64) / 32, (flags & 32) ? "x8+wait" : "x32");
    const char *nm = nmbuf; const char *unused = flags < 8 ? names[flags] : (flags == 9 ? "no MMA: 4 warps tcgen05.ld" : flags == 13 ? "no MMA: tcgen05.ld + ALU" : flags == 17 ? "MMA 50% duty + tcgen05.ld" : "no MMA: tcgen05.ld + bulk copies");
    printf("%-42s %.1f cycles/MMA (%.0f%% of peak); bulk copies %lld (%.1f B/clk), tmem loads/warp %lld (%.1f B/clk per SM) [%s]\n",
           nm, per, 100.0 * 128.0 / per, h[1], h[1] * 32768.0 / (double)h[0], h[2], h[2] * 4096.0 / (double)h[0] * ((threads - 64) / 32),
           cudaGetErrorString(e));
   }
  }
  return 0;
}
