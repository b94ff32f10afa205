// Development microbenchmark #3: tcgen05.ld round-trip cost by shape / repetition / loads in flight /
// warps per lane quarter.  No MMA running.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__device__ __forceinline__ float do_load(uint32_t taddr, int lane) {
  float r = 0.f;
  if (MODE == 0) {  // 32x32b.x32, one in flight
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    r = __uint_as_float(v[lane]);
  } else if (MODE == 1) {  // two 32x32b.x32 in flight
    uint32_t v[32], u[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(u[0]),"=r"(u[1]),"=r"(u[2]),"=r"(u[3]),"=r"(u[4]),"=r"(u[5]),"=r"(u[6]),"=r"(u[7]),"=r"(u[8]),"=r"(u[9]),"=r"(u[10]),"=r"(u[11]),"=r"(u[12]),"=r"(u[13]),"=r"(u[14]),"=r"(u[15]),"=r"(u[16]),"=r"(u[17]),"=r"(u[18]),"=r"(u[19]),"=r"(u[20]),"=r"(u[21]),"=r"(u[22]),"=r"(u[23]),"=r"(u[24]),"=r"(u[25]),"=r"(u[26]),"=r"(u[27]),"=r"(u[28]),"=r"(u[29]),"=r"(u[30]),"=r"(u[31]) : "r"(taddr + 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    r = __uint_as_float(v[lane]) + __uint_as_float(u[lane]);
  } else if (MODE == 2) {  // 32x32b.x8
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    r = __uint_as_float(v[lane & 7]);
  } else if (MODE == 3) {  // 16x256b.x8 = 32 regs, 16 lanes x 64 columns
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    r = __uint_as_float(v[lane]);
  } else if (MODE == 5) {  // 32x32b.x16
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    r = __uint_as_float(v[lane & 15]);
  } else if (MODE == 6) {  // 4 x 32x32b.x8 in flight, one wait (32 columns)
    uint32_t v[32];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[8*q+0]),"=r"(v[8*q+1]),"=r"(v[8*q+2]),"=r"(v[8*q+3]),"=r"(v[8*q+4]),"=r"(v[8*q+5]),"=r"(v[8*q+6]),"=r"(v[8*q+7]) : "r"(taddr + 8 * q));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    r = __uint_as_float(v[lane]);
  } else if (MODE == 7) {  // 8 x 32x32b.x8 in flight, one wait (64 columns)
    uint32_t v[64];
#pragma unroll
    for (int q = 0; q < 8; ++q)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[8*q+0]),"=r"(v[8*q+1]),"=r"(v[8*q+2]),"=r"(v[8*q+3]),"=r"(v[8*q+4]),"=r"(v[8*q+5]),"=r"(v[8*q+6]),"=r"(v[8*q+7]) : "r"(taddr + 8 * q));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    r = __uint_as_float(v[lane]) + __uint_as_float(v[32 + lane]);
  } else if (MODE == 8) {  // 32x32b.x32 with .pack::16b: 64 columns of 16-bit cells -> 32 registers
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    r = __uint_as_float(v[lane]);
  } else if (MODE == 4) {  // 32x32b.x1
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    r = __uint_as_float(v);
  }
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(544, 1) k(int iters, int nwarps, long long *out, float *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm = slot;
  if (warp >= 1 && warp <= nwarps) {
    uint32_t taddr = tm + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) acc += do_load<MODE>(taddr + (uint32_t)((i & 3) * 64), lane);
    long long t1 = clock64();
    if (acc == 1234.5f) sink[0] = acc;
    if (blockIdx.x == 0 && warp == 1 && lane == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int MODE>
void run(const char *name, int bytes_per_iter) {
  long long *d; float *sink;
  cudaMalloc(&d, 8); cudaMalloc(&sink, 8);
  for (int nw : {1, 4, 8, 16}) {
    int iters = 4000;
    k<MODE><<<148, 544>>>(iters, nw, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps=%2d: %7.1f cycles per iteration, %6.1f B/clk per warp, %7.1f B/clk per SM [%s]\n", name, nw,
           (double)h / iters, bytes_per_iter * (double)iters / h, nw * bytes_per_iter * (double)iters / h, cudaGetErrorString(e));
  }
  cudaFree(d); cudaFree(sink);
}

int main() {
  run<2>("32x32b.x8", 1024);
  run<5>("32x32b.x16", 2048);
  run<6>("4 x 32x32b.x8 in flight", 4096);
  run<7>("8 x 32x32b.x8 in flight", 8192);
  run<0>("32x32b.x32", 4096);
  run<8>("32x32b.x32.pack::16b (64 col)", 8192);

  return 0;
}
