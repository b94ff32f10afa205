// bucket_sort.cuh -- stable sort of (bounded 32-bit key, position) pairs in ONE persistent launch.
//
// Every sorted-occurrence walk of the training steps (train_bpr.cu, train_fm.cu) needs the occurrences of a batch
// grouped by table row, in a FIXED order (the gradient sums are reduced in that order: no float atomics, bit-
// reproducible steps).  The keys are row ids of a table (17-26 bits), the values are always the occurrence's own
// position 0..M-1, and M is 2^20..2^25: far too little work per radix pass for a library sort, which spends its time
// in launches (cub::DeviceRadixSort: 8-bit digits, a histogram kernel + one kernel per digit; 77 us for 2^20 pairs
// and 109 us for 2^21 on B200, 10 % of a cfg3 step).
//
// Here: least-significant-digit radix sort with digits of up to 12 bits (2 passes for <= 24 bits, 3 for <= 36), all
// passes inside one cooperative kernel (grid = co-resident blocks, hand-written grid barrier):
//   A. every warp counts the digits of ITS contiguous slice of the block's chunk into a warp-private shared-memory
//      histogram (match.any groups equal digits of a 32-element step; the group's first lane adds the group size:
//      no atomics), the block publishes its column of the [digit][block] histogram matrix;
//   B. (after a grid barrier) the rows of the matrix are scanned, one warp per row;
//   C. (after a grid barrier) every block scans the 4096 row totals itself, turns its warp histograms into exclusive
//      prefixes over the warps, and walks its chunk again: destination = digit base + blocks before me + warps before
//      me + equal digits earlier in my slice + equal digits in lower lanes.  Stable by construction.
// The first pass synthesises the values (position = index), so only the keys are read.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace rb2sort {

constexpr int kMaxDigitBits = 12;
constexpr int kMaxBins = 1 << kMaxDigitBits;
constexpr int kMaxBlocks = 1024;   // capacity of the histogram matrix: [kMaxBins][kMaxBlocks]

struct Plan {
  int npass;
  int shift[3], nbits[3];
};

static inline Plan make_plan(int bits) {
  Plan p{};
  if (bits < 1) bits = 1;
  if (bits > 32) bits = 32;
  p.npass = (bits + kMaxDigitBits - 1) / kMaxDigitBits;
  int left = bits, sh = 0;
  for (int i = 0; i < p.npass; ++i) {
    int nb = (left + (p.npass - i) - 1) / (p.npass - i);
    p.shift[i] = sh;
    p.nbits[i] = nb;
    sh += nb;
    left -= nb;
  }
  return p;
}

struct Ws {
  uint32_t *tmp_key, *tmp_val;   // [M]
  uint32_t *hist;                // [kMaxBins * blocks]
  uint32_t *totals;              // [kMaxBins]
  unsigned *bar;                 // [1] grid barrier counter (zeroed before every launch)
};

static inline size_t carve(Ws &w, void *base, int64_t M) {
  Carver c(base);
  w.tmp_key = c.take<uint32_t>(M);
  w.tmp_val = c.take<uint32_t>(M);
  w.hist = c.take<uint32_t>((size_t)kMaxBins * kMaxBlocks);
  w.totals = c.take<uint32_t>(kMaxBins);
  w.bar = c.take<unsigned>(64);
  return c.off;
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all blocks are co-resident (cooperative launch); `target` = number of arrivals that complete this barrier
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (ld_acquire_gpu(bar) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

struct Args {
  const uint32_t *key_in;
  uint32_t *key_out, *val_out;
  Ws w;
  int64_t M, chunk;       // chunk: elements per block, a multiple of the block size
  int val_shift;          // value of element i = i << val_shift
  Plan plan;
};

// CT: counter type of the warp-private histograms (uint16_t when a warp's slice holds < 65535 elements);
// WARPS: warps per block (one block per SM: 16 warps with 16-bit counters, 8 with 32-bit ones)
constexpr int kPrefetch = 8;      // keys loaded ahead per warp: the walk is a chain of dependent steps, the loads are not

template <typename CT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_bucket_sort(Args a) {
  constexpr int THREADS = WARPS * 32;
  constexpr int PER = kMaxBins / THREADS;                                      // digits per thread, block-wide steps
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t warp_sums[WARPS];
  CT *cnt = reinterpret_cast<CT *>(smem_raw);                                  // [WARPS][NB]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int blk = blockIdx.x, G = gridDim.x;
  const int64_t lo = (int64_t)blk * a.chunk, hi = min(lo + a.chunk, a.M);
  const int64_t wchunk = a.chunk / WARPS;
  const int64_t wlo = min(lo + warp * wchunk, hi), whi = min(wlo + wchunk, hi);
  const unsigned lt = (1u << lane) - 1u;
  unsigned arrivals = 0;

  for (int pass = 0; pass < a.plan.npass; ++pass) {
    const int nbits = pass == 0 ? a.plan.nbits[0] : (pass == 1 ? a.plan.nbits[1] : a.plan.nbits[2]);
    const int shift = pass == 0 ? a.plan.shift[0] : (pass == 1 ? a.plan.shift[1] : a.plan.shift[2]);
    const int NB = 1 << nbits;
    const uint32_t mask = (uint32_t)NB - 1u;
    uint32_t *gbase = reinterpret_cast<uint32_t *>(smem_raw + (size_t)WARPS * NB * sizeof(CT));   // [NB]
    // ping-pong so that the LAST pass writes key_out / val_out
    const bool to_out = ((a.plan.npass - 1 - pass) & 1) == 0;
    const uint32_t *src_k = pass == 0 ? a.key_in : (to_out ? a.w.tmp_key : a.key_out);
    const uint32_t *src_v = to_out ? a.w.tmp_val : a.val_out;
    uint32_t *dst_k = to_out ? a.key_out : a.w.tmp_key;
    uint32_t *dst_v = to_out ? a.val_out : a.w.tmp_val;
    CT *mine = cnt + (size_t)warp * NB;

    // ---- A: warp-private digit counts ---------------------------------------------------------------------------
    {
      uint4 *z = reinterpret_cast<uint4 *>(cnt);
      const int n16 = (int)((size_t)WARPS * NB * sizeof(CT) / 16);
      for (int i = threadIdx.x; i < n16; i += THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    for (int64_t i0 = wlo; i0 < whi; i0 += 32 * kPrefetch) {
      uint32_t kk[kPrefetch];
#pragma unroll
      for (int u = 0; u < kPrefetch; ++u) {
        const int64_t i = i0 + u * 32 + lane;
        kk[u] = i < whi ? src_k[i] : 0u;
      }
#pragma unroll
      for (int u = 0; u < kPrefetch; ++u) {
        const int64_t i = i0 + u * 32 + lane;
        const bool valid = i < whi;
        const uint32_t d = valid ? ((kk[u] >> shift) & mask) : (0x80000000u | (uint32_t)lane);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && (peers & lt) == 0) mine[d] = (CT)(mine[d] + __popc(peers));
        __syncwarp();
      }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < NB; d += THREADS) {
      uint32_t s = 0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) s += cnt[(size_t)w * NB + d];
      a.w.hist[(size_t)d * G + blk] = s;
    }
    arrivals += G;
    grid_barrier(a.w.bar, arrivals);

    // ---- B: exclusive scan of every row [digit][0..G) of the matrix, one warp per row ---------------------------
    for (int d = blk * WARPS + warp; d < NB; d += G * WARPS) {
      uint32_t *row = a.w.hist + (size_t)d * G;
      constexpr int SEG = kMaxBlocks / 32;                 // G <= kMaxBlocks: lane l owns entries [l * seg, (l + 1) * seg)
      const int seg = (G + 31) / 32;
      uint32_t x[SEG];
      uint32_t s = 0;
#pragma unroll
      for (int j = 0; j < SEG; ++j) {
        const int b = lane * seg + j;
        x[j] = (j < seg && b < G) ? __ldcg(row + b) : 0u;
        s += x[j];
      }
      uint32_t inc = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
      }
      uint32_t run = inc - s;
#pragma unroll
      for (int j = 0; j < SEG; ++j) {
        const int b = lane * seg + j;
        if (j < seg && b < G) {
          row[b] = run;
          run += x[j];
        }
      }
      if (lane == 31) a.w.totals[d] = inc;
    }
    arrivals += G;
    grid_barrier(a.w.bar, arrivals);

    // ---- C: digit bases (scan of the totals), warp prefixes, stable scatter ---------------------------------------
    {
      // block-wide exclusive scan of totals[NB]: each thread owns `per` consecutive digits
      const int per = (NB + THREADS - 1) / THREADS;
      const int d0 = threadIdx.x * per;
      uint32_t loc[PER], mycol[PER];
      uint32_t s = 0;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const bool ok = j < per && d0 + j < NB;
        loc[j] = ok ? __ldcg(a.w.totals + d0 + j) : 0u;
        mycol[j] = ok ? __ldcg(a.w.hist + (size_t)(d0 + j) * G + blk) : 0u;
        s += loc[j];
      }
      uint32_t inc = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
      }
      if (lane == 31) warp_sums[warp] = inc;
      __syncthreads();
      uint32_t before = inc - s;
      for (int w = 0; w < warp; ++w) before += warp_sums[w];
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        if (j < per && d0 + j < NB) {
          gbase[d0 + j] = before + mycol[j];
          before += loc[j];
        }
      }
      // warp histograms -> exclusive prefixes over the warps of this block
      for (int d = threadIdx.x; d < NB; d += THREADS) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
          const uint32_t c = cnt[(size_t)w * NB + d];
          cnt[(size_t)w * NB + d] = (CT)run;
          run += c;
        }
      }
      __syncthreads();
    }
    for (int64_t i0 = wlo; i0 < whi; i0 += 32 * kPrefetch) {
      uint32_t kk[kPrefetch], vv[kPrefetch];
#pragma unroll
      for (int u = 0; u < kPrefetch; ++u) {
        const int64_t i = i0 + u * 32 + lane;
        kk[u] = i < whi ? src_k[i] : 0u;
        vv[u] = pass == 0 ? ((uint32_t)i << a.val_shift) : (i < whi ? src_v[i] : 0u);
      }
#pragma unroll
      for (int u = 0; u < kPrefetch; ++u) {
        const int64_t i = i0 + u * 32 + lane;
        const bool valid = i < whi;
        const uint32_t d = valid ? ((kk[u] >> shift) & mask) : (0x80000000u | (uint32_t)lane);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        uint32_t base = 0;
        if (valid) base = mine[d];
        __syncwarp();
        if (valid) {
          if ((peers & lt) == 0) mine[d] = (CT)(base + __popc(peers));
          const uint32_t pos = gbase[d] + base + __popc(peers & lt);
          dst_k[pos] = kk[u];
          dst_v[pos] = vv[u];
        }
        __syncwarp();
      }
    }
    if (pass + 1 < a.plan.npass) {
      arrivals += G;
      grid_barrier(a.w.bar, arrivals);
    }
  }
}

template <typename CT, int WARPS>
static inline size_t smem_bytes(int nbits) { return ((size_t)WARPS * sizeof(CT) + 4) << nbits; }

template <typename CT, int WARPS>
static inline int max_blocks() {
  // co-resident blocks of this kernel on the current device (cached per device)
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
  const size_t smem = smem_bytes<CT, WARPS>(kMaxDigitBits);
  cudaFuncSetAttribute(k_bucket_sort<CT, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bucket_sort<CT, WARPS>, WARPS * 32, smem);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int n = (per_sm > 0 ? 1 : 0) * sms;         // one block per SM: fewer arrivals per grid barrier, shorter matrix rows
  if (n > kMaxBlocks) n = kMaxBlocks;
  if (n < 1) n = 1;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

template <typename CT, int WARPS>
static inline int launch(Args &a, int nbmax, cudaStream_t st) {
  constexpr int THREADS = WARPS * 32;
  int G = max_blocks<CT, WARPS>();
  const int64_t want = (a.M + 4095) / 4096;    // a block takes at least 4096 elements
  if (want < G) G = (int)std::max<int64_t>(1, want);
  a.chunk = ((a.M + G - 1) / G + THREADS - 1) / THREADS * THREADS;
  if (sizeof(CT) == 2 && a.chunk / WARPS >= 65535) return -1;       // the caller switches to 32-bit counters
  RB2_CUDA(cudaMemsetAsync(a.w.bar, 0, sizeof(unsigned), st));
  void *args[] = {&a};
  RB2_CUDA(cudaLaunchCooperativeKernel((void *)k_bucket_sort<CT, WARPS>, dim3(G), dim3(THREADS), args,
                                       smem_bytes<CT, WARPS>(nbmax), st));
  return 0;
}

static inline size_t tmp_bytes(int64_t M) {
  Ws w;
  return carve(w, nullptr, M);
}

// key_in[M] (values < 2^bits) -> key_out[M] ascending, val_out[M] = original positions (<< val_shift); equal keys keep
// their order.
static inline int sort_positions(const uint32_t *key_in, uint32_t *key_out, uint32_t *val_out, int64_t M, int bits,
                                 void *tmp, size_t tmp_size, cudaStream_t st, int val_shift = 0) {
  if (M <= 0) return 0;
  RB2_REQUIRE(M < ((int64_t)1 << 32), RB2_EINVAL, "bucket sort: too many elements");
  Args a{};
  size_t need = carve(a.w, tmp, M);
  RB2_REQUIRE(tmp_size >= need, RB2_EWORKSPACE, "bucket sort: workspace %zu < %zu", tmp_size, need);
  a.key_in = key_in;
  a.key_out = key_out;
  a.val_out = val_out;
  a.M = M;
  a.val_shift = val_shift;
  a.plan = make_plan(bits);
  int nbmax = 0;
  for (int i = 0; i < a.plan.npass; ++i) nbmax = std::max(nbmax, a.plan.nbits[i]);
  int rc = launch<uint16_t, 16>(a, nbmax, st);
  if (rc == -1) rc = launch<uint32_t, 8>(a, nbmax, st);
  return rc;
}

}  // namespace rb2sort
