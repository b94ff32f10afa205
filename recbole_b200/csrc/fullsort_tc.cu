// fullsort_tc.cu -- tensor-core full-sort scorer for sm_100a: tcgen05.mma + TMEM + TMA, with a
// certified exact top-K.
//
// Same contract as the fp32 path (fullsort.cu): replaces BPR.full_sort_predict (bpr.py:91-96) /
// SASRec.full_sort_predict (sasrec.py:152-158) + the mask of Trainer._full_sort_batch_eval
// (trainer.py:342-345) + TopKEvaluator.collect's topk (evaluators.py:68-72); the score matrix is
// never written.
//
//   k_convert_rows   fp32 rows -> 16-bit rows (queries are gathered by id), plus per-row ||u||, ||u - h(u)||
//                    and, for the item side, max ||h(v)|| / max ||v - h(v)||.  Default: fp16 after an EXACT
//                    power-of-two rescale (each query row by its own 2^-e, the item table by one 2^-e), so
//                    that every score fits the FP16 accumulator; variant 1: bf16, no rescale.
//   k_fullsort_tc    persistent, warp-specialised, 384 threads: warp 0 = TMA producer (128B-swizzled
//                    tiles, a CTA pair loads every item slot by halves and multicasts it), warp 1 =
//                    single-thread tcgen05.mma issuer (M = 128, N = 256, K = 16; two 256-column accumulator
//                    stages in TMEM), warp groups 1 and 2 = two epilogue warp sets (thread <-> TMEM lane <->
//                    query row).  Set h drains columns [128 h, +128) of EVERY stage with two
//                    tcgen05.ld ... .pack::16b (64 fp16 scores -> 32 registers each) and hands the stage
//                    back BEFORE looking at the scores; HMNMX2 max tree against the row's threshold; a
//                    passing score is appended to a small per-row buffer in shared memory and the buffers
//                    are folded into the K' (16 or 32) candidate lists held in REGISTERS by all lanes
//                    together; the two sets publish their list minima and filter with the larger one.
//                    Pad / history are checked at the append; history without a search: a set sees its
//                    items in ascending order and the CSR row is sorted, so a cursor into the row only
//                    moves forward.  setmaxnreg moves registers from the producer / MMA warp group to the
//                    epilogue warp groups.  Rows whose certificate fails are re-scored by a second pass with
//                    fp32 accumulators (same operands, ~10x tighter bound) before the exact kernel.
//   k_refine         candidates are re-scored with the canonical fp32 chain s = fmaf(q[k], v[k], s)
//                    and ordered (score desc, id asc).  Certificate per row:
//                        exact_K  >  max_part(approx K'-th score) + E,
//                        E = ||du||*max||hv|| + ||u||*max||dv||   (Cauchy-Schwarz on u.v - hu.hv = du.hv + u.dv)
//                          + 2^-11 * sum_j ||hu[0:16j]|| * max||hv||   (FP16 accumulators: every K = 16 MMA rounds
//                            the running sum to nearest fp16 -- checked bit for bit by tools/mma_f16acc_check.cu)
//                          + slack
//                    proves no non-candidate can be in the exact top-K.  Rows that fail are
//                    compacted and redone by the fp32 kernel, so the result is always exact.
//
// What bounds it (tools/tc_trace.py, tools/tc_bench.py): with K = d = 128 a tile is only 8 MMAs (1024
// cycles); the chain TMA -> MMA -> tcgen05.ld -> release has to turn around inside that, and the kernel
// runs power-limited (SM clock ~1.5-1.6 GHz under this load).  Measured 1.18 PFLOP/s at 512k x 2M x 128 =
// 0.71 of the cuBLAS bf16 peak measured on the same GPUs; with the examination switched off entirely the
// MMA/TMA chain alone reaches 1.36.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

// fp32 path (fullsort.cu), used for the rows whose certificate fails and for dims the MMA tiling
// does not cover
int rb2_fullsort_fp32(const float *query_p, const int64_t *query_ids, int64_t nq, const float *item_p,
                      int64_t n_items_local, int64_t item_base, int32_t dim, const int64_t *hist_indptr,
                      const int64_t *hist_indices, int32_t k, int64_t *out_ids, float *out_scores, void *workspace,
                      size_t workspace_bytes, cudaStream_t st, const int32_t *row_map);
size_t rb2_fullsort_fp32_workspace(int64_t nq, int64_t n_items_local, int32_t dim, int32_t k);

// Knobs, adaptive statistics and last-call counters live in an rb2_scorer_state (caller-owned, or the calling thread's
// default): rb2_cur_scorer().  variant: 0 = default (= 3); 1 = bf16 operands, fp32 accumulators, per-CTA MMAs;
// 3 = fp16 operands (rows rescaled by powers of two), FP16 accumulators drained with .pack::16b, per-CTA MMAs;
// 2 = as 3 with CTA-pair MMAs (cta_group::2).
extern "C" int rb2_fullsort_tc_set_variant(int32_t v) {
  if (v < 0 || v > 3) return RB2_EINVAL;
  rb2_scorer_state &S = rb2_cur_scorer();
  S.fail_ema = 0.f;                     // (also forgets the first-pass statistics)
  S.calls = 0;
  S.variant = v;
  return 0;
}
extern "C" int rb2_fullsort_tc_set_kprime(int32_t kp) {
  if (kp != 0 && kp != 16 && kp != 32) return RB2_EINVAL;
  rb2_cur_scorer().kprime = kp;
  return 0;
}
extern "C" int32_t rb2_fullsort_tc_last_fallback_rows(void) { return rb2_cur_scorer().last_fallback_rows; }
extern "C" int32_t rb2_fullsort_tc_last_pass2_rows(void) { return rb2_cur_scorer().last_pass2_rows; }
// diagnostics: device buffer of 16 int64 per CTA that k_fullsort_tc fills with the cycles its producer / MMA /
// epilogue roles spent waiting on each barrier (nullptr = off)
extern "C" int rb2_fullsort_tc_set_trace(void *device_buffer) {
  rb2_cur_scorer().trace = device_buffer;
  return 0;
}

namespace {

constexpr int BM = 128;       // query rows per CTA tile (= TMEM lanes)
constexpr int BN = 256;       // items per tile = per B ring slot
constexpr int HN = 128;       // columns of every accumulator stage that one epilogue warp set drains
constexpr int NACC = 2;       // accumulator stages in TMEM (2 x 256 columns)
constexpr int BK = 64;        // bf16 elements per 128-byte swizzle row
constexpr int UNIT_BYTES = BN * BK * 2;   // one B ring slot: 256 items x 64 k
constexpr int A_KB_BYTES = BM * BK * 2;
constexpr int kThreadsTc = 384;          // warp group 0: producer warp, MMA warp, 2 idle; warp groups 1, 2: epilogue sets
constexpr int CAPB = 8;                  // per-row append buffer (entries) in front of the candidate list

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TC_DONE;\n"
      "bra TC_WAIT;\n"
      "TC_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// half of a B slot, written into BOTH CTAs of the pair (same smem offset, same mbarrier offset)
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t *bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// ---- CTA-pair (cta_group::2) MMAs: one M = 256 MMA spans both SMs, each CTA keeps its own 128 query rows
// and HALF of every B slot in shared memory (half the operand traffic per SM) ----
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in CTA 0
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, int c0, int c1,
                                                uint32_t leader_bar_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(leader_bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t *bar) {   // arrives on `bar` in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)0x3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta0(uint64_t *bar) {   // arrive on CTA 0's copy of `bar`
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, 0;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address, 16-byte units
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
// instruction descriptor: c=f32, a=b=bf16, both K-major, N=256, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// (c format: F32 = 1 << 4 / F16 = 0; a and b formats: BF16 = 1 << 7 | 1 << 10 / F16 = 0; N >> 3 at bit 17; M >> 4 at bit 24)
constexpr uint32_t make_idesc(bool acc_f32, bool in_bf16, int m) {
  return (acc_f32 ? (1u << 4) : 0u) | (in_bf16 ? ((1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
// fp16 operands, FP16 accumulator (c_format = a_format = b_format = 0): one 16-bit score in the low half of
// every 32-bit TMEM column (tools/mma_f16acc_check.cu)
constexpr uint32_t kIdescH16 = ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t kIdesc2H16 = ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

// 64 columns of 16-bit cells -> 32 registers (low half = even column, high half = odd column)
#define TC_LD32P(taddr, v)                                                                                      \
  asm volatile(                                                                                                 \
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, " \
      "%13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"  \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),           \
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),           \
        "=r"(v[30]), "=r"(v[31])                                                                               \
      : "r"(taddr))

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

struct TcParams {
  int64_t nq, n_local, item_base;
  int n_ut, n_split, tiles_per_split;   // work decomposition
  const int64_t *hist_indptr, *hist_indices;
  int *cand_ids;      // [n_split * 2][nq][KP]   (x2: one list per epilogue warp set)
  float *cand_sc;     // approximate scores; slot KP-1 of a list holds its minimum (the cut-off k_refine reads)
  long long *trace;   // diagnostics (rb2_fullsort_tc_set_trace), usually nullptr
  float *lse_m, *lse_s;   // LSE kernels: per (list, row) running max and sum of exp   [n_split * 2][nq]
  const int32_t *row_map; // second pass over the rows whose certificate failed: row of this pass -> caller's row
};
#define TC_TIMED(slot, stmt)                          \
  do {                                                \
    if (TRACE) {                                      \
      long long t_ = clock64();                       \
      stmt;                                           \
      tr[slot] += clock64() - t_;                     \
    } else {                                          \
      stmt;                                           \
    }                                                 \
  } while (0)

template <int KB, int NSTAGE, bool TWO_SM>
struct TcSmem {
  static constexpr size_t A_BYTES = (size_t)KB * A_KB_BYTES;
  static constexpr size_t B_BYTES = (size_t)NSTAGE * (TWO_SM ? UNIT_BYTES / 2 : UNIT_BYTES);
  static constexpr size_t CBUF_BYTES = (size_t)2 * CAPB * BM * 8;   // two warp sets x (score, id)
  static constexpr size_t TAU_BYTES = (size_t)2 * BM * 4;            // thresholds the two warp sets publish
  static constexpr size_t TOTAL =
      1024 /*align slack*/ + A_BYTES + B_BYTES + CBUF_BYTES + TAU_BYTES + 256 /*barriers*/;
};


// minimum of the (unsorted) candidate list and the slot that holds it: (value, slot) tournament tree
template <int KP>
__device__ __forceinline__ void list_min(const float (&ls)[KP], float &mn, int &amin) {
  float m[KP / 2];
  int a[KP / 2];
#pragma unroll
  for (int i = 0; i < KP / 2; ++i) {
    const bool lt = ls[2 * i + 1] < ls[2 * i];
    m[i] = lt ? ls[2 * i + 1] : ls[2 * i];
    a[i] = lt ? 2 * i + 1 : 2 * i;
  }
#pragma unroll
  for (int w = KP / 4; w > 0; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const bool lt = m[i + w] < m[i];
      m[i] = lt ? m[i + w] : m[i];
      a[i] = lt ? a[i + w] : a[i];
    }
  mn = m[0];
  amin = a[0];
}

// bit j of `m` |= (x > tau): FSETP + predicated LOP3
#define TC_MASK_GT_F32(m, x, tau, bit) \
  asm("{\n.reg .pred p;\nsetp.gt.f32 p, %1, %2;\n@p or.b32 %0, %0, %3;\n}\n" : "+r"(m) : "r"(x), "f"(tau), "r"(bit))
// packed pair: bit of `mlo` |= (low half > tau), bit of `mhi` |= (high half > tau)
#define TC_MASK_GT_F16X2(mlo, mhi, x, tau2, bit)                                                              \
  asm("{\n.reg .pred p, q;\nsetp.gt.f16x2 p|q, %2, %3;\n@p or.b32 %0, %0, %4;\n@q or.b32 %1, %1, %4;\n}\n" \
      : "+r"(mlo), "+r"(mhi) : "r"(x), "r"(tau2), "r"(bit))

// the registers a tcgen05.ld wrote are defined only after tcgen05.wait::ld: pin their first use behind it
#define TC_REGS_AFTER_WAIT(v)                                                                                    \
  asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), \
                    "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]),      \
                    "+r"(v[15]));                                                                                 \
  asm volatile("" : "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]),    \
                    "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),    \
                    "+r"(v[30]), "+r"(v[31]))

// v[j] for a run-time j without local memory: select tree over the bits of j (31 SELs)
__device__ __forceinline__ float pick32(const uint32_t (&v)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
  const bool b0 = j & 1, b1 = j & 2, b2 = j & 4, b3 = j & 8, b4 = j & 16;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = b0 ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = b1 ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = b2 ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = b3 ? c[2 * i + 1] : c[2 * i];
  return __uint_as_float(b4 ? d[1] : d[0]);
}

__device__ __forceinline__ float max32(const uint32_t (&v)[32]) {
  float m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i)
    m[i] = fmaxf(fmaxf(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1])), __uint_as_float(v[3 * i + 2]));
  m[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
  float a = fmaxf(fmaxf(m[0], m[1]), m[2]), b = fmaxf(fmaxf(m[3], m[4]), m[5]), c = fmaxf(fmaxf(m[6], m[7]), m[8]);
  return fmaxf(fmaxf(fmaxf(a, b), c), fmaxf(m[9], m[10]));
}

__device__ __forceinline__ __half2 as_h2(uint32_t w) { return *reinterpret_cast<__half2 *>(&w); }
__device__ __forceinline__ float ex2_approx(float x) {   // MUFU.EX2; ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// max over the 64 fp16 scores packed in 32 registers (HMNMX2 tree)
__device__ __forceinline__ float max64h(const uint32_t (&v)[32]) {
  __half2 m[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = __hmax2(as_h2(v[2 * i]), as_h2(v[2 * i + 1]));
#pragma unroll
  for (int w = 8; w > 0; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) m[i] = __hmax2(m[i], m[i + w]);
  return fmaxf(__low2float(m[0]), __high2float(m[0]));
}
__device__ __forceinline__ uint32_t pick32u(const uint32_t (&v)[32], int j) {
  return __float_as_uint(pick32(v, j));
}

// One CTA = 128 query rows; CTAs run as pairs (cluster of 2) that walk the same item tiles for two query
// tiles: each loads half of every B slot and multicasts it to both.  Two 256-column accumulator stages;
// warp set h drains columns [128 h, +128) of EVERY tile and hands the stage back to the MMA issuer as soon
// as its scores sit in registers, before they are examined (barrier-wait traces, tools/tc_trace.py: when a
// set owned a whole stage and released it after the examination, the tile time was drain + signalling
// latency, 2200-2500 cycles against 1024 of MMA).
//   TWO_SM = false: cta_group::1, each CTA issues its own M = 128 MMAs on full B slots (32 KB) that the pair
//                   loads by halves and multicasts.
//   TWO_SM = true : cta_group::2, CTA 0 issues one M = 256 MMA per k-step for the pair; each CTA keeps only
//                   its half of every B slot (16 KB).  The per-CTA MMAs read and write every B byte through
//                   shared memory once per 1024 MMA-cycles (~128 B/clk, the whole shared-memory bandwidth);
//                   splitting B halves that.
template <int KB, int NSTAGE, int KP, bool H16, bool TWO_SM, bool TRACE, bool LSE = false, bool F16IN = H16>
__global__ void __launch_bounds__(kThreadsTc, 1)
k_fullsort_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
  constexpr int SLOT_BYTES = TWO_SM ? UNIT_BYTES / 2 : UNIT_BYTES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *sA = smem;                                  // [KB][128 rows][128 B]
  unsigned char *sB = sA + TcSmem<KB, NSTAGE, TWO_SM>::A_BYTES;      // [NSTAGE][256 rows][128 B]
  float *cbuf_s = reinterpret_cast<float *>(sB + TcSmem<KB, NSTAGE, TWO_SM>::B_BYTES);          // [2][CAPB][BM]
  int *cbuf_i = reinterpret_cast<int *>(cbuf_s + 2 * CAPB * BM);                                // [2][CAPB][BM]
  volatile float *tau_pub = reinterpret_cast<float *>(cbuf_i + 2 * CAPB * BM);                  // [2][BM]
  uint64_t *bars = reinterpret_cast<uint64_t *>(const_cast<float *>(tau_pub) + 2 * BM);
  uint64_t *full = bars;                 // [NSTAGE]
  uint64_t *empty = bars + NSTAGE;       // [NSTAGE]
  uint64_t *a_full = bars + 2 * NSTAGE;  // [1]
  uint64_t *a_empty = a_full + 1;        // [1]
  uint64_t *t_full = a_empty + 1;        // [NACC]
  uint64_t *t_empty = t_full + NACC;     // [NACC]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + NACC);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int crank = (int)cluster_ctarank();
  const int n_utp = (p.n_ut + 1) / 2;                 // query-tile pairs
  const int n_work = n_utp * p.n_split;               // work items per CLUSTER
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int n_tiles_all = (int)((p.n_local + BN - 1) / BN);

  if (threadIdx.x == 0) {
    // empty: both CTAs' MMA threads commit to it (in both CTAs)
    // (2-SM: one multicast commit from CTA 0; both CTAs' epilogue threads arrive on CTA 0's t_empty)
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TWO_SM ? 1 : 2); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int q = 0; q < NACC; ++q) { mbar_init(&t_full[q], 1); mbar_init(&t_empty[q], TWO_SM ? 512 : 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns (two 256-column accumulator stages)
    if (TWO_SM) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // registers move from the producer / MMA warp group to the two epilogue warp groups (candidate list +
  // two chunks of scores per thread): 4 x 56 + 8 x 224 registers per lane <= 64 K
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      long long tr[16] = {0}, t_begin = TRACE ? clock64() : 0;
      for (int w = cluster_id; w < n_work; w += n_clusters) {
        const int ut = 2 * (w % n_utp) + crank, sp = w / n_utp;
        TC_TIMED(1, mbar_wait(a_empty, a_phase ^ 1));
        if (TWO_SM) {
          // both CTAs' A tiles report to CTA 0's barrier (the MMA issuer lives there)
          if (crank == 0) mbar_expect_tx(a_full, 2 * KB * A_KB_BYTES);
          for (int kb = 0; kb < KB; ++kb)
            tma_load_2d_2sm(sA + kb * A_KB_BYTES, &tmA, kb * BK, ut * BM, smem_u32(a_full) & kPeerBitMask);
        } else {
          mbar_expect_tx(a_full, KB * A_KB_BYTES);
          for (int kb = 0; kb < KB; ++kb) tma_load_2d(sA + kb * A_KB_BYTES, &tmA, kb * BK, ut * BM, a_full);
        }
        a_phase ^= 1;
        const int t0 = sp * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, n_tiles_all);
        for (int it = t0; it < t1; ++it) {
          for (int kb = 0; kb < KB; ++kb) {
            TC_TIMED(0, mbar_wait(&empty[stage], phase ^ 1));       // both CTAs are done reading this slot
            if (TWO_SM) {
              if (crank == 0) mbar_expect_tx(&full[stage], UNIT_BYTES);   // my half + the peer's half
              tma_load_2d_2sm(sB + (size_t)stage * SLOT_BYTES, &tmB, kb * BK, it * BN + crank * (BN / 2),
                              smem_u32(&full[stage]) & kPeerBitMask);
            } else {
              mbar_expect_tx(&full[stage], UNIT_BYTES);  // my half + the peer's half
              tma_load_2d_mc(sB + (size_t)stage * UNIT_BYTES + (size_t)crank * (UNIT_BYTES / 2), &tmB, kb * BK,
                             it * BN + crank * (BN / 2), &full[stage], (uint16_t)0x3);
            }
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (TRACE) {
        long long *o = p.trace + (size_t)blockIdx.x * 16;
        o[0] = tr[0]; o[1] = tr[1]; o[11] = clock64() - t_begin;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread; 2-SM: only CTA 0's) =====================
    if (lane == 0 && (!TWO_SM || crank == 0)) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0, tcount = 0;
      long long tr[16] = {0}, t_begin = TRACE ? clock64() : 0;
      for (int w = cluster_id; w < n_work; w += n_clusters) {
        const int sp = w / n_utp;
        TC_TIMED(4, mbar_wait(a_full, a_phase));
        a_phase ^= 1;
        const int t0 = sp * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, n_tiles_all);
        for (int it = t0; it < t1; ++it, ++tcount) {
          long long *ev = (TRACE && blockIdx.x == 0 && tcount >= 1000u && tcount < 1064u)
                              ? p.trace + 148 * 16 + (size_t)(tcount - 1000u) * 16 : nullptr;
          const int q = (int)(tcount & 1u);
          TC_TIMED(3, mbar_wait(&t_empty[q], ((tcount >> 1) & 1) ^ 1));
          tc_fence_after();
          if (ev) ev[0] = clock64();
          const uint32_t tmem_d = tmem_base + (uint32_t)(q * BN);
          for (int kb = 0; kb < KB; ++kb) {
            TC_TIMED(2, mbar_wait(&full[stage], phase));
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(smem_u32(sA + kb * A_KB_BYTES));
            const uint64_t bdesc = make_smem_desc(smem_u32(sB + (size_t)stage * SLOT_BYTES));
#pragma unroll
            for (int k4 = 0; k4 < BK / 16; ++k4) {
              // advance 16 elements = 32 bytes inside the swizzle row: +2 in 16-byte units
              if (TWO_SM)
                tc_mma_2sm(tmem_d, adesc + (uint64_t)(2 * k4), bdesc + (uint64_t)(2 * k4),
                           make_idesc(!H16, !F16IN, 2 * BM), (kb | k4) ? 1u : 0u);
              else
                tc_mma_bf16(tmem_d, adesc + (uint64_t)(2 * k4), bdesc + (uint64_t)(2 * k4),
                            make_idesc(!H16, !F16IN, BM), (kb | k4) ? 1u : 0u);
            }
            // the slot is free for both producers once these MMAs have read it
            if (TWO_SM) tc_commit_2sm(&empty[stage]); else tc_commit_mc(&empty[stage], (uint16_t)0x3);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
          if (TWO_SM) tc_commit_2sm(&t_full[q]); else tc_commit(&t_full[q]);      // accumulator stage complete
          if (ev) ev[2] = clock64();
        }
        if (TWO_SM) tc_commit_2sm(a_empty); else tc_commit(a_empty);   // every MMA reading this A tile completed
      }
      if (TRACE) {
        long long *o = p.trace + (size_t)blockIdx.x * 16;
        o[2] = tr[2]; o[3] = tr[3]; o[4] = tr[4]; o[5] = clock64() - t_begin; o[10] = tcount;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
    // ===================== epilogue: 2 warp sets x 4 warps; thread <-> TMEM lane <-> query row ========
    const int ws = (warp - 4) >> 2;               // warp set = the half of every tile it drains
    const int quarter = warp & 3;                 // TMEM lanes this warp may touch: [32*quarter, +32)
    const int t = quarter * 32 + lane;            // row inside the tile
    uint32_t tcount = 0;
    long long tr[16] = {0};
    for (int w = cluster_id; w < n_work; w += n_clusters) {
      const int ut = 2 * (w % n_utp) + crank, sp = w / n_utp;
      const int64_t r = (int64_t)ut * BM + t;
      const bool active = r < p.nq;
      const int64_t *hist = nullptr;
      int64_t hlen = 0;
      if (active && p.hist_indptr) {
        const int64_t ro = p.row_map ? (int64_t)p.row_map[r] : r;
        int64_t h0 = p.hist_indptr[ro];
        hlen = p.hist_indptr[ro + 1] - h0;
        hist = p.hist_indices + h0;
      }
      // Candidate list: KP (score, id) pairs in registers, UNSORTED; `tau_list` = its minimum, held in slot
      // `amin` (-inf while the list is not full).  A score above the filter threshold `tau` is only
      // APPENDED to a small per-row buffer in shared memory (a few instructions; 32 rows share a warp, so
      // whatever one row does here stalls the other 31); the buffers are folded into the lists by all
      // lanes together when one of them is full.  tau is therefore a little stale (low) between two folds
      // -- more appends, each far cheaper than keeping the list exact.  The two warp sets hold separate
      // lists for the same rows (the two halves of every tile) and publish their list minima: k_refine's
      // certificate cuts at the MAXIMUM over a row's lists anyway, so each set filters with the larger of
      // the two and the pair behaves like one list.  Everything ever dropped had approx <= that maximum.
      float ls[KP];
      int li[KP];
#pragma unroll
      for (int j = 0; j < KP; ++j) { ls[j] = -INFINITY; li[j] = -1; }
      float tau = active ? -INFINITY : INFINITY;   // inactive rows never append
      float tau_list = -INFINITY;
      int amin = 0, cnt = 0;
      float run_m = -INFINITY, run_s = 0.f;        // LSE: online logsumexp over EVERY column of my half tiles
      float *cb_s = cbuf_s + (size_t)ws * CAPB * BM + t;   // entry e at [e * BM]
      int *cb_i = cbuf_i + (size_t)ws * CAPB * BM + t;
      volatile float *tau_mine = tau_pub + ws * BM + t, *tau_other = tau_pub + (ws ^ 1) * BM + t;
      // everyone is done with the previous work item's filter and thresholds
      asm volatile("bar.sync 1, 256;" ::: "memory");
      *tau_mine = -INFINITY;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // History exclusion (trainer.py:344-345) without a search: a set sees its items in ascending order
      // and the CSR row is sorted, so a cursor into the row (and the id under it, in a register) answers
      // "seen in training?" for every passing score; it only ever moves forward.
      int64_t hcur = 0;
      int next_h = hlen > 0 ? (int)hist[0] : 0x7fffffff;
      const int64_t item_limit = p.item_base + p.n_local;
      const int t0 = sp * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, n_tiles_all);

      auto valid = [&](int item) -> bool {
        if (item == 0 || (int64_t)item >= item_limit) return false;      // [PAD] / zero-filled rows past the table
        if (next_h < item) {
          while (hcur + 16 < hlen && (int)hist[hcur + 16] < item) hcur += 16;   // gallop over long histories
          do { ++hcur; } while (hcur < hlen && (int)hist[hcur] < item);
          next_h = hcur < hlen ? (int)hist[hcur] : 0x7fffffff;
        }
        return next_h != item;
      };
      auto fold = [&]() {
        for (int e = 0; e < cnt; ++e) {
          const float s = cb_s[e * BM];
          const int id = cb_i[e * BM];
          if (s > tau) {
#pragma unroll
            for (int j = 0; j < KP; ++j) {   // replace the minimum
              const bool hit = j == amin;
              ls[j] = hit ? s : ls[j];
              li[j] = hit ? id : li[j];
            }
            list_min<KP>(ls, tau_list, amin);
            tau = fmaxf(tau, tau_list);
          }
        }
        cnt = 0;
        *tau_mine = tau_list;
      };
      // one chunk of scores in registers: 32 fp32 columns, or 64 fp16 columns packed two per register.
      // Latency matters more than throughput here (the stage cannot be refilled before every warp has
      // fetched its share of the NEXT one): the flags of the passing scores are gathered with two
      // instructions per register over four independent accumulators, and the usual case -- exactly one
      // passing score, which is the maximum already known -- skips the register select tree.
      auto process = [&](uint32_t (&v)[32], int64_t gbase) {
        if (LSE) {
          // zero-filled rows past the end of the table are not classes of the softmax
          const int64_t n_ok = item_limit - gbase;
          if (n_ok < 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j >= n_ok) v[j] = 0xff800000u;   // -inf
          }
        }
        const float m = H16 ? max64h(v) : max32(v);
        if (LSE && m != -INFINITY) {
          constexpr float kLog2e = 1.4426950408889634f;
          if (m > run_m) {   // rare after the first tiles
            run_s *= exp2f((run_m - m) * kLog2e);
            run_m = m;
          }
          const float mb = run_m * kLog2e;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            a0 += ex2_approx(fmaf(__uint_as_float(v[j]), kLog2e, -mb));
            a1 += ex2_approx(fmaf(__uint_as_float(v[j + 1]), kLog2e, -mb));
            a2 += ex2_approx(fmaf(__uint_as_float(v[j + 2]), kLog2e, -mb));
            a3 += ex2_approx(fmaf(__uint_as_float(v[j + 3]), kLog2e, -mb));
          }
          run_s += (a0 + a1) + (a2 + a3);
        }
        if (!__any_sync(0xffffffffu, m > tau)) return;           // the common case
        // fp16: fa = registers 0-15, fb = registers 16-31; bit j = even column of register j, bit 16 + j =
        // odd column.  fp32: fa bit j = column j.
        uint32_t fa = 0u, fb = 0u;
        if (m > tau) {
          uint32_t f0 = 0u, f1 = 0u, f2 = 0u, f3 = 0u;
          if (H16) {
            const __half2 th = __float2half2_rn(tau);            // exact: tau is +-inf or an fp16 score
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              f0 |= __hgt2_mask(as_h2(v[j]), th) & (0x00010001u << j);
              f1 |= __hgt2_mask(as_h2(v[j + 1]), th) & (0x00010001u << (j + 1));
              f2 |= __hgt2_mask(as_h2(v[16 + j]), th) & (0x00010001u << j);
              f3 |= __hgt2_mask(as_h2(v[17 + j]), th) & (0x00010001u << (j + 1));
            }
            fa = f0 | f1;
            fb = f2 | f3;
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              TC_MASK_GT_F32(f0, v[j], tau, 1u << j);
              TC_MASK_GT_F32(f1, v[j + 1], tau, 1u << (j + 1));
              TC_MASK_GT_F32(f2, v[j + 2], tau, 1u << (j + 2));
              TC_MASK_GT_F32(f3, v[j + 3], tau, 1u << (j + 3));
            }
            fa = (f0 | f1) | (f2 | f3);
          }
        }
        const bool single = (__popc(fa) + __popc(fb)) == 1;
        for (;;) {
          bool overflow = false;
          while (fa | fb) {                                       // ascending item order (the history cursor needs it)
            if (cnt == CAPB) { overflow = true; break; }        // keep the flag: resumed after the fold
            const bool second = fa == 0u;
            const uint32_t f = second ? fb : fa;
            int col, bit;
            float sc;
            if (H16) {
              const uint32_t fe = f & 0xffffu, fo = f >> 16;
              const int je = fe ? __ffs(fe) - 1 : 99, jo = fo ? __ffs(fo) - 1 : 99;
              const bool odd = jo < je;                           // column 2 jo + 1 < 2 je
              const int reg = (odd ? jo : je) + (second ? 16 : 0);
              bit = (odd ? jo + 16 : je);
              col = 2 * reg + (odd ? 1 : 0);
              if (single) {
                sc = m;
              } else {
                const uint32_t wv = pick32u(v, reg);
                sc = odd ? __high2float(as_h2(wv)) : __low2float(as_h2(wv));
              }
            } else {
              bit = __ffs(f) - 1;
              col = bit;
              sc = single ? m : pick32(v, bit);
            }
            const int item = (int)(gbase + col);
            if (valid(item)) {
              cb_s[cnt * BM] = sc;
              cb_i[cnt * BM] = item;
              ++cnt;
            }
            if (second) fb &= ~(1u << bit); else fa &= ~(1u << bit);
          }
          if (!__any_sync(0xffffffffu, overflow)) break;
          fold();                                                 // every lane folds what it has
        }
      };

      for (int it = t0; it < t1; ++it, ++tcount) {
        const int q = (int)(tcount & 1u);
        TC_TIMED(6, mbar_wait(&t_full[q], (tcount >> 1) & 1));
        tc_fence_after();
        tau = fmaxf(tau, *tau_other);
        const long long t_drain = TRACE ? clock64() : 0;
        long long *ev = (TRACE && blockIdx.x == 0 && ws == 0 && quarter == 0 && lane == 0 && tcount >= 1000u && tcount < 1064u)
                            ? p.trace + 148 * 16 + (size_t)(tcount - 1000u) * 16 : nullptr;
        if (ev) ev[4] = t_drain;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(q * BN + ws * HN);
        const int64_t g0 = p.item_base + (int64_t)it * BN + ws * HN;
        constexpr int CH = H16 ? 64 : 32;      // accumulator columns per load
        constexpr int NCH = HN / CH;           // 2 (fp16: the whole stage sits in registers) or 4
        uint32_t va[32], vb[32];
        if (H16) { TC_LD32P(taddr, va); TC_LD32P(taddr + CH, vb); } else { TC_LD32(taddr, va); TC_LD32(taddr + CH, vb); }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        TC_REGS_AFTER_WAIT(va);
        TC_REGS_AFTER_WAIT(vb);
#pragma unroll 1
        for (int c = 0; c < NCH; c += 2) {
          if (c + 2 >= NCH) {
            // the rest of the stage is in registers: hand it back before looking at the scores
            tc_fence_before();
            if (TWO_SM) mbar_arrive_cta0(&t_empty[q]); else mbar_arrive(&t_empty[q]);
            if (ev) ev[5] = clock64();
          }
          process(va, g0 + c * CH);
          if (c + 2 < NCH) TC_LD32(taddr + (c + 2) * CH, va);
          process(vb, g0 + (c + 1) * CH);
          if (c + 2 < NCH) {
            TC_LD32(taddr + (c + 3) * CH, vb);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            TC_REGS_AFTER_WAIT(va);
            TC_REGS_AFTER_WAIT(vb);
          }
        }
        if (TRACE) tr[7] += clock64() - t_drain;
        if (ev) ev[6] = clock64();
      }
      __syncwarp();
      fold();
      if (LSE && active) {
        p.lse_m[((int64_t)sp * 2 + ws) * p.nq + r] = run_m;
        p.lse_s[((int64_t)sp * 2 + ws) * p.nq + r] = run_s;
      }
      if (active) {
        // slot KP-1 of the output holds the list minimum (k_refine reads the cut-off there)
        int64_t o = (((int64_t)sp * 2 + ws) * p.nq + r) * KP;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          const int pos = (j == amin) ? KP - 1 : ((j == KP - 1) ? amin : j);
          p.cand_ids[o + pos] = li[j];
          p.cand_sc[o + pos] = ls[j];
        }
      }
    }
    if (TRACE && quarter == 0 && lane == 0) {   // one thread of each epilogue warp set
      long long *o = p.trace + (size_t)blockIdx.x * 16 + 6 + 2 * ws;
      o[0] = tr[6]; o[1] = tr[7];
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody leaves while the peer may still write into this CTA's smem / barriers
  if (warp == 1) {
    tc_fence_after();
    if (TWO_SM) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 -> 16-bit rows (+ norms).  One lane group per row.
//   H16 = false: bf16, no scaling.
//   H16 = true : fp16 after an exact power-of-two rescale so that every score fits the FP16 accumulator:
//                queries by their own 2^-e (||u * scale|| in [0.5, 1)), items by the table-wide `*gscale`
//                (max ||v * scale|| in [0.5, 1)).  Norms are those of the SCALED rows.  row_acc[r] =
//                sum over the d/16 MMA k-steps of ||h(u)[0 : 16 j]|| bounds the sum of the partial
//                accumulations the tensor core rounds to fp16.
// maxima of non-negative floats, compared as integers (a NaN stays on top and fails every certificate):
// warp shuffle, shared-memory atomic, ONE global atomic per block.  Every thread of the block calls this.
__device__ __forceinline__ void block_max3(int a, int b, int c, float *ga, float *gb, float *gc) {
  __shared__ int sm[3];
  if (threadIdx.x == 0) sm[0] = sm[1] = sm[2] = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = max(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
  }
  if (threadIdx.x % 32 == 0) { atomicMax(&sm[0], a); atomicMax(&sm[1], b); atomicMax(&sm[2], c); }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (ga) atomicMax(reinterpret_cast<int *>(ga), sm[0]);
    if (gb) atomicMax(reinterpret_cast<int *>(gb), sm[1]);
    if (gc) atomicMax(reinterpret_cast<int *>(gc), sm[2]);
  }
}
constexpr int kConvertBlocks = 148 * 16;

__device__ __forceinline__ float pow2_scale(float norm) {
  if (!(norm > 0.f) || !isfinite(norm)) return 1.f;
  int e;
  frexpf(norm, &e);                 // norm = m * 2^e, m in [0.5, 1)
  e = max(min(e, 120), -120);
  return ldexpf(1.f, -e);
}

template <int D>
__global__ void __launch_bounds__(256) k_rows_max_sqnorm(const float *__restrict__ src, int64_t rows,
                                                          float *__restrict__ out_max_sq) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));
  const int64_t stride = (int64_t)gridDim.x * blockDim.x / LANES;
  int mx = 0;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES; r < rows; r += stride) {
    Row<D> x = row_ldg<D>(src, r, lane);
    float n2 = group_sum<LANES>(row_dot_lane<D>(x, x), gmask);
    mx = max(mx, __float_as_int(n2));
  }
  block_max3(mx, 0, 0, out_max_sq, nullptr, nullptr);
}
__global__ void k_item_scale(float *maxes) { maxes[2] = pow2_scale(sqrtf(maxes[3])); }

template <int D, bool H16>
__global__ void __launch_bounds__(256) k_convert_rows(const float *__restrict__ src, const int64_t *__restrict__ ids,
                                                       const int32_t *__restrict__ row_map, int64_t rows,
                                                       int64_t src_rows, uint16_t *__restrict__ dst,
                                                       float *__restrict__ row_norm, float *__restrict__ row_dnorm,
                                                       float *__restrict__ max_bnorm, float *__restrict__ max_dnorm,
                                                       const float *__restrict__ gscale, float *__restrict__ row_scale,
                                                       float *__restrict__ row_acc) {
  constexpr int LANES = RowCfg<D>::LANES;
  constexpr int VPL = RowCfg<D>::VPL;
  static_assert(!H16 || VPL == 1, "the fp16 path covers d <= 128");
  const int lane = threadIdx.x % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));
  const int64_t stride = (int64_t)gridDim.x * blockDim.x / LANES;
  int mx_b = 0, mx_d = 0;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES; r < rows; r += stride) {
  const int64_t rr = row_map ? (int64_t)row_map[r] : r;
  int64_t sr = ids ? min(max(ids[rr], (int64_t)0), src_rows - 1) : rr;
  Row<D> x = row_ldg<D>(src, sr, lane);
  float scale = 1.f;
  if (H16) scale = gscale ? *gscale : pow2_scale(sqrtf(group_sum<LANES>(row_dot_lane<D>(x, x), gmask)));
  float n2 = 0.f, d2 = 0.f, b2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float f[4] = {x.v[i].x * scale, x.v[i].y * scale, x.v[i].z * scale, x.v[i].w * scale};
    uint16_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float back;
      if (H16) {
        __half hh = __float2half_rn(f[e]);
        h[e] = __half_as_ushort(hh);
        back = __half2float(hh);
      } else {
        __nv_bfloat16 hb = __float2bfloat16_rn(f[e]);
        h[e] = __bfloat16_as_ushort(hb);
        back = __bfloat162float(hb);
      }
      n2 = fmaf(f[e], f[e], n2);
      b2 = fmaf(back, back, b2);
      d2 = fmaf(f[e] - back, f[e] - back, d2);
    }
    uint2 packed;
    packed.x = (uint32_t)h[0] | ((uint32_t)h[1] << 16);
    packed.y = (uint32_t)h[2] | ((uint32_t)h[3] << 16);
    reinterpret_cast<uint2 *>(dst + r * D)[i * LANES + lane] = packed;
  }
  float acc = 0.f;
  if (H16 && row_acc) {
    // prefix sums of ||h(u)||^2 over the lanes (4 lanes = one K=16 MMA step)
    float pre = b2;
#pragma unroll
    for (int o = 1; o < LANES; o <<= 1) {
      float t = __shfl_up_sync(gmask, pre, o, LANES);
      if (lane >= o) pre += t;
    }
    acc = group_sum<LANES>((lane % 4 == 3) ? sqrtf(pre) * 1.0001f : 0.f, gmask);
  }
  n2 = group_sum<LANES>(n2, gmask);
  d2 = group_sum<LANES>(d2, gmask);
  b2 = group_sum<LANES>(b2, gmask);
  if (lane == 0) {
    // round the norms UP so the certificate stays rigorous
    float nn = sqrtf(n2) * 1.0001f, dd = sqrtf(d2) * 1.0001f, bb = sqrtf(b2) * 1.0001f;
    if (row_norm) row_norm[r] = nn;
    if (row_dnorm) row_dnorm[r] = dd;
    if (row_scale) row_scale[r] = scale;
    if (row_acc) row_acc[r] = acc;
    mx_b = max(mx_b, __float_as_int(bb));
    mx_d = max(mx_d, __float_as_int(dd));
  }
  }
  if (max_bnorm) block_max3(mx_b, mx_d, 0, max_bnorm, max_dnorm, nullptr);   // (uniform: a kernel argument)
}

// ---------------------------------------------------------------------------------------------
// Split-precision operands for the CE head: x = hi + lo + dd with hi = bf16(x), lo = bf16(x - hi).  A row of
// d = 64 becomes three 64-wide k-blocks -- queries [hi | hi | lo], items [hi | lo | hi] -- so that ONE
// K = 192 GEMM with fp32 accumulators yields hi.hi + hi.lo + lo.hi = x.e - (lo.lo + dd.e + x.dd'), i.e.
// logits good to ~2^-16 of ||x|| ||e|| (the logsumexp needs 1e-5).  Norms for the certificate:
// row_norm = ||x||, row_lo = ||lo||, row_dd = ||x - hi - lo||; items: maxima of the same three.
template <int D>
__global__ void __launch_bounds__(256) k_convert_split(const float *__restrict__ src, int64_t rows, bool item_side,
                                                        uint16_t *__restrict__ dst, float *__restrict__ row_norm,
                                                        float *__restrict__ row_lo, float *__restrict__ row_dd,
                                                        float *__restrict__ maxes) {
  constexpr int LANES = RowCfg<D>::LANES;
  static_assert(RowCfg<D>::VPL == 1, "d <= 128");
  const int lane = threadIdx.x % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));
  const int64_t stride = (int64_t)gridDim.x * blockDim.x / LANES;
  int mx_n = 0, mx_d = 0, mx_l = 0;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES; r < rows; r += stride) {
  Row<D> x = row_ldg<D>(src, r, lane);
  float f[4] = {x.v[0].x, x.v[0].y, x.v[0].z, x.v[0].w};
  uint16_t hi[4], lo[4];
  float n2 = 0.f, l2 = 0.f, d2 = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat16 h = __float2bfloat16_rn(f[e]);
    float hb = __bfloat162float(h);
    __nv_bfloat16 l = __float2bfloat16_rn(f[e] - hb);
    float lb = __bfloat162float(l);
    float dd = (f[e] - hb) - lb;
    hi[e] = __bfloat16_as_ushort(h);
    lo[e] = __bfloat16_as_ushort(l);
    n2 = fmaf(f[e], f[e], n2);
    l2 = fmaf(lb, lb, l2);
    d2 = fmaf(dd, dd, d2);
  }
  uint2 ph, pl;
  ph.x = (uint32_t)hi[0] | ((uint32_t)hi[1] << 16); ph.y = (uint32_t)hi[2] | ((uint32_t)hi[3] << 16);
  pl.x = (uint32_t)lo[0] | ((uint32_t)lo[1] << 16); pl.y = (uint32_t)lo[2] | ((uint32_t)lo[3] << 16);
  uint2 *o = reinterpret_cast<uint2 *>(dst + r * (3 * D));
  o[lane] = ph;
  o[LANES + lane] = item_side ? pl : ph;
  o[2 * LANES + lane] = item_side ? ph : pl;
  n2 = group_sum<LANES>(n2, gmask);
  l2 = group_sum<LANES>(l2, gmask);
  d2 = group_sum<LANES>(d2, gmask);
  if (lane == 0) {
    float nn = sqrtf(n2) * 1.0001f, ll = sqrtf(l2) * 1.0001f, dd = sqrtf(d2) * 1.0001f;   // rounded up
    if (row_norm) { row_norm[r] = nn; row_lo[r] = ll; row_dd[r] = dd; }
    mx_n = max(mx_n, __float_as_int(nn));
    mx_d = max(mx_d, __float_as_int(dd));
    mx_l = max(mx_l, __float_as_int(ll));
  }
  }
  if (maxes) block_max3(mx_n, mx_d, mx_l, maxes, maxes + 1, maxes + 2);
}

// ---------------------------------------------------------------------------------------------
// exact re-score + order + certificate.  One warp per query row; `parts` sorted lists of KP
// candidates each (parts * KP <= 1024).
// EMODE: 0 = bf16 operands / fp32 accumulators, 1 = rescaled fp16 operands / FP16 accumulators,
//        3 = rescaled fp16 operands / fp32 accumulators (second pass: no accumulate term in E),
//        2 = split bf16 operands (k_convert_split): qscale = ||lo(u)||, qacc = ||dd(u)||, qdnorm unused,
//            maxes = {max ||v||, max ||dd(v)||, max ||lo(v)||}
template <int D, int KP, int EMODE, int MAXC>
__global__ void __launch_bounds__(128) k_refine(const float *__restrict__ query_p, const int64_t *__restrict__ query_ids,
                                                 int64_t nq, const float *__restrict__ item_p, int64_t item_base,
                                                 const int *__restrict__ cand_ids, const float *__restrict__ cand_sc,
                                                 int parts, int K, const float *__restrict__ qnorm,
                                                 const float *__restrict__ qdnorm, const float *__restrict__ maxes,
                                                 const float *__restrict__ qscale, const float *__restrict__ qacc,
                                                 const int32_t *__restrict__ row_map, int64_t *__restrict__ out_ids, float *__restrict__ out_scores,
                                                 int32_t *__restrict__ fail_rows, int32_t *__restrict__ fail_count) {
  // MAXC = candidates per lane (parts * KP <= 32 * MAXC)
  const int lane = threadIdx.x % 32;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  if (r >= nq) return;
  const int64_t ro = row_map ? (int64_t)row_map[r] : r;        // the caller's row (outputs, query id)
  const int64_t qrow = query_ids ? query_ids[ro] : ro;
  const float4 *q4 = reinterpret_cast<const float4 *>(query_p + qrow * D);
  const int total = parts * KP;
  float cs[MAXC];
  int ci[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) { cs[c] = -INFINITY; ci[c] = -1; }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    int idx = c * 32 + lane;
    if (idx < total) {
      int part = idx / KP, slot = idx % KP;
      int id = cand_ids[((int64_t)part * nq + r) * KP + slot];
      if (id >= 0) {
        const float4 *v4 = reinterpret_cast<const float4 *>(item_p + ((int64_t)id - item_base) * D);
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < D / 4; ++k) {
          float4 a = __ldg(q4 + k), b = __ldg(v4 + k);
          s = fmaf(a.x, b.x, s);
          s = fmaf(a.y, b.y, s);
          s = fmaf(a.z, b.z, s);
          s = fmaf(a.w, b.w, s);
        }
        if (s != -INFINITY && s == s) { cs[c] = s; ci[c] = id; }  // the exact path never selects -inf / NaN
      }
    }
  }
  // cut-off thresholds: a FULL list (tail valid) may have dropped items with approx <= its tail score
  float tau_max = -INFINITY;
  bool any_full = false;
  for (int part = lane; part < parts; part += 32) {
    int64_t o = ((int64_t)part * nq + r) * KP + (KP - 1);
    if (cand_ids[o] >= 0) { any_full = true; tau_max = fmaxf(tau_max, cand_sc[o]); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tau_max = fmaxf(tau_max, __shfl_xor_sync(0xffffffffu, tau_max, o));
  any_full = __any_sync(0xffffffffu, any_full);
  // K rounds of warp arg-max with the (score desc, id asc) order
  float kth = -INFINITY;
  int found = 0;
  for (int j = 0; j < K; ++j) {
    float bs = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      bool better = ci[c] >= 0 && (cs[c] > bs || (cs[c] == bs && ci[c] < bi));
      bs = better ? cs[c] : bs;
      bi = better ? ci[c] : bi;
    }
    float wsc = bs;
    int wi = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float os = __shfl_xor_sync(0xffffffffu, wsc, o);
      int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (oi != 0x7fffffff && (wi == 0x7fffffff || os > wsc || (os == wsc && oi < wi))) { wsc = os; wi = oi; }
    }
    const bool have = wi != 0x7fffffff;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (ci[c] == wi) ci[c] = -1;  // remove the winner everywhere (an id may sit in two lists of one row? no: lists partition the items)
    if (lane == 0) {
      out_ids[ro * K + j] = have ? (int64_t)wi : -1;
      out_scores[ro * K + j] = have ? wsc : -INFINITY;
    }
    if (have) { kth = wsc; ++found; }
  }
  if (lane == 0 && any_full) {
    // |approx - exact| <= ||du||*max||bv|| + ||u||*max||dv||  (+ fp32 accumulation slack)
    float E = (EMODE == 2) ? 0.f : qdnorm[r] * maxes[0] + qnorm[r] * maxes[1] + 1.6e-5f * qnorm[r] * maxes[0];
    float kth_s = kth;
    if (EMODE == 2) {
      // x.e - (hi.hi + hi.lo + lo.hi) = lo.lo' + dd.e + x.dd' (+ the K = 192 fp32 accumulation of the tensor core)
      E = qscale[r] * maxes[2] + qacc[r] * maxes[0] + qnorm[r] * maxes[1] + 3e-5f * qnorm[r] * maxes[0];
    }
    if (EMODE == 3) kth_s = kth * qscale[r] * maxes[2];         // rescaled fp16 operands, fp32 accumulators
    if (EMODE == 1) {
      // scores live in the rescaled domain (exact powers of two).  Every K=16 MMA rounds the running sum
      // to fp16 (round-to-nearest, bit-checked by tools/mma_f16acc_check.cu): |err_j| <= 2^-11 |acc_j|,
      // |acc_j| <= ||h(u)[0:16j]|| * max||h(v)||; the constant adds the subnormal / second-order slack
      kth_s = kth * qscale[r] * maxes[2];
      E += 1.03f * 4.8828125e-4f * qacc[r] * maxes[0] + 3e-5f;
    }
    bool ok = (found == K) && (kth_s > tau_max + E);
    if (!ok) {
      int slot = atomicAdd(fail_count, 1);
      fail_rows[slot] = (int32_t)ro;
    }
  }
}

struct TcWs {
  uint16_t *qb, *vb;              // bf16 or fp16 rows
  float *qnorm, *qdnorm, *qscale, *qacc;
  float *maxes;  // [0] = max ||16bit(v)||, [1] = max ||v - 16bit(v)||, [2] = item scale, [3] = max ||v||^2
  int *cand_ids;
  float *cand_sc;
  int32_t *fail_rows, *fail_rows2, *fail_count;
  int64_t *fb_ids;
  float *fb_sc;
  void *fp32_ws;
  size_t fp32_bytes;
  int64_t cand_lists;             // capacity of cand_ids / cand_sc in lists of KP_MAX
};

struct TcPlan {
  int n_ut, n_split, tiles_per_split, grid;
};

TcPlan make_plan(int64_t nq, int64_t n_local, int max_split = 16) {
  TcPlan pl;
  pl.n_ut = (int)((nq + BM - 1) / BM);
  const int n_tiles = (int)((n_local + BN - 1) / BN);
  const int pairs = (pl.n_ut + 1) / 2, clusters = rb2_num_sms() / 2;
  // Split the item range into `want` pieces so that the work items (query-tile pair x piece) fill the CTA
  // pairs in whole rounds: cost = rounds x (tiles per piece + the cost of starting a work item: every piece
  // warms its thresholds up again, K' ln(n / K') slow-path events per row ~ 150 tiles' worth of time).  Few
  // query tiles need many pieces; many need one.
  int best = 1;
  long best_cost = -1;
  for (int want = 1; want <= max_split && want <= n_tiles; ++want) {
    int tps = (n_tiles + want - 1) / want;
    int n_split = (n_tiles + tps - 1) / tps;
    long work = (long)pairs * n_split;
    long rounds = (work + clusters - 1) / clusters;
    long cost = rounds * (tps + 150);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = want; }
  }
  pl.tiles_per_split = (n_tiles + best - 1) / best;
  pl.n_split = (n_tiles + pl.tiles_per_split - 1) / pl.tiles_per_split;
  int work = pairs * pl.n_split;       // work items per CTA pair
  pl.grid = 2 * (work < clusters ? work : clusters);
  return pl;
}

constexpr int KP_MAX = 32;

size_t carve_tc(TcWs &w, void *base, int64_t nq, int64_t n_local, int dim, int k) {
  Carver c(base);
  TcPlan pl = make_plan(nq, n_local);
  int64_t nq_pad = (nq + 2 * BM - 1) / (2 * BM) * (2 * BM);
  w.qb = c.take<uint16_t>(nq_pad * dim);
  w.vb = c.take<uint16_t>(n_local * dim);
  w.qnorm = c.take<float>(nq);
  w.qdnorm = c.take<float>(nq);
  w.qscale = c.take<float>(nq);
  w.qacc = c.take<float>(nq);
  w.maxes = c.take<float>(4);
  w.cand_lists = (int64_t)pl.n_split * 2 * nq;
  w.cand_ids = c.take<int>((size_t)w.cand_lists * KP_MAX);
  w.cand_sc = c.take<float>((size_t)w.cand_lists * KP_MAX);
  w.fail_rows = c.take<int32_t>(nq);
  w.fail_rows2 = c.take<int32_t>(nq);
  w.fail_count = c.take<int32_t>(4);
  w.fb_ids = c.take<int64_t>(nq * k);
  w.fb_sc = c.take<float>(nq * k);
  w.fp32_bytes = rb2_fullsort_fp32_workspace(nq, n_local, dim, k);
  w.fp32_ws = c.take<char>(w.fp32_bytes);
  return c.off;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map(CUtensorMap *m, void *base, int64_t rows, int dim, int box_rows, bool half) {
  EncodeTiledFn fn = get_encode_fn();
  RB2_REQUIRE(fn != nullptr, RB2_EINVAL, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(m, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RB2_REQUIRE(rc == CUDA_SUCCESS, RB2_EINVAL, "cuTensorMapEncodeTiled failed with %d", (int)rc);
  return 0;
}

// One tensor-core pass over `nq` rows (the caller's rows row_map[0..nq) when row_map is given): conversion of the
// queries (and of the item shard unless a previous pass left it in w.vb / w.maxes), scorer, refine.  Rows whose
// certificate fails are appended (caller's numbering) to fail_rows / *fail_count.
template <int D, int KP, bool F16IN, bool H16, bool TWO_SM, bool TRACE>
int tc_pass(const TcWs &w, const float *query_p, const int64_t *query_ids, const int32_t *row_map, int64_t nq,
            bool convert_items, const float *item_p, int64_t n_local, int64_t item_base, const int64_t *hist_indptr,
            const int64_t *hist_indices, int k, int64_t *out_ids, float *out_scores, int32_t *fail_rows,
            int32_t *fail_count, cudaStream_t st) {
  constexpr int KB = D / BK;
  constexpr int NSTAGE = TWO_SM ? 10 : 5;   // B ring depth: what fits beside A, the append buffers and the barriers
  constexpr int EMODE = F16IN ? (H16 ? 1 : 3) : 0;
  static_assert(F16IN || !H16, "FP16 accumulators need the rescaled fp16 operands");
  // the candidate buffers were sized for the first pass: a later pass over fewer rows may not split finer
  int max_split = (int)(w.cand_lists / (2 * nq));
  if (max_split > 16) max_split = 16;
  if (max_split < 1) max_split = 1;
  TcPlan pl = make_plan(nq, n_local, max_split);
  constexpr int LANES = RowCfg<D>::LANES;
  const int64_t nq_pad = (nq + 2 * BM - 1) / (2 * BM) * (2 * BM);
  {
    ProfScope prof(RB2_ST_TC_CONVERT, st, 6);
    if (nq_pad > nq) RB2_CUDA(cudaMemsetAsync(w.qb + nq * D, 0, (size_t)(nq_pad - nq) * D * 2, st));
    const unsigned item_blocks = (unsigned)min((int64_t)kConvertBlocks, (n_local * LANES + 255) / 256);
    if (convert_items) {
      RB2_CUDA(cudaMemsetAsync(w.maxes, 0, 4 * sizeof(float), st));
      if (F16IN) {
        k_rows_max_sqnorm<D><<<item_blocks, 256, 0, st>>>(item_p, n_local, w.maxes + 3);
        k_item_scale<<<1, 1, 0, st>>>(w.maxes);
      }
    }
    k_convert_rows<D, F16IN><<<(unsigned)min((int64_t)kConvertBlocks, (nq * LANES + 255) / 256), 256, 0, st>>>(
        query_p, query_ids, row_map, nq, INT64_MAX, w.qb, w.qnorm, w.qdnorm, nullptr, nullptr, nullptr, w.qscale, w.qacc);
    if (convert_items)
      k_convert_rows<D, F16IN><<<item_blocks, 256, 0, st>>>(item_p, nullptr, nullptr, n_local, n_local, w.vb, nullptr,
                                                            nullptr, w.maxes, w.maxes + 1,
                                                            F16IN ? w.maxes + 2 : nullptr, nullptr, nullptr);
  }
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, w.qb, nq_pad, D, BM, F16IN);
  if (rc) return rc;
  rc = make_map(&tmB, w.vb, n_local, D, BN / 2, F16IN);   // each CTA of the pair loads half a slot
  if (rc) return rc;
  TcParams p;
  p.nq = nq; p.n_local = n_local; p.item_base = item_base;
  p.n_ut = pl.n_ut; p.n_split = pl.n_split; p.tiles_per_split = pl.tiles_per_split;
  p.hist_indptr = hist_indptr; p.hist_indices = hist_indices;
  p.cand_ids = w.cand_ids; p.cand_sc = w.cand_sc;
  p.trace = static_cast<long long *>(rb2_cur_scorer().trace);
  p.lse_m = nullptr; p.lse_s = nullptr;
  p.row_map = row_map;
  const size_t smem = TcSmem<KB, NSTAGE, TWO_SM>::TOTAL;
  {
    ProfScope prof(RB2_ST_TC_SCORE, st);
    auto kern = k_fullsort_tc<KB, NSTAGE, KP, H16, TWO_SM, TRACE, false, F16IN>;
    RB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)pl.grid);
    cfg.blockDim = dim3(kThreadsTc);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RB2_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
    RB2_CUDA(cudaGetLastError());
  }
  {
    ProfScope prof(RB2_ST_TC_REFINE, st);
    const int total = pl.n_split * 2 * KP;
#define RB2_REFINE(MAXC_)                                                                                             \
  k_refine<D, KP, EMODE, MAXC_><<<(unsigned)((nq * 32 + 127) / 128), 128, 0, st>>>(                                    \
      query_p, query_ids, nq, item_p, item_base, w.cand_ids, w.cand_sc, pl.n_split * 2, k, w.qnorm, w.qdnorm, w.maxes, \
      w.qscale, w.qacc, row_map, out_ids, out_scores, fail_rows, fail_count)
    if (total <= 32) RB2_REFINE(1);
    else if (total <= 64) RB2_REFINE(2);
    else if (total <= 128) RB2_REFINE(4);
    else if (total <= 256) RB2_REFINE(8);
    else RB2_REFINE(32);
#undef RB2_REFINE
    RB2_CUDA(cudaGetLastError());
  }
  return 0;
}

// The cascade: (1) all rows through the fast pass; (2) the rows whose certificate failed, again, with fp32
// accumulators -- same fp16 operands, an error bound ~10x tighter (the FP16-accumulate term dominates E) at
// ~85 % of the speed; (3) what still fails goes to the exact CUDA-core kernel.  Trained tables (a few popular
// items with large norms inflate max||v||) fail 2-3 % of the rows in (1) and ~10x fewer in (2).
template <int D, int KP, bool H16, bool TWO_SM, bool TRACE = false>
int run_tc(const float *query_p, const int64_t *query_ids, int64_t nq, const float *item_p, int64_t n_local,
           int64_t item_base, const int64_t *hist_indptr, const int64_t *hist_indices, int k, int64_t *out_ids,
           float *out_scores, void *workspace, size_t workspace_bytes, cudaStream_t st) {
  TcWs w;
  size_t need = carve_tc(w, workspace, nq, n_local, D, k);
  RB2_REQUIRE(workspace_bytes >= need, RB2_EWORKSPACE, "rb2_fullsort_topk(tc): workspace %zu < %zu", workspace_bytes, need);
  RB2_CUDA(cudaMemsetAsync(w.fail_count, 0, 4 * sizeof(int32_t), st));
  // When most first-pass certificates fail (well-trained tables: E is relative to max ||v||, the score gaps are
  // not) the FP16-accumulator pass is wasted work: the break-even is ~15 % failing rows.  Track the recent
  // failure fraction and start with fp32 accumulators when it is above that, probing again every 16th call.
  rb2_scorer_state &S = rb2_cur_scorer();
  const bool fast_first = H16 && (S.fail_ema < 0.15f || (S.calls++ % 16) == 15);
  int rc;
  if (fast_first || !H16)
    rc = tc_pass<D, KP, H16, H16, TWO_SM, TRACE>(w, query_p, query_ids, nullptr, nq, true, item_p, n_local, item_base,
                                                 hist_indptr, hist_indices, k, out_ids, out_scores, w.fail_rows,
                                                 w.fail_count, st);
  else
    rc = tc_pass<D, KP, true, false, false, false>(w, query_p, query_ids, nullptr, nq, true, item_p, n_local, item_base,
                                                   hist_indptr, hist_indices, k, out_ids, out_scores, w.fail_rows,
                                                   w.fail_count, st);
  if (rc) return rc;
  int32_t n_fail = 0;
  RB2_CUDA(cudaMemcpyAsync(&n_fail, w.fail_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  RB2_CUDA(cudaStreamSynchronize(st));
  S.last_pass2_rows = 0;
  const int32_t *final_rows = w.fail_rows;
  if (fast_first) {
    if (nq >= 1024) S.fail_ema = 0.5f * S.fail_ema + 0.5f * (float)n_fail / (float)nq;
    if (n_fail > 0) {
      S.last_pass2_rows = n_fail;
      rc = tc_pass<D, KP, true, false, false, false>(w, query_p, query_ids, w.fail_rows, n_fail, false, item_p, n_local,
                                                     item_base, hist_indptr, hist_indices, k, out_ids, out_scores,
                                                     w.fail_rows2, w.fail_count + 1, st);
      if (rc) return rc;
      RB2_CUDA(cudaMemcpyAsync(&n_fail, w.fail_count + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      RB2_CUDA(cudaStreamSynchronize(st));
      final_rows = w.fail_rows2;
    }
  }
  S.last_fallback_rows = n_fail;
  if (n_fail > 0) {   // redo exactly
    rc = rb2_fullsort_fp32(query_p, query_ids, n_fail, item_p, n_local, item_base, D, hist_indptr, hist_indices, k,
                           out_ids, out_scores, w.fp32_ws, w.fp32_bytes, st, final_rows);
    if (rc) return rc;
  }
  return 0;
}

// ---- CE head on the tensor cores (d = 64, k <= 16): split-precision logits, online logsumexp -------------
struct TcLseWs {
  uint16_t *qb, *vb;
  float *qnorm, *qlo, *qdd, *maxes;
  int *cand_ids;
  float *cand_sc;
  int32_t *fail_rows, *fail_count;
  void *fp32_ws;
  size_t fp32_bytes;
};
size_t carve_tc_lse(TcLseWs &w, void *base, int64_t nq, int64_t n_items, int k) {
  Carver c(base);
  TcPlan pl = make_plan(nq, n_items);
  int64_t nq_pad = (nq + 2 * BM - 1) / (2 * BM) * (2 * BM);
  w.qb = c.take<uint16_t>(nq_pad * 192);
  w.vb = c.take<uint16_t>(n_items * 192);
  w.qnorm = c.take<float>(nq);
  w.qlo = c.take<float>(nq);
  w.qdd = c.take<float>(nq);
  w.maxes = c.take<float>(4);
  w.cand_ids = c.take<int>((size_t)pl.n_split * 2 * nq * KP_MAX);
  w.cand_sc = c.take<float>((size_t)pl.n_split * 2 * nq * KP_MAX);
  w.fail_rows = c.take<int32_t>(nq);
  w.fail_count = c.take<int32_t>(4);
  w.fp32_bytes = rb2_fullsort_fp32_workspace(nq, n_items, 64, k);
  w.fp32_ws = c.take<char>(w.fp32_bytes);
  return c.off;
}

template <int KP>
int run_tc_lse(const float *x, int64_t nq, const float *item_p, int64_t n_items, int k, int64_t *out_ids,
               float *out_scores, void *workspace, size_t workspace_bytes, cudaStream_t st, float *lse_m, float *lse_s,
               int *parts_out) {
  constexpr int D = 64, KB = 3, NSTAGE = 5;
  TcLseWs w;
  size_t need = carve_tc_lse(w, workspace, nq, n_items, k);
  RB2_REQUIRE(workspace_bytes >= need, RB2_EWORKSPACE, "rb2_ce_head(tc): workspace %zu < %zu", workspace_bytes, need);
  TcPlan pl = make_plan(nq, n_items);
  RB2_REQUIRE(pl.n_split * 2 <= 64, RB2_EINVAL, "rb2_ce_head(tc): %d logsumexp parts > 64", pl.n_split * 2);
  constexpr int LANES = RowCfg<D>::LANES;
  const int64_t nq_pad = (nq + 2 * BM - 1) / (2 * BM) * (2 * BM);
  {
    ProfScope prof(RB2_ST_TC_CONVERT, st, 5);
    RB2_CUDA(cudaMemsetAsync(w.maxes, 0, 4 * sizeof(float), st));
    RB2_CUDA(cudaMemsetAsync(w.fail_count, 0, 4 * sizeof(int32_t), st));
    if (nq_pad > nq) RB2_CUDA(cudaMemsetAsync(w.qb + nq * 192, 0, (size_t)(nq_pad - nq) * 192 * 2, st));
    k_convert_split<D><<<(unsigned)min((int64_t)kConvertBlocks, (nq * LANES + 255) / 256), 256, 0, st>>>(x, nq, false, w.qb, w.qnorm, w.qlo, w.qdd, nullptr);
    k_convert_split<D><<<(unsigned)min((int64_t)kConvertBlocks, (n_items * LANES + 255) / 256), 256, 0, st>>>(item_p, n_items, true, w.vb, nullptr,
                                                                                 nullptr, nullptr, w.maxes);
  }
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, w.qb, nq_pad, 192, BM, false);
  if (rc) return rc;
  rc = make_map(&tmB, w.vb, n_items, 192, BN / 2, false);
  if (rc) return rc;
  TcParams p;
  p.nq = nq; p.n_local = n_items; p.item_base = 0;
  p.n_ut = pl.n_ut; p.n_split = pl.n_split; p.tiles_per_split = pl.tiles_per_split;
  p.hist_indptr = nullptr; p.hist_indices = nullptr;
  p.cand_ids = w.cand_ids; p.cand_sc = w.cand_sc;
  p.trace = nullptr;
  p.lse_m = lse_m; p.lse_s = lse_s;
  p.row_map = nullptr;
  *parts_out = pl.n_split * 2;
  const size_t smem = TcSmem<KB, NSTAGE, false>::TOTAL;
  {
    ProfScope prof(RB2_ST_TC_SCORE, st);
    auto kern = k_fullsort_tc<KB, NSTAGE, KP, false, false, false, true, false>;
    RB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)pl.grid);
    cfg.blockDim = dim3(kThreadsTc);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RB2_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
    RB2_CUDA(cudaGetLastError());
  }
  {
    ProfScope prof(RB2_ST_TC_REFINE, st);
    const int total = pl.n_split * 2 * KP;
#define RB2_REFINE(MAXC_)                                                                                         \
  k_refine<D, KP, 2, MAXC_><<<(unsigned)((nq * 32 + 127) / 128), 128, 0, st>>>(                                    \
      x, nullptr, nq, item_p, 0, w.cand_ids, w.cand_sc, pl.n_split * 2, k, w.qnorm, nullptr, w.maxes, w.qlo, w.qdd, \
      nullptr, out_ids, out_scores, w.fail_rows, w.fail_count)
    if (total <= 32) RB2_REFINE(1);
    else if (total <= 64) RB2_REFINE(2);
    else if (total <= 128) RB2_REFINE(4);
    else if (total <= 256) RB2_REFINE(8);
    else RB2_REFINE(32);
#undef RB2_REFINE
    RB2_CUDA(cudaGetLastError());
  }
  int32_t n_fail = 0;
  RB2_CUDA(cudaMemcpyAsync(&n_fail, w.fail_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  RB2_CUDA(cudaStreamSynchronize(st));
  rb2_cur_scorer().last_fallback_rows = n_fail;
  if (n_fail > 0) {
    rc = rb2_fullsort_fp32(x, nullptr, n_fail, item_p, n_items, 0, D, nullptr, nullptr, k, out_ids, out_scores, w.fp32_ws,
                           w.fp32_bytes, st, w.fail_rows);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace

size_t rb2_fullsort_tc_workspace_bytes(int64_t nq, int64_t n_items_local, int32_t dim, int32_t k) {
  if (dim != 64 && dim != 128) return rb2_fullsort_fp32_workspace(nq, n_items_local, dim, k) + 256;
  TcWs w;
  return carve_tc(w, nullptr, nq, n_items_local, dim, k) + 256;
}

int rb2_fullsort_tc(const float *query_p, const int64_t *query_ids, int64_t nq, const float *item_p,
                    int64_t n_items_local, int64_t item_base, int32_t dim, const int64_t *hist_indptr,
                    const int64_t *hist_indices, int32_t k, int64_t *out_ids, float *out_scores, void *workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  if ((dim != 64 && dim != 128) || k > 16) {
    // the MMA tiling covers d = 64 / 128 and K <= 16 (K' = 32 candidates); other shapes take the
    // exact CUDA-core kernel
    rb2_cur_scorer().last_fallback_rows = (int32_t)(nq > INT32_MAX ? INT32_MAX : nq);
    return rb2_fullsort_fp32(query_p, query_ids, nq, item_p, n_items_local, item_base, dim, hist_indptr,
                             hist_indices, k, out_ids, out_scores, workspace, workspace_bytes, st, nullptr);
  }
  // K' = 16 candidates per list for small K, and for short item ranges (list upkeep dominates there and a
  // row whose certificate fails is cheap to redo); 32 otherwise
  const int32_t tc_kprime = rb2_cur_scorer().kprime, tc_variant = rb2_cur_scorer().variant;
  const bool small_list = (tc_kprime == 16) || (tc_kprime == 0 && (k <= 8 || (k <= 12 && n_items_local <= 262144)));
#define RB2_TC_ARGS                                                                                           \
  (query_p, query_ids, nq, item_p, n_items_local, item_base, hist_indptr, hist_indices, k, out_ids, out_scores, \
   workspace, workspace_bytes, st)
#define RB2_TC(D_, KP_)                                                        \
  return (tc_variant == 1)   ? run_tc<D_, KP_, false, false> RB2_TC_ARGS     \
         : (tc_variant == 2) ? run_tc<D_, KP_, true, true> RB2_TC_ARGS       \
                               : run_tc<D_, KP_, true, false> RB2_TC_ARGS
  if (dim == 64) {
    if (small_list) RB2_TC(64, 16);
    RB2_TC(64, 32);
  }
  if (small_list) RB2_TC(128, 16);
  if (rb2_cur_scorer().trace) {   // the instrumented build exists for d = 128, K' = 32 only (tools/tc_trace.py)
    if (tc_variant == 2) return run_tc<128, 32, true, true, true> RB2_TC_ARGS;
    if (tc_variant != 1) return run_tc<128, 32, true, false, true> RB2_TC_ARGS;
  }
  RB2_TC(128, 32);
#undef RB2_TC
}

// CE head (fullsort.cu: rb2_ce_head) on the tensor cores: top-k as rb2_fullsort_tc plus the per-part
// (max, sum exp) pairs of the logsumexp over ALL items.  Covers dim == 64, k <= 16; returns 1 otherwise.
size_t rb2_fullsort_tc_lse_workspace_bytes(int64_t nq, int64_t n_items, int32_t dim, int32_t k) {
  if (dim != 64 || k > 16) return 0;
  TcLseWs w;
  return carve_tc_lse(w, nullptr, nq, n_items, k) + 256;
}
int rb2_fullsort_tc_lse(const float *x, int64_t nq, const float *item_p, int64_t n_items, int32_t dim, int32_t k,
                        int64_t *out_ids, float *out_scores, void *workspace, size_t workspace_bytes, cudaStream_t st,
                        float *lse_m, float *lse_s, int *parts_out) {
  if (dim != 64 || k > 16) return 1;
  if (k <= 8) return run_tc_lse<16>(x, nq, item_p, n_items, k, out_ids, out_scores, workspace, workspace_bytes, st, lse_m, lse_s, parts_out);
  return run_tc_lse<32>(x, nq, item_p, n_items, k, out_ids, out_scores, workspace, workspace_bytes, st, lse_m, lse_s, parts_out);
}
