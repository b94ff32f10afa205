// ce_backward.cu -- backward of the full-sort cross-entropy head on the 5th-generation tensor cores (sm_100a).
//
// Replaces autograd of  SASRec.calculate_loss, CE branch (recbole/model/sequential_recommender/sasrec.py:137-141):
//     logits = seq_output @ item_embedding.weight.T ;  loss = nn.CrossEntropyLoss()(logits, pos_items)
// i.e.  G = (softmax(logits) - onehot(pos)) * grad / B,   dX = G @ E  [B, H],   dE = G^T @ X  [N, H].
// The reference materialises logits AND G ([B, N] fp32 each: 2 x 16.4 GB at BASELINE config 4).  Here neither
// exists: the logits are recomputed tile by tile from the row logsumexp the forward pass kept (rb2_ce_head), G tiles
// live in shared memory only.
//
// One kernel, run twice with the roles of the two matrices swapped ("P" owns the TMEM lanes, "Q" streams):
//   pass dX: P = X (128 query rows per CTA tile), Q = E split into item ranges -> partial dX per range, summed
//            in a fixed order by k_ce_reduce;
//   pass dE: P = E (128 item rows per CTA tile), Q = X (all queries) -> dE rows, complete.
// Recomputing S in both passes costs 1/3 more MMA work than a single pass but needs no [tiles x tiles] flush of
// partial accumulators and no float atomics (results are bit-reproducible).
//
// Per (P tile, Q tile of 64 rows), every operand is a two-way split x = hi + lo:
//   MMA 1   S[128 x 64]  = P_hi Q_hi^T + P_hi Q_lo^T + P_lo Q_hi^T          (fp32 in TMEM)
//   epilogue (2 sets of 4 warps, thread <-> TMEM lane <-> P row):  g = (exp2(s*log2e - lse*log2e) - [hit]) * scale,
//            written as the K-major, 128-byte-swizzled A operand of
//   MMA 2   Out[128 x 64] += G_hi QT_hi + G_hi QT_lo + G_lo QT_hi           (K = the 64 Q rows; QT = Q transposed)
// The halves are FP16 (11 + 11 significant bits), not bf16 (8 + 8): a logit error d changes every softmax weight
// by a factor 1 + d, and with peaked rows (|x||e| ~ 25 at BASELINE config 4's scales) nothing averages out -- a
// two-way bf16 split leaves d ~ 1.4e-5 rms / 1e-4 max and gradients 2e-5 off; the fp16 split gives d ~ 2e-7 rms and
// gradients at 1e-6 (measured against torch autograd).  FP16's range is met by exact power-of-two rescales: X and E
// by 2^kx, 2^ke (from max |.|, measured on the device), g by 2^kg (|g| <= |scale|, chosen on the host), so that the
// largest magnitudes sit at 2^13..2^14; the epilogue / drain multiply by 2^-(kp+kq) and 2^-(kg+kq).
// The tensor core's fp32 accumulate truncates: a chain of thousands of accumulations into one TMEM accumulator drifts
// (measured: 7.8e-6 of the gradient's scale after 1 500 MMAs, growing linearly).  Out is therefore drained every
// FLUSH tiles (two ping-pong accumulators, so the MMAs never wait) and the chunks are summed in fp32 by CUDA cores.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 4-7 / 8-11 = the epilogue sets
// (alternate Q tiles).  TMEM: two 64-column S stages + one 64-column Out accumulator.
#include <algorithm>
#include <cmath>

#include <cuda_fp16.h>

#include "tc_helpers.cuh"

namespace {

using namespace tc;

constexpr int D = 64;                 // hidden size served by this path
constexpr int PM = 128;               // P rows per tile = TMEM lanes
constexpr int QN = 64;                // Q rows per tile
constexpr int NS = 3;                 // Q ring stages
constexpr int FLUSH = 16;             // Q tiles accumulated in TMEM before the Out accumulator is drained (even)
constexpr int kThreadsCe = 384;
constexpr int TILE_P = PM * D * 2;    // 16 KB: one bf16 [128 x 64] operand tile
constexpr int TILE_Q = QN * D * 2;    // 8 KB
constexpr int STAGE_BYTES = 4 * TILE_Q;          // Q_hi, Q_lo, QT_hi, QT_lo
constexpr int G_BYTES = 2 * TILE_P;              // G_hi, G_lo of one epilogue set
constexpr size_t kSmemCe = 1024 + 2 * TILE_P + (size_t)NS * STAGE_BYTES + 2 * G_BYTES + 2 * QN * 12 + 512;
constexpr float kLog2e = 1.4426950408889634f;

struct CeBwdParams {
  int64_t nP, nQ;                 // valid rows of P and Q
  int64_t n_items;                // classes of the softmax (rows of E)
  int n_ptiles, n_split, qtiles_per_split, n_qtiles;
  const float *omp;               // dE pass: [queries] 1 - softmax[b, target_b], from the dX pass (well-conditioned)
  const float *lse;               // [queries] row logsumexp of the forward pass
  const int64_t *target;          // [queries]
  float scale;                    // upstream gradient / number of rows of the mean, times 2^kg
  float inv_gscale;               // 2^-kg
  const float *q_inv_scale;       // device: 2^-kq of the Q-side table (written by k_ce_scale)
  const float *p_inv_scale;       // device: 2^-kp of the P-side table
  float *out;                     // ROWS_Q: [n_split][n_ptiles * 128][64] partials; else [nP][64]
  float *zpart;                   // ROWS_Q: [n_split * 2][n_ptiles * 128] partial sums of exp(logit - lse)
};

// ROWS_Q = true: P rows are queries (dX pass); false: P rows are items (dE pass)
template <bool ROWS_Q>
__global__ void __launch_bounds__(kThreadsCe, 1)
k_ce_bwd(const __grid_constant__ CUtensorMap tmPh, const __grid_constant__ CUtensorMap tmPl,
         const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
         const __grid_constant__ CUtensorMap tmQTh, const __grid_constant__ CUtensorMap tmQTl, CeBwdParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *sP = smem;                                   // P_hi, P_lo
  unsigned char *sQ = sP + 2 * TILE_P;                        // [NS][Q_hi, Q_lo, QT_hi, QT_lo]
  unsigned char *sG = sQ + (size_t)NS * STAGE_BYTES;          // [2 sets][G_hi, G_lo]
  float *col_nl = reinterpret_cast<float *>(sG + 2 * G_BYTES);          // [2][QN]  -lse * log2e of the tile's queries (dE pass)
  int *col_pos = reinterpret_cast<int *>(col_nl + 2 * QN);              // [2][QN]  their targets
  float *col_omp = reinterpret_cast<float *>(col_pos + 2 * QN);         // [2][QN]  -(1 - p[target]) * scale
  uint64_t *bars = reinterpret_cast<uint64_t *>(col_omp + 2 * QN);
  uint64_t *p_full = bars, *p_empty = bars + 1, *o_full = bars + 2, *o_empty = bars + 4;     // o_*: [2]
  uint64_t *q_full = bars + 6, *q_empty = q_full + NS;
  uint64_t *s_full = q_empty + NS, *s_empty = s_full + 2, *g_full = s_empty + 2, *g_empty = g_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(g_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n_work = p.n_ptiles * p.n_split;

  if (threadIdx.x == 0) {
    mbar_init(p_full, 1); mbar_init(p_empty, 1);
    for (int o = 0; o < 2; ++o) { mbar_init(&o_full[o], 1); mbar_init(&o_empty[o], 128); }
    for (int s = 0; s < NS; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    for (int e = 0; e < 2; ++e) {
      mbar_init(&s_full[e], 1); mbar_init(&s_empty[e], 128);
      mbar_init(&g_full[e], 128); mbar_init(&g_empty[e], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmPh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQTh) : "memory");
  }
  if (warp == 1) {   // TMEM: 256 columns (S stage 0, S stage 1, Out, spare)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // both MMAs: M = 128, N = 64, K = 16, FP16 A and B (format 0), fp32 accumulate
  constexpr uint32_t kIdescH = (1u << 4) | ((uint32_t)(QN >> 3) << 17) | ((uint32_t)(PM >> 4) << 24);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, wc = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++wc) {
        const int pt = w % p.n_ptiles, sp = w / p.n_ptiles;
        mbar_wait_wd(p_empty, (wc & 1) ^ 1);
        mbar_expect_tx(p_full, 2 * TILE_P);
        tma_load_2d(sP, &tmPh, 0, pt * PM, p_full);
        tma_load_2d(sP + TILE_P, &tmPl, 0, pt * PM, p_full);
        const int q0 = sp * p.qtiles_per_split, q1 = min(q0 + p.qtiles_per_split, p.n_qtiles);
        for (int qt = q0; qt < q1; ++qt) {
          mbar_wait_wd(&q_empty[stage], phase ^ 1);
          unsigned char *st = sQ + (size_t)stage * STAGE_BYTES;
          mbar_expect_tx(&q_full[stage], STAGE_BYTES);
          tma_load_2d(st, &tmQh, 0, qt * QN, &q_full[stage]);
          tma_load_2d(st + TILE_Q, &tmQl, 0, qt * QN, &q_full[stage]);
          tma_load_2d(st + 2 * TILE_Q, &tmQTh, qt * QN, 0, &q_full[stage]);
          tma_load_2d(st + 3 * TILE_Q, &tmQTl, qt * QN, 0, &q_full[stage]);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, wc = 0;
      uint32_t n1[2] = {0, 0}, n2[2] = {0, 0};      // MMA-1 / MMA-2 groups issued for each epilogue set
      uint32_t oc = 0, ouse[2] = {0, 0};            // Out chunks started so far; uses of each Out accumulator
      const uint64_t dPh = make_smem_desc(smem_u32(sP)), dPl = make_smem_desc(smem_u32(sP + TILE_P));

      auto mma2 = [&](int stg, int i, int T) {
        const int e = i & 1;
        const bool first = (i % FLUSH) == 0, last = (i % FLUSH) == FLUSH - 1 || i == T - 1;
        const int oi = (int)(oc & 1u);
        const uint32_t tmem_o = tmem_base + (uint32_t)(2 * QN + oi * QN);
        mbar_wait_wd(&g_full[e], n2[e] & 1);
        ++n2[e];
        if (first) {                                         // this accumulator's previous chunk has been drained
          mbar_wait_wd(&o_empty[oi], (ouse[oi] & 1) ^ 1);
          ++ouse[oi];
        }
        tc_fence_after();
        const unsigned char *st = sQ + (size_t)stg * STAGE_BYTES;
        const uint64_t dGh = make_smem_desc(smem_u32(sG + (size_t)e * G_BYTES));
        const uint64_t dGl = make_smem_desc(smem_u32(sG + (size_t)e * G_BYTES + TILE_P));
        const uint64_t dTh = make_smem_desc(smem_u32(st + 2 * TILE_Q)), dTl = make_smem_desc(smem_u32(st + 3 * TILE_Q));
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint64_t a = term == 2 ? dGl : dGh, b = term == 1 ? dTl : dTh;
#pragma unroll
          for (int k4 = 0; k4 < D / 16; ++k4)
            tc_mma_bf16(tmem_o, a + (uint64_t)(2 * k4), b + (uint64_t)(2 * k4), kIdescH, (first && term == 0 && k4 == 0) ? 0u : 1u);
        }
        tc_commit(&g_empty[e]);        // G of this set may be overwritten
        tc_commit(&q_empty[stg]);      // both MMAs that read this Q stage are complete
        if (last) {
          tc_commit(&o_full[oi]);      // the chunk is complete: the drainer may read it
          ++oc;
        }
      };

      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++wc) {
        const int sp = w / p.n_ptiles;
        const int q0 = sp * p.qtiles_per_split, q1 = min(q0 + p.qtiles_per_split, p.n_qtiles);
        const int T = q1 - q0;
        mbar_wait_wd(p_full, wc & 1);
        tc_fence_after();
        int prev_stage = 0;
        for (int i = 0; i < T; ++i) {
          const int e = i & 1;
          mbar_wait_wd(&q_full[stage], phase);
          mbar_wait_wd(&s_empty[e], (n1[e] & 1) ^ 1);
          ++n1[e];
          tc_fence_after();
          const unsigned char *st = sQ + (size_t)stage * STAGE_BYTES;
          const uint64_t dQh = make_smem_desc(smem_u32(st)), dQl = make_smem_desc(smem_u32(st + TILE_Q));
          const uint32_t tmem_s = tmem_base + (uint32_t)(e * QN);
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint64_t a = term == 2 ? dPl : dPh, b = term == 1 ? dQl : dQh;
#pragma unroll
            for (int k4 = 0; k4 < D / 16; ++k4)
              tc_mma_bf16(tmem_s, a + (uint64_t)(2 * k4), b + (uint64_t)(2 * k4), kIdescH, (term | k4) ? 1u : 0u);
          }
          tc_commit(&s_full[e]);
          if (i > 0) mma2(prev_stage, i - 1, T);
          prev_stage = stage;
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        mma2(prev_stage, T - 1, T);
        tc_commit(p_empty);      // every MMA that read this P tile is complete
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 2 sets x 4 warps; thread <-> TMEM lane <-> P row =====================
    const int e = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int t = quarter * 32 + lane;
    const int tin = threadIdx.x - 128 - e * 128;           // 0..127 inside the set
    uint32_t my_cnt = 0, wc = 0;
    unsigned char *gh = sG + (size_t)e * G_BYTES, *gl = gh + TILE_P;
    float *nl = col_nl + e * QN;
    int *cp = col_pos + e * QN;
    float *co = col_omp + e * QN;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float s2 = __ldg(p.p_inv_scale) * __ldg(p.q_inv_scale) * kLog2e;    // TMEM holds 2^(kp+kq) * logit
    const float om = p.inv_gscale * __ldg(p.q_inv_scale);                     // exact: both are powers of two
    uint32_t odone = 0;                                                        // Out chunks drained so far (set 1)
    float *out_row = nullptr;
    // Set 1 drains every chunk of FLUSH tiles (chunks end on odd tiles, i.e. its own, except possibly the last one of
    // a work item): Out -> registers -> added to the row's fp32 result in global memory (first chunk: stored).
    auto drain = [&](int last_tile) {
      const int oi = (int)(odone & 1u);
      mbar_wait_wd(&o_full[oi], (odone >> 1) & 1);
      ++odone;
      tc_fence_after();
      uint32_t va[32], vb[32];
      const uint32_t taddr = lane_addr + (uint32_t)(2 * QN + oi * QN);
      TC_LD32(taddr, va);
      TC_LD32(taddr + 32, vb);
      tmem_wait_ld();
      TC_REGS_AFTER_WAIT(va);
      TC_REGS_AFTER_WAIT(vb);
      tc_fence_before();
      mbar_arrive(&o_empty[oi]);
      if (out_row) {
        const bool first_chunk = last_tile < FLUSH;
        float4 *o4 = reinterpret_cast<float4 *>(out_row);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 a = make_float4(__uint_as_float(va[4 * j]) * om, __uint_as_float(va[4 * j + 1]) * om,
                                 __uint_as_float(va[4 * j + 2]) * om, __uint_as_float(va[4 * j + 3]) * om);
          float4 b = make_float4(__uint_as_float(vb[4 * j]) * om, __uint_as_float(vb[4 * j + 1]) * om,
                                 __uint_as_float(vb[4 * j + 2]) * om, __uint_as_float(vb[4 * j + 3]) * om);
          if (!first_chunk) {
            const float4 pa = o4[j], pb = o4[8 + j];
            a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
            b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
          }
          o4[j] = a;
          o4[8 + j] = b;
        }
      }
    };
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++wc) {
      const int pt = w % p.n_ptiles, sp = w / p.n_ptiles;
      const int64_t row = (int64_t)pt * PM + t;
      const int q0 = sp * p.qtiles_per_split, q1 = min(q0 + p.qtiles_per_split, p.n_qtiles);
      const int T = q1 - q0;
      out_row = ROWS_Q ? p.out + ((size_t)sp * p.n_ptiles * PM + row) * D : (row < p.nP ? p.out + (size_t)row * D : nullptr);
      float row_nl = 0.f, zsum = 0.f;
      int64_t row_pos = -1;
      if (ROWS_Q && row < p.nP) {
        row_nl = -p.lse[row] * kLog2e;
        row_pos = p.target[row];
      }
      for (int i = e; i < T; i += 2) {
        const int64_t qbase = (int64_t)(q0 + i) * QN;
        if (!ROWS_Q) {
          // the tile's 64 queries: -lse * log2e and target, shared by the set
          asm volatile("bar.sync %0, 128;" ::"r"(1 + e) : "memory");       // the previous tile's values are no longer read
          if (tin < QN) {
            const int64_t b = qbase + tin;
            nl[tin] = b < p.nQ ? -p.lse[b] * kLog2e : -INFINITY;
            cp[tin] = b < p.nQ ? (int)p.target[b] : -1;
            co[tin] = b < p.nQ ? -p.omp[b] * p.scale : 0.f;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + e) : "memory");
        }
        mbar_wait_wd(&s_full[e], my_cnt & 1);
        tc_fence_after();
        uint32_t va[32], vb[32];
        const uint32_t taddr = lane_addr + (uint32_t)(e * QN);
        TC_LD32(taddr, va);
        TC_LD32(taddr + 32, vb);
        tmem_wait_ld();
        TC_REGS_AFTER_WAIT(va);
        TC_REGS_AFTER_WAIT(vb);
        tc_fence_before();
        mbar_arrive(&s_empty[e]);          // the S stage may be overwritten by the tile after next
        // G of the previous tile of this set has been consumed by its MMA 2
        mbar_wait_wd(&g_empty[e], (my_cnt & 1) ^ 1);
        ++my_cnt;
#pragma unroll
        for (int c = 0; c < 8; ++c) {       // 8 columns -> one 16-byte chunk of the hi tile and one of the lo tile
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float g2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int j = c * 8 + h * 2 + u;
              const float s = __uint_as_float(j < 32 ? va[j & 31] : vb[j & 31]);
              float g;
              if (ROWS_Q) {
                // every class but the target: k_ce_reduce adds the target's own weight to the row's sum of exp (the
                // softmax normaliser against the forward pass's lse) and applies -(1 - p[target]) * E[target] in fp32
                // from that sum -- no p - 1 cancellation for confident rows
                const int64_t item = qbase + j;
                const float ex = (item < p.n_items && item != row_pos) ? ex2_approx(fmaf(s, s2, row_nl)) : 0.f;
                zsum += ex;
                g = ex * p.scale;
              } else {
                const float ex = ex2_approx(fmaf(s, s2, nl[j]));           // ex2(-inf) = 0 for padded queries
                g = ((int64_t)cp[j] == row) ? co[j] : ex * p.scale;        // target: -(1 - p) * scale from the dX pass
              }
              g2[u] = g;
            }
            // hi = g with the mantissa cut to FP16's 10 bits (exact in FP16 for normal values: no conversion back is
            // needed to form lo = g - hi); both halves leave through the packed converter (one instruction per pair;
            // the scalar F2F conversions ran on the same quarter-rate pipe as the exponentials)
            const float h0f = __uint_as_float(__float_as_uint(g2[0]) & 0xFFFFE000u);
            const float h1f = __uint_as_float(__float_as_uint(g2[1]) & 0xFFFFE000u);
            const __half2 hp = __floats2half2_rn(h0f, h1f);
            const __half2 lp = __floats2half2_rn(g2[0] - h0f, g2[1] - h1f);
            hw[h] = *reinterpret_cast<const uint32_t *>(&hp);
            lw[h] = *reinterpret_cast<const uint32_t *>(&lp);
          }
          // K-major, 128-byte swizzle: row t, 16-byte chunk c sits at chunk position c ^ (t & 7)
          const uint32_t off = (uint32_t)t * 128u + (uint32_t)((c ^ (t & 7)) << 4);
          *reinterpret_cast<uint4 *>(gh + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4 *>(gl + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        fence_proxy_async_smem();           // generic-proxy writes -> visible to the tensor core's async proxy
        mbar_arrive(&g_full[e]);
        if (e == 1 && ((i % FLUSH) == FLUSH - 1 || i == T - 1)) drain(i);
      }
      if (ROWS_Q) p.zpart[(size_t)(sp * 2 + e) * p.n_ptiles * PM + row] = zsum;
      if (e == 1 && ((T - 1) & 1) == 0) drain(T - 1);       // the last chunk ends on a tile of the other set
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

// max |x| of a tensor as the bit pattern of a non-negative float (compared as an int); then the power-of-two scale
// that puts it in [2^13, 2^14): sc[0] = 2^kq, sc[1] = 2^-kq
__global__ void __launch_bounds__(256) k_ce_absmax(const float *__restrict__ src, int64_t n4, int *__restrict__ out_bits) {
  int m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = __ldg(reinterpret_cast<const float4 *>(src) + i);
    const float a = fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w)));
    if (a < INFINITY) m = max(m, __float_as_int(a));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out_bits, m);
}
__global__ void k_ce_scale(const int *__restrict__ max_bits, float *__restrict__ sc) {
  const float m = __int_as_float(*max_bits);
  int e = 0;
  if (m > 0.f) frexpf(m, &e);            // m = f * 2^e, f in [0.5, 1)
  const int k = m > 0.f ? 14 - e : 0;
  sc[0] = ldexpf(1.f, k);
  sc[1] = ldexpf(1.f, -k);
}

// fp32 [rows, 64] -> rescaled FP16 hi / lo, row-major [rows_pad, 64] (MMA 1) and transposed [64, rows_pad] (MMA 2);
// rows >= `rows` are zero.  One block per 64-row slab.
__global__ void __launch_bounds__(256) k_ce_split_t(const float *__restrict__ src, int64_t rows, int64_t rows_pad,
                                                    const float *__restrict__ sc, uint16_t *__restrict__ hi,
                                                    uint16_t *__restrict__ lo, uint16_t *__restrict__ hiT,
                                                    uint16_t *__restrict__ loT) {
  __shared__ uint16_t th[64][66], tl[64][66];
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  const float qs = __ldg(sc);
  for (int i = threadIdx.x; i < 64 * 16; i += 256) {      // 64 rows x 16 float4
    const int r = i / 16, c4 = i % 16;
    const int64_t row = r0 + r;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows) x = __ldg(reinterpret_cast<const float4 *>(src + row * D) + c4);
    const float f[4] = {x.x, x.y, x.z, x.w};
    uint16_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float fs = f[k] * qs;                          // exact (power of two); |fs| < 2^14
      const __half hh = __float2half_rn(fs);
      const __half hl = __float2half_rn(fs - __half2float(hh));
      h[k] = __half_as_ushort(hh);
      l[k] = __half_as_ushort(hl);
      th[c4 * 4 + k][r] = h[k];
      tl[c4 * 4 + k][r] = l[k];
    }
    uint2 ph, pl;
    ph.x = (uint32_t)h[0] | ((uint32_t)h[1] << 16); ph.y = (uint32_t)h[2] | ((uint32_t)h[3] << 16);
    pl.x = (uint32_t)l[0] | ((uint32_t)l[1] << 16); pl.y = (uint32_t)l[2] | ((uint32_t)l[3] << 16);
    reinterpret_cast<uint2 *>(hi + row * D)[c4] = ph;
    reinterpret_cast<uint2 *>(lo + row * D)[c4] = pl;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {      // 64 dims x 32 pairs of rows
    const int d = i / 32, rp = i % 32;
    const uint32_t vh = (uint32_t)th[d][2 * rp] | ((uint32_t)th[d][2 * rp + 1] << 16);
    const uint32_t vl = (uint32_t)tl[d][2 * rp] | ((uint32_t)tl[d][2 * rp + 1] << 16);
    reinterpret_cast<uint32_t *>(hiT + (size_t)d * rows_pad + r0)[rp] = vh;
    reinterpret_cast<uint32_t *>(loT + (size_t)d * rows_pad + r0)[rp] = vl;
  }
}

// Per query row b (fixed summation order):
//   Z' = sum over the item ranges of the partial sums of exp(logit_j - lse_fwd), j != target;
//   Z  = Z' + exp(<x_b, e_target> - lse_fwd)   (= 1 up to the error of the forward pass's logsumexp, which used bf16
//        halves);  lse_corr = lse_fwd + log Z  and  omp = Z' / Z = 1 - softmax[b, target]  for the dE pass;
//   dX[b] = (sum of the partials) / Z - grad_scale * omp * E[target]          (partials = grad_scale * sum' exp_j E_j).
__global__ void __launch_bounds__(256) k_ce_reduce(const float *__restrict__ part, const float *__restrict__ zpart,
                                                   int n_split, int64_t rows_pad, int64_t rows,
                                                   const float *__restrict__ x, const float *__restrict__ item_p,
                                                   const int64_t *__restrict__ target, const float *__restrict__ lse,
                                                   float grad_scale, float *__restrict__ dx_out,
                                                   float *__restrict__ lse_corr, float *__restrict__ omp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // float4 index
  if (i >= rows * (D / 4)) return;
  const int64_t row = i / (D / 4);
  const int c4 = (int)(i % (D / 4));
  float zp = 0.f;
  for (int s = 0; s < 2 * n_split; ++s) zp += __ldg(zpart + (size_t)s * rows_pad + row);
  const float4 *xr = reinterpret_cast<const float4 *>(x + row * D);
  const float4 *er = reinterpret_cast<const float4 *>(item_p + target[row] * D);
  float sp = 0.f;
#pragma unroll
  for (int k = 0; k < D / 4; ++k) {
    const float4 a = __ldg(xr + k), b = __ldg(er + k);
    sp = fmaf(a.x, b.x, sp); sp = fmaf(a.y, b.y, sp); sp = fmaf(a.z, b.z, sp); sp = fmaf(a.w, b.w, sp);
  }
  const float z = zp + expf(sp - lse[row]);
  const float om = zp / z;
  if (c4 == 0) {
    lse_corr[row] = lse[row] + logf(z);
    omp[row] = om;
  }
  if (!dx_out) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < n_split; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(part + (size_t)s * rows_pad * D) + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float iz = 1.f / z, c = -grad_scale * om;
  const float4 e = __ldg(er + c4);
  reinterpret_cast<float4 *>(dx_out)[i] =
      make_float4(fmaf(acc.x, iz, c * e.x), fmaf(acc.y, iz, c * e.y), fmaf(acc.z, iz, c * e.z), fmaf(acc.w, iz, c * e.w));
}

struct CeBwdWs {
  uint16_t *xh, *xl, *xth, *xtl;      // queries
  uint16_t *eh, *el, *eth, *etl;      // items
  float *part;                        // dX partials
  float *zpart;                       // [n_split_x * 2][nq_pad]
  float *lse_corr;                    // [nq] logsumexp consistent with THIS pass's logits
  float *omp;                         // [nq] 1 - softmax[b, target_b]
  int *max_bits;                      // [2]: max |x|, max |e| (bit patterns)
  float *sc_x, *sc_e;                 // [2] each: 2^kq, 2^-kq
  int n_split_x;
};

int pick_split(int n_ptiles, int n_qtiles, int grid) {
  // work items = n_ptiles * n_split; prefer >= 2 waves of the grid with the best fill of the last wave
  int best = 1;
  double best_eff = -1.0;
  const int max_split = std::min(n_qtiles, 64);
  for (int s = 1; s <= max_split; ++s) {
    const int qps = (n_qtiles + s - 1) / s;
    const int s_eff = (n_qtiles + qps - 1) / qps;
    if (s_eff != s) continue;
    const long work = (long)n_ptiles * s;
    const long waves = (work + grid - 1) / grid;
    double eff = (double)work / (double)(waves * grid);
    if (work < 2L * grid) eff *= 0.9;             // one wave: no overlap of a work item's prologue / drain
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  return best;
}

size_t carve_ce_bwd(CeBwdWs &w, void *base, int64_t nq, int64_t n_items) {
  Carver c(base);
  const int64_t nq_pad = (nq + PM - 1) / PM * PM, ni_pad = (n_items + PM - 1) / PM * PM;
  w.xh = c.take<uint16_t>(nq_pad * D); w.xl = c.take<uint16_t>(nq_pad * D);
  w.xth = c.take<uint16_t>(nq_pad * D); w.xtl = c.take<uint16_t>(nq_pad * D);
  w.eh = c.take<uint16_t>(ni_pad * D); w.el = c.take<uint16_t>(ni_pad * D);
  w.eth = c.take<uint16_t>(ni_pad * D); w.etl = c.take<uint16_t>(ni_pad * D);
  w.n_split_x = pick_split((int)(nq_pad / PM), (int)(ni_pad / QN), rb2_num_sms());
  w.part = c.take<float>((size_t)w.n_split_x * nq_pad * D);
  w.zpart = c.take<float>((size_t)w.n_split_x * 2 * nq_pad);
  w.lse_corr = c.take<float>(nq_pad);
  w.omp = c.take<float>(nq_pad);
  w.max_bits = c.take<int>(2);
  w.sc_x = c.take<float>(2);
  w.sc_e = c.take<float>(2);
  return c.off;
}

template <bool ROWS_Q>
int launch_pass(const CeBwdParams &p, uint16_t *ph, uint16_t *pl, int64_t p_rows_pad, uint16_t *qh, uint16_t *ql,
                uint16_t *qth, uint16_t *qtl, int64_t q_rows_pad, cudaStream_t st) {
  CUtensorMap mPh, mPl, mQh, mQl, mQTh, mQTl;
  int rc;
  if ((rc = make_map_bf16(&mPh, ph, p_rows_pad, D, PM))) return rc;
  if ((rc = make_map_bf16(&mPl, pl, p_rows_pad, D, PM))) return rc;
  if ((rc = make_map_bf16(&mQh, qh, q_rows_pad, D, QN))) return rc;
  if ((rc = make_map_bf16(&mQl, ql, q_rows_pad, D, QN))) return rc;
  // [64 dims][rows] FP16 (16-bit elements move the same way): box = 64 rows (K) x 64 dims
  if ((rc = make_map_bf16(&mQTh, qth, D, q_rows_pad, D))) return rc;
  if ((rc = make_map_bf16(&mQTl, qtl, D, q_rows_pad, D))) return rc;
  auto kern = k_ce_bwd<ROWS_Q>;
  RB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCe));
  const int n_work = p.n_ptiles * p.n_split;
  const int grid = std::min(n_work, rb2_num_sms());
  kern<<<grid, kThreadsCe, kSmemCe, st>>>(mPh, mPl, mQh, mQl, mQTh, mQTl, p);
  RB2_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) k_dense_adam(float *__restrict__ P, float *__restrict__ M, float *__restrict__ V,
                                                    const float *__restrict__ G, int64_t n4, OptScalars o) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 p = reinterpret_cast<float4 *>(P)[i];
  const float4 g = __ldg(reinterpret_cast<const float4 *>(G) + i);
  if (o.kind == RB2_OPT_SGD) {
    sgd_elem(p.x, g.x, o); sgd_elem(p.y, g.y, o); sgd_elem(p.z, g.z, o); sgd_elem(p.w, g.w, o);
  } else {
    float4 m = reinterpret_cast<float4 *>(M)[i], v = reinterpret_cast<float4 *>(V)[i];
    adam_elem(p.x, m.x, v.x, g.x, o); adam_elem(p.y, m.y, v.y, g.y, o);
    adam_elem(p.z, m.z, v.z, g.z, o); adam_elem(p.w, m.w, v.w, g.w, o);
    reinterpret_cast<float4 *>(M)[i] = m;
    reinterpret_cast<float4 *>(V)[i] = v;
  }
  reinterpret_cast<float4 *>(P)[i] = p;
}

}  // namespace

extern "C" size_t rb2_ce_head_backward_workspace_bytes(int64_t nq, int64_t n_items, int32_t dim) {
  if (dim != D || nq <= 0 || n_items <= 0) return 0;
  CeBwdWs w;
  return carve_ce_bwd(w, nullptr, nq, n_items) + 256;
}

extern "C" int rb2_ce_head_backward(const float *x, int64_t nq, const float *item_p, int64_t n_items, int32_t dim,
                                    const int64_t *target, const float *lse, float grad_scale, float *dx_out,
                                    float *de_out, void *workspace, size_t workspace_bytes, void *stream) {
  RB2_REQUIRE(x && item_p && target && lse && workspace && (dx_out || de_out), RB2_EINVAL,
              "rb2_ce_head_backward: null argument");
  RB2_REQUIRE(dim == D, RB2_EINVAL, "rb2_ce_head_backward: the tensor-core path serves hidden size 64 (got %d)", (int)dim);
  RB2_REQUIRE(nq > 0 && n_items > 0 && n_items < ((int64_t)1 << 31) && nq < ((int64_t)1 << 31), RB2_EINVAL,
              "rb2_ce_head_backward: sizes out of range");
  CeBwdWs w;
  size_t need = carve_ce_bwd(w, workspace, nq, n_items);
  RB2_REQUIRE(workspace_bytes >= need, RB2_EWORKSPACE, "rb2_ce_head_backward: workspace %zu < %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nq_pad = (nq + PM - 1) / PM * PM, ni_pad = (n_items + PM - 1) / PM * PM;
  {
    ProfScope prof(RB2_ST_TC_CONVERT, st, 2);
    RB2_CUDA(cudaMemsetAsync(w.max_bits, 0, 2 * sizeof(int), st));
    const int mb = rb2_num_sms() * 8;
    k_ce_absmax<<<(unsigned)std::min<int64_t>(mb, (nq * (D / 4) + 255) / 256), 256, 0, st>>>(x, nq * (D / 4), w.max_bits);
    k_ce_absmax<<<(unsigned)std::min<int64_t>(mb, (n_items * (D / 4) + 255) / 256), 256, 0, st>>>(item_p, n_items * (D / 4),
                                                                                              w.max_bits + 1);
    k_ce_scale<<<1, 1, 0, st>>>(w.max_bits, w.sc_x);
    k_ce_scale<<<1, 1, 0, st>>>(w.max_bits + 1, w.sc_e);
    k_ce_split_t<<<(unsigned)(nq_pad / 64), 256, 0, st>>>(x, nq, nq_pad, w.sc_x, w.xh, w.xl, w.xth, w.xtl);
    k_ce_split_t<<<(unsigned)(ni_pad / 64), 256, 0, st>>>(item_p, n_items, ni_pad, w.sc_e, w.eh, w.el, w.eth, w.etl);
    RB2_CUDA(cudaGetLastError());
  }
  CeBwdParams p{};
  p.n_items = n_items;
  p.lse = lse;
  p.target = target;
  // |g| <= |grad_scale|: 2^kg puts it at 2^13..2^14
  int ge = 0;
  if (grad_scale != 0.f && std::isfinite(grad_scale)) std::frexp(std::fabs(grad_scale), &ge);
  const int kg = (grad_scale != 0.f && std::isfinite(grad_scale)) ? 14 - ge : 0;
  p.scale = std::ldexp(grad_scale, kg);
  p.inv_gscale = std::ldexp(1.f, -kg);
  {
    // pass dX: P = queries, Q = items in n_split ranges.  Always run: it also measures every row's softmax
    // normaliser against the forward pass's lse (the dE pass needs a logsumexp consistent to ~1e-6).
    ProfScope prof(RB2_ST_TC_SCORE, st, 2);
    p.nP = nq; p.nQ = n_items;
    p.n_ptiles = (int)(nq_pad / PM);
    p.n_qtiles = (int)(ni_pad / QN);
    p.n_split = w.n_split_x;
    p.qtiles_per_split = (p.n_qtiles + p.n_split - 1) / p.n_split;
    p.n_split = (p.n_qtiles + p.qtiles_per_split - 1) / p.qtiles_per_split;
    p.out = w.part;
    p.zpart = w.zpart;
    p.q_inv_scale = w.sc_e + 1;
    p.p_inv_scale = w.sc_x + 1;
    int rc = launch_pass<true>(p, w.xh, w.xl, nq_pad, w.eh, w.el, w.eth, w.etl, ni_pad, st);
    if (rc) return rc;
    // the partials hold scale' * sum_j exp_j E_j with scale' = grad_scale (the 2^kg is undone in the drain)
    k_ce_reduce<<<(unsigned)((nq * (D / 4) + 255) / 256), 256, 0, st>>>(w.part, w.zpart, p.n_split, nq_pad, nq, x, item_p,
                                                                         target, lse, grad_scale, dx_out, w.lse_corr,
                                                                         w.omp);
    RB2_CUDA(cudaGetLastError());
  }
  if (de_out) {
    // pass dE: P = items, Q = all queries
    ProfScope prof(RB2_ST_TC_REFINE, st, 1);
    p.nP = n_items; p.nQ = nq;
    p.n_ptiles = (int)(ni_pad / PM);
    p.n_qtiles = (int)(nq_pad / QN);
    p.n_split = 1;
    p.qtiles_per_split = p.n_qtiles;
    p.out = de_out;
    p.zpart = nullptr;
    p.lse = w.lse_corr;
    p.omp = w.omp;
    p.q_inv_scale = w.sc_x + 1;
    p.p_inv_scale = w.sc_e + 1;
    int rc = launch_pass<false>(p, w.eh, w.el, ni_pad, w.xh, w.xl, w.xth, w.xtl, nq_pad, st);
    if (rc) return rc;
  }
  return 0;
}

/* dense optimizer step over a whole parameter tensor (the item table after rb2_ce_head_backward: every class has a
 * gradient, so the reference's dense torch.optim step IS the row-sparse one here) */
extern "C" int rb2_dense_step(float *p, float *m, float *v, const float *grad, int64_t count, const rb2_optim *h_opt,
                              void *stream) {
  RB2_REQUIRE(p && grad && h_opt, RB2_EINVAL, "rb2_dense_step: null argument");
  RB2_REQUIRE(count >= 0 && count % 4 == 0, RB2_EINVAL, "rb2_dense_step: count must be a multiple of 4");
  OptScalars o = rb2_opt_scalars(h_opt);
  RB2_REQUIRE(o.kind == RB2_OPT_SGD || o.kind == RB2_OPT_ADAM, RB2_EINVAL, "rb2_dense_step: sgd or adam");
  if (o.kind == RB2_OPT_ADAM) RB2_REQUIRE(m && v, RB2_EINVAL, "rb2_dense_step: Adam needs m and v");
  if (count == 0) return 0;
  k_dense_adam<<<(unsigned)((count / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, m, v, grad, count / 4, o);
  RB2_CUDA(cudaGetLastError());
  return 0;
}
