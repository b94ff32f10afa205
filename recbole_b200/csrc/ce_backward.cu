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
// Per (P tile, Q tile of 64 rows), all operands bf16 split in two (x = hi + lo, |lo| <= 2^-9 |x|):
//   MMA 1   S[128 x 64]  = P_hi Q_hi^T + P_hi Q_lo^T + P_lo Q_hi^T          (fp32 in TMEM, error ~2^-17 |p||q|)
//   epilogue (2 sets of 4 warps, thread <-> TMEM lane <-> P row):  g = (exp2(s*log2e - lse*log2e) - [hit]) * scale,
//            g = g_hi + g_lo written as the K-major, 128-byte-swizzled A operand of
//   MMA 2   Out[128 x 64] += G_hi QT_hi + G_hi QT_lo + G_lo QT_hi           (K = the 64 Q rows; QT = Q transposed)
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 4-7 / 8-11 = the epilogue sets
// (alternate Q tiles).  TMEM: two 64-column S stages + one 64-column Out accumulator.
#include <algorithm>

#include "tc_helpers.cuh"

namespace {

using namespace tc;

constexpr int D = 64;                 // hidden size served by this path
constexpr int PM = 128;               // P rows per tile = TMEM lanes
constexpr int QN = 64;                // Q rows per tile
constexpr int NS = 3;                 // Q ring stages
constexpr int kThreadsCe = 384;
constexpr int TILE_P = PM * D * 2;    // 16 KB: one bf16 [128 x 64] operand tile
constexpr int TILE_Q = QN * D * 2;    // 8 KB
constexpr int STAGE_BYTES = 4 * TILE_Q;          // Q_hi, Q_lo, QT_hi, QT_lo
constexpr int G_BYTES = 2 * TILE_P;              // G_hi, G_lo of one epilogue set
constexpr size_t kSmemCe = 1024 + 2 * TILE_P + (size_t)NS * STAGE_BYTES + 2 * G_BYTES + 2 * QN * 8 + 512;
constexpr float kLog2e = 1.4426950408889634f;

struct CeBwdParams {
  int64_t nP, nQ;                 // valid rows of P and Q
  int64_t n_items;                // classes of the softmax (rows of E)
  int n_ptiles, n_split, qtiles_per_split, n_qtiles;
  const float *lse;               // [queries] row logsumexp of the forward pass
  const int64_t *target;          // [queries]
  float scale;                    // upstream gradient / number of rows of the mean
  float *out;                     // ROWS_Q: [n_split][n_ptiles * 128][64] partials; else [nP][64]
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
  uint64_t *bars = reinterpret_cast<uint64_t *>(col_pos + 2 * QN);
  uint64_t *p_full = bars, *p_empty = bars + 1, *o_full = bars + 2, *o_empty = bars + 3;
  uint64_t *q_full = bars + 4, *q_empty = q_full + NS;
  uint64_t *s_full = q_empty + NS, *s_empty = s_full + 2, *g_full = s_empty + 2, *g_empty = g_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(g_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n_work = p.n_ptiles * p.n_split;

  if (threadIdx.x == 0) {
    mbar_init(p_full, 1); mbar_init(p_empty, 1); mbar_init(o_full, 1); mbar_init(o_empty, 128);
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
  constexpr uint32_t kIdesc = idesc_bf16_f32(PM, QN);     // both MMAs are M = 128, N = 64, K = 16

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
      const uint32_t tmem_o = tmem_base + 2 * QN;
      const uint64_t dPh = make_smem_desc(smem_u32(sP)), dPl = make_smem_desc(smem_u32(sP + TILE_P));

      auto mma2 = [&](int stg, int i, bool first) {
        const int e = i & 1;
        mbar_wait_wd(&g_full[e], n2[e] & 1);
        ++n2[e];
        if (first) mbar_wait_wd(o_empty, (wc & 1) ^ 1);      // the previous work item's Out has been drained
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
            tc_mma_bf16(tmem_o, a + (uint64_t)(2 * k4), b + (uint64_t)(2 * k4), kIdesc, (first && term == 0 && k4 == 0) ? 0u : 1u);
        }
        tc_commit(&g_empty[e]);        // G of this set may be overwritten
        tc_commit(&q_empty[stg]);      // both MMAs that read this Q stage are complete
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
              tc_mma_bf16(tmem_s, a + (uint64_t)(2 * k4), b + (uint64_t)(2 * k4), kIdesc, (term | k4) ? 1u : 0u);
          }
          tc_commit(&s_full[e]);
          if (i > 0) mma2(prev_stage, i - 1, i - 1 == 0);
          prev_stage = stage;
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        mma2(prev_stage, T - 1, T == 1);
        tc_commit(o_full);       // Out of this work item is complete
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
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++wc) {
      const int pt = w % p.n_ptiles, sp = w / p.n_ptiles;
      const int64_t row = (int64_t)pt * PM + t;
      const int q0 = sp * p.qtiles_per_split, q1 = min(q0 + p.qtiles_per_split, p.n_qtiles);
      const int T = q1 - q0;
      float row_nl = 0.f;
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
                const int64_t item = qbase + j;
                const float ex = ex2_approx(fmaf(s, kLog2e, row_nl));
                g = (item < p.n_items) ? (ex - (item == row_pos ? 1.f : 0.f)) * p.scale : 0.f;
              } else {
                const float ex = ex2_approx(fmaf(s, kLog2e, nl[j]));       // ex2(-inf) = 0 for padded queries
                g = (ex - ((int64_t)cp[j] == row ? 1.f : 0.f)) * p.scale;
              }
              g2[u] = g;
            }
            const __nv_bfloat16 h0 = __float2bfloat16_rn(g2[0]), h1 = __float2bfloat16_rn(g2[1]);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(g2[0] - __bfloat162float(h0));
            const __nv_bfloat16 l1 = __float2bfloat16_rn(g2[1] - __bfloat162float(h1));
            hw[h] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lw[h] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          // K-major, 128-byte swizzle: row t, 16-byte chunk c sits at chunk position c ^ (t & 7)
          const uint32_t off = (uint32_t)t * 128u + (uint32_t)((c ^ (t & 7)) << 4);
          *reinterpret_cast<uint4 *>(gh + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4 *>(gl + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        fence_proxy_async_smem();           // generic-proxy writes -> visible to the tensor core's async proxy
        mbar_arrive(&g_full[e]);
      }
      if (e == 0) {
        // drain the Out accumulator of this work item
        mbar_wait_wd(o_full, wc & 1);
        tc_fence_after();
        uint32_t va[32], vb[32];
        const uint32_t taddr = lane_addr + (uint32_t)(2 * QN);
        TC_LD32(taddr, va);
        TC_LD32(taddr + 32, vb);
        tmem_wait_ld();
        TC_REGS_AFTER_WAIT(va);
        TC_REGS_AFTER_WAIT(vb);
        tc_fence_before();
        mbar_arrive(o_empty);
        const bool store = ROWS_Q ? true : row < p.nP;
        if (store) {
          float *o = ROWS_Q ? p.out + ((size_t)sp * p.n_ptiles * PM + row) * D : p.out + (size_t)row * D;
          float4 *o4 = reinterpret_cast<float4 *>(o);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o4[j] = make_float4(__uint_as_float(va[4 * j]), __uint_as_float(va[4 * j + 1]), __uint_as_float(va[4 * j + 2]),
                                __uint_as_float(va[4 * j + 3]));
            o4[8 + j] = make_float4(__uint_as_float(vb[4 * j]), __uint_as_float(vb[4 * j + 1]), __uint_as_float(vb[4 * j + 2]),
                                    __uint_as_float(vb[4 * j + 3]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

// fp32 [rows, 64] -> bf16 hi / lo, row-major [rows_pad, 64] and transposed [64, rows_pad]; rows >= `rows` are zero.
// One block per 64-row slab.
__global__ void __launch_bounds__(256) k_ce_split_t(const float *__restrict__ src, int64_t rows, int64_t rows_pad,
                                                    uint16_t *__restrict__ hi, uint16_t *__restrict__ lo,
                                                    uint16_t *__restrict__ hiT, uint16_t *__restrict__ loT) {
  __shared__ uint16_t th[64][66], tl[64][66];
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * 16; i += 256) {      // 64 rows x 16 float4
    const int r = i / 16, c4 = i % 16;
    const int64_t row = r0 + r;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rows) x = __ldg(reinterpret_cast<const float4 *>(src + row * D) + c4);
    const float f[4] = {x.x, x.y, x.z, x.w};
    uint16_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat16 hb = __float2bfloat16_rn(f[k]);
      const __nv_bfloat16 lb = __float2bfloat16_rn(f[k] - __bfloat162float(hb));
      h[k] = __bfloat16_as_ushort(hb);
      l[k] = __bfloat16_as_ushort(lb);
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

// dX[row] = sum over the item ranges of the partials, in range order
__global__ void __launch_bounds__(256) k_ce_reduce(const float *__restrict__ part, int n_split, int64_t rows_pad,
                                                   int64_t rows, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // float4 index
  if (i >= rows * (D / 4)) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < n_split; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(part + (size_t)s * rows_pad * D) + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  reinterpret_cast<float4 *>(out)[i] = acc;
}

struct CeBwdWs {
  uint16_t *xh, *xl, *xth, *xtl;      // queries
  uint16_t *eh, *el, *eth, *etl;      // items
  float *part;                        // dX partials
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
  if ((rc = make_map_bf16(&mQTh, qth, D, q_rows_pad, D))) return rc;     // [64 dims][rows]: box = 64 rows (K) x 64 dims
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
    k_ce_split_t<<<(unsigned)(nq_pad / 64), 256, 0, st>>>(x, nq, nq_pad, w.xh, w.xl, w.xth, w.xtl);
    k_ce_split_t<<<(unsigned)(ni_pad / 64), 256, 0, st>>>(item_p, n_items, ni_pad, w.eh, w.el, w.eth, w.etl);
    RB2_CUDA(cudaGetLastError());
  }
  CeBwdParams p{};
  p.n_items = n_items;
  p.lse = lse;
  p.target = target;
  p.scale = grad_scale;
  if (dx_out) {
    // pass dX: P = queries, Q = items in n_split ranges
    ProfScope prof(RB2_ST_TC_SCORE, st, 2);
    p.nP = nq; p.nQ = n_items;
    p.n_ptiles = (int)(nq_pad / PM);
    p.n_qtiles = (int)(ni_pad / QN);
    p.n_split = w.n_split_x;
    p.qtiles_per_split = (p.n_qtiles + p.n_split - 1) / p.n_split;
    p.out = w.part;
    int rc = launch_pass<true>(p, w.xh, w.xl, nq_pad, w.eh, w.el, w.eth, w.etl, ni_pad, st);
    if (rc) return rc;
    k_ce_reduce<<<(unsigned)((nq * (D / 4) + 255) / 256), 256, 0, st>>>(w.part, p.n_split, nq_pad, nq, dx_out);
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
