// fullsort.cu -- full-sort scorer (exact fp32 path), top-K merge and on-device metrics for sm_100a.
//
// Replaces BPR.full_sort_predict (bpr.py:91-96) + Trainer._full_sort_batch_eval masking
// (trainer.py:342-345) + TopKEvaluator.collect's flip/topk (evaluators.py:68-72) +
// TopKEvaluator.evaluate / metrics.py (evaluators.py:78-141, metrics.py:27-164).
// The [users, n_items] score matrix never exists: every score is compared against the row's
// running K-th best as soon as it is produced.
//
// k_fullsort_fp32: one thread owns one query row (its embedding lives in registers) and walks
// the item table, which streams through shared memory in double-buffered tiles loaded by the
// bulk-copy engine (cp.async.bulk + mbarrier).  All lanes of a warp read the same item element,
// so the shared-memory reads are broadcasts.  Scores are the canonical chain
// s = fmaf(q[k], v[k], s), k ascending (see oracle/csrc/oracle.c).  History / pad masking costs
// nothing per element: it is only checked for the rare element that beats the running threshold.
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int kRows = 128;  // query rows (= threads) per block

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar` (TMA engine, UBLKCP in SASS)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ bool csr_contains(const int64_t *__restrict__ a, int64_t n, int64_t x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo < n && a[lo] == x;
}

// rare path: candidate beat the running threshold.  Lists live in shared memory, element j of
// the thread's list at [j * kRows + tid] (conflict-free).
__device__ __noinline__ void topk_insert(float *sc, int *id, int K, float s, int64_t item,
                                         const int64_t *__restrict__ hist, int64_t hlen, float &tau) {
  if (item == 0) return;                          // [PAD], trainer.py:343
  if (hlen > 0 && csr_contains(hist, hlen, item)) return;  // trainer.py:344-345
  int j = K - 1;
  while (j > 0 && sc[(j - 1) * kRows] < s) {      // equal scores stay in front: lower id first
    sc[j * kRows] = sc[(j - 1) * kRows];
    id[j * kRows] = id[(j - 1) * kRows];
    --j;
  }
  sc[j * kRows] = s;
  id[j * kRows] = (int)item;
  tau = sc[(K - 1) * kRows];
}

template <int D>
struct FsCfg {
  static constexpr int TI = (4096 / D) < 8 ? 8 : (4096 / D);  // items per tile (16 KB)
  static constexpr int TILE_FLOATS = TI * D;
};

template <int D, bool LSE>
__global__ void __launch_bounds__(kRows) k_fullsort_fp32(const float *__restrict__ query_p,
                                                          const int64_t *__restrict__ query_ids, int64_t nq,
                                                          const float *__restrict__ item_p, int64_t n_local,
                                                          int64_t item_base, const int64_t *__restrict__ hist_indptr,
                                                          const int64_t *__restrict__ hist_indices, int K,
                                                          int64_t items_per_split, int64_t *__restrict__ out_ids,
                                                          float *__restrict__ out_scores,
                                                          const int32_t *__restrict__ row_map, int scatter_out,
                                                          float *__restrict__ lse_m, float *__restrict__ lse_s) {
  constexpr int TI = FsCfg<D>::TI;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *tile0 = reinterpret_cast<float *>(smem_raw);
  float *tile1 = tile0 + FsCfg<D>::TILE_FLOATS;
  float *lsc = tile1 + FsCfg<D>::TILE_FLOATS;               // [K][kRows]
  int *lid = reinterpret_cast<int *>(lsc + (size_t)K * kRows);  // [K][kRows]
  __shared__ __align__(8) uint64_t bars[2];

  const int tid = threadIdx.x;
  const int64_t rc = (int64_t)blockIdx.x * kRows + tid;   // compact row (position in row_map)
  const bool active = rc < nq;
  // row_map (optional): the rows of the caller's problem this launch redoes (tensor-core fallback)
  const int64_t r = (active && row_map) ? (int64_t)row_map[rc] : rc;
  const int64_t i_begin = (int64_t)blockIdx.y * items_per_split;
  const int64_t i_end = min(i_begin + items_per_split, n_local);
  const int64_t n_tiles = (i_end - i_begin + TI - 1) / TI;

  // query row -> registers
  float q[D];
  {
    int64_t row = active ? (query_ids ? query_ids[r] : r) : 0;
    const float4 *src = reinterpret_cast<const float4 *>(query_p + row * D);
#pragma unroll
    for (int k = 0; k < D / 4; ++k) {
      float4 v = active ? __ldg(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      q[4 * k] = v.x; q[4 * k + 1] = v.y; q[4 * k + 2] = v.z; q[4 * k + 3] = v.w;
    }
  }
  const int64_t *hist = nullptr;
  int64_t hlen = 0;
  if (active && hist_indptr) {
    int64_t h0 = hist_indptr[r];
    hlen = hist_indptr[r + 1] - h0;
    hist = hist_indices + h0;
  }
  for (int j = 0; j < K; ++j) {
    lsc[j * kRows + tid] = -INFINITY;
    lid[j * kRows + tid] = -1;
  }
  float tau = -INFINITY;
  // online logsumexp over EVERY item of the range, pad row included (nn.CrossEntropyLoss over n_items
  // classes, sasrec.py:139-140): run_m = running max, run_s = sum exp(s - run_m)
  float run_m = -INFINITY, run_s = 0.f;
  auto lse_add = [&](float sc) {
    if (sc > run_m) { run_s = run_s * __expf(run_m - sc) + 1.f; run_m = sc; }
    else run_s += __expf(sc - run_m);
  };
  float *my_sc = lsc + tid;
  int *my_id = lid + tid;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int64_t t) {
    int64_t i0 = i_begin + t * TI;
    uint32_t bytes = (uint32_t)(min((int64_t)TI, i_end - i0) * D * sizeof(float));
    uint64_t *bar = &bars[t & 1];
    mbar_expect_tx(bar, bytes);
    bulk_g2s((t & 1) ? tile1 : tile0, item_p + i0 * D, bytes, bar);
  };
  if (tid == 0 && n_tiles > 0) issue(0);

  for (int64_t t = 0; t < n_tiles; ++t) {
    if (tid == 0 && t + 1 < n_tiles) issue(t + 1);
    mbar_wait(&bars[t & 1], (uint32_t)((t >> 1) & 1));
    const float4 *vt = reinterpret_cast<const float4 *>((t & 1) ? tile1 : tile0);
    const int64_t i0 = i_begin + t * TI;
    const int cnt = (int)min((int64_t)TI, i_end - i0);
    if (active) {
      int i = 0;
      for (; i + 4 <= cnt; i += 4) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        const float4 *v0 = vt + (size_t)(i + 0) * (D / 4);
        const float4 *v1 = vt + (size_t)(i + 1) * (D / 4);
        const float4 *v2 = vt + (size_t)(i + 2) * (D / 4);
        const float4 *v3 = vt + (size_t)(i + 3) * (D / 4);
#pragma unroll
        for (int k = 0; k < D / 4; ++k) {
          float4 a = v0[k], b = v1[k], c = v2[k], d = v3[k];
          s0 = fmaf(q[4 * k], a.x, s0); s1 = fmaf(q[4 * k], b.x, s1);
          s2 = fmaf(q[4 * k], c.x, s2); s3 = fmaf(q[4 * k], d.x, s3);
          s0 = fmaf(q[4 * k + 1], a.y, s0); s1 = fmaf(q[4 * k + 1], b.y, s1);
          s2 = fmaf(q[4 * k + 1], c.y, s2); s3 = fmaf(q[4 * k + 1], d.y, s3);
          s0 = fmaf(q[4 * k + 2], a.z, s0); s1 = fmaf(q[4 * k + 2], b.z, s1);
          s2 = fmaf(q[4 * k + 2], c.z, s2); s3 = fmaf(q[4 * k + 2], d.z, s3);
          s0 = fmaf(q[4 * k + 3], a.w, s0); s1 = fmaf(q[4 * k + 3], b.w, s1);
          s2 = fmaf(q[4 * k + 3], c.w, s2); s3 = fmaf(q[4 * k + 3], d.w, s3);
        }
        const int64_t g = item_base + i0 + i;
        if (LSE) { lse_add(s0); lse_add(s1); lse_add(s2); lse_add(s3); }
        if (s0 > tau) topk_insert(my_sc, my_id, K, s0, g + 0, hist, hlen, tau);
        if (s1 > tau) topk_insert(my_sc, my_id, K, s1, g + 1, hist, hlen, tau);
        if (s2 > tau) topk_insert(my_sc, my_id, K, s2, g + 2, hist, hlen, tau);
        if (s3 > tau) topk_insert(my_sc, my_id, K, s3, g + 3, hist, hlen, tau);
      }
      for (; i < cnt; ++i) {
        float s = 0.f;
        const float4 *v0 = vt + (size_t)i * (D / 4);
#pragma unroll
        for (int k = 0; k < D / 4; ++k) {
          float4 a = v0[k];
          s = fmaf(q[4 * k], a.x, s);
          s = fmaf(q[4 * k + 1], a.y, s);
          s = fmaf(q[4 * k + 2], a.z, s);
          s = fmaf(q[4 * k + 3], a.w, s);
        }
        if (LSE) lse_add(s);
        if (s > tau) topk_insert(my_sc, my_id, K, s, item_base + i0 + i, hist, hlen, tau);
      }
    }
    __syncthreads();  // everyone is done with this buffer before it is refilled
  }

  if (active) {
    int64_t o = ((int64_t)blockIdx.y * nq + (scatter_out ? r : rc)) * K;
    for (int j = 0; j < K; ++j) {
      out_ids[o + j] = (int64_t)my_id[j * kRows];
      out_scores[o + j] = my_sc[j * kRows];
    }
    if (LSE) {
      lse_m[(int64_t)blockIdx.y * nq + rc] = run_m;
      lse_s[(int64_t)blockIdx.y * nq + rc] = run_s;
    }
  }
}

// CE head: merge the per-split (max, sum) pairs, subtract the target logit (canonical chain), and
// leave per-row loss terms for the fixed-order mean
template <int D>
__global__ void k_ce_finish(const float *__restrict__ x, const float *__restrict__ item_p, int64_t nq, int64_t n_items,
                            const int64_t *__restrict__ target, const float *__restrict__ lse_m,
                            const float *__restrict__ lse_s, int parts, float *__restrict__ lse_out,
                            float *__restrict__ row_loss) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nq) return;
  float m = -INFINITY;
  for (int p = 0; p < parts; ++p) m = fmaxf(m, lse_m[(int64_t)p * nq + r]);
  float ssum = 0.f;
  for (int p = 0; p < parts; ++p) ssum += lse_s[(int64_t)p * nq + r] * expf(lse_m[(int64_t)p * nq + r] - m);
  float lse = m + logf(ssum);
  if (lse_out) lse_out[r] = lse;
  int64_t t = target ? min(max(target[r], (int64_t)0), n_items - 1) : 0;
  const float *q = x + r * D, *v = item_p + t * D;
  float sc = 0.f;
  for (int k = 0; k < D; ++k) sc = fmaf(q[k], v[k], sc);
  row_loss[r] = target ? lse - sc : 0.f;
}

__global__ void k_mean(const float *__restrict__ v, int64_t n, float *out) {
  __shared__ double sm[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)v[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(sm[0] / (double)n);
}

// merge `parts` sorted lists per row; order (score desc, id asc); id -1 = empty slot
__global__ void k_topk_merge(const int64_t *__restrict__ ids, const float *__restrict__ scores, int parts, int64_t nq,
                             int K, int64_t *__restrict__ out_ids, float *__restrict__ out_scores,
                             const int32_t *__restrict__ row_map) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nq) return;
  const int64_t ro = row_map ? (int64_t)row_map[r] : r;
  unsigned char head[64];
  for (int p = 0; p < parts; ++p) head[p] = 0;
  for (int j = 0; j < K; ++j) {
    int best = -1;
    float bs = -INFINITY;
    int64_t bi = -1;
    for (int p = 0; p < parts; ++p) {
      if (head[p] >= K) continue;
      int64_t o = ((int64_t)p * nq + r) * K + head[p];
      int64_t id = ids[o];
      if (id < 0) continue;
      float s = scores[o];
      if (best < 0 || s > bs || (s == bs && id < bi)) { best = p; bs = s; bi = id; }
    }
    if (best >= 0) head[best]++;
    out_ids[ro * K + j] = bi;
    out_scores[ro * K + j] = best >= 0 ? bs : -INFINITY;
  }
}

// ---------------------------------------------------------------------------------------------
// metrics
__device__ __forceinline__ int64_t count_le(const int64_t *__restrict__ a, int64_t n, int64_t x) {
  int64_t lo = 0, hi = n;  // number of elements <= x
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// column the reference's swap puts item t in (general_dataloader.py:325-327, trainer.py:347-350)
__device__ int64_t ref_column(const int64_t *__restrict__ pos, int64_t p, int64_t t) {
  if (p == 0) return t;
  const int64_t c = count_le(pos, p, p - 1);  // positives already inside [0, p)
  const int64_t m = p - c;                    // length of each swap list
  const bool is_pos = csr_contains(pos, p, t);
  if (is_pos && t >= p) {
    int64_t i = count_le(pos, p, t) - c;      // t = b_i (1-based)
    int64_t want = m + 1 - i;                 // partner a_want: want-th non-positive column in [0, p)
    int64_t lo = 0, hi = p - 1;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (mid + 1 - count_le(pos, p, mid) >= want) hi = mid; else lo = mid + 1;
    }
    return lo;
  }
  if (!is_pos && t < p) {
    int64_t i = t + 1 - count_le(pos, p, t);  // t = a_i
    return pos[p - i];                        // partner b_{m+1-i} = pos[c + (m+1-i) - 1]
  }
  return t;
}

constexpr int kMetricThreads = 128;

__global__ void __launch_bounds__(kMetricThreads) k_topk_metrics(
    const int64_t *__restrict__ topk, int64_t nq, int K, int64_t n_items, const int64_t *__restrict__ pos_indptr,
    const int64_t *__restrict__ pos_indices, const double *__restrict__ discount, const double *__restrict__ idcg,
    double *__restrict__ block_part, uint8_t *__restrict__ hit_out, int64_t *__restrict__ ref_idx) {
  extern __shared__ double wsum[];  // [warps][6][K]
  const int warps = kMetricThreads / 32;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = r < nq;
  const int64_t *pos = nullptr;
  int64_t plen = 0;
  if (active) {
    int64_t p0 = pos_indptr[r];
    plen = pos_indptr[r + 1] - p0;
    pos = pos_indices + p0;
  }
  int cum = 0, first = -1;
  double dcg = 0.0, sum_pre = 0.0;
  for (int j = 0; j < K; ++j) {
    double vals[RB2_NUM_METRICS] = {0, 0, 0, 0, 0, 0};
    if (active) {
      int64_t id = topk[r * K + j];
      bool h = id >= 0 && plen > 0 && csr_contains(pos, plen, id);
      if (hit_out) hit_out[r * K + j] = h ? 1 : 0;
      if (ref_idx) ref_idx[r * (K + 1) + j] = n_items - 1 - ref_column(pos, plen, id >= 0 ? id : 0);
      if (h) {
        ++cum;
        if (first < 0) first = j;
        dcg += discount[j];
        sum_pre += (double)cum / (double)(j + 1);
      }
      if (plen > 0) {
        int64_t lim = plen < (int64_t)(j + 1) ? plen : (int64_t)(j + 1);
        vals[RB2_M_RECALL] = (double)cum / (double)plen;      // metrics.py:110
        vals[RB2_M_MRR] = first >= 0 ? 1.0 / (double)(first + 1) : 0.0;  // metrics.py:57-64
        vals[RB2_M_NDCG] = dcg / idcg[lim - 1];               // metrics.py:131-146
        vals[RB2_M_HIT] = cum > 0 ? 1.0 : 0.0;                // metrics.py:41-42
        vals[RB2_M_PRECISION] = (double)cum / (double)(j + 1);  // metrics.py:164
        vals[RB2_M_MAP] = sum_pre / (double)lim;              // metrics.py:84-92
      }
    }
#pragma unroll
    for (int m = 0; m < RB2_NUM_METRICS; ++m) {
      double v = vals[m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) wsum[(warp * RB2_NUM_METRICS + m) * K + j] = v;
    }
  }
  if (active && ref_idx) ref_idx[r * (K + 1) + K] = n_items;  // shape column, evaluators.py:69,75
  __syncthreads();
  for (int e = threadIdx.x; e < RB2_NUM_METRICS * K; e += blockDim.x) {
    double s = 0.0;
    for (int w2 = 0; w2 < warps; ++w2) s += wsum[w2 * RB2_NUM_METRICS * K + e];
    block_part[(int64_t)blockIdx.x * RB2_NUM_METRICS * K + e] = s;
  }
}

__global__ void k_metrics_reduce(const double *__restrict__ block_part, int64_t n_blocks, int n, double *sums) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double s = 0.0;
  for (int64_t b = 0; b < n_blocks; ++b) s += block_part[b * n + e];
  sums[e] = s;
}

struct FsPlan {
  int n_split;
  int64_t items_per_split;
  size_t smem;
};

template <int D>
FsPlan plan_fp32(int64_t nq, int64_t n_local, int K) {
  constexpr int TI = FsCfg<D>::TI;
  int64_t bx = (nq + kRows - 1) / kRows;
  int64_t want = ((int64_t)rb2_num_sms() * 2 + bx - 1) / bx;
  int64_t max_split = (n_local + TI - 1) / TI;
  int64_t ns = want < 1 ? 1 : want;
  if (ns > 64) ns = 64;
  if (ns > max_split) ns = max_split;
  if (ns < 1) ns = 1;
  int64_t per = (n_local + ns - 1) / ns;
  per = (per + TI - 1) / TI * TI;
  ns = (n_local + per - 1) / per;
  if (ns < 1) ns = 1;
  FsPlan p;
  p.n_split = (int)ns;
  p.items_per_split = per;
  p.smem = 2 * FsCfg<D>::TILE_FLOATS * sizeof(float) + (size_t)K * kRows * (sizeof(float) + sizeof(int));
  return p;
}

FsPlan plan_any(int dim, int64_t nq, int64_t n_local, int K) {
  switch (dim) {
    case 16: return plan_fp32<16>(nq, n_local, K);
    case 32: return plan_fp32<32>(nq, n_local, K);
    case 64: return plan_fp32<64>(nq, n_local, K);
    case 128: return plan_fp32<128>(nq, n_local, K);
    default: { FsPlan p; p.n_split = 0; p.items_per_split = 0; p.smem = 0; return p; }
  }
}

}  // namespace

size_t rb2_fullsort_fp32_workspace(int64_t nq, int64_t n_items_local, int32_t dim, int32_t k);
int rb2_fullsort_fp32_lse(const float *query_p, const int64_t *query_ids, int64_t nq, const float *item_p,
                          int64_t n_items_local, int64_t item_base, int32_t dim, const int64_t *hist_indptr,
                          const int64_t *hist_indices, int32_t k, int64_t *out_ids, float *out_scores,
                          void *workspace, size_t workspace_bytes, cudaStream_t st, const int32_t *row_map,
                          float *lse_m, float *lse_s, int *parts_out);
// implemented in fullsort_tc.cu
size_t rb2_fullsort_tc_workspace_bytes(int64_t nq, int64_t n_items_local, int32_t dim, int32_t k);
int rb2_fullsort_tc(const float *query_p, const int64_t *query_ids, int64_t nq, const float *item_p,
                    int64_t n_items_local, int64_t item_base, int32_t dim, const int64_t *hist_indptr,
                    const int64_t *hist_indices, int32_t k, int64_t *out_ids, float *out_scores, void *workspace,
                    size_t workspace_bytes, cudaStream_t st);

extern "C" size_t rb2_fullsort_workspace_bytes(int64_t nq, int64_t n_items_local, int32_t dim, int32_t k,
                                               int32_t mode) {
  if (mode == RB2_SCORER_TC) return rb2_fullsort_tc_workspace_bytes(nq, n_items_local, dim, k);
  return rb2_fullsort_fp32_workspace(nq, n_items_local, dim, k);
}

size_t rb2_fullsort_fp32_workspace(int64_t nq, int64_t n_items_local, int32_t dim, int32_t k) {
  FsPlan p = plan_any(dim, nq, n_items_local, k);
  // a later call on fewer rows (tensor-core fallback) may split the items further: size for 64 parts
  int64_t parts = p.n_split > 0 ? 64 : 0;
  int64_t rows = nq < 4096 ? nq : (nq * p.n_split + 63) / 64;  // enough for nq rows at n_split, or few rows at 64
  if (rows < 1) rows = 1;
  Carver c(nullptr);
  c.take<int64_t>((size_t)parts * rows * k);
  c.take<float>((size_t)parts * rows * k);
  return c.off + 256;
}

int rb2_fullsort_fp32(const float *query_p, const int64_t *query_ids, int64_t nq, const float *item_p,
                      int64_t n_items_local, int64_t item_base, int32_t dim, const int64_t *hist_indptr,
                      const int64_t *hist_indices, int32_t k, int64_t *out_ids, float *out_scores, void *workspace,
                      size_t workspace_bytes, cudaStream_t st, const int32_t *row_map) {
  return rb2_fullsort_fp32_lse(query_p, query_ids, nq, item_p, n_items_local, item_base, dim, hist_indptr,
                               hist_indices, k, out_ids, out_scores, workspace, workspace_bytes, st, row_map, nullptr,
                               nullptr, nullptr);
}

// same, optionally with the per-split logsumexp partials (lse_m / lse_s sized [64, nq]); *parts_out =
// number of splits used
int rb2_fullsort_fp32_lse(const float *query_p, const int64_t *query_ids, int64_t nq, const float *item_p,
                          int64_t n_items_local, int64_t item_base, int32_t dim, const int64_t *hist_indptr,
                          const int64_t *hist_indices, int32_t k, int64_t *out_ids, float *out_scores,
                          void *workspace, size_t workspace_bytes, cudaStream_t st, const int32_t *row_map,
                          float *lse_m, float *lse_s, int *parts_out) {
  FsPlan p = plan_any(dim, nq, n_items_local, k);
  RB2_REQUIRE(p.n_split > 0, RB2_EINVAL, "rb2_fullsort_topk: embedding dim %d not supported by the fp32 scorer (16, 32, 64, 128)",
              (int)dim);
  {
    // The workspace may have been sized for a different row count (the tensor-core scorer's fallback re-scores
    // n_fail < nq rows with the buffer of the full call, and fewer rows want MORE item splits): never split further
    // than the per-split lists fit -- fewer splits only cost parallelism.
    const size_t per_list = (size_t)nq * k * (sizeof(int64_t) + sizeof(float)) + 512;
    int64_t fit = (int64_t)(workspace_bytes / per_list);
    if (fit < 1) fit = 1;
    if (p.n_split > fit && !lse_m) {
      const int ti = dim == 16 ? FsCfg<16>::TI : dim == 32 ? FsCfg<32>::TI : dim == 64 ? FsCfg<64>::TI : FsCfg<128>::TI;
      int64_t per = (n_items_local + fit - 1) / fit;
      per = (per + ti - 1) / ti * ti;
      p.items_per_split = per;
      p.n_split = (int)((n_items_local + per - 1) / per);
    }
  }
  if (parts_out) *parts_out = p.n_split;
  RB2_REQUIRE(p.smem <= 200 * 1024, RB2_EINVAL, "rb2_fullsort_topk: k=%d too large", (int)k);
  Carver c(workspace);
  int64_t *part_ids = c.take<int64_t>((size_t)p.n_split * nq * k);
  float *part_sc = c.take<float>((size_t)p.n_split * nq * k);
  RB2_REQUIRE(c.off <= workspace_bytes, RB2_EWORKSPACE, "rb2_fullsort_topk: workspace %zu < %zu", workspace_bytes,
              c.off);
  const bool direct = p.n_split == 1;
  dim3 grid((unsigned)((nq + kRows - 1) / kRows), (unsigned)p.n_split);
#define RB2_FS(D_)                                                                                              \
  {                                                                                                             \
    if (lse_m) {                                                                                                 \
      RB2_CUDA(cudaFuncSetAttribute(k_fullsort_fp32<D_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                    (int)p.smem));                                                              \
      k_fullsort_fp32<D_, true><<<grid, kRows, p.smem, st>>>(                                                   \
          query_p, query_ids, nq, item_p, n_items_local, item_base, hist_indptr, hist_indices, k,               \
          p.items_per_split, direct ? out_ids : part_ids, direct ? out_scores : part_sc, row_map, direct ? 1 : 0, \
          lse_m, lse_s);                                                                                        \
    } else {                                                                                                    \
      RB2_CUDA(cudaFuncSetAttribute(k_fullsort_fp32<D_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                    (int)p.smem));                                                              \
      k_fullsort_fp32<D_, false><<<grid, kRows, p.smem, st>>>(                                                  \
          query_p, query_ids, nq, item_p, n_items_local, item_base, hist_indptr, hist_indices, k,               \
          p.items_per_split, direct ? out_ids : part_ids, direct ? out_scores : part_sc, row_map, direct ? 1 : 0, \
          nullptr, nullptr);                                                                                    \
    }                                                                                                           \
  }
  {
    ProfScope prof(RB2_ST_FULLSORT, st);
    switch (dim) {
      case 16: RB2_FS(16) break;
      case 32: RB2_FS(32) break;
      case 64: RB2_FS(64) break;
      case 128: RB2_FS(128) break;
    }
  }
#undef RB2_FS
  RB2_CUDA(cudaGetLastError());
  if (!direct) {
    ProfScope prof(RB2_ST_TOPK_MERGE, st);
    k_topk_merge<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(part_ids, part_sc, p.n_split, nq, k, out_ids,
                                                               out_scores, row_map);
    RB2_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int rb2_fullsort_topk(const float *query_p, const int64_t *query_ids, int64_t nq, const float *item_p,
                                 int64_t n_items_local, int64_t item_base, int32_t dim, const int64_t *hist_indptr,
                                 const int64_t *hist_indices, int32_t k, int32_t mode, int64_t *out_ids,
                                 float *out_scores, void *workspace, size_t workspace_bytes, void *stream) {
  RB2_REQUIRE(query_p && item_p && out_ids && out_scores && workspace, RB2_EINVAL, "rb2_fullsort_topk: null argument");
  RB2_REQUIRE(k >= 1 && k <= 128, RB2_EINVAL, "rb2_fullsort_topk: k=%d outside 1..128", (int)k);
  RB2_REQUIRE(n_items_local >= 1 && item_base >= 0, RB2_EINVAL, "rb2_fullsort_topk: bad item shard");
  RB2_REQUIRE(item_base + n_items_local < ((int64_t)1 << 31), RB2_EINVAL, "rb2_fullsort_topk: item ids must fit int32");
  RB2_REQUIRE((hist_indptr == nullptr) == (hist_indices == nullptr) || hist_indptr != nullptr, RB2_EINVAL,
              "rb2_fullsort_topk: hist_indices without hist_indptr");
  if (nq <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == RB2_SCORER_TC)
    return rb2_fullsort_tc(query_p, query_ids, nq, item_p, n_items_local, item_base, dim, hist_indptr, hist_indices,
                           k, out_ids, out_scores, workspace, workspace_bytes, st);
  RB2_REQUIRE(mode == RB2_SCORER_FP32, RB2_EINVAL, "rb2_fullsort_topk: unknown mode %d", (int)mode);
  return rb2_fullsort_fp32(query_p, query_ids, nq, item_p, n_items_local, item_base, dim, hist_indptr, hist_indices,
                           k, out_ids, out_scores, workspace, workspace_bytes, st, nullptr);
}

extern "C" int rb2_topk_merge(const int64_t *ids, const float *scores, int32_t parts, int64_t nq, int32_t k,
                              int64_t *out_ids, float *out_scores, void *stream) {
  RB2_REQUIRE(ids && scores && out_ids && out_scores, RB2_EINVAL, "rb2_topk_merge: null argument");
  RB2_REQUIRE(parts >= 1 && parts <= 64, RB2_EINVAL, "rb2_topk_merge: parts=%d outside 1..64", (int)parts);
  if (nq <= 0) return 0;
  k_topk_merge<<<(unsigned)((nq + 127) / 128), 128, 0, (cudaStream_t)stream>>>(ids, scores, parts, nq, k, out_ids,
                                                                              out_scores, nullptr);
  RB2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" size_t rb2_topk_metrics_workspace_bytes(int64_t nq, int32_t k) {
  int64_t blocks = (nq + kMetricThreads - 1) / kMetricThreads;
  return rb2_align((size_t)blocks * RB2_NUM_METRICS * k * sizeof(double)) + 256;
}

extern "C" int rb2_topk_metrics(const int64_t *topk_ids, int64_t nq, int32_t k, int64_t n_items,
                                const int64_t *pos_indptr, const int64_t *pos_indices, const double *discount,
                                const double *idcg, double *sums, uint8_t *hit, int64_t *ref_idx, void *workspace,
                                size_t workspace_bytes, void *stream) {
  RB2_REQUIRE(topk_ids && pos_indptr && pos_indices && discount && idcg && sums && workspace, RB2_EINVAL,
              "rb2_topk_metrics: null argument");
  RB2_REQUIRE(k >= 1 && k <= 128, RB2_EINVAL, "rb2_topk_metrics: k=%d outside 1..128", (int)k);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (nq + kMetricThreads - 1) / kMetricThreads;
  RB2_REQUIRE(workspace_bytes >= rb2_topk_metrics_workspace_bytes(nq, k) - 256, RB2_EWORKSPACE,
              "rb2_topk_metrics: workspace too small");
  if (nq <= 0) {
    RB2_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * RB2_NUM_METRICS * k, st));
    return 0;
  }
  double *part = reinterpret_cast<double *>(workspace);
  ProfScope prof(RB2_ST_METRICS, st, 2);
  size_t smem = (size_t)(kMetricThreads / 32) * RB2_NUM_METRICS * k * sizeof(double);
  RB2_CUDA(cudaFuncSetAttribute(k_topk_metrics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_topk_metrics<<<(unsigned)blocks, kMetricThreads, smem, st>>>(topk_ids, nq, k, n_items, pos_indptr, pos_indices,
                                                                discount, idcg, part, hit, ref_idx);
  int n = RB2_NUM_METRICS * k;
  k_metrics_reduce<<<(n + 127) / 128, 128, 0, st>>>(part, blocks, n, sums);
  RB2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb2_fullsort_topk_s(const float *query_p, const int64_t *query_ids, int64_t nq, const float *item_p,
                                   int64_t n_items_local, int64_t item_base, int32_t dim, const int64_t *hist_indptr,
                                   const int64_t *hist_indices, int32_t k, int32_t mode, int64_t *out_ids,
                                   float *out_scores, void *workspace, size_t workspace_bytes, void *stream,
                                   rb2_scorer_state *h_state) {
  ScorerScope scope(h_state);
  return rb2_fullsort_topk(query_p, query_ids, nq, item_p, n_items_local, item_base, dim, hist_indptr, hist_indices, k,
                           mode, out_ids, out_scores, workspace, workspace_bytes, stream);
}

// tensor-core form (fullsort_tc.cu), dim == 64 and k <= 16
size_t rb2_fullsort_tc_lse_workspace_bytes(int64_t nq, int64_t n_items, int32_t dim, int32_t k);
int rb2_fullsort_tc_lse(const float *x, int64_t nq, const float *item_p, int64_t n_items, int32_t dim, int32_t k,
                        int64_t *out_ids, float *out_scores, void *workspace, size_t workspace_bytes, cudaStream_t st,
                        float *lse_m, float *lse_s, int *parts_out);
extern "C" int rb2_ce_head_set_scorer(int32_t mode) {   // 0 = tensor cores where covered, 1 = CUDA-core fp32 kernel
  if (mode != 0 && mode != 1) return RB2_EINVAL;
  rb2_cur_scorer().ce_scorer = mode;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// CE head: logits = X . E^T never materialised; loss = mean(logsumexp - logit[target]); top-K of the
// same pass with the pad column masked.
extern "C" size_t rb2_ce_head_workspace_bytes(int64_t nq, int64_t n_items, int32_t dim, int32_t k) {
  Carver c(nullptr);
  c.take<float>((size_t)64 * nq);
  c.take<float>((size_t)64 * nq);
  c.take<float>(nq);
  size_t fs = rb2_fullsort_fp32_workspace(nq, n_items, dim, k), tc = rb2_fullsort_tc_lse_workspace_bytes(nq, n_items, dim, k);
  return c.off + (fs > tc ? fs : tc) + 256;
}

extern "C" int rb2_ce_head(const float *x, int64_t nq, const float *item_p, int64_t n_items, int32_t dim,
                           const int64_t *target, int32_t k, float *loss_out, float *lse_out, int64_t *topk_ids,
                           float *topk_scores, void *workspace, size_t workspace_bytes, void *stream) {
  RB2_REQUIRE(x && item_p && topk_ids && topk_scores && workspace, RB2_EINVAL, "rb2_ce_head: null argument");
  RB2_REQUIRE(k >= 1 && k <= 128, RB2_EINVAL, "rb2_ce_head: k=%d outside 1..128", (int)k);
  RB2_REQUIRE(!target || loss_out, RB2_EINVAL, "rb2_ce_head: target given without loss_out");
  if (nq <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(workspace);
  float *lse_m = c.take<float>((size_t)64 * nq);
  float *lse_s = c.take<float>((size_t)64 * nq);
  float *row_loss = c.take<float>(nq);
  size_t fs_bytes = rb2_fullsort_fp32_workspace(nq, n_items, dim, k);
  size_t tc_bytes = rb2_fullsort_tc_lse_workspace_bytes(nq, n_items, dim, k);
  const bool use_tc = rb2_cur_scorer().ce_scorer == 0 && tc_bytes > 0;
  if (use_tc && tc_bytes > fs_bytes) fs_bytes = tc_bytes;
  void *fs_ws = c.take<char>(fs_bytes);
  RB2_REQUIRE(c.off <= workspace_bytes, RB2_EWORKSPACE, "rb2_ce_head: workspace %zu < %zu", workspace_bytes, c.off);
  int parts = 0;
  int rc = use_tc ? rb2_fullsort_tc_lse(x, nq, item_p, n_items, dim, k, topk_ids, topk_scores, fs_ws, fs_bytes, st, lse_m,
                                        lse_s, &parts)
                  : rb2_fullsort_fp32_lse(x, nullptr, nq, item_p, n_items, 0, dim, nullptr, nullptr, k, topk_ids,
                                          topk_scores, fs_ws, fs_bytes, st, nullptr, lse_m, lse_s, &parts);
  if (rc) return rc;
  ProfScope prof(RB2_ST_MISC, st, 2);
  unsigned blocks = (unsigned)((nq + 127) / 128);
  switch (dim) {
    case 16: k_ce_finish<16><<<blocks, 128, 0, st>>>(x, item_p, nq, n_items, target, lse_m, lse_s, parts, lse_out, row_loss); break;
    case 32: k_ce_finish<32><<<blocks, 128, 0, st>>>(x, item_p, nq, n_items, target, lse_m, lse_s, parts, lse_out, row_loss); break;
    case 64: k_ce_finish<64><<<blocks, 128, 0, st>>>(x, item_p, nq, n_items, target, lse_m, lse_s, parts, lse_out, row_loss); break;
    case 128: k_ce_finish<128><<<blocks, 128, 0, st>>>(x, item_p, nq, n_items, target, lse_m, lse_s, parts, lse_out, row_loss); break;
  }
  if (loss_out) k_mean<<<1, 256, 0, st>>>(row_loss, nq, loss_out);
  RB2_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Compatibility path: the score matrix itself (BPR.full_sort_predict, bpr.py:91-96) for callers that insist on
// it -- an UNMODIFIED reference Trainer._full_sort_batch_eval (trainer.py:328-352) masks and top-k's the matrix
// with ATen.  The reference asks for 1-2 users per call (general_dataloader.py:330-334), so this is a stream over
// the item table: one thread per item, kScoreQ query rows per pass held in shared memory, every score the
// canonical fp32 chain s = fmaf(q[k], v[k], s), k ascending (bit-identical to the top-K kernels).
namespace {
constexpr int kScoreQ = 8;
constexpr int kScoreThreads = 256;

__global__ void __launch_bounds__(kScoreThreads) k_fullsort_scores(const float *__restrict__ query_p,
                                                                   const int64_t *__restrict__ query_ids, int64_t nq,
                                                                   int64_t n_query_rows,
                                                                   const float *__restrict__ item_p, int64_t n_items,
                                                                   int dim, float *__restrict__ out,
                                                                   int32_t *range_error) {
  extern __shared__ float qs[];                       // [kScoreQ, dim]
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int d4 = dim / 4;
  for (int64_t q0 = (int64_t)blockIdx.y * kScoreQ; q0 < nq; q0 += (int64_t)gridDim.y * kScoreQ) {
    const int nqq = (int)min((int64_t)kScoreQ, nq - q0);
    __syncthreads();
    for (int i = threadIdx.x; i < nqq * dim; i += blockDim.x) {
      int64_t row = query_ids ? query_ids[q0 + i / dim] : q0 + i / dim;
      if (row < 0 || row >= n_query_rows) {
        if (range_error) *range_error = 1;
        row = min(max(row, (int64_t)0), n_query_rows - 1);
      }
      qs[i] = query_p[row * dim + i % dim];
    }
    __syncthreads();
    if (item < n_items) {
      float s[kScoreQ];
#pragma unroll
      for (int q = 0; q < kScoreQ; ++q) s[q] = 0.f;
      const float4 *vp = reinterpret_cast<const float4 *>(item_p + item * dim);
      for (int k = 0; k < d4; ++k) {
        const float4 v = __ldg(vp + k);
#pragma unroll
        for (int q = 0; q < kScoreQ; ++q) {
          if (q < nqq) {
            const float4 u = reinterpret_cast<const float4 *>(qs + q * dim)[k];    // broadcast
            s[q] = fmaf(u.x, v.x, s[q]);
            s[q] = fmaf(u.y, v.y, s[q]);
            s[q] = fmaf(u.z, v.z, s[q]);
            s[q] = fmaf(u.w, v.w, s[q]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kScoreQ; ++q)
        if (q < nqq) out[(q0 + q) * n_items + item] = s[q];
    }
  }
}
}  // namespace

extern "C" int rb2_fullsort_scores(const float *query_p, const int64_t *query_ids, int64_t nq, int64_t n_query_rows,
                                   const float *item_p, int64_t n_items, int32_t dim, float *out_scores,
                                   void *stream) {
  RB2_REQUIRE(query_p && item_p && out_scores, RB2_EINVAL, "rb2_fullsort_scores: null argument");
  RB2_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 1024, RB2_EINVAL, "rb2_fullsort_scores: dim %d (multiple of 4, <= 1024)",
              (int)dim);
  RB2_REQUIRE(n_items > 0 && n_query_rows > 0 && nq >= 0, RB2_EINVAL, "rb2_fullsort_scores: sizes");
  if (nq == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned bx = (unsigned)((n_items + kScoreThreads - 1) / kScoreThreads);
  const int64_t passes = (nq + kScoreQ - 1) / kScoreQ;
  // enough blocks to fill the machine; one y-slice walks several query passes when there are many
  unsigned by = (unsigned)std::min<int64_t>(passes, std::max<int64_t>(1, (int64_t)rb2_num_sms() * 8 / bx));
  ProfScope prof(RB2_ST_FULLSORT, st, 1);
  k_fullsort_scores<<<dim3(bx, by), kScoreThreads, (size_t)kScoreQ * dim * sizeof(float), st>>>(
      query_p, query_ids, nq, n_query_rows, item_p, n_items, dim, out_scores, nullptr);
  RB2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb2_ce_head_s(const float *x, int64_t nq, const float *item_p, int64_t n_items, int32_t dim,
                             const int64_t *target, int32_t k, float *loss_out, float *lse_out, int64_t *topk_ids,
                             float *topk_scores, void *workspace, size_t workspace_bytes, void *stream,
                             rb2_scorer_state *h_state) {
  ScorerScope scope(h_state);
  return rb2_ce_head(x, nq, item_p, n_items, dim, target, k, loss_out, lse_out, topk_ids, topk_scores, workspace,
                     workspace_bytes, stream);
}
