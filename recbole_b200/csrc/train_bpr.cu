// train_bpr.cu -- fused BPR-MF training step for sm_100a (HBM-bound gather / row-sparse update).
//
// Replaces, per batch, the reference's  BPR.calculate_loss (bpr.py:74-83) + BPRLoss (loss.py:43-49)
// + autograd's dense embedding backward (trainer.py:170) + dense torch.optim step (trainer.py:173).
//
// Data flow:
//   k_make_keys   ids -> (row key, occurrence) pairs for the user side (B) and the item side (2B)
//   radix sort    both pair lists by row id (cub::DeviceRadixSort, only the bits the table needs)
//   k_user_side   one lane-group walks a tile of T sorted user occurrences.  For every sample it
//                 gathers the positive and negative item rows, computes x = u.(vi - vj), the loss
//                 term and g = dL/dx, accumulates du += g*(vi - vj) for the run of equal user ids,
//                 stores gu[s] = g*u (the only per-sample intermediate, 4*d bytes) and, when the run
//                 is complete inside the tile, applies the optimizer step to the user row in place.
//   k_fixup       runs that straddle tile boundaries are reduced tile-by-tile in a fixed order and
//                 stepped once (deterministic: no float atomics anywhere).
//   k_item_side   same walk over the 2B sorted item occurrences: dv += (+/-) gu[s]; step the item row.
//   k_fixup       ditto for items.
//   k_loss        fixed-order reduction of the per-tile loss partials.
// Every touched row is read (p, m, v) and written (p, m, v) exactly once per step.
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

#include <cub/device/device_scan.cuh>

#include "bucket_sort.cuh"
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr float kGamma = 1e-10f;  // BPRLoss gamma, loss.py:43
constexpr int64_t kLossParts = 131072;  // capacity of the per-group loss partial buffer (forward-only path)

struct BprWs {
  WsHeader *hdr;
  uint32_t *ukey, *uval, *ukey_s, *uval_s;  // [B]
  uint32_t *ikey, *ival, *ikey_s, *ival_s;  // [2B]
  int2 *pn;                                 // [B] (pos, neg) as int32
  float *gu;                                // [B, D]
  float *u_head, *u_tail, *i_head, *i_tail; // [tiles, D]
  uint8_t *u_fh, *u_ft, *i_fh, *i_ft;       // [tiles] flags
  double *loss_part;                        // [user tiles]
  float *u_blk, *i_blk;                     // [tiles / kChainBlk + 1, D] sums of whole blocks of chain partials
  uint8_t *u_blk_ok, *i_blk_ok;             // [tiles / kChainBlk + 1]
  // peer-memory (multi-GPU) step only
  uint32_t *alt_ukey_s, *alt_uval_s, *alt_ikey_s, *alt_ival_s;   // the OTHER batch slot (sorted keys of the next batch)
  int2 *alt_pn;
  unsigned long long *src, *dst;            // [2B] per item occurrence: where to read the row / push its gradient
  uint32_t *fetch_list;                     // [2B] remote rows that occur more than once (fetched once)
  uint32_t *fetch_count;                    // [1]
  double *loss_sum;                         // [1] this rank's un-normalised loss sum
  void *cub_tmp;
  size_t cub_bytes;
};
constexpr int kChainBlk = 16;               // tiles per pre-reduced block of a long chain (k_chain_blocks)
constexpr int kTileMax = 64;                // upper bound of pick_tile()

// tile length: long enough that few runs straddle tiles, short enough to fill the machine
inline int pick_tile(int64_t n_occ, int lanes) {
  int64_t groups_wanted = (int64_t)rb2_num_sms() * 2048 / lanes;  // one full wave of lane groups
  int64_t t = n_occ / (groups_wanted > 0 ? groups_wanted : 1);
  if (t < 8) t = 8;
  if (t > 64) t = 64;
  return (int)t;
}

inline int64_t max_tiles(int64_t n_occ) { return (n_occ + 7) / 8; }

size_t carve(BprWs &w, void *base, int64_t B, int dim, bool p2p = false) {
  Carver c(base);
  w.hdr = c.take<WsHeader>(1);
  w.ukey = c.take<uint32_t>(B);
  w.uval = c.take<uint32_t>(B);
  w.ukey_s = c.take<uint32_t>(B);
  w.uval_s = c.take<uint32_t>(B);
  w.ikey = c.take<uint32_t>(2 * B);
  w.ival = c.take<uint32_t>(2 * B);
  w.ikey_s = c.take<uint32_t>(2 * B);
  w.ival_s = c.take<uint32_t>(2 * B);
  w.pn = c.take<int2>(B);
  w.gu = c.take<float>(B * dim);
  int64_t tu = max_tiles(B), ti = max_tiles(2 * B);
  w.u_head = c.take<float>(tu * dim);
  w.u_tail = c.take<float>(tu * dim);
  w.i_head = c.take<float>(ti * dim);
  w.i_tail = c.take<float>(ti * dim);
  w.u_fh = c.take<uint8_t>(tu);
  w.u_ft = c.take<uint8_t>(tu);
  w.i_fh = c.take<uint8_t>(ti);
  w.i_ft = c.take<uint8_t>(ti);
  w.loss_part = c.take<double>(tu > kLossParts ? tu : kLossParts);
  w.u_blk = c.take<float>((tu / kChainBlk + 1) * dim);
  w.i_blk = c.take<float>((ti / kChainBlk + 1) * dim);
  w.u_blk_ok = c.take<uint8_t>(tu / kChainBlk + 1);
  w.i_blk_ok = c.take<uint8_t>(ti / kChainBlk + 1);
  w.src = w.dst = nullptr;
  w.fetch_list = w.fetch_count = nullptr;
  w.loss_sum = nullptr;
  w.alt_ukey_s = w.alt_uval_s = w.alt_ikey_s = w.alt_ival_s = nullptr;
  w.alt_pn = nullptr;
  if (p2p) {
    w.alt_ukey_s = c.take<uint32_t>(B);
    w.alt_uval_s = c.take<uint32_t>(B);
    w.alt_ikey_s = c.take<uint32_t>(2 * B);
    w.alt_ival_s = c.take<uint32_t>(2 * B);
    w.alt_pn = c.take<int2>(B);
    w.src = c.take<unsigned long long>(2 * B);
    w.dst = c.take<unsigned long long>(2 * B);
    w.fetch_list = c.take<uint32_t>(2 * B);
    w.fetch_count = c.take<uint32_t>(64);
    w.loss_sum = c.take<double>(32);
  }
  size_t b1 = 0, b2 = 0;
  b1 = rb2sort::tmp_bytes(B);
  b2 = rb2sort::tmp_bytes(2 * B);
  w.cub_bytes = b1 > b2 ? b1 : b2;
  w.cub_tmp = c.take<char>(w.cub_bytes);
  return c.off;
}

inline int bits_for(int64_t n) {
  int b = 1;
  while (b < 32 && ((int64_t)1 << b) < n) ++b;
  return b;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_make_keys(const int64_t *__restrict__ user, const int64_t *__restrict__ pos,
                            const int64_t *__restrict__ neg, int64_t B, int64_t n_users, int64_t n_items,
                            BprWs w, int64_t user_base = 0) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  int64_t u = user[s] - user_base, p = pos[s], n = neg[s];
  bool bad = (u < 0) | (u >= n_users) | (p < 0) | (p >= n_items) | (n < 0) | (n >= n_items);
  if (bad) {
    w.hdr->range_error = 1;
    u = min(max(u, (int64_t)0), n_users - 1);
    p = min(max(p, (int64_t)0), n_items - 1);
    n = min(max(n, (int64_t)0), n_items - 1);
  }
  w.ukey[s] = (uint32_t)u;
  w.uval[s] = (uint32_t)s;
  w.ikey[2 * s] = (uint32_t)p;
  w.ival[2 * s] = (uint32_t)(2 * s);
  w.ikey[2 * s + 1] = (uint32_t)n;
  w.ival[2 * s + 1] = (uint32_t)(2 * s + 1);
  w.pn[s] = make_int2((int)p, (int)n);
}

struct Tables {
  float *up, *um, *uv;
  int32_t *ul;
  float *ip, *im, *iv;
  int32_t *il;
  float *ig;  // optional [n_items, d]: the item side writes the summed gradient here instead of stepping
  int32_t *itouched;  // optional [n_items]: set to 1 for every row the item side wrote to ig
};

__device__ __forceinline__ void bpr_sample(float x, float inv_b, float &loss_term, float &g) {
  // loss.py:48   -log(gamma + sigmoid(x)) ;  d/dx = -sig*(1-sig)/(gamma+sig), times 1/B for the mean.
  // For a trained model (x >> 0) the loss term log(den) ~ -(1 - den) and the factor (1 - sig) inherit the ABSOLUTE
  // rounding of sig, in the reference too; what must not be added on top is an absolute error in the logarithm (the
  // MUFU log2 has 2^-22: 1e-4 of the loss of a converged model) or a biased reciprocal.  So: correctly rounded
  // reciprocal for sig, and near 1 the logarithm as 2 atanh((den - 1) / (den + 1)) -- den - 1 is exact, the series
  // is RELATIVELY accurate (t^8 / 9 < 1e-7 for den in [0.75, 1.34]); __expf / __fdividef only where their error is
  // relative to the quantity itself.
  const float e = __expf(-x);
  const float sig = __frcp_rn(1.f + e);
  const float den = kGamma + sig;
  const float t = __fdividef(den - 1.f, den + 1.f), t2 = t * t;
  const float near1 = 2.f * t * fmaf(t2, fmaf(t2, fmaf(t2, 1.f / 7.f, 0.2f), 1.f / 3.f), 1.f);
  loss_term = -((den > 0.75f) ? near1 : __logf(den));
  g = -inv_b * __fdividef(sig * (1.f - sig), den);
}

// ---------------------------------------------------------------------------------------------
// user side
template <int D, bool LAZY>
__global__ void __launch_bounds__(kThreads, LAZY ? 1 : (D <= 64 ? 3 : (D == 128 ? 2 : 1))) k_user_side(Tables t, BprWs w, int64_t B, int T, int64_t n_tiles,
                                                         float inv_b, OptScalars o) {
  constexpr int LANES = RowCfg<D>::LANES;
  constexpr int UNR = 4;
  const int lane = threadIdx.x % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (tile >= n_tiles) return;
  const int64_t lo = tile * T, hi = min(lo + (int64_t)T, B);
  const uint32_t *__restrict__ keys = w.ukey_s;
  const uint32_t *__restrict__ vals = w.uval_s;
  const uint32_t kInvalid = 0xffffffffu;
  const uint32_t prev_key = lo > 0 ? keys[lo - 1] : kInvalid;

  uint32_t cur = kInvalid;
  bool started_before = false;
  Row<D> u = row_zero<D>(), acc = row_zero<D>();
  float loss_local = 0.f;
  uint8_t fh = 0, ft = 0;

  auto finish_run = [&](bool continues) {
    if (cur == kInvalid) return;
    if (!started_before && !continues) {
      if (LAZY) {
        row_update_full<D, true>(t.up, t.um, t.uv, t.ul, cur, lane, acc, o);
      } else {
        row_update<D>(t.up, t.um, t.uv, cur, lane, u, acc, o);
      }
    } else if (started_before) {
      row_st<D>(w.u_head, tile, lane, acc);
      fh = continues ? 2 : 1;
    } else {
      row_st<D>(w.u_tail, tile, lane, acc);
      ft = 1;
    }
  };

  for (int64_t base = lo; base < hi; base += UNR) {
    uint32_t k[UNR], s[UNR];
    Row<D> a[UNR], b[UNR], ur[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      int64_t p = base + j;
      bool ok = p < hi;
      k[j] = ok ? keys[p] : kInvalid;
      s[j] = ok ? vals[p] : 0u;
    }
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      if (k[j] != kInvalid) {
        int2 pn = w.pn[s[j]];
        a[j] = row_ld_effective<D, LAZY>(t.ip, t.im, t.iv, t.il, pn.x, lane, o);
        b[j] = row_ld_effective<D, LAZY>(t.ip, t.im, t.iv, t.il, pn.y, lane, o);
        // the user row is only needed where a run starts (a batch hits the same user many times)
        const uint32_t before = (j == 0) ? cur : k[j > 0 ? j - 1 : 0];
        if (k[j] != before) ur[j] = row_ld_effective<D, LAZY>(t.up, t.um, t.uv, t.ul, k[j], lane, o);
      }
    }
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      if (k[j] == kInvalid) break;
      if (k[j] != cur) {
        finish_run(false);
        cur = k[j];
        u = ur[j];
        acc = row_zero<D>();
        started_before = (base + j == lo) && (cur == prev_key);
      }
      // x = <u, vi> - <u, vj>   (bpr.py:81); the two per-lane partial dots share one shuffle reduction
      float x = group_sum<LANES>(row_dot_lane<D>(u, a[j]) - row_dot_lane<D>(u, b[j]), gmask);
      float lt, g;
      bpr_sample(x, inv_b, lt, g);
      loss_local += lt;
      row_fma_diff<D>(acc, g, a[j], b[j]);    // du += g*(vi - vj)
      row_st<D>(w.gu, s[j], lane, row_scale<D>(g, u));  // dvi = g*u ; dvj = -g*u
    }
  }
  finish_run(hi < B && keys[hi] == cur);
  if (lane == 0) {
    w.u_fh[tile] = fh;
    w.u_ft[tile] = ft;
    w.loss_part[tile] = (double)loss_local;
  }
}

// ---------------------------------------------------------------------------------------------
// item side: value = 2*s + is_neg ; contribution = (+/-) gu[s]
template <int D, bool LAZY>
__global__ void __launch_bounds__(kThreads) k_item_side(Tables t, BprWs w, int64_t n_occ, int T, int64_t n_tiles,
                                                         OptScalars o) {
  constexpr int LANES = RowCfg<D>::LANES;
  constexpr int UNR = 4;
  const int lane = threadIdx.x % LANES;
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (tile >= n_tiles) return;
  const int64_t lo = tile * T, hi = min(lo + (int64_t)T, n_occ);
  const uint32_t *__restrict__ keys = w.ikey_s;
  const uint32_t *__restrict__ vals = w.ival_s;
  const uint32_t kInvalid = 0xffffffffu;
  const uint32_t prev_key = lo > 0 ? keys[lo - 1] : kInvalid;

  uint32_t cur = kInvalid;
  bool started_before = false;
  Row<D> acc = row_zero<D>();
  uint8_t fh = 0, ft = 0;

  auto finish_run = [&](bool continues) {
    if (cur == kInvalid) return;
    if (!started_before && !continues) {
      if (t.ig) {
        row_st<D>(t.ig, cur, lane, acc);
        if (t.itouched && lane == 0) t.itouched[cur] = 1;
      } else {
        row_update_full<D, LAZY>(t.ip, t.im, t.iv, t.il, cur, lane, acc, o);
      }
    } else if (started_before) {
      row_st<D>(w.i_head, tile, lane, acc);
      fh = continues ? 2 : 1;
    } else {
      row_st<D>(w.i_tail, tile, lane, acc);
      ft = 1;
    }
  };

  for (int64_t base = lo; base < hi; base += UNR) {
    uint32_t k[UNR], s[UNR];
    Row<D> c[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      int64_t p = base + j;
      bool ok = p < hi;
      k[j] = ok ? keys[p] : kInvalid;
      s[j] = ok ? vals[p] : 0u;
    }
#pragma unroll
    for (int j = 0; j < UNR; ++j)
      if (k[j] != kInvalid) c[j] = row_ld<D>(w.gu, s[j] >> 1, lane);
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      if (k[j] == kInvalid) break;
      if (k[j] != cur) {
        finish_run(false);
        cur = k[j];
        acc = row_zero<D>();
        started_before = (base + j == lo) && (cur == prev_key);
      }
      row_fma<D>(acc, (s[j] & 1u) ? -1.f : 1.f, c[j]);
    }
  }
  finish_run(hi < n_occ && keys[hi] == cur);
  if (lane == 0) {
    w.i_fh[tile] = fh;
    w.i_ft[tile] = ft;
  }
}

// ---------------------------------------------------------------------------------------------
// runs that straddle tiles: tile `t` holds the head of such a run in tail[t]; the following tiles
// hold its continuation in head[t+1..] (flag 2 = continues further).  Fixed summation order.
template <int D, bool LAZY>
__global__ void __launch_bounds__(kThreads) k_fixup(float *P, float *M, float *V, int32_t *L,
                                                     const uint32_t *__restrict__ keys_sorted,
                                                     const float *__restrict__ head, const float *__restrict__ tail,
                                                     const uint8_t *__restrict__ fh, const uint8_t *__restrict__ ft,
                                                     int64_t n_occ, int T, int64_t n_tiles, OptScalars o,
                                                     float *__restrict__ grad_out, int32_t *__restrict__ touched) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (tile >= n_tiles || !ft[tile]) return;
  int64_t last_pos = min((tile + 1) * (int64_t)T, n_occ) - 1;
  uint32_t key = keys_sorted[last_pos];
  Row<D> acc = row_ld<D>(tail, tile, lane);
  // the chain of continuation partials can be thousands of tiles long for a hot row: walk it CH tiles
  // at a time with all loads in flight (still a fixed summation order)
  constexpr int CH = 8;
  bool done = false;
  for (int64_t j0 = tile + 1; j0 < n_tiles && !done; j0 += CH) {
    uint8_t f[CH];
    Row<D> part[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) f[c] = (j0 + c < n_tiles) ? fh[j0 + c] : (uint8_t)0;
    int cnt = 0;  // partials of this chunk that belong to the chain
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (!done) {
        if (f[c]) ++cnt;
        if (f[c] != 2) done = true;
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c)
      if (c < cnt) part[c] = row_ld<D>(head, j0 + c, lane);
#pragma unroll
    for (int c = 0; c < CH; ++c)
      if (c < cnt) row_add<D>(acc, part[c]);
  }
  if (grad_out) {
    row_st<D>(grad_out, key, lane, acc);
    if (touched && lane == 0) touched[key] = 1;
  } else {
    row_update_full<D, LAZY>(P, M, V, L, key, lane, acc, o);
  }
}

#include "train_bpr_fused.cuh"

__global__ void k_loss(const double *__restrict__ part, int64_t n, double inv_b, float *loss_out,
                       double *loss_accum) {
  // single block, fixed order: thread i sums part[i], part[i+blockDim], ... then a tree
  __shared__ double sm[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float l = (float)(sm[0] * inv_b);
    loss_out[0] = l;
    if (loss_accum) loss_accum[0] += (double)l;
  }
}

// forward-only loss (calculate_loss without the step)
template <int D>
__global__ void __launch_bounds__(kThreads) k_bpr_loss(const float *__restrict__ up, const float *__restrict__ ip,
                                                        const int64_t *__restrict__ user,
                                                        const int64_t *__restrict__ pos,
                                                        const int64_t *__restrict__ neg, int64_t B, int64_t n_users,
                                                        int64_t n_items, double *part, WsHeader *hdr) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));
  int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  int64_t ngroups = (int64_t)gridDim.x * blockDim.x / LANES;
  float local = 0.f;
  for (int64_t s = gid; s < B; s += ngroups) {
    int64_t u = user[s], p = pos[s], n = neg[s];
    if ((u < 0) | (u >= n_users) | (p < 0) | (p >= n_items) | (n < 0) | (n >= n_items)) {
      hdr->range_error = 1;
      u = min(max(u, (int64_t)0), n_users - 1);
      p = min(max(p, (int64_t)0), n_items - 1);
      n = min(max(n, (int64_t)0), n_items - 1);
    }
    Row<D> ur = row_ldg<D>(up, u, lane), a = row_ldg<D>(ip, p, lane), b = row_ldg<D>(ip, n, lane);
    float x = group_sum<LANES>(row_dot_lane<D>(ur, a) - row_dot_lane<D>(ur, b), gmask);
    float lt, g;
    bpr_sample(x, 1.f, lt, g);
    local += lt;
  }
  if (lane == 0 && gid < ngroups) part[gid] = (double)local;
}

template <typename K>
int allow_smem(K kernel, size_t bytes) {
  RB2_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <int D, bool ADAM, bool P2P>
int launch_user_fused(const Tables &t, const BprWs &w, int64_t B, int Tu, int64_t ntu, float inv_b, const OptScalars &o,
                      cudaStream_t st) {
  using C = FusedCfg<D, ADAM, P2P>;
  int rc = allow_smem(k_user_fused<D, ADAM, P2P>, C::kSmem);   // per device; a host-side attribute write
  if (rc) return rc;
  k_user_fused<D, ADAM, P2P><<<(unsigned)((ntu + C::GPB - 1) / C::GPB), kThreads, C::kSmem, st>>>(t, w, B, Tu, ntu,
                                                                                                   inv_b, o);
  return 0;
}

template <int D, bool ADAM, bool P2P>
int launch_item_fused(const Tables &t, const BprWs &w, const PeerTable &pt, int64_t n_occ, int Ti, int64_t nti,
                      const OptScalars &o, cudaStream_t st) {
  using C = ItemCfg<D, ADAM, P2P>;
  int rc = allow_smem(k_item_fused<D, ADAM, P2P>, C::kSmem);
  if (rc) return rc;
  k_item_fused<D, ADAM, P2P><<<(unsigned)((nti + C::GPB - 1) / C::GPB), kThreads, C::kSmem, st>>>(t, w, pt, n_occ, Ti,
                                                                                                   nti, o);
  return 0;
}

// RB2_OPT_ADAM_LAZY on the fused path: before anything reads them, the rows this batch touches are brought to the value
// the reference's dense Adam holds after step - 1 (common.cuh row_replay), ONCE per distinct row: one lane group per
// sorted occurrence, the first occurrence of a run does the work.  After that the row-sparse kernels compute exactly
// the dense step for these rows (same m, v, bias corrections), so the row is marked as being at `step`.
template <int D>
__global__ void __launch_bounds__(kThreads) k_lazy_catchup(float *P, float *M, float *V, int32_t *L,
                                                            const uint32_t *__restrict__ keys_s, int64_t n_occ,
                                                            int64_t n_rows, OptScalars o, float sqrt_beta2) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  const int64_t pos = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (pos >= n_occ) return;
  const uint32_t key = keys_s[pos];
  if ((pos > 0 && keys_s[pos - 1] == key) || key >= n_rows) return;
  const int last = L[key];
  if (last < o.step - 1 && (last > 0 || o.wd != 0.f)) {
    Row<D> p = row_ld<D>(P, key, lane), m = row_ld<D>(M, key, lane), v = row_ld<D>(V, key, lane);
    if (o.wd == 0.f)
      row_replay_fast<D>(p, m, v, last, o.step - 1, o, sqrt_beta2);
    else
      row_replay<D>(p, m, v, last, o.step - 1, o);
    row_st<D>(P, key, lane, p);
    row_st<D>(M, key, lane, m);
    row_st<D>(V, key, lane, v);
  }
  if (lane == 0) L[key] = o.step;
}

// the row-sparse Adam / SGD step, single GPU (train_bpr_fused.cuh)
template <int D>
int launch_step_fused(Tables t, BprWs w, int64_t B, int64_t n_users, int64_t n_items, const OptScalars &o,
                      float *loss_out, double *loss_accum, cudaStream_t st, int64_t global_batch) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int Tu = std::min(pick_tile(B, LANES), FusedCfg<D, true, false>::TMAX), Ti = pick_tile(2 * B, LANES);
  const int64_t ntu = (B + Tu - 1) / Tu, nti = (2 * B + Ti - 1) / Ti;
  auto blocks = [](int64_t groups) { return (unsigned)((groups * LANES + kThreads - 1) / kThreads); };
  const PeerTable none{};
  size_t tmp = w.cub_bytes;
  {
    ProfScope prof(RB2_ST_SORT_USER, st, 2 + (bits_for(n_users) + 7) / 8);
    { int rc_ = rb2sort::sort_positions(w.ukey, w.ukey_s, w.uval_s, B, bits_for(n_users), w.cub_tmp, w.cub_bytes, st); if (rc_) return rc_; }
  }
  tmp = w.cub_bytes;
  {
    ProfScope prof(RB2_ST_SORT_ITEM, st, 2 + (bits_for(n_items) + 7) / 8);
    { int rc_ = rb2sort::sort_positions(w.ikey, w.ikey_s, w.ival_s, 2 * B, bits_for(n_items), w.cub_tmp, w.cub_bytes, st); if (rc_) return rc_; }
  }
  {
    ProfScope prof(RB2_ST_PLAN, st, o.kind == RB2_OPT_ADAM_LAZY ? 3 : 1);
    k_mark_local<<<(unsigned)((2 * B + 255) / 256), 256, 0, st>>>(w, 2 * B);
    if (o.kind == RB2_OPT_ADAM_LAZY) {
      const float sb2 = sqrtf(o.beta2);
      k_lazy_catchup<D><<<blocks(B), kThreads, 0, st>>>(t.up, t.um, t.uv, t.ul, w.ukey_s, B, n_users, o, sb2);
      k_lazy_catchup<D><<<blocks(2 * B), kThreads, 0, st>>>(t.ip, t.im, t.iv, t.il, w.ikey_s, 2 * B, n_items, o, sb2);
    }
  }
  {
    ProfScope prof(RB2_ST_USER_SIDE, st);
    int rc = (o.kind == RB2_OPT_SGD) ? launch_user_fused<D, false, false>(t, w, B, Tu, ntu, 1.f / (float)global_batch, o, st)
                                     : launch_user_fused<D, true, false>(t, w, B, Tu, ntu, 1.f / (float)global_batch, o, st);
    if (rc) return rc;
  }
  {
    ProfScope prof(RB2_ST_USER_FIXUP, st, 2);
    k_chain_blocks<D><<<blocks(ntu / kChainBlk + 1), kThreads, 0, st>>>(w.u_head, w.u_fh, w.u_blk, w.u_blk_ok, ntu);
    k_fixup_fused<D, false><<<blocks(ntu), kThreads, 0, st>>>(t.up, t.um, t.uv, w.ukey_s, w.u_head, w.u_tail, w.u_fh,
                                                               w.u_ft, w.u_blk, w.u_blk_ok, B, Tu, ntu, o, none);
  }
  {
    ProfScope prof(RB2_ST_ITEM_SIDE, st);
    int rc = (o.kind == RB2_OPT_SGD) ? launch_item_fused<D, false, false>(t, w, none, 2 * B, Ti, nti, o, st)
                                     : launch_item_fused<D, true, false>(t, w, none, 2 * B, Ti, nti, o, st);
    if (rc) return rc;
  }
  {
    ProfScope prof(RB2_ST_ITEM_FIXUP, st, 2);
    k_chain_blocks<D><<<blocks(nti / kChainBlk + 1), kThreads, 0, st>>>(w.i_head, w.i_fh, w.i_blk, w.i_blk_ok, nti);
    k_fixup_fused<D, false><<<blocks(nti), kThreads, 0, st>>>(t.ip, t.im, t.iv, w.ikey_s, w.i_head, w.i_tail, w.i_fh,
                                                               w.i_ft, w.i_blk, w.i_blk_ok, 2 * B, Ti, nti, o, none);
  }
  {
    ProfScope prof(RB2_ST_LOSS, st);
    k_loss<<<1, 256, 0, st>>>(w.loss_part, ntu, 1.0 / (double)global_batch, loss_out, loss_accum);
  }
  RB2_CUDA(cudaGetLastError());
  return 0;
}

template <int D, bool LAZY>
int launch_step(Tables t, BprWs w, int64_t B, int64_t n_users, int64_t n_items, const OptScalars &o, float *loss_out,
                double *loss_accum, cudaStream_t st, int64_t global_batch, const uint32_t *pre_ikey_s,
                const uint32_t *pre_ival_s, cudaEvent_t rows_ready) {
  constexpr int LANES = RowCfg<D>::LANES;
  // Batches that hit every row many times (BASELINE config 2 at B = 2^20: each user row ~8x, each item row ~78x, tables
  // + state resident in L2) are bound by issue slots, not by memory latency: the register-path kernels below do that
  // shape in 0.61 ms against 0.98 ms for the staged ones (measured), which win wherever rows are mostly distinct.
  const bool dense_batch = B >= n_users && 2 * B >= 4 * n_items;
  // (adam_lazy takes the same kernels behind k_lazy_catchup)
  if (!t.ig && !pre_ikey_s && !rows_ready && !dense_batch)
    return launch_step_fused<D>(t, w, B, n_users, n_items, o, loss_out, loss_accum, st, global_batch);
  const int Tu = pick_tile(B, LANES), Ti = pick_tile(2 * B, LANES);
  const int64_t ntu = (B + Tu - 1) / Tu, nti = (2 * B + Ti - 1) / Ti;
  auto blocks = [](int64_t groups) { return (unsigned)((groups * LANES + kThreads - 1) / kThreads); };

  size_t tmp = w.cub_bytes;
  {
    // cub one-sweep radix sort: histogram + scan + one kernel per 8-bit digit pass
    ProfScope prof(RB2_ST_SORT_USER, st, 2 + (bits_for(n_users) + 7) / 8);
    { int rc_ = rb2sort::sort_positions(w.ukey, w.ukey_s, w.uval_s, B, bits_for(n_users), w.cub_tmp, w.cub_bytes, st); if (rc_) return rc_; }
  }
  tmp = w.cub_bytes;
  if (pre_ikey_s) {
    // rb2_item_plan already sorted the item occurrences (by global id == by compact id)
    w.ikey_s = const_cast<uint32_t *>(pre_ikey_s);
    w.ival_s = const_cast<uint32_t *>(pre_ival_s);
  } else {
    ProfScope prof(RB2_ST_SORT_ITEM, st, 2 + (bits_for(n_items) + 7) / 8);
    { int rc_ = rb2sort::sort_positions(w.ikey, w.ikey_s, w.ival_s, 2 * B, bits_for(n_items), w.cub_tmp, w.cub_bytes, st); if (rc_) return rc_; }
  }
  // sharded step: the keys and sorts above depend on the ids only and overlap with the collective that
  // delivers the item rows on another stream; everything from here on reads them
  if (rows_ready) RB2_CUDA(cudaStreamWaitEvent(st, rows_ready, 0));
  if (LAZY) {   // single-GPU only (the sharded entry points refuse adam_lazy): catch the touched rows up, once per row
    ProfScope prof(RB2_ST_PLAN, st, 2);
    const float sb2 = sqrtf(o.beta2);
    k_lazy_catchup<D><<<blocks(B), kThreads, 0, st>>>(t.up, t.um, t.uv, t.ul, w.ukey_s, B, n_users, o, sb2);
    if (!t.ig)
      k_lazy_catchup<D><<<blocks(2 * B), kThreads, 0, st>>>(t.ip, t.im, t.iv, t.il, w.ikey_s, 2 * B, n_items, o, sb2);
  }
  {
    ProfScope prof(RB2_ST_USER_SIDE, st);
    k_user_side<D, false><<<blocks(ntu), kThreads, 0, st>>>(t, w, B, Tu, ntu, 1.f / (float)global_batch, o);
  }
  {
    ProfScope prof(RB2_ST_USER_FIXUP, st);
    k_fixup<D, false><<<blocks(ntu), kThreads, 0, st>>>(t.up, t.um, t.uv, t.ul, w.ukey_s, w.u_head, w.u_tail, w.u_fh,
                                                        w.u_ft, B, Tu, ntu, o, nullptr, nullptr);
  }
  {
    ProfScope prof(RB2_ST_ITEM_SIDE, st);
    k_item_side<D, false><<<blocks(nti), kThreads, 0, st>>>(t, w, 2 * B, Ti, nti, o);
  }
  {
    ProfScope prof(RB2_ST_ITEM_FIXUP, st);
    k_fixup<D, false><<<blocks(nti), kThreads, 0, st>>>(t.ip, t.im, t.iv, t.il, w.ikey_s, w.i_head, w.i_tail, w.i_fh,
                                                        w.i_ft, 2 * B, Ti, nti, o, t.ig, t.itouched);
  }
  {
    ProfScope prof(RB2_ST_LOSS, st);
    k_loss<<<1, 256, 0, st>>>(w.loss_part, ntu, 1.0 / (double)global_batch, loss_out, loss_accum);
  }
  RB2_CUDA(cudaGetLastError());
  return 0;
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_lazy_flush(float *P, float *M, float *V, int32_t *L, int64_t rows,
                                                          OptScalars o) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (row >= rows) return;
  int last = L[row];
  if (last >= o.step || (last == 0 && o.wd == 0.f)) {
    if (lane == 0 && last < o.step) L[row] = o.step;
    return;
  }
  Row<D> p = row_ld<D>(P, row, lane), m = row_ld<D>(M, row, lane), v = row_ld<D>(V, row, lane);
  row_replay<D>(p, m, v, last, o.step, o);
  row_st<D>(P, row, lane, p);
  row_st<D>(M, row, lane, m);
  row_st<D>(V, row, lane, v);
  if (lane == 0) L[row] = o.step;
}

}  // namespace

#define RB2_DISPATCH_DIM(dim, ...)                                                           \
  switch (dim) {                                                                              \
    case 16: { constexpr int D_ = 16; __VA_ARGS__; } break;                                          \
    case 32: { constexpr int D_ = 32; __VA_ARGS__; } break;                                          \
    case 64: { constexpr int D_ = 64; __VA_ARGS__; } break;                                          \
    case 128: { constexpr int D_ = 128; __VA_ARGS__; } break;                                        \
    case 256: { constexpr int D_ = 256; __VA_ARGS__; } break;                                        \
    default:                                                                                  \
      rb2_set_error("embedding dim %d not supported (16, 32, 64, 128, 256)", (int)(dim));     \
      return RB2_EINVAL;                                                                      \
  }

extern "C" size_t rb2_bpr_workspace_bytes(int64_t batch, int32_t dim) {
  BprWs w;
  return carve(w, nullptr, batch, dim);
}

// ---------------------------------------------------------------------------------------------
// rb2_item_plan: the id-only half of a sharded step.  De-duplicates the item ids of a batch on the
// device (radix sort + head flags + scan), rewrites pos / neg as compact indices, lists the unique
// ids (ascending = grouped by owner shard) and where each owner's range starts.  The sorted item
// occurrences stay in the plan workspace and are reused by rb2_bpr_train_step_sharded (the mapping
// global id -> compact id is monotone, so no second sort).
namespace {
struct PlanWs {
  uint32_t *key, *val, *key_s, *val_s, *ckey_s, *flag, *rank;
  void *cub_tmp;
  size_t cub_bytes;
};
size_t carve_plan(PlanWs &w, void *base, int64_t B, int want_cub) {
  Carver c(base);
  const int64_t M = 2 * B;
  w.key = c.take<uint32_t>(M);
  w.val = c.take<uint32_t>(M);
  w.key_s = c.take<uint32_t>(M);
  w.val_s = c.take<uint32_t>(M);
  w.ckey_s = c.take<uint32_t>(M);
  w.flag = c.take<uint32_t>(M);
  w.rank = c.take<uint32_t>(M);
  size_t b1 = 0, b2 = 0;
  if (want_cub) {
    b1 = rb2sort::tmp_bytes(M);
    cub::DeviceScan::InclusiveSum(nullptr, b2, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)M);
  }
  w.cub_bytes = b1 > b2 ? b1 : b2;
  w.cub_tmp = c.take<char>(w.cub_bytes);
  return c.off;
}
__global__ void k_plan_keys(const int64_t *__restrict__ pos, const int64_t *__restrict__ neg, int64_t B,
                            int64_t n_items, PlanWs w) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  int64_t p = min(max(pos[s], (int64_t)0), n_items - 1), n = min(max(neg[s], (int64_t)0), n_items - 1);
  w.key[2 * s] = (uint32_t)p;
  w.val[2 * s] = (uint32_t)(2 * s);
  w.key[2 * s + 1] = (uint32_t)n;
  w.val[2 * s + 1] = (uint32_t)(2 * s + 1);
}
__global__ void k_plan_flags(PlanWs w, int64_t M) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  w.flag[i] = (i == 0 || w.key_s[i] != w.key_s[i - 1]) ? 1u : 0u;
}
__global__ void k_plan_emit(PlanWs w, int64_t M, int64_t *__restrict__ uniq, int64_t *__restrict__ pos_c,
                            int64_t *__restrict__ neg_c) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  uint32_t c = w.rank[i] - 1u;
  w.ckey_s[i] = c;
  if (w.flag[i]) uniq[c] = (int64_t)w.key_s[i];
  uint32_t o = w.val_s[i];
  if (o & 1u) neg_c[o >> 1] = (int64_t)c; else pos_c[o >> 1] = (int64_t)c;
}
__global__ void k_plan_cuts(PlanWs w, int64_t M, const int64_t *__restrict__ uniq, const int64_t *__restrict__ bounds,
                            int world, int64_t *__restrict__ cuts) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > world + 1) return;
  const int64_t n_uniq = (int64_t)w.rank[M - 1];
  if (g == world + 1) { cuts[g] = n_uniq; return; }
  const int64_t b = bounds[g];
  int64_t lo = 0, hi = n_uniq;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (uniq[mid] < b) lo = mid + 1; else hi = mid;
  }
  cuts[g] = lo;
}
}  // namespace

extern "C" size_t rb2_item_plan_workspace_bytes(int64_t batch) {
  PlanWs w;
  return carve_plan(w, nullptr, batch, 1);
}

extern "C" int rb2_item_plan(const int64_t *pos, const int64_t *neg, int64_t batch, int64_t n_items,
                             const int64_t *shard_bounds, int32_t world, int64_t *uniq, int64_t *pos_c,
                             int64_t *neg_c, int64_t *cuts, void *plan_workspace, size_t plan_workspace_bytes,
                             void *stream) {
  RB2_REQUIRE(pos && neg && shard_bounds && uniq && pos_c && neg_c && cuts && plan_workspace, RB2_EINVAL,
              "rb2_item_plan: null argument");
  RB2_REQUIRE(batch > 0 && batch < ((int64_t)1 << 30) && world >= 1, RB2_EINVAL, "rb2_item_plan: bad sizes");
  PlanWs w;
  size_t need = carve_plan(w, plan_workspace, batch, 1);
  RB2_REQUIRE(plan_workspace_bytes >= need, RB2_EWORKSPACE, "rb2_item_plan: workspace %zu < %zu",
              plan_workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = 2 * batch;
  ProfScope prof(RB2_ST_MISC, st, 6 + (bits_for(n_items) + 7) / 8);
  k_plan_keys<<<(unsigned)((batch + 255) / 256), 256, 0, st>>>(pos, neg, batch, n_items, w);
  size_t tmp = w.cub_bytes;
  { int rc_ = rb2sort::sort_positions(w.key, w.key_s, w.val_s, M, bits_for(n_items), w.cub_tmp, w.cub_bytes, st); if (rc_) return rc_; }
  k_plan_flags<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(w, M);
  tmp = w.cub_bytes;
  RB2_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tmp, w.flag, w.rank, (int)M, st));
  k_plan_emit<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(w, M, uniq, pos_c, neg_c);
  k_plan_cuts<<<1, 64, 0, st>>>(w, M, uniq, shard_bounds, world, cuts);
  RB2_REQUIRE(world + 2 <= 64, RB2_EINVAL, "rb2_item_plan: world too large");
  RB2_CUDA(cudaGetLastError());
  return 0;
}

static int bpr_step_impl(float *user_p, float *user_m, float *user_v, int32_t *user_last, float *item_p,
                         float *item_m, float *item_v, int32_t *item_last, int64_t n_users, int64_t n_items,
                         int32_t dim, const int64_t *user, const int64_t *pos, const int64_t *neg, int64_t batch,
                         const rb2_optim *h_opt, float *loss_out, double *loss_accum, void *workspace,
                         size_t workspace_bytes, void *stream, float *item_grad_out, int64_t global_batch,
                         int32_t *item_touched, const uint32_t *pre_ikey_s = nullptr,
                         const uint32_t *pre_ival_s = nullptr, cudaEvent_t rows_ready = nullptr) {
  RB2_REQUIRE(user_p && item_p && user && pos && neg && h_opt && loss_out && workspace, RB2_EINVAL,
              "rb2_bpr_train_step: null argument");
  RB2_REQUIRE(batch > 0 && batch < ((int64_t)1 << 30), RB2_EINVAL, "rb2_bpr_train_step: batch %lld out of range",
              (long long)batch);
  RB2_REQUIRE(n_users > 0 && n_items > 0 && n_users < ((int64_t)1 << 32) - 1 && n_items < ((int64_t)1 << 31),
              RB2_EINVAL, "rb2_bpr_train_step: table sizes must fit 32 / 31 bits");
  OptScalars o = rb2_opt_scalars(h_opt);
  RB2_REQUIRE(o.kind == RB2_OPT_SGD || o.kind == RB2_OPT_ADAM || o.kind == RB2_OPT_ADAM_LAZY, RB2_EINVAL,
              "rb2_bpr_train_step: unknown optimizer kind %d", o.kind);
  if (o.kind != RB2_OPT_SGD)
    RB2_REQUIRE(user_m && user_v && (item_grad_out || (item_m && item_v)), RB2_EINVAL,
                "rb2_bpr_train_step: Adam needs m and v");
  if (o.kind == RB2_OPT_ADAM_LAZY)
    RB2_REQUIRE(user_last && (item_grad_out || item_last) && o.lazy_step_size && o.lazy_bc2_sqrt, RB2_EINVAL,
                "rb2_bpr_train_step: RB2_OPT_ADAM_LAZY needs *_last and the lazy tables");
  if (global_batch <= 0) global_batch = batch;
  BprWs w;
  size_t need = carve(w, workspace, batch, dim);
  RB2_REQUIRE(workspace_bytes >= need, RB2_EWORKSPACE, "rb2_bpr_train_step: workspace %zu < %zu", workspace_bytes,
              need);
  cudaStream_t st = (cudaStream_t)stream;
  Tables t{user_p, user_m, user_v, user_last, item_p, item_m, item_v, item_last, item_grad_out, item_touched};
  {
    ProfScope prof(RB2_ST_KEYS, st);
    k_make_keys<<<(unsigned)((batch + 255) / 256), 256, 0, st>>>(user, pos, neg, batch, n_users, n_items, w);
  }
  // with item_grad_out the item rows are read as they are (the all-gathered table).  RB2_OPT_ADAM_LAZY there: the user
  // rows are caught up here; the owners keep their item shards current with the dense zero-gradient step of
  // rb2_dense_rows_update, so the gathered rows already are what dense Adam holds.  The planned (sparse, all-to-all)
  // exchange fetches rows nobody caught up: refused.
  const bool lazy = o.kind == RB2_OPT_ADAM_LAZY;
  RB2_REQUIRE(!(lazy && pre_ikey_s), RB2_EINVAL,
              "rb2_bpr_train_step_sharded: adam_lazy is not available with an item plan (sparse exchange); use the "
              "dense or the peer-memory exchange");
  RB2_DISPATCH_DIM(dim, {
    int rc = lazy ? launch_step<D_, true>(t, w, batch, n_users, n_items, o, loss_out, loss_accum, st, global_batch,
                                          pre_ikey_s, pre_ival_s, rows_ready)
                  : launch_step<D_, false>(t, w, batch, n_users, n_items, o, loss_out, loss_accum, st, global_batch,
                                           pre_ikey_s, pre_ival_s, rows_ready);
    if (rc) return rc;
  });
  return 0;
}

extern "C" int rb2_bpr_train_step(float *user_p, float *user_m, float *user_v, int32_t *user_last, float *item_p,
                                  float *item_m, float *item_v, int32_t *item_last, int64_t n_users,
                                  int64_t n_items, int32_t dim, const int64_t *user, const int64_t *pos,
                                  const int64_t *neg, int64_t batch, const rb2_optim *h_opt, float *loss_out,
                                  double *loss_accum, void *workspace, size_t workspace_bytes, void *stream) {
  return bpr_step_impl(user_p, user_m, user_v, user_last, item_p, item_m, item_v, item_last, n_users, n_items, dim,
                       user, pos, neg, batch, h_opt, loss_out, loss_accum, workspace, workspace_bytes, stream,
                       nullptr, 0, nullptr);
}

extern "C" int rb2_bpr_train_step_sharded_ev(float *user_p, float *user_m, float *user_v, int32_t *user_last,
                                             const float *item_rows, int64_t n_users, int64_t n_item_rows, int32_t dim,
                                             const int64_t *user, const int64_t *pos, const int64_t *neg, int64_t batch,
                                             int64_t global_batch, const rb2_optim *h_opt, float *loss_out,
                                             double *loss_accum, float *item_grad_out, int32_t *item_touched,
                                             const void *item_plan, void *workspace, size_t workspace_bytes,
                                             void *stream, void *rows_ready_event) {
  RB2_REQUIRE(item_grad_out != nullptr, RB2_EINVAL, "rb2_bpr_train_step_sharded: item_grad_out is null");
  const uint32_t *pk = nullptr, *pv = nullptr;
  if (item_plan) {  // workspace filled by rb2_item_plan for THIS batch
    PlanWs pw;
    carve_plan(pw, const_cast<void *>(item_plan), batch, 1);
    pk = pw.ckey_s;
    pv = pw.val_s;
  }
  return bpr_step_impl(user_p, user_m, user_v, user_last, const_cast<float *>(item_rows), nullptr, nullptr, nullptr,
                       n_users, n_item_rows, dim, user, pos, neg, batch, h_opt, loss_out, loss_accum, workspace,
                       workspace_bytes, stream, item_grad_out, global_batch, item_touched, pk, pv,
                       (cudaEvent_t)rows_ready_event);
}

extern "C" int rb2_bpr_train_step_sharded(float *user_p, float *user_m, float *user_v, int32_t *user_last,
                                          const float *item_rows, int64_t n_users, int64_t n_item_rows, int32_t dim,
                                          const int64_t *user, const int64_t *pos, const int64_t *neg, int64_t batch,
                                          int64_t global_batch, const rb2_optim *h_opt, float *loss_out,
                                          double *loss_accum, float *item_grad_out, int32_t *item_touched,
                                          const void *item_plan, void *workspace, size_t workspace_bytes,
                                          void *stream) {
  return rb2_bpr_train_step_sharded_ev(user_p, user_m, user_v, user_last, item_rows, n_users, n_item_rows, dim, user, pos,
                                       neg, batch, global_batch, h_opt, loss_out, loss_accum, item_grad_out, item_touched,
                                       item_plan, workspace, workspace_bytes, stream, nullptr);
}

// ---------------------------------------------------------------------------------------------
// rb2_sparse_rows_update: (ids[M], grads[M, d]) with duplicate ids -> sum per row, one optimizer step
// per touched row.  The owner side of the sharded step (gradients arriving from every rank).
namespace {
struct RowsWs {
  WsHeader *hdr;
  uint32_t *key, *val, *key_s, *val_s;
  float *head, *tail;
  uint8_t *fh, *ft;
  void *cub_tmp;
  size_t cub_bytes;
};
size_t carve_rows(RowsWs &w, void *base, int64_t M, int dim) {
  Carver c(base);
  w.hdr = c.take<WsHeader>(1);
  w.key = c.take<uint32_t>(M);
  w.val = c.take<uint32_t>(M);
  w.key_s = c.take<uint32_t>(M);
  w.val_s = c.take<uint32_t>(M);
  int64_t tiles = max_tiles(M);
  w.head = c.take<float>(tiles * dim);
  w.tail = c.take<float>(tiles * dim);
  w.fh = c.take<uint8_t>(tiles);
  w.ft = c.take<uint8_t>(tiles);
  size_t b = 0;
  b = rb2sort::tmp_bytes(M);
  w.cub_bytes = b;
  w.cub_tmp = c.take<char>(b);
  return c.off;
}
__global__ void k_rows_keys(const int64_t *__restrict__ ids, int64_t M, int64_t n_rows, RowsWs w) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  int64_t r = ids[j];
  if (r < 0 || r >= n_rows) {
    w.hdr->range_error = 1;
    r = min(max(r, (int64_t)0), n_rows - 1);
  }
  w.key[j] = (uint32_t)r;
  w.val[j] = (uint32_t)(2 * j);  // item-side encoding: value = 2*source_row + is_neg
}
}  // namespace

namespace {
template <int D, bool DENSE>
__global__ void __launch_bounds__(kThreads) k_dense_rows_update(float *P, float *M, float *V, int64_t rows,
                                                                 const float *__restrict__ grads,
                                                                 const int32_t *__restrict__ touched, OptScalars o) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (row >= rows) return;
  if (touched[row] <= 0) {
    if (DENSE) row_zero_grad_step<D>(P, M, V, row, lane, gmask, o);     // RB2_OPT_ADAM_LAZY: dense Adam moves it too
    return;
  }
  Row<D> p = row_ld<D>(P, row, lane);
  Row<D> g = row_ldg<D>(grads, row, lane);
  row_update<D>(P, M, V, row, lane, p, g, o);
}
}  // namespace

// Owner side of the "replicated small table" exchange: grads [n_rows, d] is the reduce-scattered sum
// of every rank's per-row gradient, touched[row] > 0 marks the rows that occurred in some rank's batch;
// exactly those rows take one optimizer step (row-sparse semantics, as everywhere else).
extern "C" int rb2_dense_rows_update(float *p, float *m, float *v, int64_t n_rows, int32_t dim, const float *grads,
                                     const int32_t *touched, const rb2_optim *h_opt, void *stream) {
  RB2_REQUIRE(p && grads && touched && h_opt, RB2_EINVAL, "rb2_dense_rows_update: null argument");
  OptScalars o = rb2_opt_scalars(h_opt);
  RB2_REQUIRE(o.kind == RB2_OPT_SGD || o.kind == RB2_OPT_ADAM || o.kind == RB2_OPT_ADAM_LAZY, RB2_EINVAL,
              "rb2_dense_rows_update: optimizer kind %d not supported here", o.kind);
  if (o.kind != RB2_OPT_SGD) RB2_REQUIRE(m && v, RB2_EINVAL, "rb2_dense_rows_update: Adam needs m and v");
  if (n_rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(RB2_ST_ITEM_SIDE, st);
  RB2_DISPATCH_DIM(dim, {
    constexpr int LANES = RowCfg<D_>::LANES;
    unsigned blocks = (unsigned)((n_rows * LANES + kThreads - 1) / kThreads);
    if (o.kind == RB2_OPT_ADAM_LAZY) k_dense_rows_update<D_, true><<<blocks, kThreads, 0, st>>>(p, m, v, n_rows, grads, touched, o);
    else k_dense_rows_update<D_, false><<<blocks, kThreads, 0, st>>>(p, m, v, n_rows, grads, touched, o);
  });
  RB2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" size_t rb2_sparse_rows_update_workspace_bytes(int64_t m, int32_t dim) {
  RowsWs w;
  return carve_rows(w, nullptr, m, dim);
}

extern "C" int rb2_sparse_rows_update(float *p, float *m, float *v, int32_t *last, int64_t n_rows, int32_t dim,
                                      const int64_t *ids, const float *grads, int64_t count, const rb2_optim *h_opt,
                                      void *workspace, size_t workspace_bytes, void *stream) {
  RB2_REQUIRE(p && ids && grads && h_opt && workspace, RB2_EINVAL, "rb2_sparse_rows_update: null argument");
  if (count <= 0) return 0;
  RB2_REQUIRE(count < ((int64_t)1 << 30) && n_rows < ((int64_t)1 << 32) - 1, RB2_EINVAL,
              "rb2_sparse_rows_update: sizes out of range");
  OptScalars o = rb2_opt_scalars(h_opt);
  if (o.kind != RB2_OPT_SGD) RB2_REQUIRE(m && v, RB2_EINVAL, "rb2_sparse_rows_update: Adam needs m and v");
  if (o.kind == RB2_OPT_ADAM_LAZY)
    RB2_REQUIRE(last && o.lazy_step_size && o.lazy_bc2_sqrt, RB2_EINVAL, "rb2_sparse_rows_update: lazy state missing");
  RowsWs rw;
  size_t need = carve_rows(rw, workspace, count, dim);
  RB2_REQUIRE(workspace_bytes >= need, RB2_EWORKSPACE, "rb2_sparse_rows_update: workspace %zu < %zu", workspace_bytes,
              need);
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope prof(RB2_ST_KEYS, st);
    k_rows_keys<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(ids, count, n_rows, rw);
  }
  size_t tmp = rw.cub_bytes;
  {
    ProfScope prof(RB2_ST_SORT_ITEM, st, 2 + (bits_for(n_rows) + 7) / 8);
    { int rc_ = rb2sort::sort_positions(rw.key, rw.key_s, rw.val_s, count, bits_for(n_rows), rw.cub_tmp, rw.cub_bytes, st, 1); if (rc_) return rc_; }
  }
  BprWs w{};
  w.hdr = rw.hdr;
  w.ikey_s = rw.key_s;
  w.ival_s = rw.val_s;
  w.gu = const_cast<float *>(grads);
  w.i_head = rw.head;
  w.i_tail = rw.tail;
  w.i_fh = rw.fh;
  w.i_ft = rw.ft;
  Tables t{nullptr, nullptr, nullptr, nullptr, p, m, v, last, nullptr, nullptr};
  const bool lazy = o.kind == RB2_OPT_ADAM_LAZY;
  RB2_DISPATCH_DIM(dim, {
    constexpr int LANES = RowCfg<D_>::LANES;
    const int Ti = pick_tile(count, LANES);
    const int64_t nti = (count + Ti - 1) / Ti;
    unsigned blocks = (unsigned)((nti * LANES + kThreads - 1) / kThreads);
    if (lazy) {
      { ProfScope prof(RB2_ST_ITEM_SIDE, st); k_item_side<D_, true><<<blocks, kThreads, 0, st>>>(t, w, count, Ti, nti, o); }
      { ProfScope prof(RB2_ST_ITEM_FIXUP, st);
        k_fixup<D_, true><<<blocks, kThreads, 0, st>>>(p, m, v, last, rw.key_s, rw.head, rw.tail, rw.fh, rw.ft, count, Ti, nti, o, nullptr, nullptr); }
    } else {
      { ProfScope prof(RB2_ST_ITEM_SIDE, st); k_item_side<D_, false><<<blocks, kThreads, 0, st>>>(t, w, count, Ti, nti, o); }
      { ProfScope prof(RB2_ST_ITEM_FIXUP, st);
        k_fixup<D_, false><<<blocks, kThreads, 0, st>>>(p, m, v, last, rw.key_s, rw.head, rw.tail, rw.fh, rw.ft, count, Ti, nti, o, nullptr, nullptr); }
    }
  });
  RB2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb2_bpr_loss(const float *user_p, const float *item_p, int64_t n_users, int64_t n_items, int32_t dim,
                            const int64_t *user, const int64_t *pos, const int64_t *neg, int64_t batch,
                            float *loss_out, void *workspace, size_t workspace_bytes, void *stream) {
  RB2_REQUIRE(user_p && item_p && user && pos && neg && loss_out && workspace, RB2_EINVAL,
              "rb2_bpr_loss: null argument");
  RB2_REQUIRE(batch > 0, RB2_EINVAL, "rb2_bpr_loss: empty batch");
  BprWs w;
  size_t need = carve(w, workspace, batch, dim);
  RB2_REQUIRE(workspace_bytes >= need, RB2_EWORKSPACE, "rb2_bpr_loss: workspace %zu < %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  RB2_DISPATCH_DIM(dim, {
    constexpr int LANES = RowCfg<D_>::LANES;
    int64_t want = (batch * LANES + kThreads - 1) / kThreads;
    unsigned blocks = (unsigned)std::min<int64_t>(want, kLossParts * LANES / kThreads);
    int64_t ngroups = (int64_t)blocks * kThreads / LANES;  // <= kLossParts
    k_bpr_loss<D_><<<blocks, kThreads, 0, st>>>(user_p, item_p, user, pos, neg, batch, n_users, n_items,
                                                w.loss_part, w.hdr);
    k_loss<<<1, 256, 0, st>>>(w.loss_part, ngroups, 1.0 / (double)batch, loss_out, nullptr);
  });
  RB2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb2_adam_lazy_flush(float *p, float *m, float *v, int32_t *last, int64_t rows, int32_t dim,
                                   const rb2_optim *h_opt, void *stream) {
  RB2_REQUIRE(p && m && v && last && h_opt, RB2_EINVAL, "rb2_adam_lazy_flush: null argument");
  OptScalars o = rb2_opt_scalars(h_opt);
  RB2_REQUIRE(o.lazy_step_size && o.lazy_bc2_sqrt, RB2_EINVAL, "rb2_adam_lazy_flush: lazy tables missing");
  cudaStream_t st = (cudaStream_t)stream;
  RB2_DISPATCH_DIM(dim, {
    constexpr int LANES = RowCfg<D_>::LANES;
    unsigned blocks = (unsigned)((rows * LANES + kThreads - 1) / kThreads);
    k_lazy_flush<D_><<<blocks, kThreads, 0, st>>>(p, m, v, last, rows, o);
  });
  RB2_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Peer-memory step (include/recbole_b200.h (1d)): the item table is row-sharded over the GPUs of one NVLink
// domain; rows are read from and gradients written to the owners' memory by the kernels themselves.
// The id-only work for the NEXT batch (keys, two sorts) runs on a side stream of its own: issued right after this
// rank's barrier-B signal, it overlaps with the wait for the slower peers, the owner update and barrier A instead of
// delaying this rank's owner update -- the rank that reaches barrier B LAST used to hold everybody up by its 0.2 ms of
// sorts (measured on 8 GPUs: 0.2 ms of barrier-A wait on every other rank).  One side stream + two events per
// (device, caller stream), created on first use and kept for the life of the process.
namespace {
struct P2pSide {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, sorted = nullptr;
  bool pending = false;          // `sorted` has been recorded and not yet waited for
};
std::mutex g_side_mu;
std::map<std::pair<int, cudaStream_t>, P2pSide> g_side;

int p2p_side(cudaStream_t st, P2pSide **out) {
  int dev = 0;
  RB2_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_side_mu);
  P2pSide &s = g_side[std::make_pair(dev, st)];
  if (!s.stream) {
    RB2_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    RB2_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    RB2_CUDA(cudaEventCreateWithFlags(&s.sorted, cudaEventDisableTiming));
  }
  *out = &s;
  return 0;
}
}  // namespace

extern "C" size_t rb2_bpr_p2p_workspace_bytes(int64_t batch, int32_t dim) {
  BprWs w;
  return carve(w, nullptr, batch, dim, true);
}

extern "C" int rb2_bpr_train_step_p2p(float *user_p, float *user_m, float *user_v, float *item_m, float *item_v,
                                      int64_t n_users_local, int64_t n_items, int32_t dim, const int64_t *user,
                                      int64_t user_base, const int64_t *pos, const int64_t *neg, int64_t batch,
                                      int64_t global_batch,
                                      const rb2_optim *h_opt, const rb2_peers *h_peers, float *item_cache,
                                      float *loss_out, double *loss_accum, void *workspace, size_t workspace_bytes,
                                      void *stream, int32_t prepared, const int64_t *next_user,
                                      const int64_t *next_pos, const int64_t *next_neg, int32_t *user_last) {
  RB2_REQUIRE(user_p && user && pos && neg && h_opt && h_peers && item_cache && loss_out && workspace, RB2_EINVAL,
              "rb2_bpr_train_step_p2p: null argument");
  RB2_REQUIRE((next_user != nullptr) == (next_pos != nullptr) && (next_user != nullptr) == (next_neg != nullptr),
              RB2_EINVAL, "rb2_bpr_train_step_p2p: next_user / next_pos / next_neg go together");
  RB2_REQUIRE(batch > 0 && batch < ((int64_t)1 << 30) && global_batch >= batch, RB2_EINVAL,
              "rb2_bpr_train_step_p2p: bad batch sizes");
  const rb2_peers &hp = *h_peers;
  RB2_REQUIRE(hp.world >= 1 && hp.world <= RB2_MAX_PEERS && hp.me >= 0 && hp.me < hp.world && hp.item_block > 0 &&
                  hp.item_block * hp.world >= n_items,
              RB2_EINVAL, "rb2_bpr_train_step_p2p: bad peer table (world %d, me %d, block %lld)", hp.world, hp.me,
              (long long)hp.item_block);
  RB2_REQUIRE(n_users_local > 0 && n_items > 0 && n_users_local < ((int64_t)1 << 32) - 1 && n_items < ((int64_t)1 << 31),
              RB2_EINVAL, "rb2_bpr_train_step_p2p: table sizes must fit 32 / 31 bits");
  OptScalars o = rb2_opt_scalars(h_opt);
  RB2_REQUIRE(o.kind == RB2_OPT_SGD || o.kind == RB2_OPT_ADAM || o.kind == RB2_OPT_ADAM_LAZY, RB2_EINVAL,
              "rb2_bpr_train_step_p2p: optimizer kind %d not supported (sgd, adam, adam_lazy)", o.kind);
  // RB2_OPT_ADAM_LAZY = the trajectory of the reference's dense Adam: the (local) user rows of the batch are caught
  // up once per row before the step (k_lazy_catchup, user_last), the owner takes the zero-gradient step of every
  // untouched row of its item shard (k_owner_update<DENSE>)
  const bool lazy = o.kind == RB2_OPT_ADAM_LAZY;
  if (lazy)
    RB2_REQUIRE(user_last && o.lazy_step_size && o.lazy_bc2_sqrt, RB2_EINVAL,
                "rb2_bpr_train_step_p2p: RB2_OPT_ADAM_LAZY needs user_last and the lazy tables");
  if (o.kind != RB2_OPT_SGD)
    RB2_REQUIRE(user_m && user_v && item_m && item_v, RB2_EINVAL, "rb2_bpr_train_step_p2p: Adam needs m and v");
  RB2_REQUIRE(hp.seq >= 1 && hp.seq < ((int64_t)1 << 31), RB2_EINVAL,
              "rb2_bpr_train_step_p2p: h_peers->seq must count 1, 2, 3, ... (it is the barrier sequence)");
  const uint32_t seq = (uint32_t)hp.seq;
  PeerTable pt{};
  PeerSync ps{};
  for (int r = 0; r < hp.world; ++r) {
    RB2_REQUIRE(hp.item_p[r] && hp.grad_slots[r] && hp.stamps[r] && hp.flags[r] && hp.loss_slots[r], RB2_EINVAL,
                "rb2_bpr_train_step_p2p: peer %d has a null buffer", r);
    pt.V[r] = hp.item_p[r];
    pt.G[r] = hp.grad_slots[r];
    pt.stamp[r] = hp.stamps[r];
    ps.flags[r] = hp.flags[r];
    ps.loss[r] = hp.loss_slots[r];
  }
  pt.cache = item_cache;
  pt.i_block = hp.item_block;
  pt.me = ps.me = hp.me;
  pt.world = ps.world = hp.world;
  pt.step = (int32_t)seq;
  BprWs w;
  size_t need = carve(w, workspace, batch, dim, true);
  RB2_REQUIRE(workspace_bytes >= need, RB2_EWORKSPACE, "rb2_bpr_train_step_p2p: workspace %zu < %zu", workspace_bytes,
              need);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t B = batch;
  float *item_local = const_cast<float *>(hp.item_p[hp.me]);
  Tables t{user_p, user_m, user_v, nullptr, item_local, item_m, item_v, nullptr, nullptr, nullptr};
  const unsigned long long timeout_ns = 30ull * 1000ull * 1000ull * 1000ull;
  int64_t n_local = n_items - (int64_t)hp.me * hp.item_block;
  if (n_local > hp.item_block) n_local = hp.item_block;
  if (n_local < 0) n_local = 0;
  // Two batch slots (sorted keys + (pos, neg) pairs): the call with sequence number seq works on slot seq & 1 and,
  // given the NEXT batch's ids, fills the other slot while it waits for its peers in barrier B -- the id-only work
  // (keys, two sorts: ~0.2 ms of a 2.5 ms step on 8 GPUs) then costs nothing.  `prepared` says the previous call did
  // that for THIS batch.
  if (seq & 1u) {
    std::swap(w.ukey_s, w.alt_ukey_s); std::swap(w.uval_s, w.alt_uval_s);
    std::swap(w.ikey_s, w.alt_ikey_s); std::swap(w.ival_s, w.alt_ival_s);
    std::swap(w.pn, w.alt_pn);
  }
  auto keys_and_sorts = [&](BprWs &ws, const int64_t *u_, const int64_t *p_, const int64_t *n_, cudaStream_t ks) -> int {
    {
      ProfScope prof(RB2_ST_KEYS, ks, 1);
      k_make_keys<<<(unsigned)((B + 255) / 256), 256, 0, ks>>>(u_, p_, n_, B, n_users_local, n_items, ws, user_base);
    }
    {
      ProfScope prof(RB2_ST_SORT_USER, ks, 2);
      int rc_ = rb2sort::sort_positions(ws.ukey, ws.ukey_s, ws.uval_s, B, bits_for(n_users_local), ws.cub_tmp, ws.cub_bytes, ks);
      if (rc_) return rc_;
    }
    {
      ProfScope prof(RB2_ST_SORT_ITEM, ks, 2);
      int rc_ = rb2sort::sort_positions(ws.ikey, ws.ikey_s, ws.ival_s, 2 * B, bits_for(n_items), ws.cub_tmp, ws.cub_bytes, ks);
      if (rc_) return rc_;
    }
    return 0;
  };
  P2pSide *side = nullptr;
  { int rc_ = p2p_side(st, &side); if (rc_) return rc_; }
  if (side->pending) {     // the previous call's side-stream sorts (this batch's, if `prepared`): done before anything reads them
    RB2_CUDA(cudaStreamWaitEvent(st, side->sorted, 0));
    side->pending = false;
  }
  RB2_CUDA(cudaMemsetAsync(w.fetch_count, 0, sizeof(uint32_t), st));
  if (!prepared) {
    int rc_ = keys_and_sorts(w, user, pos, neg, st);
    if (rc_) return rc_;
  }
  {
    // barrier A: every owner has finished the previous step's update; from here on peers' rows may be read and
    // their slots / stamps written
    // (the signal half was issued at the end of the previous step, right after the owner update: a rank does not
    // hold the others up with its own keys / sorts; step 1 has no predecessor and does both halves here)
    ProfScope prof(RB2_ST_BARRIER, st);
    k_peer_barrier<<<1, 32, 0, st>>>(ps, seq, 0, seq == 1 ? (kBarSignal | kBarWait) : kBarWait, nullptr,
                                     0.0, nullptr, nullptr, w.hdr, timeout_ns);
  }
  RB2_DISPATCH_DIM(dim, {
    constexpr int LANES = RowCfg<D_>::LANES;
    const int Tu = std::min(pick_tile(B, LANES), FusedCfg<D_, true, true>::TMAX), Ti = pick_tile(2 * B, LANES);
    const int64_t ntu = (B + Tu - 1) / Tu, nti = (2 * B + Ti - 1) / Ti;
    auto blocks = [](int64_t groups) { return (unsigned)((groups * LANES + kThreads - 1) / kThreads); };
    {
      ProfScope prof(RB2_ST_PLAN, st, lazy ? 2 : 1);
      k_plan_p2p<D_><<<(unsigned)((2 * B + 255) / 256), 256, 0, st>>>(w, pt, 2 * B);
      if (lazy)
        k_lazy_catchup<D_><<<blocks(B), kThreads, 0, st>>>(user_p, user_m, user_v, user_last, w.ukey_s, B,
                                                           n_users_local, o, sqrtf(o.beta2));
    }
    {
      ProfScope prof(RB2_ST_FETCH, st, 1);
      k_fetch_rows<D_><<<(unsigned)rb2_num_sms() * 8, kThreads, 0, st>>>(w, pt);
    }
    {
      ProfScope prof(RB2_ST_USER_SIDE, st);
      const float inv_b = 1.f / (float)global_batch;
      int rc = (o.kind == RB2_OPT_SGD) ? launch_user_fused<D_, false, true>(t, w, B, Tu, ntu, inv_b, o, st)
                                       : launch_user_fused<D_, true, true>(t, w, B, Tu, ntu, inv_b, o, st);
      if (rc) return rc;
    }
    {
      ProfScope prof(RB2_ST_USER_FIXUP, st, 2);
      k_chain_blocks<D_><<<blocks(ntu / kChainBlk + 1), kThreads, 0, st>>>(w.u_head, w.u_fh, w.u_blk, w.u_blk_ok, ntu);
      k_fixup_fused<D_, false><<<blocks(ntu), kThreads, 0, st>>>(t.up, t.um, t.uv, w.ukey_s, w.u_head, w.u_tail, w.u_fh,
                                                                  w.u_ft, w.u_blk, w.u_blk_ok, B, Tu, ntu, o, pt);
    }
    {
      ProfScope prof(RB2_ST_ITEM_SIDE, st);
      int rc = launch_item_fused<D_, false, true>(t, w, pt, 2 * B, Ti, nti, o, st);
      if (rc) return rc;
    }
    {
      ProfScope prof(RB2_ST_ITEM_FIXUP, st, 2);
      k_chain_blocks<D_><<<blocks(nti / kChainBlk + 1), kThreads, 0, st>>>(w.i_head, w.i_fh, w.i_blk, w.i_blk_ok, nti);
      k_fixup_fused<D_, true><<<blocks(nti), kThreads, 0, st>>>(nullptr, nullptr, nullptr, w.ikey_s, w.i_head, w.i_tail,
                                                                 w.i_fh, w.i_ft, w.i_blk, w.i_blk_ok, 2 * B, Ti, nti, o, pt);
    }
    {
      ProfScope prof(RB2_ST_LOSS, st);
      k_loss_sum<<<1, 256, 0, st>>>(w.loss_part, ntu, w.loss_sum);
    }
    {
      // barrier B: every rank's gradient rows have landed in the owners' slots; the loss sums travel with it
      ProfScope prof(RB2_ST_BARRIER_B, st, 2);
      // signal half: my gradient rows are in the owners' slots (and my loss sum in their loss slots) ...
      k_peer_barrier<<<1, 32, 0, st>>>(ps, seq, 1, kBarSignal, w.loss_sum, 1.0 / (double)global_batch, loss_out,
                                       loss_accum, w.hdr, timeout_ns);
    }
    if (next_user) {
      // ... the next batch's keys and sorts start now, on the side stream: they overlap with the wait below, the owner
      // update and barrier A, and never delay them (the other batch slot was last read by the previous call's kernels,
      // which precede the fork event on this stream) ...
      BprWs wn = w;
      wn.ukey_s = w.alt_ukey_s; wn.uval_s = w.alt_uval_s; wn.ikey_s = w.alt_ikey_s; wn.ival_s = w.alt_ival_s;
      wn.pn = w.alt_pn;
      RB2_CUDA(cudaEventRecord(side->fork, st));
      RB2_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
      int rc_ = keys_and_sorts(wn, next_user, next_pos, next_neg, side->stream);
      if (rc_) return rc_;
      RB2_CUDA(cudaEventRecord(side->sorted, side->stream));
      side->pending = true;
    }
    {
      // ... wait half (forms the global mean loss)
      ProfScope prof(RB2_ST_BARRIER_B, st, 0);
      k_peer_barrier<<<1, 32, 0, st>>>(ps, seq, 1, kBarWait, w.loss_sum, 1.0 / (double)global_batch, loss_out,
                                       loss_accum, w.hdr, timeout_ns);
    }
    {
      ProfScope prof(RB2_ST_OWNER, st, 2);
      if (n_local > 0) {
        if (lazy) k_owner_update<D_, true><<<blocks(n_local), kThreads, 0, st>>>(item_local, item_m, item_v, pt, n_local, o);
        else k_owner_update<D_, false><<<blocks(n_local), kThreads, 0, st>>>(item_local, item_m, item_v, pt, n_local, o);
      }
      // signal half of the NEXT step's barrier A: my rows are up to date, peers may read them and reuse my slots
      k_peer_barrier<<<1, 32, 0, st>>>(ps, seq + 1u, 0, kBarSignal, nullptr, 0.0, nullptr, nullptr, w.hdr,
                                       timeout_ns);
    }
  });
  RB2_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// The grouping primitive of every step above, exported for tests and callers that build their own walks.
extern "C" size_t rb2_sort_positions_workspace_bytes(int64_t count) { return rb2sort::tmp_bytes(count) + 256; }

extern "C" int rb2_sort_positions(const uint32_t *keys, int64_t count, int32_t key_bits, uint32_t *keys_sorted,
                                  uint32_t *positions, void *workspace, size_t workspace_bytes, void *stream) {
  RB2_REQUIRE(keys && keys_sorted && positions && workspace, RB2_EINVAL, "rb2_sort_positions: null argument");
  RB2_REQUIRE(count >= 0 && count < ((int64_t)1 << 32) && key_bits >= 1 && key_bits <= 32, RB2_EINVAL,
              "rb2_sort_positions: count / key_bits out of range");
  return rb2sort::sort_positions(keys, keys_sorted, positions, count, key_bits, workspace, workspace_bytes,
                                 (cudaStream_t)stream);
}
