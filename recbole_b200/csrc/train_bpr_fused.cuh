// train_bpr_fused.cuh -- the row-sparse Adam / SGD step of train_bpr.cu, second generation (included by
// train_bpr.cu inside its anonymous namespace).
//
// What changed against k_user_side / k_item_side (which stay for RB2_OPT_ADAM_LAZY):
//   * rows are staged through shared memory by the bulk-copy engine (cp.async.bulk + mbarrier, one stage per
//     sample, S stages per lane group): the loads in flight no longer live in registers, a group keeps
//     S samples x up to 9 rows outstanding and never stalls on a dependent global load inside the walk;
//   * an item that occurs exactly ONCE in the batch ("single", about half of the occurrences of a Zipf batch)
//     is finished where its gradient is born: single-GPU step -> the user-side kernel takes its optimizer
//     step in place (no gu round trip, no second read of p); peer-memory step -> the user-side kernel pushes
//     +-g*u straight into the owner's gradient slot over NVLink.  The item-side walk skips singles;
//   * long chains of tile partials (a hot item spans thousands of tiles) are pre-reduced kChainBlk tiles at
//     a time by k_chain_blocks, so the fixed-order chain walk is short;
//   * peer-memory (multi-GPU) form: item rows are read from / gradients are written to the owners' memory
//     directly (cudaIpc-mapped peers over NVLink), two flag barriers per step, no NCCL call, no host sync.

struct PeerTable {
  const float *V[RB2_MAX_PEERS];   // every rank's item shard (p), [i_block, D]
  float *G[RB2_MAX_PEERS];         // every rank's gradient slots [world, i_block, D]: slot s = what rank s sent
  int32_t *stamp[RB2_MAX_PEERS];   // every rank's [world, i_block]: == step when the slot row is valid
  float *cache;                    // local [world * i_block, D]: remote rows that occur more than once
  int64_t i_block;
  int32_t me, world, step;
};

struct PeerSync {
  uint32_t *flags[RB2_MAX_PEERS];  // every rank's [2, RB2_MAX_PEERS] barrier flags (A, B), indexed by sender
  double *loss[RB2_MAX_PEERS];     // every rank's [2, RB2_MAX_PEERS] loss partials (parity, sender)
  int32_t me, world;
};

constexpr uint32_t kSingleBit = 0x80000000u;

// ---- single-GPU: mark single occurrences in pn (bit 31 of the item id) --------------------------------------
__global__ void k_mark_local(BprWs w, int64_t M) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const uint32_t k = w.ikey_s[i];
  const bool single = (i == 0 || w.ikey_s[i - 1] != k) && (i + 1 == M || w.ikey_s[i + 1] != k);
  reinterpret_cast<uint32_t *>(w.pn)[w.ival_s[i]] = k | (single ? kSingleBit : 0u);
}

// ---- peer-memory step: per occurrence, where its row is read from and where its gradient goes ------------------
template <int D>
__global__ void k_plan_p2p(BprWs w, PeerTable pt, int64_t M) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < M;
  bool fetch = false;
  uint32_t k = 0;
  if (valid) {
    k = w.ikey_s[i];
    const uint32_t o = w.ival_s[i];
    const bool head = (i == 0 || w.ikey_s[i - 1] != k), tail = (i + 1 == M || w.ikey_s[i + 1] != k);
    const int owner = (int)(k / pt.i_block);
    const int64_t row = (int64_t)k - (int64_t)owner * pt.i_block;
    unsigned long long src, dst = 0ull;
    if (head && tail) {            // single: read the owner's row directly, push the gradient directly
      const int64_t slot_row = (int64_t)pt.me * pt.i_block + row;
      src = (unsigned long long)(pt.V[owner] + row * D);
      dst = (unsigned long long)(pt.G[owner] + slot_row * D);
      pt.stamp[owner][slot_row] = pt.step;
    } else if (owner == pt.me) {
      src = (unsigned long long)(pt.V[owner] + row * D);
    } else {                       // remote and repeated: fetched once into the local cache by k_fetch_rows
      src = (unsigned long long)(pt.cache + (int64_t)k * D);
      fetch = head;
    }
    w.src[o] = src;
    w.dst[o] = dst;
  }
  const unsigned m = __ballot_sync(0xffffffffu, fetch);
  if (m) {
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(w.fetch_count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (fetch) w.fetch_list[base + __popc(m & ((1u << lane) - 1u))] = k;
  }
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_fetch_rows(BprWs w, PeerTable pt) {
  constexpr int LANES = RowCfg<D>::LANES;
  constexpr int UNR = 8;      // rows in flight per lane group: the loads cross NVLink (microseconds of latency)
  const int lane = threadIdx.x % LANES;
  const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / LANES;
  const int64_t n = (int64_t)*w.fetch_count;
  for (int64_t j0 = gid * UNR; j0 < n; j0 += ngroups * UNR) {
    uint32_t k[UNR];
    Row<D> r[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) k[j] = (j0 + j < n) ? w.fetch_list[j0 + j] : 0xffffffffu;
#pragma unroll
    for (int j = 0; j < UNR; ++j)
      if (k[j] != 0xffffffffu) {
        const int owner = (int)(k[j] / pt.i_block);
        r[j] = row_ld<D>(pt.V[owner], (int64_t)k[j] - (int64_t)owner * pt.i_block, lane);
      }
#pragma unroll
    for (int j = 0; j < UNR; ++j)
      if (k[j] != 0xffffffffu) row_st<D>(pt.cache, k[j], lane, r[j]);
  }
}

// ---- rows through shared memory --------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ Row<D> row_lds(const float *sp, int lane) {
  const float4 *p = reinterpret_cast<const float4 *>(sp);
  Row<D> r;
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) r.v[i] = p[i * RowCfg<D>::LANES + lane];
  return r;
}
template <int D>
__device__ __forceinline__ void row_st_ptr(float *dst, int lane, const Row<D> &r) {
  float4 *p = reinterpret_cast<float4 *>(dst);
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) p[i * RowCfg<D>::LANES + lane] = r.v[i];
}

template <int D, bool ADAM, bool P2P>
struct FusedCfg {
  static constexpr int LANES = RowCfg<D>::LANES;
  static constexpr int GPB = kThreads / LANES;                 // lane groups (tiles) per block
  static constexpr int ROWS = ADAM ? (P2P ? 5 : 9) : 3;        // u a b | mu vu | ma va mb vb
  static constexpr int S = ADAM ? (P2P ? 4 : 2) : 4;           // stages (samples in flight) per group
  static constexpr int REC = P2P ? 40 : 16;                    // bytes of ids per sample
  static constexpr int TMAX = GPB <= 16 ? kTileMax : (GPB == 32 ? 32 : 16);   // tile length cap (ids live in smem)
  static constexpr size_t kStageBytes = (size_t)GPB * S * ROWS * D * sizeof(float);
  static constexpr size_t kIdBytes = (size_t)GPB * TMAX * REC;
  static constexpr size_t kSmem = kStageBytes + kIdBytes + (size_t)GPB * S * sizeof(uint64_t);
};

enum { R_U = 0, R_A = 1, R_B = 2, R_MU = 3, R_VU = 4, R_MA = 5, R_VA = 6, R_MB = 7, R_VB = 8 };

// user side.  One lane group walks a tile of T sorted user occurrences; sample i of the tile lives in stage i % S.
template <int D, bool ADAM, bool P2P>
__global__ void __launch_bounds__(kThreads, 2) k_user_fused(Tables t, BprWs w, int64_t B, int T, int64_t n_tiles,
                                                            float inv_b, OptScalars o) {
  using C = FusedCfg<D, ADAM, P2P>;
  constexpr int LANES = C::LANES, S = C::S, ROWS = C::ROWS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x % LANES, gib = threadIdx.x / LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));
  const int64_t tile = (int64_t)blockIdx.x * C::GPB + gib;
  if (tile >= n_tiles) return;                                   // whole groups leave; nothing below is block-wide
  float *stage = reinterpret_cast<float *>(smem_raw) + (size_t)gib * S * ROWS * D;
  unsigned char *idb = smem_raw + C::kStageBytes + (size_t)gib * C::TMAX * C::REC;
  uint4 *ids = reinterpret_cast<uint4 *>(idb);                   // local: (key, val, pos | single, neg | single)
  uint2 *kv = reinterpret_cast<uint2 *>(idb + (size_t)C::TMAX * 32);    // p2p: (key, val) after the pointers
  ulonglong4 *ptrs = reinterpret_cast<ulonglong4 *>(idb);        // p2p: (src pos, src neg, dst pos, dst neg)
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + C::kStageBytes + C::kIdBytes) + gib * S;

  const int64_t lo = tile * T, hi = min(lo + (int64_t)T, B);
  const int n = (int)(hi - lo);
  const uint32_t kInvalid = 0xffffffffu;
  const uint32_t prev_key = lo > 0 ? w.ukey_s[lo - 1] : kInvalid;
  const uint32_t next_key = hi < B ? w.ukey_s[hi] : kInvalid;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) rb2_mbar_init(&bars[s], 1);
    rb2_mbar_init_fence();
  }
  for (int j = lane; j < n; j += LANES) {
    const uint32_t k = w.ukey_s[lo + j], v = w.uval_s[lo + j];
    if (P2P) {
      kv[j] = make_uint2(k, v);
      ptrs[j] = make_ulonglong4(w.src[2 * (int64_t)v], w.src[2 * (int64_t)v + 1], w.dst[2 * (int64_t)v],
                                w.dst[2 * (int64_t)v + 1]);
    } else {
      const int2 pn = w.pn[v];
      ids[j] = make_uint4(k, v, (uint32_t)pn.x, (uint32_t)pn.y);
    }
  }
  __syncwarp(gmask);

  auto key_of = [&](int i) -> uint32_t { return P2P ? kv[i].x : ids[i].x; };

  auto issue = [&](int i) {
    float *sp = stage + (size_t)(i % S) * ROWS * D;
    uint64_t *bar = &bars[i % S];
    const uint32_t key = key_of(i);
    const uint32_t pk = (i == 0) ? prev_key : key_of(i - 1);
    const bool need_u = (i == 0) || key != pk;
    const bool need_mv = ADAM && key != pk;
    uint32_t pos = 0, neg = 0;
    unsigned long long sa = 0, sb = 0;
    if (P2P) {
      sa = ptrs[i].x;
      sb = ptrs[i].y;
    } else {
      pos = ids[i].z;
      neg = ids[i].w;
    }
    const bool ps = !P2P && ADAM && (pos & kSingleBit), ns = !P2P && ADAM && (neg & kSingleBit);
    const int rows = 2 + (need_u ? 1 : 0) + (need_mv ? 2 : 0) + (ps ? 2 : 0) + (ns ? 2 : 0);
    if (lane == 0) rb2_mbar_expect_tx(bar, (uint32_t)(rows * D * sizeof(float)));
    const int64_t prow = (int64_t)(pos & ~kSingleBit) * D, nrow = (int64_t)(neg & ~kSingleBit) * D;
    for (int r = lane; r < ROWS; r += LANES) {
      const float *src = nullptr;
      switch (r) {
        case R_U: if (need_u) src = t.up + (int64_t)key * D; break;
        case R_A: src = P2P ? reinterpret_cast<const float *>(sa) : t.ip + prow; break;
        case R_B: src = P2P ? reinterpret_cast<const float *>(sb) : t.ip + nrow; break;
        case R_MU: if (need_mv) src = t.um + (int64_t)key * D; break;
        case R_VU: if (need_mv) src = t.uv + (int64_t)key * D; break;
        case R_MA: if (ps) src = t.im + prow; break;
        case R_VA: if (ps) src = t.iv + prow; break;
        case R_MB: if (ns) src = t.im + nrow; break;
        case R_VB: if (ns) src = t.iv + nrow; break;
      }
      if (src) rb2_bulk_g2s(sp + r * D, src, (uint32_t)(D * sizeof(float)), bar);
    }
  };

  for (int i = 0; i < S && i < n; ++i) issue(i);

  uint32_t cur = kInvalid;
  bool started_before = false;
  Row<D> u = row_zero<D>(), mu = row_zero<D>(), vu = row_zero<D>(), acc = row_zero<D>();
  float loss_local = 0.f;
  uint8_t fh = 0, ft = 0;

  auto finish_run = [&](bool continues) {
    if (cur == kInvalid) return;
    if (!started_before && !continues) {
      row_step_regs<D>(u, mu, vu, acc, o);          // u is the row's value on entry, (mu, vu) came with it
      row_st<D>(t.up, cur, lane, u);
      if (ADAM) {
        row_st<D>(t.um, cur, lane, mu);
        row_st<D>(t.uv, cur, lane, vu);
      }
    } else if (started_before) {
      row_st<D>(w.u_head, tile, lane, acc);
      fh = continues ? 2 : 1;
    } else {
      row_st<D>(w.u_tail, tile, lane, acc);
      ft = 1;
    }
  };

  for (int i = 0; i < n; ++i) {
    const float *sp = stage + (size_t)(i % S) * ROWS * D;
    rb2_mbar_wait(&bars[i % S], (uint32_t)((i / S) & 1));
    const uint32_t key = key_of(i);
    const uint32_t val = P2P ? kv[i].y : ids[i].y;
    if (key != cur) {
      finish_run(false);
      cur = key;
      u = row_lds<D>(sp + R_U * D, lane);
      started_before = (i == 0) && (cur == prev_key);
      if (ADAM && !started_before) {
        mu = row_lds<D>(sp + R_MU * D, lane);
        vu = row_lds<D>(sp + R_VU * D, lane);
      }
      acc = row_zero<D>();
    }
    Row<D> a = row_lds<D>(sp + R_A * D, lane), b = row_lds<D>(sp + R_B * D, lane);
    // x = <u, vi> - <u, vj>   (bpr.py:81); the two per-lane partial dots share one shuffle reduction
    const float x = group_sum<LANES>(row_dot_lane<D>(u, a) - row_dot_lane<D>(u, b), gmask);
    float lt, g;
    bpr_sample(x, inv_b, lt, g);
    loss_local += lt;
    row_fma_diff<D>(acc, g, a, b);      // du += g*(vi - vj)
    const Row<D> gu = row_scale<D>(g, u);          // dvi = g*u ; dvj = -g*u
    if (P2P) {
      const unsigned long long da = ptrs[i].z, db = ptrs[i].w;
      if (da) row_st_ptr<D>(reinterpret_cast<float *>(da), lane, gu);
      if (db) row_st_ptr<D>(reinterpret_cast<float *>(db), lane, row_scale<D>(-1.f, gu));
      if (!(da && db)) row_st<D>(w.gu, val, lane, gu);
    } else {
      const uint32_t pos = ids[i].z, neg = ids[i].w;
      const bool ps = pos & kSingleBit, ns = neg & kSingleBit;
      if (ps) {                 // the only occurrence of this item in the batch: its whole step, here
        Row<D> m = row_zero<D>(), v = row_zero<D>();
        if (ADAM) {
          m = row_lds<D>(sp + R_MA * D, lane);
          v = row_lds<D>(sp + R_VA * D, lane);
        }
        row_step_regs<D>(a, m, v, gu, o);
        const int64_t r = pos & ~kSingleBit;
        row_st<D>(t.ip, r, lane, a);
        if (ADAM) {
          row_st<D>(t.im, r, lane, m);
          row_st<D>(t.iv, r, lane, v);
        }
      }
      if (ns) {
        Row<D> m = row_zero<D>(), v = row_zero<D>();
        if (ADAM) {
          m = row_lds<D>(sp + R_MB * D, lane);
          v = row_lds<D>(sp + R_VB * D, lane);
        }
        row_step_regs<D>(b, m, v, row_scale<D>(-1.f, gu), o);
        const int64_t r = neg & ~kSingleBit;
        row_st<D>(t.ip, r, lane, b);
        if (ADAM) {
          row_st<D>(t.im, r, lane, m);
          row_st<D>(t.iv, r, lane, v);
        }
      }
      if (!(ps && ns)) row_st<D>(w.gu, val, lane, gu);
    }
    __syncwarp(gmask);          // every lane has read stage i % S: it may be refilled
    if (i + S < n) issue(i + S);
  }
  finish_run(next_key == cur);
  if (lane == 0) {
    w.u_fh[tile] = fh;
    w.u_ft[tile] = ft;
    w.loss_part[tile] = (double)loss_local;
  }
}

// item side: the walk of k_item_side over the occurrences that are NOT single, with the same staging as the user
// side: the gradient row gu[s] of every occurrence, and (single GPU) the row's (p, m, v) at the start of a run,
// arrive through the bulk-copy engine; the loop body exists once (one optimizer-step site: small code).
// A finished run is stepped in place (single GPU) or written into the owner's gradient slot (peer-memory step).
template <int D, bool ADAM, bool P2P>
struct ItemCfg {
  static constexpr int LANES = RowCfg<D>::LANES;
  static constexpr int GPB = kThreads / LANES;
  static constexpr int ROWS = P2P ? 1 : (ADAM ? 4 : 2);        // gu | p | m v
  static constexpr int S = 4;
  static constexpr size_t kStageBytes = (size_t)GPB * S * ROWS * D * sizeof(float);
  static constexpr size_t kIdBytes = (size_t)GPB * (kTileMax + 2) * sizeof(uint2);
  static constexpr size_t kListBytes = (size_t)GPB * kTileMax;
  static constexpr size_t kSmem = kStageBytes + kIdBytes + kListBytes + (size_t)GPB * S * sizeof(uint64_t);
};
enum { I_G = 0, I_P = 1, I_M = 2, I_V = 3 };

template <int D, bool ADAM, bool P2P>
__global__ void __launch_bounds__(kThreads, 2) k_item_fused(Tables t, BprWs w, PeerTable pt, int64_t n_occ, int T,
                                                             int64_t n_tiles, OptScalars o) {
  using C = ItemCfg<D, ADAM, P2P>;
  constexpr int LANES = C::LANES, S = C::S, ROWS = C::ROWS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x % LANES, gib = threadIdx.x / LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));
  const int64_t tile = (int64_t)blockIdx.x * C::GPB + gib;
  if (tile >= n_tiles) return;
  float *stage = reinterpret_cast<float *>(smem_raw) + (size_t)gib * S * ROWS * D;
  // (key, value) of the tile's occurrences, of the one before it [0] and of the one after it [n + 1]
  uint2 *kv = reinterpret_cast<uint2 *>(smem_raw + C::kStageBytes) + (size_t)gib * (kTileMax + 2);
  uint8_t *list = smem_raw + C::kStageBytes + C::kIdBytes + (size_t)gib * kTileMax;   // tile positions that are not single
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + C::kStageBytes + C::kIdBytes + C::kListBytes) + gib * S;

  const int64_t lo = tile * T, hi = min(lo + (int64_t)T, n_occ);
  const int n = (int)(hi - lo);
  const uint32_t kInvalid = 0xffffffffu;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) rb2_mbar_init(&bars[s], 1);
    rb2_mbar_init_fence();
  }
  for (int j = lane; j < n + 2; j += LANES) {
    const int64_t p = lo - 1 + j;
    kv[j] = (p >= 0 && p < n_occ) ? make_uint2(w.ikey_s[p], w.ival_s[p]) : make_uint2(kInvalid, 0u);
  }
  __syncwarp(gmask);
  // compact the non-single positions (order kept): every lane flags its positions, a ballot ranks them
  int m = 0;
  for (int j0 = 0; j0 < n; j0 += LANES) {
    const int j = j0 + lane;
    const bool keep = j < n && !(kv[j + 1].x != kv[j].x && kv[j + 1].x != kv[j + 2].x);
    const unsigned bal = __ballot_sync(gmask, keep) & gmask;
    if (keep) list[m + __popc(bal & ((1u << (threadIdx.x % 32)) - 1u))] = (uint8_t)j;
    m += __popc(bal);
  }
  __syncwarp(gmask);
  const uint32_t prev_key = kv[0].x, next_key = kv[n + 1].x;

  auto issue = [&](int q) {
    const int j = list[q];
    const uint2 e = kv[j + 1];
    float *sp = stage + (size_t)(q % S) * ROWS * D;
    uint64_t *bar = &bars[q % S];
    // the row itself is wanted where a run starts INSIDE this tile (a run that began earlier only adds a partial)
    const bool need_p = !P2P && e.x != kv[j].x;
    const int rows = 1 + (need_p ? ROWS - 1 : 0);
    if (lane == 0) rb2_mbar_expect_tx(bar, (uint32_t)(rows * D * sizeof(float)));
    for (int r = lane; r < ROWS; r += LANES) {
      const float *src = nullptr;
      switch (r) {
        case I_G: src = w.gu + (int64_t)(e.y >> 1) * D; break;
        case I_P: if (need_p) src = t.ip + (int64_t)e.x * D; break;
        case I_M: if (need_p) src = t.im + (int64_t)e.x * D; break;
        case I_V: if (need_p) src = t.iv + (int64_t)e.x * D; break;
      }
      if (src) rb2_bulk_g2s(sp + r * D, src, (uint32_t)(D * sizeof(float)), bar);
    }
  };
  for (int q = 0; q < S && q < m; ++q) issue(q);

  uint32_t cur = kInvalid;
  bool started_before = false;
  Row<D> acc = row_zero<D>(), p = row_zero<D>(), mm = row_zero<D>(), vv = row_zero<D>();
  uint8_t fh = 0, ft = 0;

  auto finish_run = [&](bool continues) {
    if (cur == kInvalid) return;
    if (!started_before && !continues) {
      if (P2P) {
        const int owner = (int)(cur / pt.i_block);
        const int64_t slot_row = (int64_t)pt.me * pt.i_block + ((int64_t)cur - (int64_t)owner * pt.i_block);
        row_st_ptr<D>(pt.G[owner] + slot_row * D, lane, acc);
        if (lane == 0) pt.stamp[owner][slot_row] = pt.step;
      } else {
        row_step_regs<D>(p, mm, vv, acc, o);
        row_st<D>(t.ip, cur, lane, p);
        if (ADAM) {
          row_st<D>(t.im, cur, lane, mm);
          row_st<D>(t.iv, cur, lane, vv);
        }
      }
    } else if (started_before) {
      row_st<D>(w.i_head, tile, lane, acc);
      fh = continues ? 2 : 1;
    } else {
      row_st<D>(w.i_tail, tile, lane, acc);
      ft = 1;
    }
  };

#pragma unroll 1
  for (int q = 0; q < m; ++q) {
    const float *sp = stage + (size_t)(q % S) * ROWS * D;
    rb2_mbar_wait(&bars[q % S], (uint32_t)((q / S) & 1));
    const int j = list[q];
    const uint2 e = kv[j + 1];
    if (e.x != cur) {
      finish_run(false);
      cur = e.x;
      started_before = (cur == kv[j].x);          // only possible at j == 0: the run began in an earlier tile
      if (!P2P && !started_before) {
        p = row_lds<D>(sp + I_P * D, lane);
        if (ADAM) {
          mm = row_lds<D>(sp + I_M * D, lane);
          vv = row_lds<D>(sp + I_V * D, lane);
        }
      }
      acc = row_zero<D>();
    }
    const Row<D> g = row_lds<D>(sp + I_G * D, lane);
    row_fma<D>(acc, (e.y & 1u) ? -1.f : 1.f, g);
    __syncwarp(gmask);
    if (q + S < m) issue(q + S);
  }
  finish_run(next_key == cur);
  if (lane == 0) {
    w.i_fh[tile] = fh;
    w.i_ft[tile] = ft;
  }
}

// every aligned block of kChainBlk tiles that lies entirely inside one run (all of its tiles hold a head partial
// that continues) is summed by one lane group, in tile order
template <int D>
__global__ void __launch_bounds__(kThreads) k_chain_blocks(const float *__restrict__ head, const uint8_t *__restrict__ fh,
                                                            float *__restrict__ blk, uint8_t *__restrict__ blk_ok,
                                                            int64_t n_tiles) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int64_t t0 = b * kChainBlk;
  if (t0 >= n_tiles) return;
  bool ok = t0 + kChainBlk <= n_tiles;
  if (ok) {
    const uint4 f = *reinterpret_cast<const uint4 *>(fh + t0);     // kChainBlk == 16 flag bytes
    ok = f.x == 0x02020202u && f.y == 0x02020202u && f.z == 0x02020202u && f.w == 0x02020202u;
  }
  if (ok) {
    Row<D> acc = row_zero<D>();
#pragma unroll
    for (int h = 0; h < kChainBlk; h += 8) {        // 8 loads in flight, tile order
      Row<D> part[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) part[j] = row_ld<D>(head, t0 + h + j, lane);
#pragma unroll
      for (int j = 0; j < 8; ++j) row_add<D>(acc, part[j]);
    }
    row_st<D>(blk, b, lane, acc);
  }
  if (lane == 0) blk_ok[b] = ok ? 1 : 0;
}

// runs that straddle tiles (see k_fixup); whole pre-reduced blocks are taken in one load
template <int D, bool P2P>
__global__ void __launch_bounds__(kThreads) k_fixup_fused(float *P, float *M, float *V,
                                                           const uint32_t *__restrict__ keys_sorted,
                                                           const float *__restrict__ head, const float *__restrict__ tail,
                                                           const uint8_t *__restrict__ fh, const uint8_t *__restrict__ ft,
                                                           const float *__restrict__ blk, const uint8_t *__restrict__ blk_ok,
                                                           int64_t n_occ, int T, int64_t n_tiles, OptScalars o,
                                                           PeerTable pt) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (tile >= n_tiles || !ft[tile]) return;
  const int64_t last_pos = min((tile + 1) * (int64_t)T, n_occ) - 1;
  const uint32_t key = keys_sorted[last_pos];
  Row<D> acc = row_ld<D>(tail, tile, lane);
  constexpr int CH = 8;
  bool done = false;
  int64_t j0 = tile + 1;
  while (j0 < n_tiles && !done) {
    if ((j0 % kChainBlk) == 0 && blk_ok[j0 / kChainBlk]) {          // kChainBlk interior tiles at once
      Row<D> bsum = row_ld<D>(blk, j0 / kChainBlk, lane);
      row_add<D>(acc, bsum);
      j0 += kChainBlk;                                               // the block's last tile continues (flag 2)
      continue;
    }
    // up to CH single tiles, never crossing into a pre-reduced block
    uint8_t f[CH];
    Row<D> part[CH];
    int lim = CH;
    const int64_t to_blk = kChainBlk - (j0 % kChainBlk);             // tiles until the next block boundary
    if (to_blk < lim) lim = (int)to_blk;
#pragma unroll
    for (int c = 0; c < CH; ++c) f[c] = (c < lim && j0 + c < n_tiles) ? fh[j0 + c] : (uint8_t)0;
    int cnt = 0;
    bool stop = false;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (c < lim && !stop) {
        if (f[c]) ++cnt;
        if (f[c] != 2) stop = true;
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c)
      if (c < cnt) part[c] = row_ld<D>(head, j0 + c, lane);
#pragma unroll
    for (int c = 0; c < CH; ++c)
      if (c < cnt) row_add<D>(acc, part[c]);
    done = stop;
    j0 += lim;
  }
  if (P2P) {
    const int owner = (int)(key / pt.i_block);
    const int64_t slot_row = (int64_t)pt.me * pt.i_block + ((int64_t)key - (int64_t)owner * pt.i_block);
    row_st_ptr<D>(pt.G[owner] + slot_row * D, lane, acc);
    if (lane == 0) pt.stamp[owner][slot_row] = pt.step;
  } else {
    row_update_full<D, false>(P, M, V, nullptr, key, lane, acc, o);
  }
}

// ---- peer-memory step: barrier, owner update, loss -----------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Flag barrier over the peers' memory: rank `me` writes `seq` into flag [which][me] of every rank (mode & 1), then
// waits until every rank has written `seq` (or more) into its own (mode & 2).  Everything the calling stream wrote
// to peer memory before the signalling launch is visible to the peers after their wait.  The two halves may be
// separate launches: a rank signals "my owner update is done" at the END of a step and waits for the others only
// after the id-only work of the next one.  With my_loss != NULL the ranks also exchange their loss sums and every
// rank forms the same global mean (fixed rank order).
enum { kBarSignal = 1, kBarWait = 2 };
__global__ void k_peer_barrier(PeerSync ps, uint32_t seq, int which, int mode, const double *my_loss, double inv_b,
                               float *loss_out, double *loss_accum, WsHeader *hdr, unsigned long long timeout_ns) {
  const int r = threadIdx.x;
  const int parity = (int)(seq & 1u);
  if (r < ps.world) {
    if (mode & kBarSignal) {
      if (my_loss) ps.loss[r][parity * RB2_MAX_PEERS + ps.me] = *my_loss;
      __threadfence_system();
      st_release_sys(ps.flags[r] + which * RB2_MAX_PEERS + ps.me, seq);
    }
    if (mode & kBarWait) {
      const uint32_t *mine = ps.flags[ps.me] + which * RB2_MAX_PEERS + r;
      const unsigned long long t0 = globaltimer_ns();
      while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) {
        __nanosleep(64);
        if (globaltimer_ns() - t0 > timeout_ns) {
          hdr->peer_timeout = 1;
          break;
        }
      }
    }
  }
  __syncwarp();
  if (r == 0 && my_loss && (mode & kBarWait)) {
    __threadfence_system();
    double s = 0.0;
    for (int q = 0; q < ps.world; ++q) s += ps.loss[ps.me][parity * RB2_MAX_PEERS + q];
    const float l = (float)(s * inv_b);
    loss_out[0] = l;
    if (loss_accum) loss_accum[0] += (double)l;
  }
}

__global__ void k_loss_sum(const double *__restrict__ part, int64_t n, double *out) {
  __shared__ double sm[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sm[0];
}

// owner side: a local row takes one step with the sum (fixed sender order) of the slots stamped with this step.
// DENSE (RB2_OPT_ADAM_LAZY on the sharded paths): the reference's dense Adam also moves the rows nobody touched (their
// exp_avg keeps decaying into the parameter); an owner's shard is small enough (n_items / world rows) to simply take
// that zero-gradient step for every row with non-zero moments, so item rows are always current and need no `last`.
template <int D>
__device__ __forceinline__ void row_zero_grad_step(float *P, float *M, float *V, int64_t row, int lane, unsigned gmask,
                                                   const OptScalars &o) {
  Row<D> m = row_ld<D>(M, row, lane), v = row_ld<D>(V, row, lane);
  bool nz = o.wd != 0.f;
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i)
    nz |= (m.v[i].x != 0.f) | (m.v[i].y != 0.f) | (m.v[i].z != 0.f) | (m.v[i].w != 0.f) | (v.v[i].x != 0.f) |
          (v.v[i].y != 0.f) | (v.v[i].z != 0.f) | (v.v[i].w != 0.f);
  if (!(__ballot_sync(gmask, nz) & gmask)) return;       // never stepped: zero moments, the row does not move
  Row<D> p = row_ld<D>(P, row, lane);
  const Row<D> g = row_zero<D>();
  row_step_regs<D>(p, m, v, g, o);
  row_st<D>(P, row, lane, p);
  row_st<D>(M, row, lane, m);
  row_st<D>(V, row, lane, v);
}

template <int D, bool DENSE>
__global__ void __launch_bounds__(kThreads) k_owner_update(float *P, float *M, float *V, PeerTable pt, int64_t n_local,
                                                            OptScalars o) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (row >= n_local) return;
  const int32_t *__restrict__ stp = pt.stamp[pt.me];
  const float *__restrict__ G = pt.G[pt.me];
  bool on[RB2_MAX_PEERS];
  Row<D> part[RB2_MAX_PEERS];
  bool any = false;
#pragma unroll
  for (int s = 0; s < RB2_MAX_PEERS; ++s) {
    on[s] = s < pt.world && stp[(int64_t)s * pt.i_block + row] == pt.step;
    any |= on[s];
  }
  if (!any) {
    if (DENSE) row_zero_grad_step<D>(P, M, V, row, lane, gmask, o);
    return;
  }
#pragma unroll
  for (int s = 0; s < RB2_MAX_PEERS; ++s)
    if (on[s]) part[s] = row_ld<D>(G, (int64_t)s * pt.i_block + row, lane);
  Row<D> acc = row_zero<D>();
#pragma unroll
  for (int s = 0; s < RB2_MAX_PEERS; ++s)
    if (on[s]) row_add<D>(acc, part[s]);
  row_update_full<D, false>(P, M, V, nullptr, row, lane, acc, o);
}
