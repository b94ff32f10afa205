// train_fm.cu -- fused FM (factorization machine, token fields) training step for sm_100a:
// multi-field embedding bag forward, BCE loss, backward as a de-duplicated row-sparse update.
//
// Replaces, per batch: ContextRecommender.embed_input_fields + FMEmbedding (id + per-field offset
// into ONE table, recbole/model/abstract_recommender.py:220-224,361-412, recbole/model/layers.py:141-144),
// BaseFactorizationMachine 0.5*sum_k[(sum_f v)^2 - sum_f v^2] (layers.py:164-171), FMFirstOrderLinear
// (d = 1 table + bias, layers.py:1021-1061), FM.forward/calculate_loss: sigmoid + nn.BCELoss
// (recbole/model/context_aware_recommender/fm.py:47-56), loss.backward() and the dense Adam step
// (trainer.py:170-173).
//
//   k_fm_forward  one warp per sample; its lane groups (d/4 lanes each) take the fields round-robin:
//                 gather v_f (second-order row) and w_f (first-order scalar), S = sum_f v_f,
//                 z = sum_f w_f + b + 0.5*sum_k(S_k^2 - sum_f v_fk^2), y = sigmoid(z), BCE term,
//                 gz = dLoss/dz.  Stores gs[s] = gz*S (d floats) and gz[s]: the only intermediates.
//   radix sort    B*F (row, occurrence) pairs by row id.
//   k_fm_rows     walks tiles of sorted occurrences; for the run of one row r:
//                     dE_r = sum_s gs[s] - (sum_s gz[s]) * v_r        (d z/d v_f = S - v_f)
//                     dW_r = sum_s gz[s]
//                 complete runs are stepped in place (p, m, v of the row and of its first-order scalar),
//                 straddling runs go through the fixed-order fix-up (no float atomics).
//   k_fm_bias     db = sum_s gz[s], dense Adam on the scalar bias; loss reduction.
// Algorithmic bytes per sample: F*(24*d + 24) + 8*F + 4.

#include "bucket_sort.cuh"
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct FmWs {
  WsHeader *hdr;
  uint32_t *key, *val, *key_s, *val_s;   // [B*F]
  float *gs;                             // [B, D]
  float *gz;                             // [B]
  float *head, *tail;                    // [tiles, D]
  float *head_z, *tail_z;                // [tiles]
  uint8_t *fh, *ft;
  float *blk_head, *blk_z;               // [tiles / 64, D], [tiles / 64]: sums of 64 interior head partials
  uint8_t *blk_ok;                       // [tiles / 64]
  double *loss_part, *gz_part;           // [warps blocks]
  float *float_part;                     // [kMaxFloat][chunks][D + 2] partial sums of the FLOAT fields' gradients
  void *cub_tmp;
  size_t cub_bytes;
  int64_t n_parts;
};

size_t carve(FmWs &w, void *base, int64_t B, int F, int dim) {
  Carver c(base);
  const int64_t M = B * F;
  w.hdr = c.take<WsHeader>(1);
  w.key = c.take<uint32_t>(M);
  w.val = c.take<uint32_t>(M);
  w.key_s = c.take<uint32_t>(M);
  w.val_s = c.take<uint32_t>(M);
  w.gs = c.take<float>(B * dim);
  w.gz = c.take<float>(B);
  int64_t tiles = rb2_max_tiles(M);
  w.head = c.take<float>(tiles * dim);
  w.tail = c.take<float>(tiles * dim);
  w.head_z = c.take<float>(tiles);
  w.tail_z = c.take<float>(tiles);
  w.fh = c.take<uint8_t>(tiles);
  w.ft = c.take<uint8_t>(tiles);
  const int64_t blks = tiles / 64 + 1;
  w.blk_head = c.take<float>(blks * dim);
  w.blk_z = c.take<float>(blks);
  w.blk_ok = c.take<uint8_t>(blks);
  w.n_parts = (B + 7) / 8 + 64;
  w.loss_part = c.take<double>(w.n_parts);
  w.gz_part = c.take<double>(w.n_parts);
  w.float_part = c.take<float>((size_t)RB2_FM_MAX_FLOAT * ((B + 4095) / 4096) * (dim + 2));
  size_t b = 0;
  b = rb2sort::tmp_bytes(M);
  w.cub_bytes = b;
  w.cub_tmp = c.take<char>(b);
  return c.off;
}

// internal optimizer kind of rb2_fm_grad_step: the row's summed gradient REPLACES the row (the tables are the
// fetched copies of a sharded step; every row is processed exactly once, after the forward has read it)
constexpr int kOptGradOut = 99;
constexpr int kOptLossOnly = 98;   // rb2_fm_loss: the reduction writes the mean loss and nothing else

struct FmTables {
  float *E, *mE, *vE;   // [rows, D]
  float *W, *mW, *vW;   // [rows]
  float *bias;          // [3]: b, m, v
  int32_t *last;        // [rows] RB2_OPT_ADAM_LAZY: the step at which row r (of E and of W) was last brought up to date
  // FLOAT fields (abstract_recommender.py:236-258, layers.py:947-966): field f owns ONE row Ef[f] (and one scalar
  // Wf[f]) that every sample scales by its value x[s, f]
  const float *fx;      // [B, n_float] values, or nullptr
  int n_float;
  float *Ef, *mEf, *vEf;   // [n_float, D]
  float *Wf, *mWf, *vWf;   // [n_float]
  // TOKEN_SEQ fields (abstract_recommender.py:277-314, mean pooling): the id columns [n_tok, n_cols) of a sample are the
  // padded sequences of n_seq fields (field j: columns [seq_start[j], seq_start[j + 1])); id 0 = padding = masked.
  // Field j's vector is e_j = c_j * sum over its unmasked ids of the row, c_j = 1 / (count + 1e-8).  Its tables are
  // rows >= seq_row_base of E / W (ids + offsets like any token column).
  int n_tok, n_seq;
  const int32_t *seq_start;   // [n_seq + 1] column ranges (device)
  const int32_t *col_seq;     // [n_cols] seq field of a column, -1 for token columns (device)
  int64_t seq_row_base, n_rows;
  float *ef;                  // [B, n_seq, D] the pooled vectors (kept for the backward)
  float *coef;                // [B, n_seq]    c_j
};

// ---------------------------------------------------------------------------------------------
// forward (+ intermediates for the backward when TRAIN)
// STORE = false with TRAIN = true: loss only (rb2_fm_loss) -- nothing is kept for a backward
template <int D, bool TRAIN, bool STORE = TRAIN>
__global__ void __launch_bounds__(kThreads) k_fm_forward(FmTables t, const int64_t *__restrict__ ids,
                                                          const int64_t *__restrict__ offsets,
                                                          const float *__restrict__ label, int64_t B, int F,
                                                          int64_t n_rows, float inv_b, FmWs w,
                                                          float *__restrict__ y_out, OptScalars o) {
  constexpr int LANES = RowCfg<D>::LANES;
  constexpr int GROUPS = RowCfg<D>::GROUPS;
  static_assert(RowCfg<D>::VPL == 1, "FM path supports d <= 128");
  const int lane = threadIdx.x % 32, gl = lane % LANES, g = lane / LANES;
  const int warp_in_block = threadIdx.x / 32;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int64_t n_warps = (int64_t)gridDim.x * blockDim.x / 32;
  const float bias = t.bias[0];
  // dense-Adam parity mode: a row is read as the reference's dense optimizer holds it after step - 1 (rows the
  // previous batches did not touch kept moving on their momentum / weight decay; common.cuh row_replay)
  const bool lazy = o.kind == RB2_OPT_ADAM_LAZY;
  float loss_local = 0.f, gz_local = 0.f;
  for (int64_t s = warp_global; s < B; s += n_warps) {
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
    float sq = 0.f, first = 0.f, cross = 0.f;
    auto fetch = [&](int f, float4 &v) {
      int64_t row = ids[s * F + f] + offsets[f];
      if (row < 0 || row >= n_rows) {
        w.hdr->range_error = 1;
        row = min(max(row, (int64_t)0), n_rows - 1);
      }
      if (lazy) {
        v = row_ld_effective<D, true>(t.E, t.mE, t.vE, t.last, row, gl, o).v[0];
        if (gl == 0) {
          float pw = t.W[row];
          const int last = t.last[row];
          if (last < o.step - 1 && (last > 0 || o.wd != 0.f)) {
            float mw = t.mW[row], vw = t.vW[row];
            adam_replay_elem(pw, mw, vw, last + 1, o.step - 1, o);
          }
          first += pw;
        }
      } else {
        v = __ldg(reinterpret_cast<const float4 *>(t.E + row * D) + gl);
        if (gl == 0) first += __ldg(t.W + row);
      }
      if (STORE && gl == 0) {
        int64_t o = s * F + f;
        w.key[o] = (uint32_t)row;
        w.val[o] = (uint32_t)o;
      }
    };
    if (F == 2 && t.n_float == 0 && t.n_seq == 0) {
      // two fields = the point-wise "dot" model (fork's MFSimple, mfsimple.py:39-46:
      // sigmoid(<u,v> + b_u + b_i + b)): take the product directly instead of the
      // 0.5*[(u+v)^2 - u^2 - v^2] identity, which cancels badly in fp32
      if (g == 0) {
        float4 a, b;
        fetch(0, a);
        fetch(1, b);
        S = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        cross = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
      }
    } else {
      const int n_tok = t.n_seq > 0 ? t.n_tok : F;
      for (int f = g; f < n_tok; f += GROUPS) {
        float4 v;
        fetch(f, v);
        S.x += v.x; S.y += v.y; S.z += v.z; S.w += v.w;
        sq = fmaf(v.x, v.x, sq); sq = fmaf(v.y, v.y, sq); sq = fmaf(v.z, v.z, sq); sq = fmaf(v.w, v.w, sq);
      }
      // TOKEN_SEQ fields: one lane group pools a whole field (masked mean, abstract_recommender.py:293-309); the pooled
      // vector is the field's.  First order: the plain masked SUM of the scalars (layers.py:1000-1012), which fetch()
      // already adds to `first`.
      for (int j = g; j < t.n_seq; j += GROUPS) {
        const int c0 = t.seq_start[j], c1 = t.seq_start[j + 1];
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        int cnt = 0;
        for (int col = c0; col < c1; ++col) {
          if (ids[s * F + col] != 0) {
            float4 v;
            fetch(col, v);                         // also records (row, occurrence)
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            ++cnt;
          } else if (STORE && gl == 0) {           // masked: an occurrence of no row (sorted behind every real one)
            const int64_t oo = s * F + col;
            w.key[oo] = (uint32_t)n_rows;
            w.val[oo] = (uint32_t)oo;
          }
        }
        const float c = 1.f / ((float)cnt + 1e-8f);
        a.x *= c; a.y *= c; a.z *= c; a.w *= c;
        S.x += a.x; S.y += a.y; S.z += a.z; S.w += a.w;
        sq = fmaf(a.x, a.x, sq); sq = fmaf(a.y, a.y, sq); sq = fmaf(a.z, a.z, sq); sq = fmaf(a.w, a.w, sq);
        if (STORE) {
          reinterpret_cast<float4 *>(t.ef + ((size_t)s * t.n_seq + j) * D)[gl] = a;
          if (gl == 0) t.coef[(size_t)s * t.n_seq + j] = c;
        }
      }
    }
    if (t.n_float > 0) {
      // FLOAT fields: e_f = x * Ef[f] joins the sum and the sum of squares like any other field's vector
      for (int f = g; f < t.n_float; f += GROUPS) {
        const float x = __ldg(t.fx + s * t.n_float + f);
        float4 v = __ldg(reinterpret_cast<const float4 *>(t.Ef + (size_t)f * D) + gl);
        v.x *= x; v.y *= x; v.z *= x; v.w *= x;
        S.x += v.x; S.y += v.y; S.z += v.z; S.w += v.w;
        sq = fmaf(v.x, v.x, sq); sq = fmaf(v.y, v.y, sq); sq = fmaf(v.z, v.z, sq); sq = fmaf(v.w, v.w, sq);
        if (gl == 0) first = fmaf(x, __ldg(t.Wf + f), first);
      }
    }
    // across the lane groups of the warp: S (per lane-in-group), sq / cross / first (everything)
#pragma unroll
    for (int o = LANES; o < 32; o <<= 1) {
      S.x += __shfl_xor_sync(0xffffffffu, S.x, o);
      S.y += __shfl_xor_sync(0xffffffffu, S.y, o);
      S.z += __shfl_xor_sync(0xffffffffu, S.z, o);
      S.w += __shfl_xor_sync(0xffffffffu, S.w, o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      cross += __shfl_xor_sync(0xffffffffu, cross, o);
      first += __shfl_xor_sync(0xffffffffu, first, o);
    }
    float second;
    if (F == 2 && t.n_float == 0 && t.n_seq == 0) {
      second = cross;
    } else {
      // sum_k S_k^2 over the d elements: every group holds the full S after the reduction above
      float ss = S.x * S.x + S.y * S.y + S.z * S.z + S.w * S.w;
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      second = 0.5f * (ss - sq);                          // layers.py:164-171
    }
    const float z = first + bias + second;                // fm.py:49, layers.py:1061
    const float y = 1.f / (1.f + expf(-z));
    if (y_out && lane == 0) y_out[s] = y;
    if (TRAIN) {
      const float lab = label[s];
      // nn.BCELoss: log terms clamped at -100; backward divides by max(y(1-y), 1e-12)
      const float ly = fmaxf(logf(y), -100.f), l1y = fmaxf(logf(1.f - y), -100.f);
      const float yy = y * (1.f - y);
      const float gz = inv_b * (y - lab) * (yy / fmaxf(yy, 1e-12f));
      if (lane == 0) {
        loss_local += -(lab * ly + (1.f - lab) * l1y);
        gz_local += gz;
        if (STORE) w.gz[s] = gz;
      }
      if (STORE && g == 0)
        reinterpret_cast<float4 *>(w.gs + s * D)[gl] = make_float4(gz * S.x, gz * S.y, gz * S.z, gz * S.w);
    }
  }
  if (TRAIN && lane == 0 && warp_global < w.n_parts) {
    w.loss_part[warp_global] = (double)loss_local;
    w.gz_part[warp_global] = (double)gz_local;
  }
  (void)warp_in_block;
}

// ---------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void fm_row_step(const FmTables &t, int64_t row, int lane, const Row<D> &gsum, float zsum,
                                            const OptScalars &o) {
  // dE = sum gs - zsum * v ;  dW = zsum      (rows of TOKEN_SEQ tables: gsum is complete, no -zsum * v term)
  if (t.n_seq > 0 && row >= t.n_rows) return;          // the masked sequence entries' pseudo row
  const float zv = (t.n_seq > 0 && row >= t.seq_row_base) ? 0.f : zsum;
  if (o.kind == RB2_OPT_ADAM_LAZY) {
    Row<D> p = row_ld<D>(t.E, row, lane), m = row_ld<D>(t.mE, row, lane), v = row_ld<D>(t.vE, row, lane);
    const int last = t.last[row];
    row_replay<D>(p, m, v, last, o.step - 1, o);       // the value the forward read
    Row<D> g = gsum;
    row_fma<D>(g, -zv, p);
#pragma unroll
    for (int i = 0; i < RowCfg<D>::VPL; ++i) {
      adam_elem(p.v[i].x, m.v[i].x, v.v[i].x, g.v[i].x, o);
      adam_elem(p.v[i].y, m.v[i].y, v.v[i].y, g.v[i].y, o);
      adam_elem(p.v[i].z, m.v[i].z, v.v[i].z, g.v[i].z, o);
      adam_elem(p.v[i].w, m.v[i].w, v.v[i].w, g.v[i].w, o);
    }
    row_st<D>(t.E, row, lane, p);
    row_st<D>(t.mE, row, lane, m);
    row_st<D>(t.vE, row, lane, v);
    if (lane == 0) {
      float pw = t.W[row], mw = t.mW[row], vw = t.vW[row];
      if (last < o.step - 1 && (last > 0 || o.wd != 0.f)) adam_replay_elem(pw, mw, vw, last + 1, o.step - 1, o);
      adam_elem(pw, mw, vw, zsum, o);
      t.W[row] = pw;
      t.mW[row] = mw;
      t.vW[row] = vw;
      t.last[row] = o.step;
    }
    return;
  }
  Row<D> p = row_ld<D>(t.E, row, lane);
  Row<D> g = gsum;
  row_fma<D>(g, -zv, p);
  if (o.kind == kOptGradOut) {
    row_st<D>(t.E, row, lane, g);
    if (lane == 0) t.W[row] = zsum;
    return;
  }
  row_update<D>(t.E, t.mE, t.vE, row, lane, p, g, o);
  if (lane == 0) {
    float pw = t.W[row];
    if (o.kind == RB2_OPT_SGD) {
      sgd_elem(pw, zsum, o);
    } else {
      float m = t.mW[row], v = t.vW[row];
      adam_elem(pw, m, v, zsum, o);
      t.mW[row] = m;
      t.vW[row] = v;
    }
    t.W[row] = pw;
  }
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_fm_rows(FmTables t, FmWs w, int64_t n_occ, int F, int T,
                                                       int64_t n_tiles, OptScalars o) {
  constexpr int LANES = RowCfg<D>::LANES;
  constexpr int UNR = 4;
  const int lane = threadIdx.x % LANES;
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (tile >= n_tiles) return;
  const int64_t lo = tile * T, hi = min(lo + (int64_t)T, n_occ);
  const uint32_t *__restrict__ keys = w.key_s;
  const uint32_t *__restrict__ vals = w.val_s;
  const uint32_t kInvalid = 0xffffffffu;
  const uint32_t prev_key = lo > 0 ? keys[lo - 1] : kInvalid;
  uint32_t cur = kInvalid;
  bool started_before = false;
  Row<D> acc = row_zero<D>();
  float zacc = 0.f;
  uint8_t fh = 0, ft = 0;

  auto finish_run = [&](bool continues) {
    if (cur == kInvalid) return;
    if (!started_before && !continues) {
      fm_row_step<D>(t, cur, lane, acc, zacc, o);
    } else if (started_before) {
      row_st<D>(w.head, tile, lane, acc);
      if (lane == 0) w.head_z[tile] = zacc;
      fh = continues ? 2 : 1;
    } else {
      row_st<D>(w.tail, tile, lane, acc);
      if (lane == 0) w.tail_z[tile] = zacc;
      ft = 1;
    }
  };

  for (int64_t base = lo; base < hi; base += UNR) {
    uint32_t k[UNR], s[UNR];
    Row<D> c[UNR];
    float z[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      int64_t p = base + j;
      bool ok = p < hi;
      k[j] = ok ? keys[p] : kInvalid;
      if (t.n_seq > 0 && k[j] >= (uint32_t)t.n_rows) k[j] = kInvalid;      // masked sequence entries (sorted last)
      s[j] = ok ? vals[p] / (uint32_t)F : 0u;
    }
#pragma unroll
    for (int j = 0; j < UNR; ++j)
      if (k[j] != kInvalid) {
        c[j] = row_ld<D>(w.gs, s[j], lane);
        z[j] = w.gz[s[j]];
        if (t.n_seq > 0) {
          const int col = (int)(vals[base + j] % (uint32_t)F);
          const int sj = col >= t.n_tok ? t.col_seq[col] : -1;
          if (sj >= 0) {          // an entry of a pooled field: c_j * (gz * S - gz * e_j) for the row, gz for W (a sum)
            const size_t q = (size_t)s[j] * t.n_seq + sj;
            const float cf = t.coef[q];
            const Row<D> e = row_ld<D>(t.ef, (int64_t)q, lane);
            row_fma<D>(c[j], -z[j], e);
            c[j] = row_scale<D>(cf, c[j]);
          }
        }
      }
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      if (k[j] == kInvalid) break;
      if (k[j] != cur) {
        finish_run(false);
        cur = k[j];
        acc = row_zero<D>();
        zacc = 0.f;
        started_before = (base + j == lo) && (cur == prev_key);
      }
      row_add<D>(acc, c[j]);
      zacc += z[j];
    }
  }
  finish_run(hi < n_occ && keys[hi] == cur);
  if (lane == 0) {
    w.fh[tile] = fh;
    w.ft[tile] = ft;
  }
}

// A hot row (a field with a handful of values) spans thousands of tiles; walking its partials one chain at a
// time left the GPU idle for longer than the main pass (ncu: 333 us at 1.7 % SM throughput).  First level, in
// parallel: every aligned block of 64 tiles that lies ENTIRELY inside one run (all heads "continue") is summed
// by one warp, in a fixed order; the chain walk below then advances 64 tiles per load over such blocks.
template <int D>
__global__ void __launch_bounds__(kThreads) k_fm_block_reduce(FmWs w, int64_t n_tiles) {
  constexpr int LANES = RowCfg<D>::LANES;
  constexpr int GROUPS = 32 / LANES;
  const int lane32 = threadIdx.x % 32, lane = lane32 % LANES, g = lane32 / LANES;
  const int64_t blk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int64_t t0 = blk * 64;
  if (t0 + 64 > n_tiles) {                    // (whole warps leave together)
    if (t0 < n_tiles + 64 && lane32 == 0 && blk <= n_tiles / 64) w.blk_ok[blk] = 0;
    return;
  }
  const bool in0 = w.fh[t0 + lane32] == 2, in1 = w.fh[t0 + 32 + lane32] == 2;
  const bool ok = __all_sync(0xffffffffu, in0 && in1);
  if (!ok) {
    if (lane32 == 0) w.blk_ok[blk] = 0;
    return;
  }
  Row<D> acc = row_zero<D>();
  float z = 0.f;
#pragma unroll 4
  for (int j = g; j < 64; j += GROUPS) {      // group g: tiles g, g + GROUPS, ...
    Row<D> r = row_ld<D>(w.head, t0 + j, lane);
    row_add<D>(acc, r);
    if (lane == 0) z += w.head_z[t0 + j];
  }
#pragma unroll
  for (int o = LANES; o < 32; o <<= 1) {      // groups -> group 0
#pragma unroll
    for (int i = 0; i < RowCfg<D>::VPL; ++i) {
      acc.v[i].x += __shfl_xor_sync(0xffffffffu, acc.v[i].x, o);
      acc.v[i].y += __shfl_xor_sync(0xffffffffu, acc.v[i].y, o);
      acc.v[i].z += __shfl_xor_sync(0xffffffffu, acc.v[i].z, o);
      acc.v[i].w += __shfl_xor_sync(0xffffffffu, acc.v[i].w, o);
    }
    z += __shfl_xor_sync(0xffffffffu, z, o);
  }
  if (g == 0) {
    row_st<D>(w.blk_head, blk, lane, acc);
    if (lane == 0) { w.blk_z[blk] = z; w.blk_ok[blk] = 1; }
  }
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_fm_fixup(FmTables t, FmWs w, int64_t n_occ, int T, int64_t n_tiles,
                                                        OptScalars o) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (tile >= n_tiles || !w.ft[tile]) return;
  int64_t last_pos = min((tile + 1) * (int64_t)T, n_occ) - 1;
  uint32_t key = w.key_s[last_pos];
  Row<D> acc = row_ld<D>(w.tail, tile, lane);
  float zacc = w.tail_z[tile];
  // hot rows (a field with a handful of values) have chains of thousands of partials: walk them CH at
  // a time with all loads in flight (fixed summation order)
  constexpr int CH = 16;
  bool done = false;
  for (int64_t j0 = tile + 1; j0 < n_tiles && !done; j0 += CH) {
    while ((j0 & 63) == 0 && j0 + 64 <= n_tiles && w.blk_ok[j0 >> 6]) {   // 64 interior tiles at once
      Row<D> b = row_ld<D>(w.blk_head, j0 >> 6, lane);
      row_add<D>(acc, b);
      zacc += w.blk_z[j0 >> 6];
      j0 += 64;
    }
    if (j0 >= n_tiles) break;
    uint8_t f[CH];
    Row<D> part[CH];
    float pz[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) f[c] = (j0 + c < n_tiles) ? w.fh[j0 + c] : (uint8_t)0;
    int cnt = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (!done) {
        if (f[c]) ++cnt;
        if (f[c] != 2) done = true;
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c)
      if (c < cnt) { part[c] = row_ld<D>(w.head, j0 + c, lane); pz[c] = w.head_z[j0 + c]; }
#pragma unroll
    for (int c = 0; c < CH; ++c)
      if (c < cnt) { row_add<D>(acc, part[c]); zacc += pz[c]; }
  }
  fm_row_step<D>(t, key, lane, acc, zacc, o);
}

// bias step + loss reduction (single block, fixed order)
__global__ void k_fm_bias_loss(FmTables t, FmWs w, int64_t n_parts, double inv_b, OptScalars o, float *loss_out,
                               double *loss_accum) {
  __shared__ double sl[256], sg[256];
  double l = 0.0, g = 0.0;
  for (int64_t i = threadIdx.x; i < n_parts; i += blockDim.x) { l += w.loss_part[i]; g += w.gz_part[i]; }
  sl[threadIdx.x] = l;
  sg[threadIdx.x] = g;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { sl[threadIdx.x] += sl[threadIdx.x + s]; sg[threadIdx.x] += sg[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float loss = (float)(sl[0] * inv_b);
    loss_out[0] = loss;
    if (loss_accum) loss_accum[0] += (double)loss;
    if (o.kind == kOptLossOnly) return;
    float b = t.bias[0], gb = (float)sg[0];
    if (o.kind == kOptGradOut) {
      loss_out[1] = gb;          // this rank's share of d loss / d bias
      return;
    }
    if (o.kind == RB2_OPT_SGD) {
      sgd_elem(b, gb, o);
    } else {                     // Adam (the bias is dense: stepped at every step, so lazy == plain)
      float m = t.bias[1], v = t.bias[2];
      adam_elem(b, m, v, gb, o);
      t.bias[1] = m;
      t.bias[2] = v;
    }
    t.bias[0] = b;
  }
}

// ---- FLOAT fields: every sample touches every float row, so their gradients are dense reductions over the batch:
//   dEf[f] = sum_s x_sf * gs[s] - (sum_s x_sf^2 gz[s]) * Ef[f]        (e_f = x Ef[f]; d z / d e_f = S - e_f)
//   dWf[f] = sum_s x_sf * gz[s]
// k_fm_float_reduce: one block per (chunk of kFloatChunk samples, field), fixed-order sums; k_fm_float_update: one
// block per field sums the chunk partials in chunk order and takes the optimizer step (dense: stepped every step).
constexpr int kFloatChunk = 4096;
template <int D>
__global__ void __launch_bounds__(kThreads) k_fm_float_reduce(FmTables t, FmWs w, int64_t B, float *__restrict__ part) {
  constexpr int LANES = RowCfg<D>::LANES, GR = kThreads / LANES;
  __shared__ float sm[GR][D + 2];
  const int lane = threadIdx.x % LANES, grp = threadIdx.x / LANES;
  const int f = blockIdx.y;
  const int64_t lo = (int64_t)blockIdx.x * kFloatChunk, hi = min(lo + (int64_t)kFloatChunk, B);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float z2 = 0.f, z1 = 0.f;
  for (int64_t s = lo + grp; s < hi; s += GR) {
    const float x = __ldg(t.fx + s * t.n_float + f);
    const float4 gv = __ldg(reinterpret_cast<const float4 *>(w.gs + s * D) + lane);
    acc.x = fmaf(x, gv.x, acc.x); acc.y = fmaf(x, gv.y, acc.y); acc.z = fmaf(x, gv.z, acc.z); acc.w = fmaf(x, gv.w, acc.w);
    if (lane == 0) {
      const float gz = __ldg(w.gz + s);
      z2 = fmaf(x * x, gz, z2);
      z1 = fmaf(x, gz, z1);
    }
  }
  sm[grp][4 * lane] = acc.x; sm[grp][4 * lane + 1] = acc.y; sm[grp][4 * lane + 2] = acc.z; sm[grp][4 * lane + 3] = acc.w;
  if (lane == 0) { sm[grp][D] = z2; sm[grp][D + 1] = z1; }
  __syncthreads();
  if (threadIdx.x < D + 2) {
    float a = 0.f;
    for (int gI = 0; gI < GR; ++gI) a += sm[gI][threadIdx.x];
    part[((size_t)f * gridDim.x + blockIdx.x) * (D + 2) + threadIdx.x] = a;
  }
}
template <int D>
__global__ void __launch_bounds__(256) k_fm_float_update(FmTables t, const float *__restrict__ part, int n_chunks,
                                                         OptScalars o) {
  const int f = blockIdx.x, k = threadIdx.x;
  if (k >= D + 2) return;
  __shared__ float zz[2];
  float a = 0.f;
  for (int c = 0; c < n_chunks; ++c) a += part[((size_t)f * n_chunks + c) * (D + 2) + k];
  if (k >= D) zz[k - D] = a;
  __syncthreads();
  if (k < D) {
    float p = t.Ef[(size_t)f * D + k];
    const float g = fmaf(-zz[0], p, a);
    if (o.kind == RB2_OPT_SGD) {
      sgd_elem(p, g, o);
    } else {
      float m = t.mEf[(size_t)f * D + k], v = t.vEf[(size_t)f * D + k];
      adam_elem(p, m, v, g, o);
      t.mEf[(size_t)f * D + k] = m;
      t.vEf[(size_t)f * D + k] = v;
    }
    t.Ef[(size_t)f * D + k] = p;
  } else if (k == D + 1) {
    float p = t.Wf[f];
    if (o.kind == RB2_OPT_SGD) {
      sgd_elem(p, a, o);
    } else {
      float m = t.mWf[f], v = t.vWf[f];
      adam_elem(p, m, v, a, o);
      t.mWf[f] = m;
      t.vWf[f] = v;
    }
    t.Wf[f] = p;
  }
}

__global__ void k_zero_parts(FmWs w) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w.n_parts) { w.loss_part[i] = 0.0; w.gz_part[i] = 0.0; }
}

}  // namespace

#define RB2_FM_DIM(dim, ...)                                                                  \
  switch (dim) {                                                                              \
    case 16: { constexpr int D_ = 16; __VA_ARGS__; } break;                                   \
    case 32: { constexpr int D_ = 32; __VA_ARGS__; } break;                                   \
    case 64: { constexpr int D_ = 64; __VA_ARGS__; } break;                                   \
    case 128: { constexpr int D_ = 128; __VA_ARGS__; } break;                                 \
    default:                                                                                  \
      rb2_set_error("FM embedding dim %d not supported (16, 32, 64, 128)", (int)(dim));       \
      return RB2_EINVAL;                                                                      \
  }

static int fm_set_float(FmTables &t, const rb2_fm_float *f, bool train, const char *who) {
  if (!f || f->n_float <= 0) return 0;
  RB2_REQUIRE(f->n_float <= RB2_FM_MAX_FLOAT, RB2_EINVAL, "%s: %d float fields (max %d)", who, (int)f->n_float,
              (int)RB2_FM_MAX_FLOAT);
  RB2_REQUIRE(f->values && f->Ef && f->Wf, RB2_EINVAL, "%s: float fields need values, Ef and Wf", who);
  t.fx = f->values;
  t.n_float = f->n_float;
  t.Ef = f->Ef; t.mEf = f->mEf; t.vEf = f->vEf;
  t.Wf = f->Wf; t.mWf = f->mWf; t.vWf = f->vWf;
  (void)train;
  return 0;
}

static int fm_set_seq(FmTables &t, const rb2_fm_seq *q, int64_t n_rows, int32_t n_fields, bool train, const char *who) {
  t.n_rows = n_rows;
  if (!q || q->n_seq <= 0) return 0;
  RB2_REQUIRE(q->n_seq <= RB2_FM_MAX_SEQ && q->n_token_cols >= 0 && q->n_token_cols < n_fields, RB2_EINVAL,
              "%s: bad TOKEN_SEQ description (n_seq %d, token columns %d of %d)", who, (int)q->n_seq,
              (int)q->n_token_cols, (int)n_fields);
  RB2_REQUIRE(q->seq_start && q->col_seq && q->seq_row_base >= 0 && q->seq_row_base <= n_rows, RB2_EINVAL,
              "%s: TOKEN_SEQ fields need seq_start, col_seq and seq_row_base", who);
  if (train) RB2_REQUIRE(q->pooled && q->coef, RB2_EINVAL, "%s: TOKEN_SEQ training needs the pooled / coef buffers", who);
  t.n_tok = q->n_token_cols;
  t.n_seq = q->n_seq;
  t.seq_start = q->seq_start;
  t.col_seq = q->col_seq;
  t.seq_row_base = q->seq_row_base;
  t.ef = q->pooled;
  t.coef = q->coef;
  return 0;
}

extern "C" size_t rb2_fm_workspace_bytes(int64_t batch, int32_t n_fields, int32_t dim) {
  FmWs w;
  return carve(w, nullptr, batch, n_fields, dim);
}

static int fm_step(FmTables t, int64_t n_rows, int32_t dim, const int64_t *ids, const int64_t *offsets, int32_t n_fields,
                   const float *label, int64_t batch, double norm_batch, OptScalars o, float *loss_out,
                   double *loss_accum, void *workspace, size_t workspace_bytes, cudaStream_t st, const char *who) {
  RB2_REQUIRE(batch > 0 && n_fields > 0 && batch * (int64_t)n_fields < ((int64_t)1 << 31), RB2_EINVAL,
              "%s: batch*fields out of range", who);
  RB2_REQUIRE(n_rows > 0 && n_rows < ((int64_t)1 << 32) - 1, RB2_EINVAL, "%s: table too large", who);
  FmWs w;
  size_t need = carve(w, workspace, batch, n_fields, dim);
  RB2_REQUIRE(workspace_bytes >= need, RB2_EWORKSPACE, "%s: workspace %zu < %zu", who, workspace_bytes, need);
  const int64_t M = batch * n_fields;
  k_zero_parts<<<(unsigned)((w.n_parts + 255) / 256), 256, 0, st>>>(w);
  RB2_FM_DIM(dim, {
    constexpr int LANES = RowCfg<D_>::LANES;
    int64_t warps = std::min<int64_t>(batch, w.n_parts);
    warps = std::min<int64_t>(warps, (int64_t)rb2_num_sms() * 64);
    unsigned fblocks = (unsigned)((warps * 32 + kThreads - 1) / kThreads);
    if ((int64_t)fblocks * (kThreads / 32) > w.n_parts) fblocks = (unsigned)std::max<int64_t>(1, w.n_parts / (kThreads / 32));
    {
      ProfScope prof(RB2_ST_FM_FWD, st, 2);
      k_fm_forward<D_, true><<<fblocks, kThreads, 0, st>>>(t, ids, offsets, label, batch, n_fields, n_rows,
                                                           (float)(1.0 / norm_batch), w, nullptr, o);
    }
    if (t.n_float > 0 && o.kind != kOptGradOut) {
      ProfScope prof(RB2_ST_FM_UPDATE, st, 2);
      const int n_chunks = (int)((batch + kFloatChunk - 1) / kFloatChunk);
      k_fm_float_reduce<D_><<<dim3((unsigned)n_chunks, (unsigned)t.n_float), kThreads, 0, st>>>(t, w, batch, w.float_part);
      k_fm_float_update<D_><<<(unsigned)t.n_float, 256, 0, st>>>(t, w.float_part, n_chunks, o);
    }
    size_t tmp = w.cub_bytes;
    {
      ProfScope prof(RB2_ST_SORT_ITEM, st, 2 + (rb2_bits_for(n_rows) + 7) / 8);
      { int rc_ = rb2sort::sort_positions(w.key, w.key_s, w.val_s, M, rb2_bits_for(n_rows + (t.n_seq > 0 ? 1 : 0)), w.cub_tmp, w.cub_bytes, st); if (rc_) return rc_; }
    }
    const int T = rb2_pick_tile(M, LANES);
    const int64_t nt = (M + T - 1) / T;
    unsigned blocks = (unsigned)((nt * LANES + kThreads - 1) / kThreads);
    {
      ProfScope prof(RB2_ST_FM_UPDATE, st, 4);
      k_fm_rows<D_><<<blocks, kThreads, 0, st>>>(t, w, M, n_fields, T, nt, o);
      const int64_t nblk = nt / 64 + 1;
      k_fm_block_reduce<D_><<<(unsigned)((nblk * 32 + kThreads - 1) / kThreads), kThreads, 0, st>>>(w, nt);
      k_fm_fixup<D_><<<blocks, kThreads, 0, st>>>(t, w, M, T, nt, o);
      k_fm_bias_loss<<<1, 256, 0, st>>>(t, w, w.n_parts, 1.0 / norm_batch, o, loss_out, loss_accum);
    }
  });
  RB2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb2_fm_train_step(float *E, float *mE, float *vE, float *W, float *mW, float *vW, float *bias3,
                                 int32_t *row_last, int64_t n_rows, int32_t dim, const int64_t *ids,
                                 const int64_t *offsets, int32_t n_fields, const float *label, int64_t batch,
                                 const rb2_optim *h_opt, float *loss_out, double *loss_accum, void *workspace,
                                 size_t workspace_bytes, void *stream, const rb2_fm_float *h_float,
                                 const rb2_fm_seq *h_seq) {
  RB2_REQUIRE(E && W && bias3 && ids && offsets && label && h_opt && loss_out && workspace, RB2_EINVAL,
              "rb2_fm_train_step: null argument");
  OptScalars o = rb2_opt_scalars(h_opt);
  RB2_REQUIRE(o.kind == RB2_OPT_SGD || o.kind == RB2_OPT_ADAM || o.kind == RB2_OPT_ADAM_LAZY, RB2_EINVAL,
              "rb2_fm_train_step: optimizer kind %d not supported (sgd, adam, adam_lazy)", o.kind);
  if (o.kind != RB2_OPT_SGD) RB2_REQUIRE(mE && vE && mW && vW, RB2_EINVAL, "rb2_fm_train_step: Adam needs m and v");
  if (o.kind == RB2_OPT_ADAM_LAZY)
    RB2_REQUIRE(row_last && o.lazy_step_size && o.lazy_bc2_sqrt, RB2_EINVAL,
                "rb2_fm_train_step: adam_lazy needs row_last and the bias-correction tables");
  FmTables t{E, mE, vE, W, mW, vW, bias3, row_last};
  if (int rc = fm_set_float(t, h_float, true, "rb2_fm_train_step")) return rc;
  if (int rc = fm_set_seq(t, h_seq, n_rows, n_fields, true, "rb2_fm_train_step")) return rc;
  if (t.n_float > 0 && o.kind != RB2_OPT_SGD)
    RB2_REQUIRE(t.mEf && t.vEf && t.mWf && t.vWf, RB2_EINVAL, "rb2_fm_train_step: Adam needs the float fields' m and v");
  return fm_step(t, n_rows, dim, ids, offsets, n_fields, label, batch, (double)batch, o, loss_out, loss_accum, workspace,
                 workspace_bytes, (cudaStream_t)stream, "rb2_fm_train_step");
}

// Sharded step, the local part: rows_e [n_rows, dim] / rows_w [n_rows] are the FETCHED copies of the rows this
// rank's samples touch (ids index them, offsets may be all zero); on return every row holds its summed
// gradient (d loss / d row with loss = sum over the GLOBAL batch / global_batch), loss2[0] = this rank's share of
// the loss and loss2[1] = its share of d loss / d bias.  Nothing is stepped here.
extern "C" int rb2_fm_grad_step(float *rows_e, float *rows_w, const float *bias3, int64_t n_rows, int32_t dim,
                                const int64_t *ids, const int64_t *offsets, int32_t n_fields, const float *label,
                                int64_t batch, int64_t global_batch, float *loss2, void *workspace,
                                size_t workspace_bytes, void *stream) {
  RB2_REQUIRE(rows_e && rows_w && bias3 && ids && offsets && label && loss2 && workspace, RB2_EINVAL,
              "rb2_fm_grad_step: null argument");
  RB2_REQUIRE(global_batch >= batch, RB2_EINVAL, "rb2_fm_grad_step: global_batch < batch");
  OptScalars o = {};
  o.kind = kOptGradOut;
  FmTables t{rows_e, nullptr, nullptr, rows_w, nullptr, nullptr, const_cast<float *>(bias3), nullptr};
  return fm_step(t, n_rows, dim, ids, offsets, n_fields, label, batch, (double)global_batch, o, loss2, nullptr, workspace,
                 workspace_bytes, (cudaStream_t)stream, "rb2_fm_grad_step");
}

// ---- owner side of the d = 1 table: (ids[M], grads[M]) with duplicates -> sum per row (fixed order), one step
namespace {
__global__ void k_scalar_keys(const int64_t *__restrict__ ids, int64_t M, int64_t n_rows, uint32_t *key, uint32_t *val) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  int64_t r = ids[i];
  key[i] = (uint32_t)min(max(r, (int64_t)0), n_rows - 1);
  val[i] = (uint32_t)i;
}
__global__ void k_scalar_rows(float *P, float *Mm, float *V, const uint32_t *__restrict__ key_s,
                              const uint32_t *__restrict__ val_s, const float *__restrict__ grads, int64_t M,
                              OptScalars o) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const uint32_t k = key_s[i];
  if (i > 0 && key_s[i - 1] == k) return;          // not the head of its run
  float g = 0.f;
  for (int64_t j = i; j < M && key_s[j] == k; ++j) g += grads[val_s[j]];   // runs are short (<= senders)
  float p = P[k];
  if (o.kind == RB2_OPT_SGD) {
    sgd_elem(p, g, o);
  } else {
    float m = Mm[k], v = V[k];
    adam_elem(p, m, v, g, o);
    Mm[k] = m;
    V[k] = v;
  }
  P[k] = p;
}
__global__ void k_scalar_step(float *p3, const float *grad, OptScalars o) {
  float p = p3[0], g = grad[0];
  if (o.kind == RB2_OPT_SGD) {
    sgd_elem(p, g, o);
  } else {
    float m = p3[1], v = p3[2];
    adam_elem(p, m, v, g, o);
    p3[1] = m;
    p3[2] = v;
  }
  p3[0] = p;
}
}  // namespace

extern "C" size_t rb2_scalar_rows_update_workspace_bytes(int64_t m) {
  Carver c(nullptr);
  c.take<uint32_t>(m); c.take<uint32_t>(m); c.take<uint32_t>(m); c.take<uint32_t>(m);
  size_t b = 0;
  b = rb2sort::tmp_bytes(m);
  c.take<char>(b);
  return c.off + 256;
}
extern "C" int rb2_scalar_rows_update(float *p, float *m, float *v, int64_t n_rows, const int64_t *ids,
                                      const float *grads, int64_t M, const rb2_optim *h_opt, void *workspace,
                                      size_t workspace_bytes, void *stream) {
  RB2_REQUIRE(p && ids && grads && h_opt && workspace, RB2_EINVAL, "rb2_scalar_rows_update: null argument");
  RB2_REQUIRE(M >= 0 && M < ((int64_t)1 << 31) && n_rows > 0 && n_rows < ((int64_t)1 << 32) - 1, RB2_EINVAL,
              "rb2_scalar_rows_update: sizes out of range");
  if (M == 0) return 0;
  OptScalars o = rb2_opt_scalars(h_opt);
  RB2_REQUIRE(o.kind == RB2_OPT_SGD || o.kind == RB2_OPT_ADAM, RB2_EINVAL, "rb2_scalar_rows_update: sgd or adam");
  if (o.kind != RB2_OPT_SGD) RB2_REQUIRE(m && v, RB2_EINVAL, "rb2_scalar_rows_update: Adam needs m and v");
  RB2_REQUIRE(workspace_bytes >= rb2_scalar_rows_update_workspace_bytes(M) - 256, RB2_EWORKSPACE,
              "rb2_scalar_rows_update: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(workspace);
  uint32_t *key = c.take<uint32_t>(M), *key_s = c.take<uint32_t>(M), *val = c.take<uint32_t>(M), *val_s = c.take<uint32_t>(M);
  const size_t b = rb2sort::tmp_bytes(M);
  char *tmp = c.take<char>(b);
  unsigned blocks = (unsigned)((M + 255) / 256);
  k_scalar_keys<<<blocks, 256, 0, st>>>(ids, M, n_rows, key, val);
  { int rc_ = rb2sort::sort_positions(key, key_s, val_s, M, rb2_bits_for(n_rows), tmp, b, st); if (rc_) return rc_; }
  k_scalar_rows<<<blocks, 256, 0, st>>>(p, m, v, key_s, val_s, grads, M, o);
  RB2_CUDA(cudaGetLastError());
  return 0;
}
/* one optimizer step of a 3-float parameter block (value, m, v) with the gradient read from the device */
extern "C" int rb2_scalar_step(float *p3, const float *grad, const rb2_optim *h_opt, void *stream) {
  RB2_REQUIRE(p3 && grad && h_opt, RB2_EINVAL, "rb2_scalar_step: null argument");
  OptScalars o = rb2_opt_scalars(h_opt);
  RB2_REQUIRE(o.kind == RB2_OPT_SGD || o.kind == RB2_OPT_ADAM, RB2_EINVAL, "rb2_scalar_step: sgd or adam");
  k_scalar_step<<<1, 1, 0, (cudaStream_t)stream>>>(p3, grad, o);
  RB2_CUDA(cudaGetLastError());
  return 0;
}

/* RB2_OPT_ADAM_LAZY: bring every row of E and of W to step h_opt->step (the zero-gradient / weight-decay steps it
 * missed) before the tables are read by predict / loss / a checkpoint. */
namespace {
template <int D>
__global__ void __launch_bounds__(kThreads) k_fm_lazy_flush(FmTables t, int64_t n_rows, OptScalars o) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (row >= n_rows) return;
  const int last = t.last[row];
  if (last >= o.step) return;
  if (last > 0 || o.wd != 0.f) {
    Row<D> p = row_ld<D>(t.E, row, lane), m = row_ld<D>(t.mE, row, lane), v = row_ld<D>(t.vE, row, lane);
    row_replay<D>(p, m, v, last, o.step, o);
    row_st<D>(t.E, row, lane, p);
    row_st<D>(t.mE, row, lane, m);
    row_st<D>(t.vE, row, lane, v);
    if (lane == 0) {
      float pw = t.W[row], mw = t.mW[row], vw = t.vW[row];
      adam_replay_elem(pw, mw, vw, last + 1, o.step, o);
      t.W[row] = pw;
      t.mW[row] = mw;
      t.vW[row] = vw;
    }
  }
  if (lane == 0) t.last[row] = o.step;
}
}  // namespace

extern "C" int rb2_fm_lazy_flush(float *E, float *mE, float *vE, float *W, float *mW, float *vW, int32_t *row_last,
                                 int64_t n_rows, int32_t dim, const rb2_optim *h_opt, void *stream) {
  RB2_REQUIRE(E && mE && vE && W && mW && vW && row_last && h_opt, RB2_EINVAL, "rb2_fm_lazy_flush: null argument");
  OptScalars o = rb2_opt_scalars(h_opt);
  RB2_REQUIRE(o.kind == RB2_OPT_ADAM_LAZY && o.lazy_step_size && o.lazy_bc2_sqrt, RB2_EINVAL,
              "rb2_fm_lazy_flush: needs an adam_lazy optimizer description");
  if (n_rows <= 0 || o.step <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  FmTables t{E, mE, vE, W, mW, vW, nullptr, row_last};
  RB2_FM_DIM(dim, {
    constexpr int LANES = RowCfg<D_>::LANES;
    k_fm_lazy_flush<D_><<<(unsigned)((n_rows * LANES + kThreads - 1) / kThreads), kThreads, 0, st>>>(t, n_rows, o);
  });
  RB2_CUDA(cudaGetLastError());
  return 0;
}

/* forward + mean BCE only (FM.calculate_loss, fm.py:52-56, without a backward): loss_out[0] = the batch's loss */
extern "C" int rb2_fm_loss(const float *E, const float *W, const float *bias3, int64_t n_rows, int32_t dim,
                           const int64_t *ids, const int64_t *offsets, int32_t n_fields, const float *label,
                           int64_t batch, float *loss_out, void *workspace, size_t workspace_bytes, void *stream,
                           const rb2_fm_float *h_float, const rb2_fm_seq *h_seq) {
  RB2_REQUIRE(E && W && bias3 && ids && offsets && label && loss_out && workspace, RB2_EINVAL,
              "rb2_fm_loss: null argument");
  RB2_REQUIRE(batch > 0 && n_fields > 0, RB2_EINVAL, "rb2_fm_loss: empty batch");
  FmWs w;
  size_t need = carve(w, workspace, batch, n_fields, dim);
  RB2_REQUIRE(workspace_bytes >= need, RB2_EWORKSPACE, "rb2_fm_loss: workspace %zu < %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  FmTables t{const_cast<float *>(E), nullptr, nullptr, const_cast<float *>(W), nullptr, nullptr,
             const_cast<float *>(bias3), nullptr};
  if (int rc = fm_set_float(t, h_float, false, "rb2_fm_loss")) return rc;
  if (int rc = fm_set_seq(t, h_seq, n_rows, n_fields, false, "rb2_fm_loss")) return rc;
  OptScalars o = {};
  o.kind = kOptLossOnly;
  k_zero_parts<<<(unsigned)((w.n_parts + 255) / 256), 256, 0, st>>>(w);
  RB2_FM_DIM(dim, {
    int64_t warps = std::min<int64_t>(std::min<int64_t>(batch, w.n_parts), (int64_t)rb2_num_sms() * 64);
    unsigned fblocks = (unsigned)((warps * 32 + kThreads - 1) / kThreads);
    if ((int64_t)fblocks * (kThreads / 32) > w.n_parts) fblocks = (unsigned)std::max<int64_t>(1, w.n_parts / (kThreads / 32));
    ProfScope prof(RB2_ST_FM_FWD, st, 2);
    k_fm_forward<D_, true, false><<<fblocks, kThreads, 0, st>>>(t, ids, offsets, label, batch, n_fields, n_rows,
                                                                (float)(1.0 / (double)batch), w, nullptr, o);
    k_fm_bias_loss<<<1, 256, 0, st>>>(t, w, w.n_parts, 1.0 / (double)batch, o, loss_out, nullptr);
  });
  RB2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb2_fm_predict(const float *E, const float *W, const float *bias3, int64_t n_rows, int32_t dim,
                              const int64_t *ids, const int64_t *offsets, int32_t n_fields, int64_t batch,
                              float *y_out, void *workspace, size_t workspace_bytes, void *stream,
                              const rb2_fm_float *h_float, const rb2_fm_seq *h_seq) {
  RB2_REQUIRE(E && W && bias3 && ids && offsets && y_out && workspace, RB2_EINVAL, "rb2_fm_predict: null argument");
  if (batch <= 0) return 0;
  FmWs w;
  size_t need = carve(w, workspace, batch, n_fields, dim);
  RB2_REQUIRE(workspace_bytes >= need, RB2_EWORKSPACE, "rb2_fm_predict: workspace %zu < %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  FmTables t{const_cast<float *>(E), nullptr, nullptr, const_cast<float *>(W), nullptr, nullptr,
             const_cast<float *>(bias3), nullptr};
  if (int rc = fm_set_float(t, h_float, false, "rb2_fm_predict")) return rc;
  if (int rc = fm_set_seq(t, h_seq, n_rows, n_fields, false, "rb2_fm_predict")) return rc;
  RB2_FM_DIM(dim, {
    int64_t warps = std::min<int64_t>(batch, (int64_t)rb2_num_sms() * 64);
    unsigned fblocks = (unsigned)((warps * 32 + kThreads - 1) / kThreads);
    ProfScope prof(RB2_ST_FM_FWD, st);
    OptScalars none = {};
    k_fm_forward<D_, false><<<fblocks, kThreads, 0, st>>>(t, ids, offsets, nullptr, batch, n_fields, n_rows, 1.f, w,
                                                          y_out, none);
  });
  RB2_CUDA(cudaGetLastError());
  return 0;
}
