// common.cuh -- shared device helpers for the sm_100a kernels of recbole_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/recbole_b200.h"

// ---- host-side error plumbing (api.cu) ------------------------------------------------------
void rb2_set_error(const char *fmt, ...);

#define RB2_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      rb2_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return (int)e__;                                                                        \
    }                                                                                         \
  } while (0)

#define RB2_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) {                   \
      rb2_set_error(__VA_ARGS__);    \
      return (code);                 \
    }                                \
  } while (0)

// optional per-stage device timing + launch counting (api.cu); off by default, costs two event
// records per stage when on.  Stage ids are part of the ABI (include/recbole_b200.h).
void rb2_prof_begin(int stage, cudaStream_t st);
void rb2_prof_end(int stage, cudaStream_t st, int launches);

// scorer state (include/recbole_b200.h rb2_scorer_state): the state of the call in flight on this thread -- the caller's
// (entry points *_s) or the thread's default (api.cu)
rb2_scorer_state &rb2_cur_scorer();
struct ScorerScope {
  rb2_scorer_state *prev;
  explicit ScorerScope(rb2_scorer_state *s);
  ~ScorerScope();
};

struct ProfScope {
  int stage, launches;
  cudaStream_t st;
  ProfScope(int stage_, cudaStream_t st_, int launches_ = 1) : stage(stage_), launches(launches_), st(st_) {
    rb2_prof_begin(stage, st);
  }
  ~ProfScope() { rb2_prof_end(stage, st, launches); }
};

static inline size_t rb2_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Carves a workspace buffer into aligned pieces; `p == nullptr` just measures.
struct Carver {
  char *base;
  size_t off = 0;
  explicit Carver(void *p) : base(reinterpret_cast<char *>(p)) {}
  template <typename T>
  T *take(size_t n) {
    size_t o = off;
    off = rb2_align(off + n * sizeof(T));
    return base ? reinterpret_cast<T *>(base + o) : nullptr;
  }
};

// workspace header: sticky flags the host wrapper checks whenever it synchronises anyway
struct WsHeader {
  int32_t range_error;  // an id was outside its table (clamped on device; reference raises IndexError)
  int32_t peer_timeout; // a cross-GPU barrier of the peer-memory step gave up waiting (a rank died or fell out of step)
  int32_t pad[62];
};

// tile length for the sorted-occurrence walks: long enough that few runs straddle tiles, short
// enough to fill the machine
static inline int rb2_num_sms();
static inline int64_t rb2_max_tiles(int64_t n_occ) { return (n_occ + 7) / 8; }
static inline int rb2_bits_for(int64_t n) {
  int b = 1;
  while (b < 32 && ((int64_t)1 << b) < n) ++b;
  return b;
}

// SM count of the CURRENT device (cached per device: a process may drive several)
static inline int rb2_num_sms() {
  constexpr int kMaxDev = 64;
  static int cache[kMaxDev] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return 148;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

static inline int rb2_pick_tile(int64_t n_occ, int lanes) {
  int64_t groups_wanted = (int64_t)rb2_num_sms() * 2048 / lanes;  // one full wave of lane groups
  int64_t t = n_occ / (groups_wanted > 0 ? groups_wanted : 1);
  if (t < 8) t = 8;
  if (t > 64) t = 64;
  return (int)t;
}

// ---- row vectors: one embedding row spread over a "group" of LANES lanes ---------------------
// D = 16, 32, 64, 128, 256.  Lane l of the group holds float4 chunks l, l+LANES, ... so that a
// row load is one fully coalesced 128-bit access per lane.
template <int D>
struct RowCfg {
  static_assert(D == 8 || D == 16 || D == 32 || D == 64 || D == 128 || D == 256, "unsupported dim");
  static constexpr int VPL = (D >= 128) ? D / 128 : 1;     // float4 per lane
  static constexpr int LANES = (D >= 128) ? 32 : D / 4;    // lanes per row
  static constexpr int GROUPS = 32 / LANES;                // rows processed side by side in a warp
};

template <int D>
struct Row {
  float4 v[RowCfg<D>::VPL];
};

template <int D>
__device__ __forceinline__ Row<D> row_zero() {
  Row<D> r;
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) r.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  return r;
}

template <int D>
__device__ __forceinline__ Row<D> row_ld(const float *__restrict__ base, int64_t row, int lane) {
  const float4 *p = reinterpret_cast<const float4 *>(base + row * D);
  Row<D> r;
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) r.v[i] = p[i * RowCfg<D>::LANES + lane];
  return r;
}

// read-only path (tables not written by the running kernel)
template <int D>
__device__ __forceinline__ Row<D> row_ldg(const float *__restrict__ base, int64_t row, int lane) {
  const float4 *p = reinterpret_cast<const float4 *>(base + row * D);
  Row<D> r;
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) r.v[i] = __ldg(p + i * RowCfg<D>::LANES + lane);
  return r;
}

template <int D>
__device__ __forceinline__ void row_st(float *__restrict__ base, int64_t row, int lane, const Row<D> &r) {
  float4 *p = reinterpret_cast<float4 *>(base + row * D);
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) p[i * RowCfg<D>::LANES + lane] = r.v[i];
}

// acc += s * (a - b).  The user-row gradient g*vi - g*vj is formed this way: when vi == vj (pos == neg, or equal
// elements) the reference's two products cancel EXACTLY, and so does this; two chained fmas would leave the rounding
// error of g*vi behind (~1e-11), which Adam's normalised step lr * g / (|g| + eps) turns into ~1e-5 of a parameter.
template <int D>
__device__ __forceinline__ void row_fma_diff(Row<D> &acc, float s, const Row<D> &a, const Row<D> &b) {
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) {
    acc.v[i].x = fmaf(s, a.v[i].x - b.v[i].x, acc.v[i].x);
    acc.v[i].y = fmaf(s, a.v[i].y - b.v[i].y, acc.v[i].y);
    acc.v[i].z = fmaf(s, a.v[i].z - b.v[i].z, acc.v[i].z);
    acc.v[i].w = fmaf(s, a.v[i].w - b.v[i].w, acc.v[i].w);
  }
}

// acc += s * a   (fused multiply-add per element)
template <int D>
__device__ __forceinline__ void row_fma(Row<D> &acc, float s, const Row<D> &a) {
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) {
    acc.v[i].x = fmaf(s, a.v[i].x, acc.v[i].x);
    acc.v[i].y = fmaf(s, a.v[i].y, acc.v[i].y);
    acc.v[i].z = fmaf(s, a.v[i].z, acc.v[i].z);
    acc.v[i].w = fmaf(s, a.v[i].w, acc.v[i].w);
  }
}

template <int D>
__device__ __forceinline__ void row_add(Row<D> &acc, const Row<D> &a) {
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) {
    acc.v[i].x += a.v[i].x;
    acc.v[i].y += a.v[i].y;
    acc.v[i].z += a.v[i].z;
    acc.v[i].w += a.v[i].w;
  }
}

template <int D>
__device__ __forceinline__ Row<D> row_scale(float s, const Row<D> &a) {
  Row<D> r;
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i)
    r.v[i] = make_float4(s * a.v[i].x, s * a.v[i].y, s * a.v[i].z, s * a.v[i].w);
  return r;
}

// lanes of one group hold the mask `gmask`; all of them call this together
template <int LANES>
__device__ __forceinline__ float group_sum(float x, unsigned gmask) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) x += __shfl_xor_sync(gmask, x, o);
  return x;
}

// partial dot of two rows on this lane (u . (a - b) style callers pass the difference)
template <int D>
__device__ __forceinline__ float row_dot_lane(const Row<D> &a, const Row<D> &b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) {
    s = fmaf(a.v[i].x, b.v[i].x, s);
    s = fmaf(a.v[i].y, b.v[i].y, s);
    s = fmaf(a.v[i].z, b.v[i].z, s);
    s = fmaf(a.v[i].w, b.v[i].w, s);
  }
  return s;
}

// ---- bulk-copy engine helpers (cp.async.bulk + mbarrier; UBLKCP / SYNCS in SASS) ----------------
__device__ __forceinline__ uint32_t rb2_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rb2_mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rb2_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void rb2_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void rb2_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rb2_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rb2_mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RB2_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RB2_WAIT_DONE;\n"
      "bra RB2_WAIT_LOOP;\n"
      "RB2_WAIT_DONE:\n"
      "}\n" ::"r"(rb2_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size a multiple of 16), completion counted in bytes on `bar`
__device__ __forceinline__ void rb2_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   rb2_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(rb2_smem_u32(bar))
               : "memory");
}

// ---- optimizer arithmetic (torch/optim/adam.py single-tensor path, see oracle/optim.py) -------
struct OptScalars {
  int kind;
  int step;
  float lr, wd, beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt;
  const float *lazy_step_size;  // [step+1] lr/(1-beta1^j), j = 1..step   (RB2_OPT_ADAM_LAZY)
  const float *lazy_bc2_sqrt;   // [step+1] sqrt(1-beta2^j)
};

static inline OptScalars rb2_opt_scalars(const rb2_optim *o) {
  OptScalars s;
  s.kind = o->kind;
  s.step = o->step;
  s.lr = o->lr;
  s.wd = o->weight_decay;
  s.beta1 = o->beta1;
  s.beta2 = o->beta2;
  s.omb1 = o->one_minus_beta1;
  s.omb2 = o->one_minus_beta2;
  s.eps = o->eps;
  s.step_size = o->step_size;
  s.bc2_sqrt = o->bc2_sqrt;
  s.lazy_step_size = o->lazy_step_size;
  s.lazy_bc2_sqrt = o->lazy_bc2_sqrt;
  return s;
}

__device__ __forceinline__ void adam_elem(float &p, float &m, float &v, float g, const OptScalars &o) {
  if (o.wd != 0.f) g = fmaf(o.wd, p, g);                 // grad.add(param, alpha=wd)
  m = fmaf(o.omb1, g - m, m);                            // exp_avg.lerp_(grad, 1-beta1)
  v = fmaf(o.omb2 * g, g, v * o.beta2);                  // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
  float denom = __fdiv_rn(__fsqrt_rn(v), o.bc2_sqrt) + o.eps;
  p = fmaf(-o.step_size, __fdiv_rn(m, denom), p);        // param.addcdiv_(exp_avg, denom, -step_size)
}

__device__ __forceinline__ void sgd_elem(float &p, float g, const OptScalars &o) {
  if (o.wd != 0.f) g = fmaf(o.wd, p, g);
  p = fmaf(-o.lr, g, p);                                 // p.add_(grad, alpha=-lr)
}

// One optimizer step on row `row` of (P, M, V) with summed gradient g; p is the row's value on
// entry (already loaded by the caller).
template <int D>
__device__ __forceinline__ void row_update(float *P, float *M, float *V, int64_t row, int lane, Row<D> p,
                                           const Row<D> &g, const OptScalars &o) {
  if (o.kind == RB2_OPT_SGD) {
#pragma unroll
    for (int i = 0; i < RowCfg<D>::VPL; ++i) {
      sgd_elem(p.v[i].x, g.v[i].x, o);
      sgd_elem(p.v[i].y, g.v[i].y, o);
      sgd_elem(p.v[i].z, g.v[i].z, o);
      sgd_elem(p.v[i].w, g.v[i].w, o);
    }
    row_st<D>(P, row, lane, p);
  } else {
    Row<D> m = row_ld<D>(M, row, lane);
    Row<D> v = row_ld<D>(V, row, lane);
#pragma unroll
    for (int i = 0; i < RowCfg<D>::VPL; ++i) {
      adam_elem(p.v[i].x, m.v[i].x, v.v[i].x, g.v[i].x, o);
      adam_elem(p.v[i].y, m.v[i].y, v.v[i].y, g.v[i].y, o);
      adam_elem(p.v[i].z, m.v[i].z, v.v[i].z, g.v[i].z, o);
      adam_elem(p.v[i].w, m.v[i].w, v.v[i].w, g.v[i].w, o);
    }
    row_st<D>(P, row, lane, p);
    row_st<D>(M, row, lane, m);
    row_st<D>(V, row, lane, v);
  }
}

// The same step on a row whose (p, m, v) are already in registers; the caller stores them.
template <int D>
__device__ __forceinline__ void row_step_regs(Row<D> &p, Row<D> &m, Row<D> &v, const Row<D> &g, const OptScalars &o) {
  if (o.kind == RB2_OPT_SGD) {
#pragma unroll
    for (int i = 0; i < RowCfg<D>::VPL; ++i) {
      sgd_elem(p.v[i].x, g.v[i].x, o);
      sgd_elem(p.v[i].y, g.v[i].y, o);
      sgd_elem(p.v[i].z, g.v[i].z, o);
      sgd_elem(p.v[i].w, g.v[i].w, o);
    }
  } else {
#pragma unroll
    for (int i = 0; i < RowCfg<D>::VPL; ++i) {
      adam_elem(p.v[i].x, m.v[i].x, v.v[i].x, g.v[i].x, o);
      adam_elem(p.v[i].y, m.v[i].y, v.v[i].y, g.v[i].y, o);
      adam_elem(p.v[i].z, m.v[i].z, v.v[i].z, g.v[i].z, o);
      adam_elem(p.v[i].w, m.v[i].w, v.v[i].w, g.v[i].w, o);
    }
  }
}

// ---- RB2_OPT_ADAM_LAZY: replay the zero-gradient steps a row missed ----------------------------
// The reference's dense Adam moves every row at every step, also rows whose gradient is zero
// (their exp_avg keeps decaying into the parameter).  A row last stepped at `last` is brought to
// step `upto` by replaying steps last+1..upto with g = 0 (+ wd*p), which is exactly what dense
// Adam computed for it.  Rows never touched (last == 0, m = v = 0) do not move.
__device__ __forceinline__ void adam_replay_elem(float &p, float &m, float &v, int from, int upto,
                                                 const OptScalars &o) {
  if (m == 0.f && v == 0.f && o.wd == 0.f) return;
  for (int j = from; j <= upto; ++j) {
    float g = (o.wd != 0.f) ? o.wd * p : 0.f;
    m = fmaf(o.omb1, g - m, m);
    v = fmaf(o.omb2 * g, g, v * o.beta2);
    float denom = __fdiv_rn(__fsqrt_rn(v), __ldg(o.lazy_bc2_sqrt + j)) + o.eps;
    p = fmaf(-__ldg(o.lazy_step_size + j), __fdiv_rn(m, denom), p);
  }
}

// The same replay for weight_decay == 0 at one reciprocal per element and step: with g = 0 the moments only decay
// (m_j = beta1^j m, v_j = beta2^j v), so sqrt(v_j) follows by multiplying with sqrt(beta2) and
//     step_size_j * m_j / (sqrt(v_j) / bc2_j + eps)  =  (step_size_j * bc2_j) * m_j / (sqrt(v_j) + eps * bc2_j).
// Per-step relative error ~1e-7 of an update that is itself <= lr per step: far inside the 1e-5 parity bound.
__device__ __forceinline__ void adam_replay4_fast(float4 &p, float4 &m, float4 &v, int from, int upto,
                                                  const OptScalars &o, float sqrt_beta2) {
  float sx = sqrtf(v.x), sy = sqrtf(v.y), sz = sqrtf(v.z), sw = sqrtf(v.w);
  for (int j = from; j <= upto; ++j) {
    const float bc = __ldg(o.lazy_bc2_sqrt + j);
    const float a = -__ldg(o.lazy_step_size + j) * bc, e = o.eps * bc;
    m.x *= o.beta1; m.y *= o.beta1; m.z *= o.beta1; m.w *= o.beta1;
    v.x *= o.beta2; v.y *= o.beta2; v.z *= o.beta2; v.w *= o.beta2;
    sx *= sqrt_beta2; sy *= sqrt_beta2; sz *= sqrt_beta2; sw *= sqrt_beta2;
    p.x = fmaf(a * m.x, __frcp_rn(sx + e), p.x);
    p.y = fmaf(a * m.y, __frcp_rn(sy + e), p.y);
    p.z = fmaf(a * m.z, __frcp_rn(sz + e), p.z);
    p.w = fmaf(a * m.w, __frcp_rn(sw + e), p.w);
  }
}

template <int D>
__device__ __forceinline__ void row_replay_fast(Row<D> &p, Row<D> &m, Row<D> &v, int last, int upto,
                                                const OptScalars &o, float sqrt_beta2) {
  if (last >= upto || last == 0) return;     // (wd == 0: a row never stepped has zero moments and does not move)
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) adam_replay4_fast(p.v[i], m.v[i], v.v[i], last + 1, upto, o, sqrt_beta2);
}

template <int D>
__device__ __forceinline__ void row_replay(Row<D> &p, Row<D> &m, Row<D> &v, int last, int upto,
                                           const OptScalars &o) {
  if (last >= upto) return;
  if (last == 0 && o.wd == 0.f) return;  // never touched: zero state, nothing moves
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) {
    adam_replay_elem(p.v[i].x, m.v[i].x, v.v[i].x, last + 1, upto, o);
    adam_replay_elem(p.v[i].y, m.v[i].y, v.v[i].y, last + 1, upto, o);
    adam_replay_elem(p.v[i].z, m.v[i].z, v.v[i].z, last + 1, upto, o);
    adam_replay_elem(p.v[i].w, m.v[i].w, v.v[i].w, last + 1, upto, o);
  }
}

// value of row `row` as the reference's dense optimizer would hold it after step o.step-1
template <int D, bool LAZY>
__device__ __forceinline__ Row<D> row_ld_effective(const float *P, const float *M, const float *V,
                                                   const int32_t *LAST, int64_t row, int lane,
                                                   const OptScalars &o) {
  Row<D> p = row_ld<D>(P, row, lane);
  if (LAZY) {
    int last = LAST[row];
    if (last < o.step - 1 && (last > 0 || o.wd != 0.f)) {
      Row<D> m = row_ld<D>(M, row, lane);
      Row<D> v = row_ld<D>(V, row, lane);
      row_replay<D>(p, m, v, last, o.step - 1, o);
    }
  }
  return p;
}

// one optimizer step with catch-up: (p, m, v) are brought to step-1 first when LAZY
template <int D, bool LAZY>
__device__ __forceinline__ void row_update_full(float *P, float *M, float *V, int32_t *LAST, int64_t row,
                                                int lane, const Row<D> &g, const OptScalars &o) {
  if (!LAZY) {
    Row<D> p = row_ld<D>(P, row, lane);
    row_update<D>(P, M, V, row, lane, p, g, o);
    return;
  }
  Row<D> p = row_ld<D>(P, row, lane);
  Row<D> m = row_ld<D>(M, row, lane);
  Row<D> v = row_ld<D>(V, row, lane);
  int last = LAST[row];
  row_replay<D>(p, m, v, last, o.step - 1, o);
#pragma unroll
  for (int i = 0; i < RowCfg<D>::VPL; ++i) {
    adam_elem(p.v[i].x, m.v[i].x, v.v[i].x, g.v[i].x, o);
    adam_elem(p.v[i].y, m.v[i].y, v.v[i].y, g.v[i].y, o);
    adam_elem(p.v[i].z, m.v[i].z, v.v[i].z, g.v[i].z, o);
    adam_elem(p.v[i].w, m.v[i].w, v.v[i].w, g.v[i].w, o);
  }
  row_st<D>(P, row, lane, p);
  row_st<D>(M, row, lane, m);
  row_st<D>(V, row, lane, v);
  if (lane == 0) LAST[row] = o.step;
}
