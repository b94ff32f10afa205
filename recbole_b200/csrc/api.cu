// api.cu -- error string + ABI version + gather-dot (BPR.predict, bpr.py:85-89).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void rb2_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *rb2_last_error(void) { return g_err; }

// ---- profiling hooks ---------------------------------------------------------------------------
#include <mutex>
#include <vector>
namespace {
struct StageProf {
  std::vector<cudaEvent_t> beg, end;  // pool, grown on demand
  size_t used = 0;
  int64_t calls = 0, launches = 0;
};
bool g_prof_on = false;
StageProf g_prof[RB2_NUM_STAGES];
std::mutex g_prof_mu;
}  // namespace

void rb2_prof_begin(int stage, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  StageProf &p = g_prof[stage];
  if (p.used == p.beg.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    p.beg.push_back(a);
    p.end.push_back(b);
  }
  cudaEventRecord(p.beg[p.used], st);
}

void rb2_prof_end(int stage, cudaStream_t st, int launches) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  StageProf &p = g_prof[stage];
  cudaEventRecord(p.end[p.used], st);
  p.used++;
  p.calls++;
  p.launches += launches;
}

extern "C" int rb2_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  return 0;
}

extern "C" int rb2_profile_read(float *h_ms, int64_t *h_calls, int64_t *h_launches) {
  RB2_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int s = 0; s < RB2_NUM_STAGES; ++s) {
    StageProf &p = g_prof[s];
    float tot = 0.f;
    for (size_t i = 0; i < p.used; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, p.beg[i], p.end[i]) == cudaSuccess) tot += ms;
    }
    if (h_ms) h_ms[s] = tot;
    if (h_calls) h_calls[s] = p.calls;
    if (h_launches) h_launches[s] = p.launches;
    p.used = 0;
    p.calls = 0;
    p.launches = 0;
  }
  return 0;
}
extern "C" int rb2_abi_version(void) { return RB2_ABI_VERSION; }

// ---- peer mapping (cudaIpc) ------------------------------------------------------------------------------
#include <dlfcn.h>
#include <map>
#include <string>
namespace {
std::mutex g_ipc_mu;
std::map<std::string, void *> g_ipc_open;   // handle bytes -> mapped base (one mapping per allocation)

// base address of the allocation containing p (driver API, resolved at run time: no link dependency)
int alloc_base(const void *p, void **base) {
  typedef int (*fn_t)(unsigned long long *, size_t *, unsigned long long);
  static fn_t fn = nullptr;
  if (!fn) {
    void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (h) fn = (fn_t)dlsym(h, "cuMemGetAddressRange_v2");
  }
  RB2_REQUIRE(fn != nullptr, RB2_EINVAL, "rb2_ipc_export: cuMemGetAddressRange_v2 not found in libcuda.so.1");
  unsigned long long b = 0;
  size_t sz = 0;
  int rc = fn(&b, &sz, (unsigned long long)(uintptr_t)p);
  RB2_REQUIRE(rc == 0, RB2_EINVAL, "rb2_ipc_export: cuMemGetAddressRange failed (%d)", rc);
  *base = (void *)(uintptr_t)b;
  return 0;
}
}  // namespace

extern "C" int rb2_ipc_export(const void *dev_ptr, void *h_handle, int64_t *h_offset) {
  RB2_REQUIRE(dev_ptr && h_handle && h_offset, RB2_EINVAL, "rb2_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == RB2_IPC_HANDLE_BYTES, "handle size");
  void *base = nullptr;
  int rc = alloc_base(dev_ptr, &base);
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, base);
  if (e != cudaSuccess) {
    rb2_set_error("rb2_ipc_export: cudaIpcGetMemHandle failed: %s (memory from cudaMallocAsync / expandable segments "
                  "cannot be exported; unset PYTORCH_CUDA_ALLOC_CONF=expandable_segments)", cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  memcpy(h_handle, &h, sizeof(h));
  *h_offset = (int64_t)((const char *)dev_ptr - (const char *)base);
  return 0;
}

extern "C" int rb2_ipc_open(const void *h_handle, int64_t offset, void **h_mapped) {
  RB2_REQUIRE(h_handle && h_mapped && offset >= 0, RB2_EINVAL, "rb2_ipc_open: bad argument");
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  std::string key((const char *)h_handle, RB2_IPC_HANDLE_BYTES);
  auto it = g_ipc_open.find(key);
  void *base = nullptr;
  if (it != g_ipc_open.end()) {
    base = it->second;
  } else {
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      rb2_set_error("rb2_ipc_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return (int)e;
    }
    g_ipc_open[key] = base;
  }
  *h_mapped = (char *)base + offset;
  return 0;
}

extern "C" int rb2_ipc_close_all(void) {
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  for (auto &kv : g_ipc_open) cudaIpcCloseMemHandle(kv.second);
  g_ipc_open.clear();
  return 0;
}

namespace {
// One lane-group per (user, item) pair: two coalesced row gathers and a shuffle reduction.
template <int D>
__global__ void __launch_bounds__(256) k_gather_dot(const float *__restrict__ up, const float *__restrict__ ip,
                                                     const int64_t *__restrict__ user,
                                                     const int64_t *__restrict__ item, int64_t n, int64_t n_users,
                                                     int64_t n_items, float *__restrict__ out) {
  constexpr int LANES = RowCfg<D>::LANES;
  const int lane = threadIdx.x % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));
  int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  int64_t ngroups = (int64_t)gridDim.x * blockDim.x / LANES;
  for (int64_t j = gid; j < n; j += ngroups) {
    int64_t u = min(max(user[j], (int64_t)0), n_users - 1);
    int64_t i = min(max(item[j], (int64_t)0), n_items - 1);
    float s = group_sum<LANES>(row_dot_lane<D>(row_ldg<D>(up, u, lane), row_ldg<D>(ip, i, lane)), gmask);
    if (lane == 0) out[j] = s;
  }
}
}  // namespace

extern "C" int rb2_gather_dot(const float *user_p, const float *item_p, int64_t n_users, int64_t n_items,
                              int32_t dim, const int64_t *user, const int64_t *item, int64_t n, float *out,
                              void *stream) {
  RB2_REQUIRE(user_p && item_p && user && item && out, RB2_EINVAL, "rb2_gather_dot: null argument");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(RB2_ST_GATHER_DOT, st);
#define RB2_GD(D_)                                                                                     \
  {                                                                                                    \
    constexpr int LANES = RowCfg<D_>::LANES;                                                           \
    int64_t want = (n * LANES + 255) / 256;                                                            \
    unsigned blocks = (unsigned)(want < (int64_t)rb2_num_sms() * 16 ? want : (int64_t)rb2_num_sms() * 16); \
    k_gather_dot<D_><<<blocks, 256, 0, st>>>(user_p, item_p, user, item, n, n_users, n_items, out);   \
  }
  switch (dim) {
    case 16: RB2_GD(16) break;
    case 32: RB2_GD(32) break;
    case 64: RB2_GD(64) break;
    case 128: RB2_GD(128) break;
    case 256: RB2_GD(256) break;
    default:
      rb2_set_error("rb2_gather_dot: embedding dim %d not supported (16, 32, 64, 128, 256)", (int)dim);
      return RB2_EINVAL;
  }
#undef RB2_GD
  RB2_CUDA(cudaGetLastError());
  return 0;
}


// ---- scorer state plumbing -------------------------------------------------------------------------------------
namespace {
thread_local rb2_scorer_state t_default_scorer = {};
thread_local rb2_scorer_state *t_cur_scorer = nullptr;
}  // namespace
rb2_scorer_state &rb2_cur_scorer() { return t_cur_scorer ? *t_cur_scorer : t_default_scorer; }
ScorerScope::ScorerScope(rb2_scorer_state *s) : prev(t_cur_scorer) { if (s) t_cur_scorer = s; }
ScorerScope::~ScorerScope() { t_cur_scorer = prev; }
