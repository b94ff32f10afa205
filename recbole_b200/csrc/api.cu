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
extern "C" int rb2_abi_version(void) { return RB2_ABI_VERSION; }

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
