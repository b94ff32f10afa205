// sampler.cu -- negative sampler that rejects used ids from a device-resident CSR (sm_100a).
//
// Replaces Sampler.sample_by_user_ids -> AbstractSampler.sample_by_key_ids
// (recbole/sampler/sampler.py:103-154,246-265), whose inner loop is a Python list comprehension
// with per-user `set` lookups.
//
// rb2_neg_sample_ref reproduces the reference's stream exactly: slot j of round r takes
// random_list[(pr + j) % L] where j is the slot's rank among the slots still pending (in slot
// order), rejected slots stay pending for the next round and the pointer advances by the number
// of pending slots (sampler.py:82-101,144-153).  The rank is an order-preserving compaction
// (cub::DeviceSelect::Flagged); the pending count comes back to the host once per round, exactly
// the reference's `while len(check_list) > 0`.
// rb2_neg_sample_hash is the device-resident variant: one launch, splitmix64 counter stream
// (oracle/sampler.py:hash_sample defines it bit for bit).
#include <cub/device/device_select.cuh>

#include "common.cuh"

namespace {

__device__ __forceinline__ bool used_contains(const int64_t *__restrict__ indptr, const int64_t *__restrict__ idx,
                                              int64_t row, int64_t x) {
  int64_t lo = indptr[row], hi = indptr[row + 1];
  const int64_t end = hi;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (idx[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo < end && idx[lo] == x;
}

// round kernel: pending[j] (slot index) draws random_list[(pr + j) % L]; flag = still rejected
__global__ void k_ref_round(const int64_t *__restrict__ key_ids, int64_t n_keys, const int64_t *__restrict__ rl,
                            int64_t L, int64_t pr, const uint32_t *__restrict__ pending, int64_t n_pending,
                            const int64_t *__restrict__ indptr, const int64_t *__restrict__ idx, int64_t n_rows,
                            int64_t *__restrict__ out, uint8_t *__restrict__ flag, WsHeader *hdr) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pending) return;
  uint32_t slot = pending ? pending[j] : (uint32_t)j;
  int64_t v = rl[(pr + j) % L];
  int64_t key = key_ids[slot % n_keys];
  if (key < 0 || key >= n_rows) {  // reference: IndexError -> ValueError (sampler.py:260-265)
    hdr->range_error = 1;
    out[slot] = v;
    flag[j] = 0;
    return;
  }
  out[slot] = v;
  flag[j] = used_contains(indptr, idx, key, v) ? 1 : 0;
}

__global__ void k_iota(uint32_t *a, int64_t n) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) a[j] = (uint32_t)j;
}

// ---- counter-based stream (oracle/sampler.py) --------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t step, uint64_t slot, uint64_t attempt) {
  uint64_t x = seed + 0x9E3779B97F4A7C15ull * (slot + 1ull);
  x ^= (attempt + 1ull) * 0xBF58476D1CE4E5B9ull;
  x += step * 0x94D049BB133111EBull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

constexpr int kMaxAttempts = 64;

__global__ void k_hash_sample(const int64_t *__restrict__ key_ids, int64_t n_keys, int64_t total, int64_t n_items,
                              const int64_t *__restrict__ indptr, const int64_t *__restrict__ idx, int64_t n_rows,
                              uint64_t seed, uint64_t step, int64_t *__restrict__ out) {
  int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= total) return;
  int64_t key = key_ids[slot % n_keys];
  key = min(max(key, (int64_t)0), n_rows - 1);
  int64_t v = 0;
  bool ok = false;
  for (int a = 0; a < kMaxAttempts && !ok; ++a) {
    v = 1 + (int64_t)__umul64hi(mix64(seed, step, (uint64_t)slot, (uint64_t)a), (uint64_t)(n_items - 1));
    ok = !used_contains(indptr, idx, key, v);
  }
  for (int64_t tries = 0; !ok && tries < n_items - 1; ++tries) {  // users who used almost everything
    v = (v + 1 < n_items) ? v + 1 : 1;
    ok = !used_contains(indptr, idx, key, v);
  }
  out[slot] = v;
}

struct SamplerWs {
  WsHeader *hdr;
  uint32_t *pend_a, *pend_b;
  uint8_t *flag;
  int64_t *count;
  void *cub_tmp;
  size_t cub_bytes;
};

size_t carve(SamplerWs &w, void *base, int64_t total) {
  Carver c(base);
  w.hdr = c.take<WsHeader>(1);
  w.pend_a = c.take<uint32_t>(total);
  w.pend_b = c.take<uint32_t>(total);
  w.flag = c.take<uint8_t>(total);
  w.count = c.take<int64_t>(1);
  size_t b = 0;
  cub::DeviceSelect::Flagged(nullptr, b, (uint32_t *)nullptr, (uint8_t *)nullptr, (uint32_t *)nullptr,
                             (int64_t *)nullptr, (int)total);
  w.cub_bytes = b;
  w.cub_tmp = c.take<char>(b);
  return c.off;
}

}  // namespace

extern "C" size_t rb2_neg_sample_workspace_bytes(int64_t n_keys, int32_t num) {
  SamplerWs w;
  return carve(w, nullptr, n_keys * num);
}

extern "C" int rb2_neg_sample_ref(const int64_t *key_ids, int64_t n_keys, int32_t num, const int64_t *random_list,
                                  int64_t random_list_length, int64_t *h_random_pr, const int64_t *used_indptr,
                                  const int64_t *used_indices, int64_t n_rows, int64_t *out, void *workspace,
                                  size_t workspace_bytes, void *stream) {
  RB2_REQUIRE(key_ids && random_list && h_random_pr && used_indptr && used_indices && out && workspace, RB2_EINVAL,
              "rb2_neg_sample_ref: null argument");
  RB2_REQUIRE(random_list_length > 0, RB2_EINVAL, "rb2_neg_sample_ref: empty random_list");
  const int64_t total = n_keys * (int64_t)num;
  if (total <= 0) return 0;
  RB2_REQUIRE(total < ((int64_t)1 << 31), RB2_EINVAL, "rb2_neg_sample_ref: too many slots");
  SamplerWs w;
  size_t need = carve(w, workspace, total);
  RB2_REQUIRE(workspace_bytes >= need, RB2_EWORKSPACE, "rb2_neg_sample_ref: workspace %zu < %zu", workspace_bytes,
              need);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t pr = *h_random_pr % random_list_length;  // sampler.py:91
  int64_t n_pending = total;
  const uint32_t *pending = nullptr;  // round 0: identity
  uint32_t *next = w.pend_a;
  k_iota<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w.pend_b, total);
  const uint32_t *cur_list = w.pend_b;
  (void)pending;
  int rounds = 0;
  while (n_pending > 0) {
    k_ref_round<<<(unsigned)((n_pending + 255) / 256), 256, 0, st>>>(key_ids, n_keys, random_list,
                                                                    random_list_length, pr, cur_list, n_pending,
                                                                    used_indptr, used_indices, n_rows, out, w.flag,
                                                                    w.hdr);
    size_t tmp = w.cub_bytes;
    RB2_CUDA(cub::DeviceSelect::Flagged(w.cub_tmp, tmp, cur_list, w.flag, next, w.count, (int)n_pending, st));
    int64_t h_count = 0;
    RB2_CUDA(cudaMemcpyAsync(&h_count, w.count, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    RB2_CUDA(cudaStreamSynchronize(st));
    pr = (pr + n_pending) % random_list_length;  // random_num advances by the number drawn
    n_pending = h_count;
    const uint32_t *t = cur_list;
    cur_list = next;
    next = const_cast<uint32_t *>(t);
    if (++rounds > 100000) {
      rb2_set_error("rb2_neg_sample_ref: no progress (a key has used every item)");
      return RB2_EINVAL;
    }
  }
  // the reference leaves random_pr un-wrapped after the last draw (sampler.py:93-94) only when the
  // draw did not wrap; both forms are congruent mod L and the next call reduces mod L first.
  *h_random_pr = pr;
  RB2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int rb2_neg_sample_hash(const int64_t *key_ids, int64_t n_keys, int32_t num, int64_t n_items,
                                   const int64_t *used_indptr, const int64_t *used_indices, int64_t n_rows,
                                   uint64_t seed, uint64_t step, int64_t *out, void *stream) {
  RB2_REQUIRE(key_ids && used_indptr && used_indices && out, RB2_EINVAL, "rb2_neg_sample_hash: null argument");
  RB2_REQUIRE(n_items >= 2 && n_items < ((int64_t)1 << 32), RB2_EINVAL, "rb2_neg_sample_hash: n_items out of range");
  const int64_t total = n_keys * (int64_t)num;
  if (total <= 0) return 0;
  k_hash_sample<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      key_ids, n_keys, total, n_items, used_indptr, used_indices, n_rows, seed, step, out);
  RB2_CUDA(cudaGetLastError());
  return 0;
}
