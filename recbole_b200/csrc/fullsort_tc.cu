// fullsort_tc.cu -- tcgen05/TMA tensor-core full-sort scorer (placeholder until the kernel lands).
#include "common.cuh"

size_t rb2_fullsort_tc_workspace_bytes(int64_t, int64_t, int32_t, int32_t) { return 256; }

int rb2_fullsort_tc(const float *, const int64_t *, int64_t, const float *, int64_t, int64_t, int32_t,
                    const int64_t *, const int64_t *, int32_t, int64_t *, float *, void *, size_t, cudaStream_t) {
  rb2_set_error("rb2_fullsort_topk: RB2_SCORER_TC is not built into this library yet");
  return RB2_EINVAL;
}
