/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * CPU restatement, in plain C, of the two pieces of the full-sort path whose
 * result depends on floating-point evaluation order:
 *
 *   - the score  S[u,i] = <U[u,:], V[i,:]>   (reference: torch.matmul in
 *     recbole/model/general_recommender/bpr.py:91-96 and
 *     recbole/model/sequential_recommender/sasrec.py:137-141,152-158)
 *   - the masked top-K selection over one score row (reference:
 *     recbole/trainer/trainer.py:342-345 mask + recbole/evaluator/evaluators.py:68-72 topk)
 *
 * The reference leaves the fp32 summation order to the BLAS it links and the
 * tie order to torch.topk ("unspecified").  BASELINE.json:north_star fixes the
 * tie rule (lower item id first); this file fixes the summation order as the
 * canonical one used by every parity test:
 *
 *      s = 0.0f;  for k = 0 .. d-1:  s = fmaf(u[k], v[k], s)
 *
 * i.e. one correctly-rounded fp32 fused multiply-add per k, ascending k.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may call this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* out[u*ni + i] = fma-chain dot of U[u,:] and V[i,:]  (row-major fp32) */
void oracle_scores_fma(const float *U, const float *V, int64_t nu, int64_t ni,
                       int64_t d, float *out) {
  for (int64_t u = 0; u < nu; ++u) {
    const float *ur = U + u * d;
    for (int64_t i = 0; i < ni; ++i) {
      const float *vr = V + i * d;
      float s = 0.0f;
      for (int64_t k = 0; k < d; ++k) s = fmaf(ur[k], vr[k], s);
      out[u * ni + i] = s;
    }
  }
}

/* pairwise gather-dot, same chain (reference: bpr.py:85-89 predict) */
void oracle_pair_scores_fma(const float *U, const float *V, const int64_t *uid,
                            const int64_t *iid, int64_t n, int64_t d, float *out) {
  for (int64_t j = 0; j < n; ++j) {
    const float *ur = U + uid[j] * d, *vr = V + iid[j] * d;
    float s = 0.0f;
    for (int64_t k = 0; k < d; ++k) s = fmaf(ur[k], vr[k], s);
    out[j] = s;
  }
}

/*
 * Masked top-K of one row of scores.  Candidates are all items except id 0
 * (trainer.py:343) and the sorted `hist` ids (trainer.py:344-345); order is
 * score descending, ties by ascending item id (north_star).  NaN scores are
 * never selected.  Slots beyond the number of candidates get id -1, score -inf.
 */
void oracle_topk_row(const float *scores, int64_t ni, const int64_t *hist,
                     int64_t nhist, int k, int64_t *out_ids, float *out_scores) {
  for (int j = 0; j < k; ++j) { out_ids[j] = -1; out_scores[j] = -INFINITY; }
  int64_t h = 0;
  for (int64_t i = 1; i < ni; ++i) {
    while (h < nhist && hist[h] < i) ++h;
    if (h < nhist && hist[h] == i) continue;
    float s = scores[i];
    if (!(s == s)) continue;
    if (s == -INFINITY) continue;
    /* strictly better than the current worst, or a free slot: ids arrive in
       ascending order so an equal score never displaces an earlier id */
    if (out_ids[k - 1] >= 0 && !(s > out_scores[k - 1])) continue;
    int j = k - 1;
    while (j > 0 && (out_ids[j - 1] < 0 || s > out_scores[j - 1])) {
      out_ids[j] = out_ids[j - 1];
      out_scores[j] = out_scores[j - 1];
      --j;
    }
    out_ids[j] = i;
    out_scores[j] = s;
  }
}

/* batched: hist is a CSR over the rows of `scores` */
void oracle_topk(const float *scores, int64_t nu, int64_t ni,
                 const int64_t *hist_indptr, const int64_t *hist_indices, int k,
                 int64_t *out_ids, float *out_scores) {
  for (int64_t u = 0; u < nu; ++u)
    oracle_topk_row(scores + u * ni, ni, hist_indices + hist_indptr[u],
                    hist_indptr[u + 1] - hist_indptr[u], k, out_ids + u * k,
                    out_scores + u * k);
}

/* fused score+topk for rows too large to materialise (bench cpu leg) */
void oracle_fullsort_topk(const float *U, const float *V, const int64_t *users,
                          int64_t nu, int64_t ni, int64_t d,
                          const int64_t *hist_indptr, const int64_t *hist_indices,
                          int k, int64_t *out_ids, float *out_scores) {
  float *row = (float *)malloc(sizeof(float) * (size_t)ni);
  for (int64_t u = 0; u < nu; ++u) {
    oracle_scores_fma(U + users[u] * d, V, 1, ni, d, row);
    oracle_topk_row(row, ni, hist_indices + hist_indptr[u],
                    hist_indptr[u + 1] - hist_indptr[u], k, out_ids + u * k,
                    out_scores + u * k);
  }
  free(row);
}
