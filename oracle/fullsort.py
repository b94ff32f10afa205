"""Full-sort evaluation restated on CPU (TEST INFRASTRUCTURE).

Reference path:
  * ``BPR.full_sort_predict``          recbole/model/general_recommender/bpr.py:91-96
  * ``Trainer._full_sort_batch_eval``  recbole/trainer/trainer.py:328-352   (pad/history mask, swap)
  * ``GeneralFullDataLoader``          recbole/data/dataloader/general_dataloader.py:294-364
  * ``Sampler.get_used_ids``           recbole/sampler/sampler.py:206-227
  * ``TopKEvaluator.collect/evaluate`` recbole/evaluator/evaluators.py:53-141

Restated semantics (verified against the reference in
tests/test_oracle_golden.py): for each evaluated user, candidates are all items
except id 0 and except ``used_ids[phase][u] - positives[phase][u]``; rank by
score descending, ties by ascending item id; ``hit[k] = k-th item in positives``.
"""
import numpy as np

from . import metrics as _metrics
from ._clib import lib as _clib


# ---- index construction (general_dataloader.py:294-328, sampler.py:206-227) --------------------

def build_csr(n_rows, rows, cols):
    """Sorted, de-duplicated CSR (indptr int64[n_rows+1], indices int64) of (row, col) pairs."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    if len(rows):
        key = np.unique(rows * (int(cols.max()) + 1) + cols)
        base = int(cols.max()) + 1
        rows, cols = key // base, key % base
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    return np.cumsum(indptr), cols.astype(np.int64)


def eval_index(n_users, phase_inters, phase):
    """History / positives CSRs for evaluating ``phase`` (index into phase_inters).

    ``phase_inters`` is a list of (user_ids, item_ids) per phase in order
    (train, valid, test).  used_ids[phase] = union of phases 0..phase
    (sampler.py:213-218); positives = this phase's items; history = used - positives
    (general_dataloader.py:319-321).  Returns (uid_list, hist_csr, pos_csr) with the
    CSRs indexed by position in uid_list (users with >= 1 positive, ascending id,
    general_dataloader.py:299-313).
    """
    pu, pi = (np.asarray(a, dtype=np.int64) for a in phase_inters[phase])
    uid_list = np.unique(pu)
    remap = -np.ones(n_users, dtype=np.int64)
    remap[uid_list] = np.arange(len(uid_list))
    pos = build_csr(len(uid_list), remap[pu], pi)
    hu = np.concatenate([np.asarray(phase_inters[p][0], dtype=np.int64) for p in range(phase + 1)])
    hi = np.concatenate([np.asarray(phase_inters[p][1], dtype=np.int64) for p in range(phase + 1)])
    keep = remap[hu] >= 0
    used = build_csr(len(uid_list), remap[hu[keep]], hi[keep])
    # history = used - positives
    hp, hidx = [0], []
    for r in range(len(uid_list)):
        u = used[1][used[0][r]:used[0][r + 1]]
        p = pos[1][pos[0][r]:pos[0][r + 1]]
        h = np.setdiff1d(u, p, assume_unique=True)
        hidx.append(h)
        hp.append(hp[-1] + len(h))
    hist = (np.asarray(hp, dtype=np.int64), np.concatenate(hidx).astype(np.int64) if hidx else np.zeros(0, np.int64))
    return uid_list, hist, pos


def reference_swap_index(positives, n_pos=None):
    """``_set_user_property`` general_dataloader.py:319-328: (swap_idx, rev_swap_idx)."""
    positives = set(int(x) for x in positives)
    p = len(positives) if n_pos is None else n_pos
    swap = np.array(sorted(set(range(p)) ^ positives), dtype=np.int64)
    return swap, swap[::-1].copy()


# ---- the reference's own matrix formulation (trainer.py:342-350, evaluators.py:68-75) ----------

def reference_batch_eval(scores, hist_rows, hist_cols, swap_row, swap_after, swap_before):
    """trainer.py:342-350 on a numpy [users, N] matrix (copy)."""
    s = np.array(scores, dtype=np.float32, copy=True)
    s[:, 0] = -np.inf
    if len(hist_rows):
        s[hist_rows, hist_cols] = -np.inf
    s[swap_row, swap_after] = s[swap_row, swap_before]
    return s


def reference_topk_idx(topk_ids, pos_indptr, pos_indices, n_items):
    """Map top-K ITEM IDS to the reference's ``topk_idx`` coordinates (SURVEY.md 8a end).

    The reference moves the positives into columns 0..p-1 (swap) and flips the
    columns, so that "is a positive" becomes ``idx >= N - p`` (evaluators.py:134).
    For item t: idx = N-1-col(t); col(t) = t if untouched by the swap, else the
    swap partner (general_dataloader.py:325-327).  Slots with id -1 (fewer than K
    candidates) map to the flipped pad column, N-1-col(0).
    """
    topk_ids = np.asarray(topk_ids, dtype=np.int64)
    out = np.empty((topk_ids.shape[0], topk_ids.shape[1] + 1), dtype=np.int64)
    out[:, -1] = n_items
    for r in range(topk_ids.shape[0]):
        pos = pos_indices[pos_indptr[r]:pos_indptr[r + 1]]
        swap, rev = reference_swap_index(pos)
        col = {int(b): int(a) for a, b in zip(swap, rev)}  # new[a] = old[b]  => item b sits in column a
        for j, t in enumerate(topk_ids[r]):
            t = int(t) if t >= 0 else 0
            out[r, j] = n_items - 1 - col.get(t, t)
    return out


# ---- restated semantics ---------------------------------------------------------------------

def full_sort_topk(U, V, users, hist_indptr, hist_indices, k, scores=None):
    """Top-k item ids / canonical scores per user.  ``scores`` may be supplied (e.g. the
    reference's own torch.matmul output) to isolate the selection from the dot product."""
    if scores is None:
        return _clib.fullsort_topk(U, V, users, hist_indptr, hist_indices, k)
    return _clib.topk(scores, hist_indptr, hist_indices, k)


def hits(topk_ids, pos_indptr, pos_indices):
    """bool[U, K]: k-th recommended item is one of the user's positives (evaluators.py:134)."""
    topk_ids = np.asarray(topk_ids)
    out = np.zeros(topk_ids.shape, dtype=bool)
    for r in range(topk_ids.shape[0]):
        out[r] = np.isin(topk_ids[r], pos_indices[pos_indptr[r]:pos_indptr[r + 1]])
    return out


def evaluate(topk_ids, pos_indptr, pos_indices, metrics, topk, precision=4):
    pos_len = np.diff(pos_indptr)
    return _metrics.evaluate(hits(topk_ids, pos_indptr, pos_indices), pos_len, metrics, topk, precision)
