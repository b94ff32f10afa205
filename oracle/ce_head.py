"""SASRec-style full-sort CE head restated on CPU (TEST INFRASTRUCTURE).

Reference: ``SASRec.calculate_loss`` CE branch
recbole/model/sequential_recommender/sasrec.py:137-141
(``logits = seq_output @ item_emb.weight.T ; nn.CrossEntropyLoss()(logits, pos)``)
and ``SASRec.full_sort_predict`` sasrec.py:152-158; sequential full-sort batches carry no
history mask and one positive (recbole/data/dataloader/sequential_dataloader.py:269-280),
``Trainer._full_sort_batch_eval`` still masks column 0 (trainer.py:343).
"""
import numpy as np

from ._clib import lib as _clib

F32 = np.float32


def logits(X, E):
    return _clib.scores_fma(X, E)


def ce_loss(X, E, target):
    """mean over rows of logsumexp(logits) - logits[target]  (fp32, max-shifted)."""
    L = logits(X, E)
    mx = L.max(axis=1)
    lse = mx + np.log(np.exp(L - mx[:, None], dtype=F32).sum(axis=1, dtype=F32), dtype=F32)
    per_row = lse - L[np.arange(len(target)), target]
    return F32(per_row.mean(dtype=F32)), lse.astype(F32)


def full_sort_topk(X, E, k):
    nu = X.shape[0]
    return _clib.fullsort_topk(X, E, np.arange(nu), np.zeros(nu + 1, np.int64), np.zeros(0, np.int64), k)


def ce_backward(X, E, target, grad_scale=None):
    """Gradients of mean CE w.r.t. X and E (autograd of sasrec.py:137-141), float64 ground truth:
    G = (softmax(X E^T) - onehot(target)) * grad_scale; dX = G E; dE = G^T X."""
    X64, E64 = X.astype(np.float64), E.astype(np.float64)
    L = X64 @ E64.T
    L -= L.max(axis=1, keepdims=True)
    P = np.exp(L)
    P /= P.sum(axis=1, keepdims=True)
    P[np.arange(len(target)), target] -= 1.0
    P *= (1.0 / X.shape[0]) if grad_scale is None else grad_scale
    return (P @ E64).astype(F32), (P.T @ X64).astype(F32)
