"""Top-K metric arithmetic restated in numpy float64 (TEST INFRASTRUCTURE).

Follows recbole/evaluator/metrics.py:27-164 (hit_/mrr_/map_/recall_/ndcg_/precision_)
and ``TopKEvaluator.evaluate/_calculate_metrics`` recbole/evaluator/evaluators.py:78-105,122-141.
The reference's ``np.float`` alias is float64.  All functions take ``pos_index``
bool[U, K] (hit at rank k) and ``pos_len`` int[U] and return float64[U, K].
"""
import numpy as np


def hit_(pos_index, pos_len):  # metrics.py:27-42
    return (np.cumsum(pos_index, axis=1) > 0).astype(int)


def mrr_(pos_index, pos_len):  # metrics.py:45-65
    idxs = pos_index.argmax(axis=1)
    result = np.zeros(pos_index.shape, dtype=np.float64)
    for row, idx in enumerate(idxs):
        result[row, idx:] = 1 / (idx + 1) if pos_index[row, idx] > 0 else 0
    return result


def precision_(pos_index, pos_len):  # metrics.py:149-164
    return pos_index.cumsum(axis=1) / np.arange(1, pos_index.shape[1] + 1)


def recall_(pos_index, pos_len):  # metrics.py:96-110
    return np.cumsum(pos_index, axis=1) / pos_len.reshape(-1, 1)


def map_(pos_index, pos_len):  # metrics.py:68-93
    pre = precision_(pos_index, pos_len)
    sum_pre = np.cumsum(pre * pos_index.astype(np.float64), axis=1)
    K = pos_index.shape[1]
    actual_len = np.minimum(pos_len, K)
    result = np.zeros(pos_index.shape, dtype=np.float64)
    for row, lens in enumerate(actual_len):
        ranges = np.arange(1, K + 1)
        ranges[lens:] = ranges[lens - 1]
        result[row] = sum_pre[row] / ranges
    return result


def ndcg_(pos_index, pos_len):  # metrics.py:113-146
    K = pos_index.shape[1]
    idcg_len = np.minimum(pos_len, K)
    iranks = np.zeros(pos_index.shape, dtype=np.float64)
    iranks[:, :] = np.arange(1, K + 1)
    idcg = np.cumsum(1.0 / np.log2(iranks + 1), axis=1)
    for row, idx in enumerate(idcg_len):
        idcg[row, idx:] = idcg[row, idx - 1]
    ranks = np.zeros(pos_index.shape, dtype=np.float64)
    ranks[:, :] = np.arange(1, K + 1)
    dcg = 1.0 / np.log2(ranks + 1)
    dcg = np.cumsum(np.where(pos_index, dcg, 0), axis=1)
    return dcg / idcg


metrics_dict = {"hit": hit_, "mrr": mrr_, "map": map_, "recall": recall_, "ndcg": ndcg_, "precision": precision_}

# order used by the CUDA metric reducer (recbole_b200) -- keep in sync with include/recbole_b200.h
METRIC_ORDER = ("recall", "mrr", "ndcg", "hit", "precision", "map")


def calculate_metrics(pos_index, pos_len, metrics):
    """evaluators.py:122-141 -> float64[len(metrics), K] (mean over users)."""
    pos_len = np.asarray(pos_len)
    res = [metrics_dict[m.lower()](pos_index, pos_len) for m in metrics]
    return np.stack(res, axis=0).mean(axis=1)


def evaluate(pos_index, pos_len, metrics, topk, precision=4):
    """evaluators.py:96-105 -> {'recall@10': 0.1234, ...}."""
    value = calculate_metrics(pos_index, pos_len, metrics)
    out = {}
    for m, row in zip(metrics, value):
        for k in topk:
            out["{}@{}".format(m, k)] = round(float(row[k - 1]), precision)
    return out
