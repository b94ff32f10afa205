"""Top-K evaluation on the device; mirrors ``TopKEvaluator`` (recbole/evaluator/evaluators.py:37-149)
and ``ProxyEvaluator.evaluate``'s result format (proxy_evaluator.py:79-95).

The reference collects an int64 [users, K+1] matrix per batch on the device, concatenates,
copies it to the host and runs numpy metric functions with Python row loops
(evaluators.py:89-141, metrics.py).  Here the hits and all six metric sums are reduced on the
device (rb2_topk_metrics); only 6*K doubles come back.
"""
import numpy as np

from . import _lib, ops

TOPK_METRICS = set(_lib.METRIC_ORDER)


class FusedTopKEvaluator:
    def __init__(self, config):
        # proxy_evaluator.py:97-109 / evaluators.py:106-120: names are case-insensitive, topk int or list
        metrics = config["metrics"]
        self.metrics = [metrics] if isinstance(metrics, str) else list(metrics)
        for m in self.metrics:
            if m.lower() not in TOPK_METRICS:
                raise ValueError("metric %r is not a top-k metric of the fused evaluator %s"
                                 % (m, sorted(TOPK_METRICS)))
        topk = config["topk"]
        self.topk = [topk] if isinstance(topk, int) else list(topk)
        for k in self.topk:
            if not isinstance(k, int) or k <= 0:
                raise ValueError("topk must be a positive integer or a list of positive integers, "
                                 "but get `{}`".format(k))
        self.precision = config["metric_decimal_place"] if config["metric_decimal_place"] is not None else 4

    @property
    def max_k(self):
        return max(self.topk)

    def sums(self, topk_ids, index):
        """Device float64 [6, K] metric sums over the evaluated users (add across shards / ranks)."""
        return ops.topk_metrics(topk_ids, index.pos_indptr, index.pos_indices, index.n_items)["sums"]

    def result(self, sums, n_users):
        """evaluators.py:96-105: mean over users, round(., precision), keys '{metric}@{k}'
        with the metric spelled as configured (lower-cased by the reference at
        proxy_evaluator.py:42)."""
        mean = sums.detach().cpu().numpy() / float(n_users)
        out = {}
        for m in self.metrics:
            row = mean[_lib.METRIC_ORDER.index(m.lower())]
            for k in self.topk:
                out["{}@{}".format(m.lower(), k)] = round(float(row[k - 1]), self.precision)
        return out

    def evaluate(self, topk_ids, index):
        return self.result(self.sums(topk_ids, index), topk_ids.shape[0])

    def reference_matrix(self, topk_ids, index):
        """int64 [users, K+1] in the reference's own swapped+flipped coordinates: feed it to an
        unmodified TopKEvaluator.evaluate (evaluators.py:78-105) as `batch_matrix_list=[m]`."""
        return ops.topk_metrics(topk_ids, index.pos_indptr, index.pos_indices, index.n_items,
                                want_ref_idx=True)["ref_idx"]
