"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT.

A CPU restatement (numpy + a small C file) of the reference's algorithm for the
embedding hot path of ghazalehnt/RecBole.  Every function
cites the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; ``recbole_b200`` (the
product) never does and fails loudly when its CUDA library is missing.

Pinning: the oracle is checked in ``tests/test_oracle_*.py``
against

* the reference's own known-answer vectors
  (``tests/metrics/test_topk_metrics.py:15-79``,
  ``tests/data/test_dataloader.py:192-233``), and
* outputs of the reference itself (its ``BPR``, ``BPRLoss``, ``FM``, ``Sampler``,
  ``GeneralFullDataLoader``, ``Trainer._full_sort_batch_eval``, ``TopKEvaluator``
  and ``torch.optim.Adam/SGD``), generated in the build container by
  ``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``.

Arithmetic that the reference delegates to PyTorch (Adam/SGD update, BCELoss,
CrossEntropyLoss) is restated from torch 2.11.0's single-tensor code path
(``torch/optim/adam.py``), see ``oracle/optim.py``.
"""
from . import bpr, ce_head, fm, fullsort, metrics, optim, sampler  # noqa: F401
from ._clib import lib as clib  # noqa: F401
