"""CPU-side checks: the C-ABI library loads and exports every symbol include/recbole_b200.h
declares, the host logic (optimizer scalars, index construction, evaluator result format) is
right, and the product refuses to run without a CUDA device (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "recbole_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rb2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import recbole_b200._lib as L
    names = _declared_symbols()
    assert len(names) >= 15
    raw = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), n
    assert set(names) == set(L.SIGNATURES), set(names) ^ set(L.SIGNATURES)
    assert L.lib.rb2_abi_version() == L.ABI_VERSION


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "recbole_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f


def test_no_cpu_fallback():
    from recbole_b200 import ops
    U = torch.zeros(4, 16)
    with pytest.raises(ValueError, match="CUDA"):
        ops._ptr(U, torch.float32)


def test_optim_scalars_match_torch_formulas():
    from recbole_b200 import ops
    from oracle import optim as ooptim
    o = ops.Optim("adam", lr=1e-3)
    for t in (1, 2, 10, 1000):
        c = o.c_struct(step=t)
        h = ooptim.adam_hparams(t)
        assert c.step_size == np.float32(h["step_size"])
        assert c.bc2_sqrt == np.float32(h["bc2_sqrt"])
        assert c.one_minus_beta1 == np.float32(1 - 0.9) and c.one_minus_beta2 == np.float32(1 - 0.999)
    with pytest.raises(ValueError):
        ops.Optim("rmsprop")


def test_eval_index_matches_oracle(golden):
    from oracle import fullsort
    from recbole_b200.data import EvalIndex
    rng = np.random.default_rng(0)
    n_users, n_items = 40, 30
    pairs = []
    for _ in range(3):
        pairs.append((rng.integers(1, n_users, 200), rng.integers(1, n_items, 200)))
    for phase in (1, 2):
        uid, hist, pos = fullsort.eval_index(n_users, pairs, phase)
        idx = EvalIndex.from_phase_pairs(n_users, n_items, pairs, phase, "cpu")
        # the oracle keeps an item that is both earlier-used and a positive out of the history too
        np.testing.assert_array_equal(idx.uid_list.numpy(), uid)
        np.testing.assert_array_equal(idx.pos_indptr.numpy(), pos[0])
        np.testing.assert_array_equal(idx.pos_indices.numpy(), pos[1])
        np.testing.assert_array_equal(idx.hist_indptr.numpy(), hist[0])
        np.testing.assert_array_equal(idx.hist_indices.numpy(), hist[1])


def test_evaluator_result_format():
    from recbole_b200.evaluator import FusedTopKEvaluator

    class Cfg(dict):
        def __getitem__(self, k):
            return self.get(k)

    ev = FusedTopKEvaluator(Cfg(metrics=["Recall", "NDCG"], topk=[1, 3], metric_decimal_place=4))
    sums = torch.arange(18, dtype=torch.float64).reshape(6, 3)
    res = ev.result(sums, 7)
    assert list(res) == ["recall@1", "recall@3", "ndcg@1", "ndcg@3"]
    assert res["recall@3"] == round(2 / 7, 4) and res["ndcg@1"] == round(6 / 7, 4)
    with pytest.raises(ValueError):
        FusedTopKEvaluator(Cfg(metrics=["AUC"], topk=[10]))
    with pytest.raises(ValueError):
        FusedTopKEvaluator(Cfg(metrics=["Recall"], topk=[0]))


def test_model_mirror_api_surface():
    """FusedBPR exposes the reference's plugin surface and state-dict keys."""
    from recbole_b200 import FusedBPR

    class Cfg(dict):
        def __getitem__(self, k):
            return self.get(k)

    class DS:
        def num(self, f):
            return {"user_id": 11, "item_id": 7}[f]

    m = FusedBPR(Cfg(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device="cpu",
                     embedding_size=16), DS())
    assert sorted(m.state_dict()) == ["item_embedding.weight", "user_embedding.weight"]
    assert (m.n_users, m.n_items, m.NEG_ITEM_ID) == (11, 7, "neg_item_id")
    for name in ("calculate_loss", "predict", "full_sort_predict"):
        assert callable(getattr(m, name))
    with pytest.raises(ValueError, match="no CPU path"):        # the product path fails loudly off the GPU
        m.full_sort_predict({"user_id": torch.tensor([1])})
    # xavier-normal std (init.py:27): sqrt(2 / (rows + d))
    big = FusedBPR(Cfg(USER_ID_FIELD="u", ITEM_ID_FIELD="i", NEG_PREFIX="neg_", device="cpu", embedding_size=64),
                   type("D", (), {"num": lambda self, f: 4000})())
    assert abs(big.user_embedding.weight.std().item() - (2 / 4064) ** 0.5) < 2e-3


def test_bench_reference_arm_cfg4_prints_one_json_line():
    """`bench.py --impl reference` needs no GPU: one JSON line with the contract's keys (smallest workload arm)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "cfg4",
                          "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_bench_reference_arm_default_workload_prints_one_json_line():
    """The DEFAULT workload's reference arm (cfg3 shape, shrunk by --scale so that the CPU finishes in seconds): same
    metric / unit / config keys as the GPU arm's line, full batches, e2e block with zero transfer bytes."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--scale", "0.002", "--batch", "65536"], capture_output=True, text=True,
                         timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bpr_train_samples_per_s" and d["unit"] == "samples/s"
    assert d["config"]["workload"] == "cfg3" and d["config"]["dim"] == 128 and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["eval"]["metric"] == "fullsort_eval_users_per_s" and d["eval"]["value"] > 0


def test_fused_fm_token_seq_host_layout_cpu():
    """FusedFM with TOKEN + FLOAT + TOKEN_SEQ fields, host side only (no kernel runs on the CPU): the token table and the
    sequence tables are views of ONE zero-padded row range, the state dict is compact and carries the reference's names,
    and the per-batch column layout (offsets, column -> sequence field, column ranges) follows the padded lengths."""
    import numpy as np
    import torch
    from recbole_b200 import FusedFM

    class Cfg(dict):
        def __getitem__(self, k):
            return self.get(k)

    class DS:
        field2type = {"t0": "token", "t1": "token", "x0": "float", "q0": "token_seq", "q1": "token_seq", "label": "float"}
        _num = {"t0": 30, "t1": 200, "x0": 1, "q0": 25, "q1": 7, "label": 1}

        def fields(self):
            return list(self._num)

        def num(self, f):
            return self._num[f]

    m = FusedFM(Cfg(LABEL_FIELD="label", embedding_size=10, device="cpu"), DS())
    assert (m.n_seq, m.n_float, m._dpad, m._tok_rows, m._seq_bases, m._all_rows) == (2, 1, 16, 230, [230, 255], 262)
    w0 = m.token_seq_embedding_table[1].weight.detach().clone()
    E, W = m._tables()
    assert tuple(E.shape) == (262, 16) and tuple(W.shape) == (262,)
    assert float(E[:, 10:].abs().max()) == 0.0                                   # the padding columns
    assert torch.equal(E[255:262, :10], w0)                                      # values carried over ...
    for p, base in ((m.token_embedding_table.embedding.weight, 0), (m.token_seq_embedding_table[0].weight, 230),
                    (m.token_seq_embedding_table[1].weight, 255)):
        assert p.data.data_ptr() == E.data_ptr() + 4 * 16 * base and p.data.stride(0) == 16      # ... and views now
    E2, _ = m._tables()
    assert E2.data_ptr() == E.data_ptr()                                          # stable: not rebuilt per call
    sd = m.state_dict()
    assert tuple(sd["token_seq_embedding_table.0.weight"].shape) == (25, 10)
    assert sd["token_seq_embedding_table.0.weight"].is_contiguous()
    assert tuple(sd["first_order_linear.token_seq_embedding_table.1.weight"].shape) == (7, 1)
    B = 5
    inter = {"t0": torch.zeros(B, dtype=torch.int64), "t1": torch.ones(B, dtype=torch.int64),
             "x0": torch.rand(B), "q0": torch.zeros((B, 6), dtype=torch.int64),
             "q1": torch.zeros((B, 3), dtype=torch.int64), "label": torch.zeros(B)}
    ids, offsets, seq = m._batch(inter, train=True)
    assert tuple(ids.shape) == (B, 2 + 6 + 3)
    assert offsets.tolist() == [0, 30] + [230] * 6 + [255] * 3
    assert seq["col_seq"].tolist() == [-1, -1] + [0] * 6 + [1] * 3 and seq["seq_start"].tolist() == [2, 8, 11]
    assert seq["n_token_cols"] == 2 and seq["seq_row_base"] == 230
    assert tuple(seq["pooled"].shape) == (B, 2, 16) and tuple(seq["coef"].shape) == (B, 2)
    inter["q0"] = torch.zeros((B, 9), dtype=torch.int64)                          # another padded length: its own layout
    ids2, offsets2, seq2 = m._batch(inter, train=False)
    assert tuple(ids2.shape) == (B, 2 + 9 + 3) and seq2["seq_start"].tolist() == [2, 11, 14] and "pooled" not in seq2
    # the optimizer-state entries follow the reference's parameter order (9 tensors)
    m._state = dict(mE=torch.zeros_like(E), vE=torch.zeros_like(E), mW=torch.zeros_like(W), vW=torch.zeros_like(W),
                    mEf=torch.zeros(1, 16), vEf=torch.zeros(1, 16), mWf=torch.zeros(1), vWf=torch.zeros(1))
    m._bias3 = torch.zeros(3)
    shapes = [tuple(a.shape) for a, _ in m._opt_entries()]
    assert shapes == [(230, 10), (1, 10), (25, 10), (7, 10), (1,), (230,), (1,), (25,), (7,)]
    assert [tuple(p.shape) for _, p in m.named_parameters()] == [(230, 10), (1, 10), (25, 10), (7, 10), (1,), (230, 1),
                                                                  (1, 1), (25, 1), (7, 1)]
