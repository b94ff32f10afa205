"""GPU parity: fused full-sort CE head (logits never materialised) vs the reference golden
(sasrec.py:137-141,152-158 expressions) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import ce_head as oce

pytestmark = pytest.mark.gpu


def test_ce_head_vs_reference_golden(golden):
    from recbole_b200 import ops
    from gpu_util import rel_err, t
    g = golden("ce_head.npz")
    out = ops.ce_head(t(g["X"]), t(g["E"]), t(g["pos"]), k=10)
    assert abs(out["loss"].item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel_err(out["lse"].cpu().numpy(), g["lse"]) < 1e-5
    o_ids, o_sc = oce.full_sort_topk(g["X"], g["E"], 10)
    np.testing.assert_array_equal(out["ids"].cpu().numpy(), o_ids)          # bit-exact vs the oracle
    np.testing.assert_array_equal(out["scores"].cpu().numpy(), o_sc)
    assert (out["ids"].cpu().numpy() != g["topk_ids"]).sum() <= 2           # reference: fp32 near-ties only


@pytest.mark.parametrize("nq,N,d", [(4096, 100001, 64), (37, 5000, 128), (1, 3, 16), (300, 70000, 32)])
def test_ce_head_random_vs_oracle(nq, N, d):
    from recbole_b200 import ops
    from gpu_util import rel_err, t
    rng = np.random.default_rng(nq + N)
    X = rng.standard_normal((nq, d)).astype(np.float32)
    X = (X - X.mean(1, keepdims=True)) / X.std(1, keepdims=True)          # layer-normed rows
    E = (rng.standard_normal((N, d)) * 0.02 * 20).astype(np.float32)
    E[0] = 0
    tgt = rng.integers(1, N, nq) if N > 1 else np.zeros(nq, np.int64)
    out = ops.ce_head(t(X), t(E), t(tgt), k=min(10, N))
    sub = np.arange(min(nq, 64))
    loss_sub, lse_sub = oce.ce_loss(X[sub], E, tgt[sub])
    assert rel_err(out["lse"].cpu().numpy()[sub], lse_sub) < 1e-5
    o_ids, o_sc = oce.full_sort_topk(X[sub], E, min(10, N))
    np.testing.assert_array_equal(out["ids"].cpu().numpy()[sub], o_ids)
    if nq <= 64:
        assert abs(out["loss"].item() - loss_sub) <= 1e-5 * abs(loss_sub)
    else:  # loss over all rows = mean(lse - target logit); check through its definition
        lse = out["lse"].cpu().numpy().astype(np.float64)
        tl = oracle_pair(X, E, tgt)
        assert abs(out["loss"].item() - float((lse - tl).mean())) <= 1e-5 * abs(float((lse - tl).mean()))


def oracle_pair(X, E, tgt):
    import oracle
    return oracle.clib.pair_scores_fma(X, E, np.arange(len(tgt)), tgt).astype(np.float64)


@pytest.mark.parametrize("scale", [0.02, 1.0])
def test_ce_head_tensor_core_vs_fp32_kernel_at_cfg4_size(scale):
    """BASELINE config 4 shape (4096 x 1,000,001 x 64): the tensor-core CE head (bf16 hi/lo split
    operands, K = 192, online logsumexp in the epilogue) against the exact CUDA-core kernel: identical
    top-k, logsumexp and loss within 1e-5 (measured ~1e-6); `scale` 1.0 gives logits of +-40."""
    from recbole_b200 import ops
    from recbole_b200._lib import lib
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev)
    gen.manual_seed(4)
    nq, N, d = 4096, 1_000_001, 64
    X = torch.randn(nq, d, device=dev, generator=gen)
    X = (X - X.mean(1, keepdim=True)) / X.std(1, keepdim=True)
    E = torch.randn(N, d, device=dev, generator=gen) * scale
    E[0] = 0
    tgt = torch.randint(1, N, (nq,), device=dev, generator=gen)
    a = ops.ce_head(X, E, tgt, 10, scorer="tc")
    fb = lib.rb2_fullsort_tc_last_fallback_rows()
    b = ops.ce_head(X, E, tgt, 10, scorer="fp32")
    assert torch.equal(a["ids"], b["ids"]) and torch.equal(a["scores"], b["scores"])
    assert fb <= nq // 100
    rel = ((a["lse"] - b["lse"]).abs() / b["lse"].abs()).max().item()
    assert rel < 1e-5, rel
    assert abs(a["loss"].item() - b["loss"].item()) <= 1e-5 * abs(b["loss"].item())
    # a random subset of the rows against the ORACLE at full item count: top-k bit-exact, logsumexp within 1e-5
    rows = np.sort(np.random.default_rng(7).choice(nq, 32, replace=False))
    Xc, Ec, tc_ = X.cpu().numpy()[rows], E.cpu().numpy(), tgt.cpu().numpy()[rows]
    o_ids, o_sc = oce.full_sort_topk(Xc, Ec, 10)
    np.testing.assert_array_equal(a["ids"].cpu().numpy()[rows], o_ids)
    np.testing.assert_array_equal(a["scores"].cpu().numpy()[rows], o_sc)
    L = Xc.astype(np.float64) @ Ec.astype(np.float64).T
    mx = L.max(axis=1)
    lse64 = mx + np.log(np.exp(L - mx[:, None]).sum(axis=1))
    got = a["lse"].cpu().numpy()[rows].astype(np.float64)
    assert np.abs(got - lse64).max() <= 1e-5 * np.abs(lse64).max() and (np.abs(got - lse64) <= 1e-5 * np.abs(lse64) + 1e-6).all()
    ops.ce_head(X[:8], E[:100], tgt[:8] % 100, 10, scorer="auto")   # leaves the knob at its default


def _elementwise_ok(a, b, rtol=1e-5):
    """Two criteria.  (1) north_star's: max |a - b| <= 1e-5 * max |b|.  (2) element-wise:
        |a - b| <= rtol * |b| + rtol * rms(row of b) + 1e-6 * rms(b)
    -- relative, with an absolute floor at the scale of the element's own gradient ROW (the rows of target items are
    orders of magnitude larger than the others; a tensor-wide floor of 1e-5 would say nothing about either kind) plus
    1e-6 of the tensor's scale: softmax weights below ~1e-6 sit in FP16's subnormal range after the rescale, their
    ABSOLUTE error is 2^-39 of the largest weight, and rows made only of such weights are not relatively accurate.
    Returns (number of elements failing (2), or -1 when (1) fails; max error / max |b|)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    glob = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
    floor = rtol * np.sqrt((b * b).mean(axis=-1, keepdims=True)) + 1e-6 * np.sqrt((b * b).mean())
    bad = np.abs(a - b) > rtol * np.abs(b) + floor
    return (int(bad.sum()) if glob <= 1e-5 else -1), glob


def test_ce_backward_vs_torch_autograd_golden(golden):
    """rb2_ce_head_backward vs torch autograd of the reference's CE branch (sasrec.py:137-141): dX, dE element-wise
    to 1e-5; one dense Adam step on the item table == torch.optim.Adam's."""
    from recbole_b200 import ops
    from gpu_util import t
    g = golden("ce_backward.npz")
    X, E, pos = t(g["X"]), t(g["E"]), t(g["pos"])
    out = ops.ce_head(X, E, pos, k=1)
    assert abs(out["loss"].item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    dx, de = ops.ce_head_backward(X, E, pos, out["lse"])
    nbad, rel = _elementwise_ok(dx.cpu().numpy(), g["dX"])
    assert nbad == 0, (nbad, rel)
    nbad, rel = _elementwise_ok(de.cpu().numpy(), g["dE"])
    assert nbad == 0, (nbad, rel)
    opt = ops.Optim("adam", lr=1e-3)
    m, v = torch.zeros_like(E), torch.zeros_like(E)
    ops.dense_step(E, m, v, de, opt, step=1)
    # Adam's first step moves every element by ~lr * sign(g): elements whose gradient is ~eps are ill-conditioned
    d = np.abs(E.cpu().numpy() - g["E1"])
    assert (d > 1e-5 * np.abs(g["E1"]).max()).sum() <= 3 and d.max() <= 2e-3


def test_ce_head_autograd_function(golden):
    """ops.ce_head_loss is one autograd node: loss.backward() fills .grad of the sequence output and of the item table."""
    from recbole_b200 import ops
    from gpu_util import t
    g = golden("ce_backward.npz")
    X = t(g["X"]).requires_grad_(True)
    W = torch.nn.Parameter(t(g["E"]))
    pre = (X * 2.0)                                   # an ordinary autograd op below the head
    loss = ops.ce_head_loss(pre * 0.5, W, t(g["pos"]))
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    loss.backward()
    assert _elementwise_ok(X.grad.cpu().numpy(), g["dX"])[0] == 0
    assert _elementwise_ok(W.grad.cpu().numpy(), g["dE"])[0] == 0


@pytest.mark.parametrize("nq,N", [(1, 3), (130, 127), (700, 40000), (4096, 300001)])
def test_ce_backward_random_vs_oracle(nq, N):
    """Shapes that do not fill tiles, many item ranges, and (last) a BASELINE-config-4-sized batch against a table
    large enough that the float64 oracle is only run on subsets: all of dX for 64 rows, dE for 512 item rows."""
    from recbole_b200 import ops
    from gpu_util import t
    rng = np.random.default_rng(nq + N)
    d = 64
    X = rng.standard_normal((nq, d)).astype(np.float32)
    X = (X - X.mean(1, keepdims=True)) / X.std(1, keepdims=True)
    E = (rng.standard_normal((N, d)) * 0.02 * 20).astype(np.float32)
    E[0] = 0
    tgt = rng.integers(1, N, nq) if N > 1 else np.zeros(nq, np.int64)
    Xd, Ed, td = t(X), t(E), t(tgt)
    out = ops.ce_head(Xd, Ed, td, k=1)
    dx, de = ops.ce_head_backward(Xd, Ed, td, out["lse"])
    dx, de = dx.cpu().numpy(), de.cpu().numpy()
    if nq * N <= 40_000_000:
        o_dx, o_de = oce.ce_backward(X, E, tgt)
        assert _elementwise_ok(dx, o_dx)[0] == 0, _elementwise_ok(dx, o_dx)
        assert _elementwise_ok(de, o_de)[0] == 0, _elementwise_ok(de, o_de)
    else:
        # float64 softmax of every row is needed for dE anyway: do it in row blocks, keep only what is compared
        rows = np.sort(rng.choice(nq, 64, replace=False))
        items = np.unique(np.concatenate([rng.choice(N, 500, replace=False), tgt[:12]]))
        E64 = E.astype(np.float64)
        o_de = np.zeros((len(items), d))
        o_dx = np.zeros((len(rows), d))
        for lo in range(0, nq, 256):
            L = X[lo:lo + 256].astype(np.float64) @ E64.T
            L -= L.max(axis=1, keepdims=True)
            P = np.exp(L)
            P /= P.sum(axis=1, keepdims=True)
            P[np.arange(P.shape[0]), tgt[lo:lo + 256]] -= 1.0
            P /= nq
            o_de += P[:, items].T @ X[lo:lo + 256].astype(np.float64)
            sel = (rows >= lo) & (rows < lo + 256)
            if sel.any():
                o_dx[sel] = P[rows[sel] - lo] @ E64
        assert _elementwise_ok(dx[rows], o_dx)[0] == 0, _elementwise_ok(dx[rows], o_dx)
        assert _elementwise_ok(de[items], o_de)[0] == 0, _elementwise_ok(de[items], o_de)
    # determinism: bit-identical reruns (no float atomics)
    dx2, de2 = ops.ce_head_backward(Xd, Ed, td, out["lse"])
    assert np.array_equal(dx2.cpu().numpy(), dx) and np.array_equal(de2.cpu().numpy(), de)
