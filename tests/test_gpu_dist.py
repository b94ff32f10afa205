"""GPU parity of the sharded (multi-GPU) path: 2 ranks sharing ONE GPU (collectives staged through
gloo), compared with the oracle's single-device step on the union of the ranks' batches."""
import numpy as np
import pytest
import torch

from dist_util import run_ranks

pytestmark = pytest.mark.gpu


def _case(seed=0, n_users=400, n_items=301, d=64, B=3000, steps=2):
    rng = np.random.default_rng(seed)
    U0 = (rng.standard_normal((n_users, d)) * 0.3).astype(np.float32)
    V0 = (rng.standard_normal((n_items, d)) * 0.3).astype(np.float32)
    batches = []
    for _ in range(steps):
        batches.append((rng.integers(1, n_users, B), np.minimum(np.exp(rng.random(B) * np.log(n_items - 1)).astype(np.int64),
                                                                n_items - 1).clip(1), rng.integers(1, n_items, B)))
    # evaluation data
    pairs = [(rng.integers(1, n_users, 4000), rng.integers(1, n_items, 4000)) for _ in range(3)]
    return U0, V0, batches, pairs


def _case_for(kind):
    # adam_lazy (dense-Adam trajectory): small batches over several steps, so that most rows are NOT touched in a step
    # and keep moving on their momentum
    return _case(B=150, steps=5) if kind == "adam_lazy" else _case()


def _rank_fn(rank, world, kind, mode, exchange):
    from oracle import fullsort as ofs
    from recbole_b200.dist import Comm, ShardedBPR, ShardedEvalIndex
    from recbole_b200.evaluator import FusedTopKEvaluator
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    U0, V0, batches, pairs = _case_for(kind)
    n_users, n_items, d = U0.shape[0], V0.shape[0], U0.shape[1]
    comm = Comm(staged=True)
    m = ShardedBPR(n_users, n_items, d, comm, dev, U_full=U0, V_full=V0, exchange=exchange)
    m.build_optimizer(kind, lr=0.05 if kind == "sgd" else 2e-3)
    losses = []
    for (u, p, n) in batches:
        mine = (u >= m.u_lo) & (u < m.u_hi)            # users are partitioned: a rank trains its own users' samples
        t = lambda a: torch.from_numpy(a[mine]).to(dev)  # noqa: E731
        lo = m.train_step(t(u), t(p), t(n), global_batch=len(u))
        losses.append(float(lo.item()))
    m.check_flags()
    m.flush()
    if exchange == "p2p":
        comm.barrier()          # peers may still be reading this rank's shard through their mappings

    class Cfg(dict):
        def __getitem__(self, k):
            return self.get(k)

    ev = FusedTopKEvaluator(Cfg(metrics=["Recall", "MRR", "NDCG", "Hit", "Precision", "MAP"], topk=[1, 5, 10],
                                metric_decimal_place=4))
    uid, hist, pos = ofs.eval_index(n_users, pairs, 2)
    idx = ShardedEvalIndex.from_global(uid, hist, pos, m.user_bounds, m.item_bounds, rank, dev)
    res = m.evaluate(idx, ev, mode=mode, layout="sharded")
    res2 = m.evaluate(idx, ev, mode=mode, layout="replicate")
    assert res2 == res
    return dict(losses=losses, U=m.U[: m.u_hi - m.u_lo].cpu().numpy(), V=m.V[: m.i_hi - m.i_lo].cpu().numpy(), res=res,
                topk=m.last_topk.cpu().numpy(), u_lo=m.u_lo, i_lo=m.i_lo)


@pytest.mark.parametrize("kind,mode,exchange", [("adam", "tc", "sparse"), ("sgd", "fp32", "sparse"),
                                                ("adam", "fp32", "dense"), ("sgd", "tc", "dense"),
                                                ("adam", "fp32", "p2p"), ("sgd", "tc", "p2p"),
                                                ("adam_lazy", "fp32", "p2p"), ("adam_lazy", "tc", "dense")])
def test_two_ranks_equal_single_device_oracle(kind, mode, exchange):
    """(adam_lazy: the oracle runs DENSE Adam -- every row moves at every step -- on the union batch.)"""
    from oracle import bpr as obpr
    from oracle import fullsort as ofs
    out = run_ranks(_rank_fn, 2, kind, mode, exchange, timeout=300)
    U0, V0, batches, pairs = _case_for(kind)
    st = obpr.new_state(U0, V0)
    lr = 0.05 if kind == "sgd" else 2e-3
    for s, (u, p, n) in enumerate(batches):
        lo = obpr.bpr_train_step(st, u, p, n, s + 1, optimizer="sgd" if kind == "sgd" else "adam", lr=lr,
                                 dense=(kind == "adam_lazy"))
        for r in range(2):
            assert abs(out[r]["losses"][s] - lo) <= 1e-5 * abs(lo)
    U = np.concatenate([out[0]["U"], out[1]["U"]])
    V = np.concatenate([out[0]["V"], out[1]["V"]])
    for got_t, want_t in ((U, st["U"]), (V, st["V"])):
        diff = np.abs(got_t - want_t)
        bad = diff > 1e-5 * np.abs(want_t).max()
        if kind == "adam_lazy":
            # eps-conditioned elements: a gradient element of ~1e-8 (|g| ~ eps) is known to ~1e-3 relative only (fp32
            # summation order), and Adam's normalised step lr * g / (|g| + eps) turns that into ~1e-3 of a full step,
            # again at every zero-gradient step that follows.  (The samples with pos == neg in this random data have
            # an EXACTLY zero user gradient in the reference; the kernels form g * (vi - vj) and keep it zero.)
            assert bad.sum() <= 1 and diff.max() <= 2e-4, (int(bad.sum()), float(diff.max()))
        else:
            assert not bad.any(), float(diff.max())
    # evaluation on the tables the ranks actually hold (so that ids can be compared bit for bit)
    uid, hist, pos = ofs.eval_index(U0.shape[0], pairs, 2)
    o_ids, _ = ofs.full_sort_topk(U, V, uid, hist[0], hist[1], 10)
    got = np.concatenate([out[0]["topk"], out[1]["topk"]])   # last_topk of the "replicate" layout: own users
    np.testing.assert_array_equal(got, o_ids)
    ref = ofs.evaluate(o_ids, pos[0], pos[1], ["recall", "mrr", "ndcg", "hit", "precision", "map"], [1, 5, 10])
    assert out[0]["res"] == ref and out[1]["res"] == ref


# ---- row-sharded FM (BASELINE config 5 across GPUs) ------------------------------------------------------------
FM_DIMS = [7, 300, 3, 41, 1000, 2, 90, 513]


def _fm_case(seed=5, B=2048, steps=3, d=16, dims=None):
    global FM_DIMS
    if dims is not None:
        FM_DIMS = dims
    rng = np.random.default_rng(seed)
    rows = int(sum(FM_DIMS))
    E0 = (rng.standard_normal((rows, d)) * 0.1).astype(np.float32)
    W0 = (rng.standard_normal(rows) * 0.1).astype(np.float32)
    batches = []
    for _ in range(steps):
        ids = np.stack([np.minimum(np.exp(rng.random(B) * np.log(n)).astype(np.int64), n - 1) for n in FM_DIMS], axis=1)
        batches.append((ids, (rng.random(B) < 0.3).astype(np.float32)))
    return E0, W0, batches


def _fm_rank_fn(rank, world, kind, B=2048, dims=None):
    from recbole_b200.dist import Comm, ShardedFM
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    E0, W0, batches = _fm_case(B=B, dims=dims)
    comm = Comm(staged=True)
    m = ShardedFM(FM_DIMS, E0.shape[1], comm, dev, E_full=E0, W_full=W0, bias=0.05)
    m.build_optimizer(kind, lr=0.05 if kind == "sgd" else 2e-3)
    losses = []
    for ids, lab in batches:
        B = ids.shape[0]
        lo, hi = rank * B // world, (rank + 1) * B // world          # the batch is split by rows
        loss = m.train_step(torch.from_numpy(ids[lo:hi]).to(dev), torch.from_numpy(lab[lo:hi]).to(dev), global_batch=B)
        losses.append(float(loss.item()))
    E, W = m.gather_tables()
    return dict(losses=losses, E=E.cpu().numpy(), W=W.cpu().numpy(), b=float(m.bias3[0].item()))


@pytest.mark.parametrize("world,kind,B,dims", [(2, "adam", 2048, None), (3, "adam", 2048, None), (2, "sgd", 2048, None),
                                               (2, "adam", 1001, [5, 300, 3, 41, 77, 2, 90])])   # odd B * F
def test_sharded_fm_equals_single_device_oracle(world, kind, B, dims):
    """The table row-sharded over `world` ranks, the batch split by rows: losses, both tables and the bias match the
    oracle's single-device row-sparse step on the whole batch to 1e-5."""
    from oracle import fm as ofm
    global FM_DIMS
    saved = list(FM_DIMS)
    try:
        _check_sharded_fm(ofm, world, kind, B, dims)
    finally:
        FM_DIMS = saved


def _check_sharded_fm(ofm, world, kind, B, dims):
    out = run_ranks(_fm_rank_fn, world, kind, B, dims, timeout=300)
    E0, W0, batches = _fm_case(B=B, dims=dims)
    st = ofm.new_state(E0, W0, 0.05)
    off = np.concatenate([[0], np.cumsum(FM_DIMS)[:-1]]).astype(np.int64)
    lr = 0.05 if kind == "sgd" else 2e-3
    for s, (ids, lab) in enumerate(batches):
        rows = ids + off[None, :]
        if kind == "adam":
            lo = ofm.fm_train_step(st, rows, lab, s + 1, dense=False, lr=lr)
        else:
            lo, dE, dW, db, _ = ofm.fm_grads(st["E"], st["W"], st["b"][0], rows, lab)
            st["E"] -= np.float32(lr) * dE
            st["W"] -= np.float32(lr) * dW
            st["b"] -= np.float32(lr) * db
        for r in range(world):
            assert abs(out[r]["losses"][s] - lo) <= 1e-5 * abs(lo), (s, out[r]["losses"][s], lo)
    for r in range(world):
        assert np.abs(out[r]["E"] - st["E"]).max() <= 1e-5 * np.abs(st["E"]).max()
        assert np.abs(out[r]["W"] - st["W"]).max() <= 1e-5 * np.abs(st["W"]).max()
        assert abs(out[r]["b"] - float(st["b"][0])) <= 1e-5 * abs(float(st["b"][0]))
