"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: id de-duplication and bucketing by
owner, all-to-all #1 (rows out) and #2 (gradients back), uneven all-gather, the eval index split.
The compute kernels are stood in for by torch index ops here -- this file checks the PLUMBING; the
kernels themselves are checked on the GPU (tests/test_gpu_dist.py)."""
import numpy as np
import torch

from dist_util import run_ranks


def _exchange_rank(rank, world, seed):
    from recbole_b200.dist import Comm, fetch_rows, plan_item_exchange, return_grads, shard_bounds
    comm = Comm()
    n_items, d = 1001, 8
    bounds = shard_bounds(n_items, world)
    rng = np.random.default_rng(seed)
    V_full = torch.from_numpy(rng.standard_normal((n_items, d)).astype(np.float32))
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    V_local = V_full[lo:hi].clone()
    items = torch.from_numpy(np.random.default_rng(seed + 10 + rank).integers(1, n_items, 5000))
    uniq, inv, send_counts = plan_item_exchange(items, bounds)
    assert torch.equal(uniq[inv], items) and sum(send_counts) == uniq.numel()
    C, local_idx, recv_counts = fetch_rows(comm, uniq, send_counts, lambda idx: V_local.index_select(0, idx), lo)
    assert torch.equal(C, V_full[uniq])                       # every rank got exactly the rows it asked for
    assert (local_idx >= 0).all() and (local_idx < hi - lo).all()
    # gradients: row value = item id * (rank + 1); the owner must receive the sum over ranks
    G = uniq.to(torch.float32).unsqueeze(1).repeat(1, d) * (rank + 1)
    grads = return_grads(comm, G, send_counts, recv_counts)
    acc = torch.zeros_like(V_local)
    acc.index_add_(0, local_idx, grads)                       # stand-in for rb2_sparse_rows_update
    # expected: sum over ranks that requested the item
    exp = torch.zeros_like(V_local)
    for r in range(world):
        it = torch.from_numpy(np.random.default_rng(seed + 10 + r).integers(1, n_items, 5000)).unique()
        it = it[(it >= lo) & (it < hi)]
        exp[it - lo] += it.to(torch.float32).unsqueeze(1) * (r + 1)
    assert torch.equal(acc, exp)
    # uneven all-gather
    rows = torch.full((3 + rank, 2), float(rank))
    out = comm.all_gather_rows(rows, [3 + r for r in range(world)])
    assert out.shape[0] == sum(3 + r for r in range(world)) and out[-1, 0] == world - 1
    t = torch.tensor([1.0 + rank])
    comm.all_reduce_sum(t)
    assert t.item() == sum(1.0 + r for r in range(world))
    return True


def test_exchange_plumbing_world2():
    assert run_ranks(_exchange_rank, 2, 123) == [True, True]


def test_exchange_plumbing_world3():
    assert run_ranks(_exchange_rank, 3, 7) == [True, True, True]


def test_sharded_eval_index_split():
    from oracle import fullsort as ofs
    from recbole_b200.dist import ShardedEvalIndex, shard_bounds
    rng = np.random.default_rng(0)
    n_users, n_items, world = 50, 40, 3
    pairs = [(rng.integers(1, n_users, 300), rng.integers(1, n_items, 300)) for _ in range(3)]
    uid, hist, pos = ofs.eval_index(n_users, pairs, 2)
    ub, ib = shard_bounds(n_users, world), shard_bounds(n_items, world)
    seen_hist, seen_pos, total_own = [], [], 0
    for r in range(world):
        idx = ShardedEvalIndex.from_global(uid, hist, pos, ub, ib, r, "cpu")
        assert sum(idx.owner_counts) == len(uid)
        total_own += idx.n_own
        hp, hi = idx.hist_indptr.numpy(), idx.hist_indices.numpy()
        assert ((hi >= ib[r]) & (hi < ib[r + 1])).all()
        rows = np.repeat(np.arange(len(uid)), np.diff(hp))
        seen_hist += list(zip(rows.tolist(), hi.tolist()))
        a = sum(idx.owner_counts[:r])
        pp, pi = idx.pos_indptr.numpy(), idx.pos_indices.numpy()
        rows = np.repeat(np.arange(idx.n_own), np.diff(pp)) + a
        seen_pos += list(zip(rows.tolist(), pi.tolist()))
        assert ((uid[a:a + idx.n_own] >= ub[r]) & (uid[a:a + idx.n_own] < ub[r + 1])).all()
    assert total_own == len(uid)
    all_hist = list(zip(np.repeat(np.arange(len(uid)), np.diff(hist[0])).tolist(), hist[1].tolist()))
    all_pos = list(zip(np.repeat(np.arange(len(uid)), np.diff(pos[0])).tolist(), pos[1].tolist()))
    assert sorted(seen_hist) == sorted(all_hist) and sorted(seen_pos) == sorted(all_pos)


def _ckpt_rank(rank, world):
    """Checkpoint interop of the sharded model on CPU tensors (no kernel runs here): the gathered state dicts use
    the reference's parameter names and torch.optim.Adam's layout, and loading them back on a different world
    layout reproduces every shard."""
    from recbole_b200.dist import Comm, ShardedBPR
    comm = Comm()
    dev = torch.device("cpu")
    rng = np.random.default_rng(3)
    n_users, n_items, d = 103, 57, 8                     # not multiples of the world size: padded tail blocks
    U0 = rng.standard_normal((n_users, d)).astype(np.float32)
    V0 = rng.standard_normal((n_items, d)).astype(np.float32)
    m = ShardedBPR(n_users, n_items, d, comm, dev, U_full=U0, V_full=V0)
    m.build_optimizer("adam", lr=3e-3, weight_decay=1e-4)
    m.optim.step = 7
    for k in ("mU", "vU", "mV", "vV"):                   # recognisable moments: row index + a per-tensor offset
        full = (U0 if k.endswith("U") else V0) * 0 + np.arange((n_users if k.endswith("U") else n_items))[:, None] + ord(k[0])
        lo, hi = (m.u_lo, m.u_hi) if k.endswith("U") else (m.i_lo, m.i_hi)
        m.state[k][: hi - lo] = torch.from_numpy(full[lo:hi].astype(np.float32))
    sd, osd = m.state_dict(), m.optimizer_state_dict()
    assert sorted(sd) == ["item_embedding.weight", "user_embedding.weight"]
    assert np.array_equal(sd["user_embedding.weight"].numpy(), U0) and np.array_equal(sd["item_embedding.weight"].numpy(), V0)
    assert float(osd["state"][0]["step"]) == 7 and osd["param_groups"][0]["lr"] == 3e-3
    assert tuple(osd["state"][1]["exp_avg"].shape) == (n_items, d)
    assert float(osd["state"][0]["exp_avg_sq"][50, 0]) == 50 + ord("v")
    # a plain torch.optim.Adam accepts the optimizer dict (same layout as the reference's checkpoint)
    pu, pv = torch.nn.Parameter(sd["user_embedding.weight"].clone()), torch.nn.Parameter(sd["item_embedding.weight"].clone())
    ref = torch.optim.Adam([pu, pv], lr=1.0)
    ref.load_state_dict({k: v for k, v in osd.items() if k != "fused_kind"})
    assert ref.param_groups[0]["lr"] == 3e-3 and ref.param_groups[0]["weight_decay"] == 1e-4
    assert torch.equal(ref.state[pv]["exp_avg"], osd["state"][1]["exp_avg"])
    # round trip into a fresh model
    m2 = ShardedBPR(n_users, n_items, d, comm, dev, U_full=U0 * 0, V_full=V0 * 0)
    m2.build_optimizer("adam")
    m2.load_state_dict(sd)
    m2.load_optimizer_state_dict(osd)
    assert torch.equal(m2.U, m.U) and torch.equal(m2.V, m.V) and m2.optim.step == 7 and m2.optim.lr == 3e-3
    for k in ("mU", "vU", "mV", "vV"):
        assert torch.equal(m2.state[k], m.state[k])
    return True


def test_sharded_checkpoint_interop_world2():
    assert run_ranks(_ckpt_rank, 2) == [True, True]


def test_sharded_checkpoint_interop_world3():
    assert run_ranks(_ckpt_rank, 3) == [True, True, True]


def _lazy_rank(rank, world):
    """adam_lazy on the sharded model: the exchanges that can keep item rows current accept it (and get a `last` array
    for the local user rows only), the sparse all-to-all exchange refuses it; an optimizer state loaded from a
    checkpoint counts as current at its step."""
    import pytest
    from recbole_b200.dist import Comm, ShardedBPR
    comm = Comm()
    dev = torch.device("cpu")
    m = ShardedBPR(90, 70, 16, comm, dev, exchange="dense")
    m.build_optimizer("adam_lazy", lr=1e-3)
    assert m.state["lastU"].shape == (m.U.shape[0],) and m.state["lastU"].dtype == torch.int32 and "lastV" not in m.state
    sd = m.optimizer_state_dict()             # nothing stepped yet: flush() is a no-op, the gathers run over gloo
    assert sd["fused_kind"] == "adam_lazy" and tuple(sd["state"][1]["exp_avg"].shape) == (70, 16)
    sd["state"][0]["step"] = torch.tensor(7.0)
    sd["param_groups"][0]["lr"] = 5e-4
    m.load_optimizer_state_dict(sd)
    assert m.optim.step == 7 and m.optim.lr == 5e-4 and int(m.state["lastU"].min()) == 7
    s = ShardedBPR(90, 70, 16, comm, dev, exchange="sparse")
    with pytest.raises(ValueError, match="sparse"):
        s.build_optimizer("adam_lazy")
    s.build_optimizer("adam")
    return True


def test_sharded_adam_lazy_host_logic_world2():
    assert run_ranks(_lazy_rank, 2) == [True, True]
