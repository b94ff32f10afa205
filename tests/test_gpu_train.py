"""GPU parity: fused BPR training step (through the C ABI) vs the oracle and the reference golden
vectors.  Tolerances: loss and updated embeddings within 1e-5 relative in fp32 (north_star)."""
import numpy as np
import pytest
import torch

from oracle import bpr as obpr
from oracle import optim as ooptim

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _run_steps(g, dim, opt_name, kind, lr, wd, steps=3):
    from recbole_b200 import ops
    from gpu_util import adam_state, t
    key = "d%d_%s_" % (dim, opt_name)
    U, V = t(g[key + "U0"]), t(g[key + "V0"])
    st = adam_state(U, V, lazy=(kind == "adam_lazy")) if kind != "sgd" else {}
    opt = ops.Optim(kind, lr=lr, weight_decay=wd)
    loss = torch.zeros(1, device=U.device)
    acc = torch.zeros(1, dtype=torch.float64, device=U.device)
    ws = ops.bpr_workspace(len(g["d%d_user_id0" % dim]), dim, U.device)
    out = []
    for s in range(steps):
        u, p, n = (t(g["d%d_%s%d" % (dim, f, s)]) for f in ("user_id", "item_id", "neg_item_id"))
        ops.bpr_train_step(U, V, st, u, p, n, opt, loss, acc, ws)
        if kind == "adam_lazy":
            # read the tables as evaluation would: flush a COPY so the lazy state keeps running
            Uc, Vc = U.clone(), V.clone()
            stc = {k: v.clone() for k, v in st.items()}
            ops.adam_lazy_flush(Uc, stc["mU"], stc["vU"], stc["lastU"], opt)
            ops.adam_lazy_flush(Vc, stc["mV"], stc["vV"], stc["lastV"], opt)
            out.append((float(loss.item()), Uc.cpu().numpy(), Vc.cpu().numpy()))
        else:
            out.append((float(loss.item()), U.cpu().numpy(), V.cpu().numpy()))
    ws.check_flags()
    return out, float(acc.item()), st


@pytest.mark.parametrize("dim", [16, 64, 128])
def test_sgd_vs_reference_golden(golden, dim):
    """Row-sparse SGD == the reference's dense SGD (zero-gradient rows do not move)."""
    from gpu_util import rel_err
    g = golden("bpr_steps.npz")
    key = "d%d_sgd_" % dim
    out, acc, _ = _run_steps(g, dim, "sgd", "sgd", 0.5, 0.0)
    tot = 0.0
    for s, (loss, U, V) in enumerate(out):
        ref = float(g[key + "loss%d" % s])
        assert abs(loss - ref) <= TOL * abs(ref)
        assert rel_err(U, g[key + "U%d" % (s + 1)]) < TOL
        assert rel_err(V, g[key + "V%d" % (s + 1)]) < TOL
        tot += loss
    assert abs(acc - tot) < 1e-6


@pytest.mark.parametrize("dim", [16, 64, 128])
@pytest.mark.parametrize("opt_name,wd", [("adam", 0.0), ("adam_wd", 1e-3)])
def test_adam_lazy_vs_reference_dense_adam(golden, dim, opt_name, wd):
    """adam_lazy replays the zero-gradient steps a row missed => same trajectory as the
    reference's dense torch.optim.Adam over 3 steps (rows drop in and out of the batches)."""
    from gpu_util import rel_err
    g = golden("bpr_steps.npz")
    key = "d%d_%s_" % (dim, opt_name)
    out, _, st = _run_steps(g, dim, opt_name, "adam_lazy", 1e-2, wd)
    for s, (loss, U, V) in enumerate(out):
        ref = float(g[key + "loss%d" % s])
        assert abs(loss - ref) <= TOL * abs(ref), (s, loss, ref)
        assert rel_err(U, g[key + "U%d" % (s + 1)]) < TOL, s
        assert rel_err(V, g[key + "V%d" % (s + 1)]) < TOL, s


@pytest.mark.parametrize("dim", [16, 64, 128])
def test_adam_first_step_is_dense_adam(golden, dim):
    """From zero moments the row-sparse Adam step IS the reference's dense step."""
    from gpu_util import rel_err
    g = golden("bpr_steps.npz")
    key = "d%d_adam_" % dim
    out, _, _ = _run_steps(g, dim, "adam", "adam", 1e-2, 0.0, steps=1)
    loss, U, V = out[0]
    assert abs(loss - float(g[key + "loss0"])) <= TOL * abs(float(g[key + "loss0"]))
    assert rel_err(U, g[key + "U1"]) < TOL and rel_err(V, g[key + "V1"]) < TOL


@pytest.mark.parametrize("dim,B,n_users,n_items", [(64, 20000, 3000, 2000), (128, 4096, 100000, 50000),
                                                   (32, 777, 50, 40), (256, 3000, 500, 400), (64, 1, 10, 10)])
@pytest.mark.parametrize("kind", ["adam", "sgd", "adam_lazy"])
def test_random_batches_vs_oracle(dim, B, n_users, n_items, kind):
    """Heavy duplication (Zipf ids), runs that straddle tiles, several steps; oracle = dense
    semantics for sgd / adam_lazy, row-sparse semantics for adam."""
    from recbole_b200 import ops
    from gpu_util import adam_state, rel_err, t
    rng = np.random.default_rng(dim * 7 + B)
    U0 = (rng.standard_normal((n_users, dim)) * 0.3).astype(np.float32)
    V0 = (rng.standard_normal((n_items, dim)) * 0.3).astype(np.float32)
    st_o = obpr.new_state(U0, V0)
    U, V = t(U0), t(V0)
    st = adam_state(U, V, lazy=(kind == "adam_lazy")) if kind != "sgd" else {}
    lr = 0.05 if kind == "sgd" else 2e-3  # error of the normalised Adam step scales with lr
    opt = ops.Optim(kind, lr=lr)
    loss = torch.zeros(1, device=U.device)
    ws = ops.bpr_workspace(B, dim, U.device)

    def zipf(n, size):
        return np.minimum((np.exp(rng.random(size) * np.log(n - 1))).astype(np.int64), n - 1).clip(1)

    for s in range(3):
        u, p, n = zipf(n_users, B), zipf(n_items, B), rng.integers(1, n_items, B)
        ops.bpr_train_step(U, V, st, t(u), t(p), t(n), opt, loss, None, ws)
        lo = obpr.bpr_train_step(st_o, u, p, n, s + 1, optimizer="sgd" if kind == "sgd" else "adam", lr=lr,
                                 dense=(kind != "adam"))
        assert abs(float(loss.item()) - lo) <= TOL * abs(lo)
    if kind == "adam_lazy":
        ops.adam_lazy_flush(U, st["mU"], st["vU"], st["lastU"], opt)
        ops.adam_lazy_flush(V, st["mV"], st["vV"], st["lastV"], opt)
    ws.check_flags()
    assert rel_err(U.cpu().numpy(), st_o["U"]) < TOL
    assert rel_err(V.cpu().numpy(), st_o["V"]) < TOL
    # element-wise (every element against its own magnitude, floored at 1e-3 of the table's largest)
    from gpu_util import elem_rel_err
    assert elem_rel_err(U.cpu().numpy(), st_o["U"]) < 10 * TOL
    assert elem_rel_err(V.cpu().numpy(), st_o["V"]) < 10 * TOL
    if kind != "sgd":
        assert rel_err(st["mV"].cpu().numpy(), st_o["mV"]) < 1e-4
        assert rel_err(st["vV"].cpu().numpy(), st_o["vV"]) < 1e-4


def test_step_is_deterministic():
    """No float atomics: two runs from the same state are bit-identical."""
    from recbole_b200 import ops
    from gpu_util import adam_state, t
    rng = np.random.default_rng(1)
    n_users, n_items, dim, B = 2000, 300, 64, 50000
    U0 = rng.standard_normal((n_users, dim)).astype(np.float32) * 0.2
    V0 = rng.standard_normal((n_items, dim)).astype(np.float32) * 0.2
    u, p, n = rng.integers(1, n_users, B), rng.integers(1, 20, B), rng.integers(1, n_items, B)
    res = []
    for _ in range(2):
        U, V = t(U0), t(V0)
        st = adam_state(U, V)
        opt = ops.Optim("adam", lr=1e-2)
        loss = torch.zeros(1, device=U.device)
        ops.bpr_train_step(U, V, st, t(u), t(p), t(n), opt, loss, None, ops.bpr_workspace(B, dim, U.device))
        res.append((U.cpu().numpy(), V.cpu().numpy(), loss.item()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]


def test_out_of_range_id_is_reported():
    from recbole_b200 import ops
    from gpu_util import t
    U, V = torch.zeros(10, 16, device="cuda"), torch.zeros(10, 16, device="cuda")
    ws = ops.bpr_workspace(4, 16, U.device)
    loss = torch.zeros(1, device="cuda")
    ops.bpr_train_step(U, V, {}, t(np.array([1, 2, 3, 10])), t(np.array([1, 2, 3, 4])), t(np.array([1, 2, 3, 4])),
                       ops.Optim("sgd"), loss, None, ws)
    with pytest.raises(IndexError):
        ws.check_flags()


def test_unsupported_dim_fails_loudly():
    from recbole_b200 import ops
    from recbole_b200._lib import RB2Error
    from gpu_util import t
    U, V = torch.zeros(10, 10, device="cuda"), torch.zeros(10, 10, device="cuda")
    ws = ops.bpr_workspace(4, 16, U.device)
    with pytest.raises(RB2Error, match="not supported"):
        ops.bpr_train_step(U, V, {}, t(np.array([1])), t(np.array([1])), t(np.array([1])), ops.Optim("sgd"),
                           torch.zeros(1, device="cuda"), None, ws)


@pytest.mark.parametrize("dim", [16, 64, 128])
def test_loss_and_predict_vs_reference_golden(golden, dim):
    from recbole_b200 import ops
    from gpu_util import rel_err, t
    g = golden("bpr_steps.npz")
    key = "d%d_adam_" % dim
    U, V = t(g[key + "U0"]), t(g[key + "V0"])
    u, p, n = (t(g["d%d_%s0" % (dim, f)]) for f in ("user_id", "item_id", "neg_item_id"))
    loss = torch.zeros(1, device=U.device)
    ops.bpr_loss(U, V, u, p, n, loss, ops.bpr_workspace(u.numel(), dim, U.device))
    assert abs(loss.item() - float(g[key + "loss0"])) <= TOL * abs(float(g[key + "loss0"]))
    # predict after 3 reference steps
    U3, V3 = t(g[key + "U3"]), t(g[key + "V3"])
    pred = ops.gather_dot(U3, V3, u, p).cpu().numpy()
    assert rel_err(pred, g[key + "pred"]) < TOL


def test_full_size_cfg2_step_properties():
    """BASELINE cfg2 shape (138 494 x 26 745, d=64, B=2^20): size-independent properties --
    finite loss near log 2 at init scale, untouched rows bit-identical, touched rows all moved,
    and a sampled subset of rows equal to the oracle's row-sparse Adam."""
    from recbole_b200 import ops
    from gpu_util import adam_state, rel_err
    n_users, n_items, dim, B = 138494, 26745, 64, 1 << 20
    gen = torch.Generator(device="cuda")
    gen.manual_seed(2020)
    U = torch.randn(n_users, dim, device="cuda", generator=gen) * (2.0 / (n_users + dim)) ** 0.5
    V = torch.randn(n_items, dim, device="cuda", generator=gen) * (2.0 / (n_items + dim)) ** 0.5
    U0, V0 = U.clone(), V.clone()
    u = torch.randint(1, n_users, (B,), device="cuda", generator=gen)
    p = torch.randint(1, n_items, (B,), device="cuda", generator=gen)
    n = torch.randint(1, n_items, (B,), device="cuda", generator=gen)
    u[: B // 2] = u[: B // 2] % 5000 + 1  # half of the batch hits 5000 users: long runs
    st = adam_state(U, V)
    opt = ops.Optim("adam", lr=1e-3)
    loss = torch.zeros(1, device="cuda")
    ws = ops.bpr_workspace(B, dim, U.device)
    ops.bpr_train_step(U, V, st, u, p, n, opt, loss, None, ws)
    ws.check_flags()
    assert abs(loss.item() - np.log(2.0)) < 1e-2
    touched_u = torch.zeros(n_users, dtype=torch.bool, device="cuda")
    touched_u[u] = True
    moved = (U != U0).any(dim=1)
    assert torch.equal(moved, touched_u)
    touched_i = torch.zeros(n_items, dtype=torch.bool, device="cuda")
    touched_i[p] = True
    touched_i[n] = True
    assert torch.equal((V != V0).any(dim=1), touched_i)
    # oracle on the CPU for the whole step (dense grads are small at this shape)
    st_o = obpr.new_state(U0.cpu().numpy(), V0.cpu().numpy())
    lo = obpr.bpr_train_step(st_o, u.cpu().numpy(), p.cpu().numpy(), n.cpu().numpy(), 1, optimizer="adam", lr=1e-3,
                             dense=False)
    assert abs(loss.item() - lo) <= TOL * abs(lo)
    assert rel_err(U.cpu().numpy(), st_o["U"]) < TOL
    assert rel_err(V.cpu().numpy(), st_o["V"]) < TOL


@pytest.mark.parametrize("dim,B,n_users,n_items", [(64, 20000, 3000, 2000), (128, 4096, 100000, 50000),
                                                   (16, 777, 50, 40), (256, 3000, 500, 400)])
@pytest.mark.parametrize("kind", ["adam", "sgd"])
def test_peer_memory_step_one_rank_vs_oracle(dim, B, n_users, n_items, kind):
    """rb2_bpr_train_step_p2p with a world of one (every "peer" pointer is local): the plan, the cache fetch, the
    pushes into the gradient slots, both barriers and the owner update against the oracle's row-sparse step."""
    from recbole_b200.dist import Comm, ShardedBPR
    from gpu_util import rel_err, t, dev
    rng = np.random.default_rng(dim * 11 + B)
    U0 = (rng.standard_normal((n_users, dim)) * 0.3).astype(np.float32)
    V0 = (rng.standard_normal((n_items, dim)) * 0.3).astype(np.float32)
    st_o = obpr.new_state(U0, V0)
    m = ShardedBPR(n_users, n_items, dim, Comm(), dev(), U_full=U0, V_full=V0, exchange="p2p")
    lr = 0.05 if kind == "sgd" else 2e-3
    m.build_optimizer(kind, lr=lr)

    def zipf(n, size):
        return np.minimum((np.exp(rng.random(size) * np.log(n - 1))).astype(np.int64), n - 1).clip(1)

    for s in range(3):
        u, p, n = zipf(n_users, B), zipf(n_items, B), rng.integers(1, n_items, B)
        loss = m.train_step(t(u), t(p), t(n))
        lo = obpr.bpr_train_step(st_o, u, p, n, s + 1, optimizer=kind, lr=lr, dense=False)
        assert abs(float(loss.item()) - lo) <= TOL * abs(lo)
    m.check_flags()
    assert m.last_exchange == "p2p"
    assert rel_err(m.U.cpu().numpy(), st_o["U"]) < TOL
    assert rel_err(m.V.cpu().numpy()[:n_items], st_o["V"]) < TOL


def test_peer_memory_step_prepares_the_next_batch(golden):
    """next_batch: the following batch's keys and sorts are computed inside the current call (between the two halves
    of barrier B) into the workspace's other slot, and the next call skips them.  Five chained steps == the same five
    steps without the hand-over, bit for bit; a batch that was NOT announced is still sorted by its own call."""
    from recbole_b200.dist import Comm, ShardedBPR
    from gpu_util import t, dev
    rng = np.random.default_rng(5)
    n_users, n_items, dim, B = 5000, 3000, 128, 8192
    U0 = (rng.standard_normal((n_users, dim)) * 0.3).astype(np.float32)
    V0 = (rng.standard_normal((n_items, dim)) * 0.3).astype(np.float32)
    batches = [tuple(t(rng.integers(1, hi, B)) for hi in (n_users, n_items, n_items)) for _ in range(5)]

    def run(chain):
        m = ShardedBPR(n_users, n_items, dim, Comm(), dev(), U_full=U0, V_full=V0, exchange="p2p")
        m.build_optimizer("adam", lr=2e-3)
        m.ids_ready = True
        losses = []
        for s, b in enumerate(batches):
            nxt = batches[s + 1] if (chain and s + 1 < len(batches) and s != 2) else None     # step 2 announces nothing
            losses.append(float(m.train_step(*b, next_batch=nxt).item()))
        m.check_flags()
        return losses, m.U.clone(), m.V.clone(), m

    l0, U_a, V_a, _ = run(False)
    l1, U_b, V_b, m = run(True)
    assert l0 == l1
    assert torch.equal(U_a, U_b) and torch.equal(V_a, V_b)
    assert m.arena.prepared_for is None            # the last call announced nothing


@pytest.mark.parametrize("kind", ["adam", "adam_lazy"])
def test_trained_tables_loss_and_step_vs_oracle(kind):
    """A converged model: x = u.(v_pos - v_neg) is 2 ... 6 for nearly every sample, so the loss terms -log(sigmoid(x)) are
    1e-1 ... 2e-3 and the gradient factor (1 - sigmoid) is small -- the regime where approximate exp / log (absolute error
    2^-21) would cost 1e-4 of the loss.  Loss within 1e-5 of the oracle's fp32 chain at every step, tables element-wise."""
    from recbole_b200 import ops
    from gpu_util import adam_state, elem_rel_err, rel_err, t
    rng = np.random.default_rng(77)
    n_users, n_items, dim, B = 5000, 4000, 64, 16384
    e = rng.standard_normal(dim)
    e /= np.linalg.norm(e)
    U0 = (1.4 * e + 0.05 * rng.standard_normal((n_users, dim))).astype(np.float32)
    V0 = (0.05 * rng.standard_normal((n_items, dim))).astype(np.float32)
    half = n_items // 2
    V0[:half] += (1.5 * e).astype(np.float32)          # the items users like ...
    V0[half:] -= (1.5 * e).astype(np.float32)          # ... and the ones they do not
    st_o = obpr.new_state(U0, V0)
    U, V = t(U0), t(V0)
    st = adam_state(U, V, lazy=(kind == "adam_lazy"))
    opt = ops.Optim(kind, lr=1e-3)
    loss = torch.zeros(1, device=U.device)
    ws = ops.bpr_workspace(B, dim, U.device)
    for s in range(3):
        u, p, n = rng.integers(1, n_users, B), rng.integers(1, half, B), rng.integers(half, n_items, B)
        ops.bpr_train_step(U, V, st, t(u), t(p), t(n), opt, loss, None, ws)
        lo = obpr.bpr_train_step(st_o, u, p, n, s + 1, optimizer="adam", lr=1e-3, dense=(kind == "adam_lazy"))
        assert 1e-3 < lo < 0.1, lo                      # the regime this test is about
        assert abs(float(loss.item()) - lo) <= TOL * abs(lo), (s, float(loss.item()), lo)
    if kind == "adam_lazy":
        ops.adam_lazy_flush(U, st["mU"], st["vU"], st["lastU"], opt)
        ops.adam_lazy_flush(V, st["mV"], st["vV"], st["lastV"], opt)
    for a, b in ((U, st_o["U"]), (V, st_o["V"])):
        assert rel_err(a.cpu().numpy(), b) < TOL
        assert elem_rel_err(a.cpu().numpy(), b) < 10 * TOL


@pytest.mark.parametrize("dim", [64, 128])
def test_adam_lazy_long_gaps_vs_dense_oracle(dim):
    """30 steps of small batches: most rows are untouched for many steps in a row, so every catch-up path runs --
    short gaps replayed in registers inside the fused user-side kernel (user rows, single items), long gaps and repeated
    items in the separate catch-up pass, the final flush -- against the oracle's DENSE Adam (every row at every step)."""
    from recbole_b200 import ops
    from gpu_util import adam_state, t
    rng = np.random.default_rng(dim)
    n_users, n_items, B, steps = 900, 700, 96, 30
    U0 = (rng.standard_normal((n_users, dim)) * 0.3).astype(np.float32)
    V0 = (rng.standard_normal((n_items, dim)) * 0.3).astype(np.float32)
    st_o = obpr.new_state(U0, V0)
    U, V = t(U0), t(V0)
    st = adam_state(U, V, lazy=True)
    opt = ops.Optim("adam_lazy", lr=2e-3)
    loss = torch.zeros(1, device=U.device)
    ws = ops.bpr_workspace(B, dim, U.device)
    for s in range(steps):
        u, p = rng.integers(1, n_users, B), rng.integers(1, n_items, B)
        n = (p + rng.integers(1, n_items - 1, B) - 1) % (n_items - 1) + 1          # never equal to p
        p[: B // 4] = p[0]                                                          # a repeated item in every batch
        n = np.where(n == p, n % (n_items - 1) + 1, n)
        ops.bpr_train_step(U, V, st, t(u), t(p), t(n), opt, loss, None, ws)
        lo = obpr.bpr_train_step(st_o, u, p, n, s + 1, optimizer="adam", lr=2e-3, dense=True)
        assert abs(float(loss.item()) - lo) <= TOL * abs(lo), (s, float(loss.item()), lo)
    ops.adam_lazy_flush(U, st["mU"], st["vU"], st["lastU"], opt)
    ops.adam_lazy_flush(V, st["mV"], st["vV"], st["lastV"], opt)
    ws.check_flags()
    for got, want in ((U, st_o["U"]), (V, st_o["V"]), (st["mU"], st_o["mU"]), (st["vV"], st_o["vV"])):
        d = np.abs(got.cpu().numpy() - want)
        bad = d > TOL * np.abs(want).max()
        assert bad.sum() <= 2 and d.max() <= 20 * TOL * np.abs(want).max(), (int(bad.sum()), float(d.max()))
