"""GPU parity: full-sort scorer + top-K + metrics + sampler (through the C ABI) vs the oracle and
the reference golden vectors.  Bar: top-K ids and metrics bit-exact (ties -> lower item id)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import fullsort as ofs
from oracle import metrics as ometrics
from oracle import sampler as osampler

pytestmark = pytest.mark.gpu


def _csrs(g):
    uid = g["uid_list"]
    n_users = int(max(g["used_user"].max(), uid.max())) + 1
    remap = -np.ones(n_users, dtype=np.int64)
    remap[uid] = np.arange(len(uid))
    pos = ofs.build_csr(len(uid), remap[g["pos_user"]], g["pos_item"])
    keep = remap[g["used_user"]] >= 0
    used = ofs.build_csr(len(uid), remap[g["used_user"][keep]], g["used_item"][keep])
    hp, hi = [0], []
    for r in range(len(uid)):
        h = np.setdiff1d(used[1][used[0][r]:used[0][r + 1]], pos[1][pos[0][r]:pos[0][r + 1]])
        hi.append(h)
        hp.append(hp[-1] + len(h))
    return (np.array(hp), np.concatenate(hi).astype(np.int64)), pos


@pytest.mark.parametrize("name", ["fullsort_small.npz", "fullsort_ml100k.npz"])
def test_fullsort_vs_reference_golden(golden, name):
    from recbole_b200 import ops
    from gpu_util import t
    g = golden(name)
    hist, pos = _csrs(g)
    K, N = int(g["topk"].max()), int(g["n_items"])
    U, V = t(g["U"]), t(g["V"])
    ids, sc = ops.fullsort_topk(U, t(g["uid_list"]), V, K, t(hist[0]), t(hist[1]))
    o_ids, o_sc = ofs.full_sort_topk(g["U"], g["V"], g["uid_list"], hist[0], hist[1], K)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)          # bit-exact ids vs the oracle
    np.testing.assert_array_equal(sc.cpu().numpy(), o_sc)            # bit-exact canonical scores
    m = ops.topk_metrics(ids, t(pos[0]), t(pos[1]), N, want_hit=True, want_ref_idx=True)
    np.testing.assert_array_equal(m["hit"].cpu().numpy().astype(bool), ofs.hits(o_ids, pos[0], pos[1]))
    # the reference's own matrix (TopKEvaluator.collect) in its swapped+flipped coordinates
    np.testing.assert_array_equal(m["ref_idx"].cpu().numpy(), g["topk_matrix"])
    # metric dict == the reference's Trainer.evaluate output
    names = [str(x) for x in g["metrics"]]
    sums = m["sums"].cpu().numpy() / len(g["uid_list"])
    res = {}
    for nme in names:
        for k in g["topk"].tolist():
            res["%s@%d" % (nme, k)] = round(float(sums[ometrics.METRIC_ORDER.index(nme)][k - 1]), 4)
    assert res == dict(zip(g["result_keys"].tolist(), g["result_vals"].tolist()))
    # per-metric sums bit-close to numpy's float64 (same per-user values, different sum order)
    ref_full = ometrics.calculate_metrics(ofs.hits(o_ids, pos[0], pos[1]), np.diff(pos[0]), list(ometrics.METRIC_ORDER))
    np.testing.assert_allclose(sums, ref_full, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("dim", [16, 32, 64, 128])
@pytest.mark.parametrize("nq,N,K", [(300, 5000, 10), (1, 37, 5), (1000, 257, 20), (130, 9, 10)])
def test_fullsort_ties_and_edges(dim, nq, N, K):
    """Embeddings on a dyadic grid => every dot product is exact => many exact ties; the order
    must be (score desc, item id asc).  N < K + masked exercises the empty-slot path."""
    from recbole_b200 import ops
    from gpu_util import t
    rng = np.random.default_rng(dim + nq)
    Q = (rng.integers(-4, 5, (nq, dim)) / 8.0).astype(np.float32)
    V = (rng.integers(-4, 5, (N, dim)) / 8.0).astype(np.float32)
    hp = [0]
    hi = []
    for r in range(nq):
        h = np.sort(rng.choice(np.arange(1, N), size=min(rng.integers(0, 6), N - 1), replace=False))
        hi.append(h)
        hp.append(hp[-1] + len(h))
    hp, hi = np.array(hp, dtype=np.int64), np.concatenate(hi).astype(np.int64)
    ids, sc = ops.fullsort_topk(t(Q), None, t(V), K, t(hp), t(hi))
    o_ids, o_sc = ofs.full_sort_topk(Q, V, np.arange(nq), hp, hi, K)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), o_sc)
    # no history at all (sequential full-sort batches, sequential_dataloader.py:269-280)
    ids2, _ = ops.fullsort_topk(t(Q), None, t(V), K)
    o2, _ = ofs.full_sort_topk(Q, V, np.arange(nq), np.zeros(nq + 1, np.int64), np.zeros(0, np.int64), K)
    np.testing.assert_array_equal(ids2.cpu().numpy(), o2)


def test_fullsort_shards_and_merge():
    """Row-sharded item table: per-shard top-K + merge == single-table top-K."""
    from recbole_b200 import ops
    from gpu_util import t
    rng = np.random.default_rng(5)
    nq, N, dim, K, G = 500, 4001, 64, 10, 4
    Q = rng.standard_normal((nq, dim)).astype(np.float32)
    V = (rng.integers(-8, 9, (N, dim)) / 16.0).astype(np.float32)
    hp = np.arange(0, 3 * nq + 1, 3, dtype=np.int64)
    hi = np.sort(rng.integers(1, N, (nq, 3)), axis=1).reshape(-1).astype(np.int64)
    full_ids, full_sc = ops.fullsort_topk(t(Q), None, t(V), K, t(hp), t(hi))
    bounds = np.linspace(0, N, G + 1).astype(int)
    parts_i, parts_s = [], []
    for gidx in range(G):
        a, b = bounds[gidx], bounds[gidx + 1]
        i, s = ops.fullsort_topk(t(Q), None, t(V[a:b]), K, t(hp), t(hi), item_base=int(a))
        parts_i.append(i)
        parts_s.append(s)
    m_ids, m_sc = ops.topk_merge(torch.stack(parts_i), torch.stack(parts_s))
    assert torch.equal(m_ids, full_ids) and torch.equal(m_sc, full_sc)
    o_ids, _ = ofs.full_sort_topk(Q, V, np.arange(nq), hp, hi, K)
    np.testing.assert_array_equal(full_ids.cpu().numpy(), o_ids)


def test_metrics_known_answer_on_device():
    """reference tests/metrics/test_topk_metrics.py:15-79 pushed through the device reducer."""
    from recbole_b200 import ops
    from gpu_util import t
    pos_len = [1, 3, 4, 2]
    hit = np.array([[0, 0, 0], [1, 1, 1], [1, 0, 1], [0, 0, 1]])
    # build top-k ids / positives that realise this hit pattern: positives are ids 100.., misses 1..
    indptr, indices, topk = [0], [], []
    for r in range(4):
        p = list(range(100 + 10 * r, 100 + 10 * r + pos_len[r]))
        indices += p
        indptr.append(len(indices))
        row, nxt = [], 0
        for j in range(3):
            if hit[r, j]:
                row.append(p[nxt]); nxt += 1
            else:
                row.append(1 + j)
        topk.append(row)
    m = ops.topk_metrics(t(np.array(topk, dtype=np.int64)), t(np.array(indptr, dtype=np.int64)),
                         t(np.array(indices, dtype=np.int64)), 1000, want_hit=True)
    np.testing.assert_array_equal(m["hit"].cpu().numpy(), hit)
    ref = ometrics.calculate_metrics(hit.astype(bool), np.array(pos_len), list(ometrics.METRIC_ORDER))
    np.testing.assert_allclose(m["sums"].cpu().numpy() / 4.0, ref, rtol=1e-15, atol=0)


def test_sampler_ref_stream_vs_reference_golden(golden):
    from recbole_b200 import ops
    from gpu_util import t
    g = golden("sampler.npz")
    used = ofs.build_csr(int(g["n_users"]), g["used_user"], g["used_item"])
    rl, ip, ix = t(g["random_list"]), t(used[0]), t(used[1])
    L = len(g["random_list"])
    for c in range(4):
        out, pr = ops.neg_sample_ref(t(g["users%d" % c]), int(g["num%d" % c]), rl, int(g["pr_before%d" % c]), ip, ix)
        np.testing.assert_array_equal(out.cpu().numpy(), g["out%d" % c])
        assert pr % L == int(g["pr_after%d" % c]) % L


def test_sampler_hash_stream_vs_oracle(golden):
    from recbole_b200 import ops
    from gpu_util import t
    g = golden("sampler.npz")
    n_items = int(g["n_items"])
    used = ofs.build_csr(int(g["n_users"]), g["used_user"], g["used_item"])
    users = g["users1"]
    out = ops.neg_sample_hash(t(users), 3, n_items, t(used[0]), t(used[1]), 2021, 7).cpu().numpy()
    np.testing.assert_array_equal(out, osampler.hash_sample(users, 3, n_items, used[0], used[1], 2021, 7))
    # dense user: deterministic scan path
    n_items = 12
    ui = np.array([i for i in range(1, n_items) if i not in (4, 9)], dtype=np.int64)
    ip = np.array([0, 0, len(ui)], dtype=np.int64)
    keys = np.ones(50, dtype=np.int64)
    out = ops.neg_sample_hash(t(keys), 1, n_items, t(ip), t(ui), 1, 1).cpu().numpy()
    np.testing.assert_array_equal(out, osampler.hash_sample(keys, 1, n_items, ip, ui, 1, 1))


def test_full_size_cfg2_eval_properties():
    """BASELINE cfg2 shape (138 494 users x 26 745 items, d=64, K=10): size-independent
    properties on all users + bit-exact ids on a random subset vs the oracle."""
    from recbole_b200 import ops
    from gpu_util import t
    n_users, n_items, dim, K = 138494, 26745, 64, 10
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7)
    U = torch.randn(n_users, dim, device="cuda", generator=gen) * 0.1
    V = torch.randn(n_items, dim, device="cuda", generator=gen) * 0.1
    users = torch.arange(1, n_users, device="cuda")
    nq = users.numel()
    hist_cols = torch.sort(torch.randint(1, n_items, (nq, 20), device="cuda", generator=gen), dim=1).values
    hp = torch.arange(0, 20 * nq + 1, 20, device="cuda", dtype=torch.int64)
    hi = hist_cols.reshape(-1).contiguous()
    ids, sc = ops.fullsort_topk(U, users, V, K, hp, hi)
    assert (ids >= 1).all() and (ids < n_items).all()
    assert (sc[:, :-1] >= sc[:, 1:]).all()                                    # sorted
    assert not (ids.unsqueeze(2) == hist_cols.unsqueeze(1)).any()            # history masked
    assert (torch.sort(ids, dim=1).values[:, 1:] != torch.sort(ids, dim=1).values[:, :-1]).all()  # distinct
    # idempotence: scoring only the winners returns the same order
    sub = torch.randperm(nq, device="cuda", generator=gen)[:64]
    o_ids, o_sc = ofs.full_sort_topk(U.cpu().numpy(), V.cpu().numpy(), users[sub].cpu().numpy(),
                                     np.arange(0, 20 * 64 + 1, 20), hist_cols[sub].cpu().numpy().reshape(-1), K)
    np.testing.assert_array_equal(ids[sub].cpu().numpy(), o_ids)
    np.testing.assert_array_equal(sc[sub].cpu().numpy(), o_sc)


# ---- tensor-core scorer (tcgen05 + TMA), certified exact ------------------------------------------

def _tc_case(dim, nq, N, K, seed, dyadic=False, hist_per_row=4, scale=1.0):
    rng = np.random.default_rng(seed)
    if dyadic:
        Q = (rng.integers(-4, 5, (nq, dim)) / 8.0).astype(np.float32)
        V = (rng.integers(-4, 5, (N, dim)) / 8.0).astype(np.float32)
    else:
        Q = (rng.standard_normal((nq, dim)) * scale).astype(np.float32)
        V = (rng.standard_normal((N, dim)) * scale).astype(np.float32)
    hp, hi = [0], []
    for r in range(nq):
        h = np.sort(rng.choice(np.arange(1, N), size=min(hist_per_row, N - 1), replace=False))
        hi.append(h)
        hp.append(hp[-1] + len(h))
    return Q, V, np.array(hp, dtype=np.int64), np.concatenate(hi).astype(np.int64)


@pytest.mark.parametrize("dim", [64, 128])
@pytest.mark.parametrize("nq,N,K", [(300, 5000, 10), (128, 256, 10), (1, 1000, 5), (1000, 70001, 10), (77, 300, 16)])
def test_fullsort_tc_random_exact(dim, nq, N, K):
    """bf16 tensor-core filter + fp32 re-score + certificate == the oracle, bit for bit."""
    from recbole_b200 import ops
    from recbole_b200._lib import lib
    from gpu_util import t
    Q, V, hp, hi = _tc_case(dim, nq, N, K, seed=dim + nq + N)
    ids, sc = ops.fullsort_topk(t(Q), None, t(V), K, t(hp), t(hi), mode="tc")
    o_ids, o_sc = ofs.full_sort_topk(Q, V, np.arange(nq), hp, hi, K)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), o_sc)
    # on gaussian data the certificate must hold for (almost) every row: the tensor path did the work
    assert lib.rb2_fullsort_tc_last_fallback_rows() <= max(1, nq // 50)


@pytest.mark.parametrize("dim", [64, 128])
def test_fullsort_tc_ties_fall_back_and_stay_exact(dim):
    """Dyadic grid => massive exact ties around the K'-th candidate => certificates fail => those
    rows are redone in fp32; the answer is still the oracle's."""
    from recbole_b200 import ops
    from gpu_util import t
    Q, V, hp, hi = _tc_case(dim, 200, 3000, 10, seed=3, dyadic=True)
    ids, sc = ops.fullsort_topk(t(Q), None, t(V), 10, t(hp), t(hi), mode="tc")
    o_ids, o_sc = ofs.full_sort_topk(Q, V, np.arange(200), hp, hi, 10)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), o_sc)


def test_fullsort_tc_query_ids_shards_and_golden(golden):
    from recbole_b200 import ops
    from gpu_util import t
    g = golden("fullsort_ml100k.npz")
    hist, pos = _csrs(g)
    ids, sc = ops.fullsort_topk(t(g["U"]), t(g["uid_list"]), t(g["V"]), 10, t(hist[0]), t(hist[1]), mode="tc")
    o_ids, o_sc = ofs.full_sort_topk(g["U"], g["V"], g["uid_list"], hist[0], hist[1], 10)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), o_sc)
    # item shards with a base offset, merged
    N = g["V"].shape[0]
    parts_i, parts_s = [], []
    for a, b in ((0, 700), (700, N)):
        i, s = ops.fullsort_topk(t(g["U"]), t(g["uid_list"]), t(g["V"][a:b]), 10, t(hist[0]), t(hist[1]),
                                 item_base=a, mode="tc")
        parts_i.append(i)
        parts_s.append(s)
    m_ids, m_sc = ops.topk_merge(torch.stack(parts_i), torch.stack(parts_s))
    np.testing.assert_array_equal(m_ids.cpu().numpy(), o_ids)


def test_fullsort_tc_unsupported_shapes_use_fp32_kernel():
    """d = 32 or K > 16 are outside the MMA tiling: the call is served by the exact CUDA-core
    kernel (still on the GPU, still exact)."""
    from recbole_b200 import ops
    from gpu_util import t
    Q, V, hp, hi = _tc_case(32, 50, 500, 10, seed=9)
    ids, _ = ops.fullsort_topk(t(Q), None, t(V), 10, t(hp), t(hi), mode="tc")
    o_ids, _ = ofs.full_sort_topk(Q, V, np.arange(50), hp, hi, 10)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
    Q, V, hp, hi = _tc_case(64, 50, 500, 20, seed=9)
    ids, _ = ops.fullsort_topk(t(Q), None, t(V), 20, t(hp), t(hi), mode="tc")
    o_ids, _ = ofs.full_sort_topk(Q, V, np.arange(50), hp, hi, 20)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)


@pytest.fixture
def tc_variant():
    """Sets the MMA variant of the tensor-core scorer for one test and restores the default."""
    from recbole_b200._lib import lib

    def set_(v):
        assert lib.rb2_fullsort_tc_set_variant(int(v)) == 0
    yield set_
    lib.rb2_fullsort_tc_set_variant(0)


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("kind", ["gauss", "row_scales", "tiny", "low_rank"])
def test_fullsort_tc_variants_exact(tc_variant, variant, kind):
    """1 = bf16 operands / fp32 accumulators, 3 = fp16 operands rescaled by powers of two / FP16
    accumulators (default), 2 = 3 with CTA-pair MMAs: all three return the oracle's top-K bit for bit,
    whatever the magnitudes of the rows (the fp16 path rescales every query row and the item table)."""
    from recbole_b200 import ops
    from gpu_util import t
    tc_variant(variant)
    rng = np.random.default_rng(17 + variant)
    nq, N, d, K = 700, 40001, 128, 10
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    V = rng.standard_normal((N, d)).astype(np.float32)
    if kind == "gauss":
        Q *= 0.1
        V *= 0.1
    elif kind == "row_scales":
        Q *= np.exp(rng.standard_normal((nq, 1)) * 3).astype(np.float32)
        V *= (np.exp(rng.standard_normal((N, 1))) * 1e-3).astype(np.float32)
    elif kind == "tiny":
        Q *= 1e-20
        V *= 1e-12
    else:
        Bm = rng.standard_normal((8, d)).astype(np.float32)
        Q = (rng.standard_normal((nq, 8)).astype(np.float32) @ Bm + 0.05 * Q).astype(np.float32)
        V = (rng.standard_normal((N, 8)).astype(np.float32) @ Bm + 0.05 * V).astype(np.float32)
    hp = np.arange(0, 5 * nq + 1, 5, dtype=np.int64)
    hi = np.sort(rng.integers(1, N, (nq, 5)), axis=1).reshape(-1).astype(np.int64)
    ids, sc = ops.fullsort_topk(t(Q), None, t(V), K, t(hp), t(hi), mode="tc")
    o_ids, o_sc = ofs.full_sort_topk(Q, V, np.arange(nq), hp, hi, K)
    np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
    np.testing.assert_array_equal(sc.cpu().numpy(), o_sc)


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_fullsort_tc_variants_agree_with_fp32_kernel_at_size(tc_variant, variant):
    """A shape the oracle cannot finish (20k x 300k x 128, 30 history items per row): the tensor-core
    variants and the exact CUDA-core kernel agree bit for bit, and the certificate holds for every row."""
    from recbole_b200 import ops
    from recbole_b200._lib import lib
    tc_variant(variant)
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    nq, N, d, h = 20000, 300001, 128, 30
    Q = torch.randn(nq, d, device=dev, generator=gen) * 0.1
    V = torch.randn(N, d, device=dev, generator=gen) * 0.1
    hp = torch.arange(0, h * nq + 1, h, device=dev, dtype=torch.int64)
    hi = torch.sort(torch.randint(1, N, (nq, h), device=dev, generator=gen), dim=1).values.reshape(-1).contiguous()
    ids_t, sc_t = ops.fullsort_topk(Q, None, V, 10, hp, hi, mode="tc")
    assert lib.rb2_fullsort_tc_last_fallback_rows() == 0
    ids_f, sc_f = ops.fullsort_topk(Q, None, V, 10, hp, hi, mode="fp32")
    assert torch.equal(ids_t, ids_f) and torch.equal(sc_t, sc_f)
    # ... and a random subset of the rows against the ORACLE (all 300k items, the rows' own history lists)
    rows = np.sort(np.random.default_rng(variant).choice(nq, 48, replace=False))
    hi_c = hi.cpu().numpy().reshape(nq, h)[rows]
    o_ids, o_sc = ofs.full_sort_topk(Q.cpu().numpy()[rows], V.cpu().numpy(), np.arange(len(rows)),
                                     np.arange(0, h * len(rows) + 1, h, dtype=np.int64), hi_c.reshape(-1), 10)
    np.testing.assert_array_equal(ids_t.cpu().numpy()[rows], o_ids)
    np.testing.assert_array_equal(sc_t.cpu().numpy()[rows], o_sc)


def test_fullsort_tc_vs_oracle_on_row_subset_at_cfg3_item_count():
    """BASELINE config 3's item side (2,000,001 x 128) against 16 384 query rows with 40 history items each: the
    tensor-core scorer's ids and canonical scores for 32 random rows equal the oracle's bit for bit (the oracle scores
    those rows against all 2 M items on the CPU)."""
    from recbole_b200 import ops
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev)
    gen.manual_seed(11)
    nq, N, d, h = 16384, 2_000_001, 128, 40
    Q = torch.randn(nq, d, device=dev, generator=gen) * 0.1
    V = torch.randn(N, d, device=dev, generator=gen) * 0.1
    hp = torch.arange(0, h * nq + 1, h, device=dev, dtype=torch.int64)
    hi = torch.sort(torch.randint(1, N, (nq, h), device=dev, generator=gen), dim=1).values.reshape(-1).contiguous()
    st = ops.ScorerState()
    ids_t, sc_t = ops.fullsort_topk(Q, None, V, 10, hp, hi, mode="tc", state=st)
    assert st.last_fallback_rows == 0
    rows = np.sort(np.random.default_rng(3).choice(nq, 32, replace=False))
    hi_c = hi.cpu().numpy().reshape(nq, h)[rows]
    o_ids, o_sc = ofs.full_sort_topk(Q.cpu().numpy()[rows], V.cpu().numpy(), np.arange(len(rows)),
                                     np.arange(0, h * len(rows) + 1, h, dtype=np.int64), hi_c.reshape(-1), 10)
    np.testing.assert_array_equal(ids_t.cpu().numpy()[rows], o_ids)
    np.testing.assert_array_equal(sc_t.cpu().numpy()[rows], o_sc)


def test_fullsort_tc_cascade_and_adaptive_first_pass(tc_variant):
    """One item with a norm 300x the others inflates max ||v|| and with it every error bound: the
    FP16-accumulator certificate fails for (nearly) all rows, they are re-scored by the fp32-accumulator pass
    (and what still fails by the exact kernel), and from the second call on the scorer starts with fp32
    accumulators.  Exact every time."""
    from recbole_b200 import ops
    from recbole_b200._lib import lib
    from gpu_util import t
    tc_variant(3)
    nq, N, d, K = 2048, 3000, 128, 10
    Q, V, hp, hi = _tc_case(d, nq, N, K, seed=3)
    V[7] *= 300.0
    o_ids, o_sc = ofs.full_sort_topk(Q, V, np.arange(nq), hp, hi, K)
    pass2 = []
    for _ in range(3):
        ids, sc = ops.fullsort_topk(t(Q), None, t(V), K, t(hp), t(hi), mode="tc")
        pass2.append(int(lib.rb2_fullsort_tc_last_pass2_rows()))
        np.testing.assert_array_equal(ids.cpu().numpy(), o_ids)
        np.testing.assert_array_equal(sc.cpu().numpy(), o_sc)
    assert pass2[0] > nq // 4          # most rows needed the second pass ...
    assert pass2[-1] == 0              # ... so later calls start with fp32 accumulators
    # easy data brings the fast pass back (the failure estimate decays on the probing calls)
    Q2, V2, hp2, hi2 = _tc_case(d, nq, N, K, seed=4)
    o2, _ = ofs.full_sort_topk(Q2, V2, np.arange(nq), hp2, hi2, K)
    for _ in range(40):
        ids, _ = ops.fullsort_topk(t(Q2), None, t(V2), K, t(hp2), t(hi2), mode="tc")
    np.testing.assert_array_equal(ids.cpu().numpy(), o2)


@pytest.mark.parametrize("nq", [4097, 40000])
def test_tc_every_certificate_fails_fallback_fits_its_workspace(nq):
    """All item rows equal: every score ties, every tensor-core certificate fails and ALL rows are re-scored by the exact
    kernel with the workspace of the full call (fewer rows want more item splits than that buffer was sized for).  The
    result is the oracle's: the k lowest ids."""
    from recbole_b200 import ops
    from gpu_util import t
    rng = np.random.default_rng(nq)
    d, N, k = 64, 3001, 10
    U = rng.standard_normal((nq, d)).astype(np.float32)
    V = np.tile(rng.standard_normal((1, d)).astype(np.float32), (N, 1))
    st = ops.ScorerState()
    ids, sc = ops.fullsort_topk(t(U), None, t(V), k, mode="tc", state=st)
    assert st.last_fallback_rows + st.last_pass2_rows > 0
    want = np.tile(np.arange(1, k + 1, dtype=np.int64), (nq, 1))          # id 0 is [PAD]
    np.testing.assert_array_equal(ids.cpu().numpy(), want)
