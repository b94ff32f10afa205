"""Synthetic workloads of BASELINE.json, generated on the CPU with numpy from fixed
seeds so that the CUDA arm, the CPU-baseline leg and `bench.py --impl reference` all consume
IDENTICAL inputs and identical pre-sampled negative ids.  Bench / test support, not product.
"""
import numpy as np

WORKLOADS = {
    # name: n_users, n_items (both incl. the [PAD] row 0), dim, interactions, default train batch
    "cfg2": dict(n_users=138_494, n_items=26_745, dim=64, n_inter=20_000_263, batch=1 << 20,
                 desc="BPR-MF synthetic ml-20m shape (138k users x 27k items, 20M inters, d=64)"),
    "cfg2-small": dict(n_users=20_000, n_items=5_000, dim=64, n_inter=1_000_000, batch=1 << 16,
                       desc="1/20 scale of cfg2 (CI / CPU smoke)"),
    "cfg3": dict(n_users=10_000_001, n_items=2_000_001, dim=128, n_inter=1_000_000_000, batch=1 << 20,
                 desc="BPR-MF synthetic 10M users x 2M items x d=128"),
}


def xavier_tables(n_users, n_items, dim, seed=2020):
    """xavier-normal fp32 tables, std = sqrt(2 / (rows + d))  (recbole/model/init.py:27)."""
    rng = np.random.default_rng(seed)
    U = (rng.standard_normal((n_users, dim), dtype=np.float32) * np.float32(np.sqrt(2.0 / (n_users + dim))))
    V = (rng.standard_normal((n_items, dim), dtype=np.float32) * np.float32(np.sqrt(2.0 / (n_items + dim))))
    return U, V


def _sorted_unique(keys):
    keys = np.sort(keys)
    keep = np.ones(len(keys), dtype=bool)
    keep[1:] = keys[1:] != keys[:-1]
    return keys[keep]


def interactions(n_users, n_items, n_inter, seed=2020):
    """Unique (user, item) pairs: log-normal user activity, Zipf(1) item popularity (popularity
    rank decoupled from the id by a permutation).  Returns int64 arrays in random order."""
    rng = np.random.default_rng(seed)
    act = np.exp(rng.standard_normal(n_users - 1))
    prob = act / act.sum()
    perm = rng.permutation(n_items - 1) + 1
    keys = np.zeros(0, dtype=np.int64)
    while len(keys) < n_inter:  # duplicates of hot (user, item) pairs are dropped; top up until exact
        n = int((n_inter - len(keys)) * 1.5) + 1024
        counts = np.floor(prob * n).astype(np.int64) + (rng.random(n_users - 1) < (prob * n) % 1.0)
        user = np.repeat(np.arange(1, n_users, dtype=np.int64), counts)
        rank = np.exp(rng.random(len(user), dtype=np.float32) * np.float32(np.log(n_items - 1))).astype(np.int64)
        item = perm[rank.clip(1, n_items - 1) - 1]
        keys = _sorted_unique(np.concatenate([keys, user * n_items + item]))
    keys = keys[rng.permutation(len(keys))[:n_inter]]
    return keys // n_items, keys % n_items


def split(user, item, seed=2020, ratios=(0.8, 0.1, 0.1)):
    rng = np.random.default_rng(seed + 1)
    r = rng.random(len(user))
    a, b = ratios[0], ratios[0] + ratios[1]
    masks = (r < a, (r >= a) & (r < b), r >= b)
    return [(user[m], item[m]) for m in masks]


def sample_negatives(users, n_items, used_keys_sorted, seed=2021):
    """One uniform negative per entry of `users`, never one of the user's used (train) items and
    never the pad id (the contract of sampler.py:103-154).  Vectorised rejection."""
    rng = np.random.default_rng(seed)
    neg = rng.integers(1, n_items, len(users))
    pending = np.arange(len(users))
    while len(pending):
        k = users[pending].astype(np.int64) * n_items + neg[pending]
        pos = np.searchsorted(used_keys_sorted, k)
        pos[pos >= len(used_keys_sorted)] = len(used_keys_sorted) - 1
        bad = used_keys_sorted[pos] == k
        pending = pending[bad]
        neg[pending] = rng.integers(1, n_items, len(pending))
    return neg


def csr_from_pairs(n_rows, rows, cols, n_cols):
    key = _sorted_unique(rows.astype(np.int64) * n_cols + cols)
    r, c = key // n_cols, key % n_cols
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(np.bincount(r, minlength=n_rows))
    return indptr, c


class BprWorkload:
    """Everything both arms need for one named workload."""

    def __init__(self, name, batch=None, n_batches=8, eval_users=None, max_inter=None, seed=2020):
        w = WORKLOADS[name]
        self.name, self.desc = name, w["desc"]
        self.n_users, self.n_items, self.dim = w["n_users"], w["n_items"], w["dim"]
        self.batch = int(batch or w["batch"])
        n_inter = w["n_inter"] if max_inter is None else min(w["n_inter"], max_inter)
        self.n_inter = n_inter
        user, item = interactions(self.n_users, self.n_items, n_inter, seed)
        self.phases = split(user, item, seed)
        tu, ti = self.phases[0]
        self.train_keys = np.sort(tu * self.n_items + ti)
        need = self.batch * n_batches
        reps = (need + len(tu) - 1) // len(tu)
        idx = np.concatenate([np.random.default_rng(seed + 7 + r).permutation(len(tu)) for r in range(reps)])[:need]
        bu, bp = tu[idx], ti[idx]
        bn = sample_negatives(bu, self.n_items, self.train_keys, seed + 1)
        self.batches = [(bu[i * self.batch:(i + 1) * self.batch], bp[i * self.batch:(i + 1) * self.batch],
                         bn[i * self.batch:(i + 1) * self.batch]) for i in range(n_batches)]
        # evaluation of the test phase: history = train + valid, positives = test
        eu, ei = self.phases[2]
        uid = _sorted_unique(eu)
        if eval_users is not None and eval_users < len(uid):
            uid = np.sort(np.random.default_rng(seed + 3).choice(uid, eval_users, replace=False))
        remap = -np.ones(self.n_users, dtype=np.int64)
        remap[uid] = np.arange(len(uid))
        m = remap[eu] >= 0
        self.uid_list = uid
        self.pos = csr_from_pairs(len(uid), remap[eu[m]], ei[m], self.n_items)
        hu = np.concatenate([self.phases[0][0], self.phases[1][0]])
        hi = np.concatenate([self.phases[0][1], self.phases[1][1]])
        m = remap[hu] >= 0
        self.hist = csr_from_pairs(len(uid), remap[hu[m]], hi[m], self.n_items)
        self.U0, self.V0 = xavier_tables(self.n_users, self.n_items, self.dim, seed)

    def rank_batches(self, rank, world, n_batches, seed=2020):
        """Per-rank batches for the user-partitioned multi-GPU run: `batch` triples per rank and step,
        drawn from the interactions of the users this rank owns (weak scaling)."""
        lo, hi = (rank * self.n_users) // world, ((rank + 1) * self.n_users) // world
        tu, ti = self.phases[0]
        m = (tu >= lo) & (tu < hi)
        tu, ti = tu[m], ti[m]
        need = self.batch * n_batches
        reps = (need + len(tu) - 1) // len(tu)
        idx = np.concatenate([np.random.default_rng(seed + 100 + 17 * rank + r).permutation(len(tu))
                              for r in range(reps)])[:need]
        bu, bp = tu[idx], ti[idx]
        bn = sample_negatives(bu, self.n_items, self.train_keys, seed + 31 + rank)
        B = self.batch
        return [(bu[i * B:(i + 1) * B], bp[i * B:(i + 1) * B], bn[i * B:(i + 1) * B]) for i in range(n_batches)]

    def describe(self):
        return dict(workload=self.name, desc=self.desc, n_users=self.n_users, n_items=self.n_items, dim=self.dim,
                    interactions=self.n_inter, train_batch=self.batch, eval_users=int(len(self.uid_list)), topk=10)


class Cfg3Device:
    """BASELINE config 3 shape (10M users x 2M items x d=128), generated ON THE DEVICE for the users
    one rank owns -- never through pandas / numpy.  Per user a sorted list of
    `per_user` Zipf(1) train items (the CSR the sampler rejects against and the evaluation history),
    batches = (user, one of its train items, a negative from rb2_neg_sample_hash), test positives =
    `n_test` further items per user."""

    def __init__(self, rank, world, device, n_users=10_000_001, n_items=2_000_001, dim=128, per_user=64, n_test=8,
                 batch=1 << 20, n_batches=4, seed=2020, scale=1.0):
        import torch
        from recbole_b200 import ops
        self.name = "cfg3"
        self.desc = "BPR-MF synthetic 10M users x 2M items x d=128 (device-generated, %d train items/user)" % per_user
        self.n_users, self.n_items, self.dim, self.batch = int(n_users * scale) | 1, int(n_items * scale) | 1, dim, batch
        n_users, n_items = self.n_users, self.n_items
        blk = (n_users + world - 1) // world
        self.u_lo, self.u_hi = min(rank * blk, n_users), min((rank + 1) * blk, n_users)
        if rank == 0:
            self.u_lo_eval = 1          # user 0 is [PAD]
        nu = self.u_hi - self.u_lo
        g = torch.Generator(device=device)
        g.manual_seed(seed + 1000 * rank)

        def zipf_items(shape):
            u = torch.rand(shape, device=device, generator=g)
            rank_ = torch.exp(u * float(np.log(n_items - 1))).long().clamp_(1, n_items - 1)
            return (rank_ * 2654435761 % (n_items - 1)) + 1       # decouple popularity from id order

        # train lists: sorted, duplicates inside a row collapsed by pushing them to distinct ids is not
        # needed -- duplicates are harmless for a CSR used as a rejection set
        items = torch.sort(zipf_items((nu, per_user)), dim=1).values
        self.used_indptr = torch.arange(0, (nu + 1) * per_user, per_user, device=device, dtype=torch.int64)
        self.used_indices = items.reshape(-1).contiguous()
        self.per_user = per_user
        # batches
        self.batches = []
        for b in range(n_batches):
            ul = torch.randint(1 if rank == 0 else 0, nu, (batch,), device=device, generator=g)
            slot = torch.randint(0, per_user, (batch,), device=device, generator=g)
            pos = items[ul, slot].contiguous()
            neg = ops.neg_sample_hash(ul, 1, n_items, self.used_indptr, self.used_indices, seed, b + 1)
            self.batches.append(((ul + self.u_lo).contiguous(), pos, neg))
        # evaluation: own users, positives = n_test fresh items
        self.test_items = torch.sort(zipf_items((nu, n_test)), dim=1).values
        self.n_test = n_test

    def describe(self):
        return dict(workload="cfg3", desc=self.desc, n_users=self.n_users, n_items=self.n_items, dim=self.dim,
                    interactions=int((self.n_users - 1) * self.per_user), train_batch=self.batch, topk=10)


class Cfg3Host:
    """The cfg3 shape on the HOST for the reference's CPU arm (`cpu_baseline`, `--impl reference`): the same
    distributions as Cfg3Device (users uniform, positives Zipf(1) with the popularity order decoupled from the id
    by a multiplicative hash, uniform negatives), numpy with a fixed seed.  Evaluation data for a bounded sample of
    users: `per_user` history items and `n_test` positives each."""

    def __init__(self, n_users=10_000_001, n_items=2_000_001, dim=128, batch=1 << 20, n_batches=4, per_user=64,
                 n_test=8, eval_users=1024, seed=2020, scale=1.0):
        self.name = "cfg3"
        self.n_users, self.n_items, self.dim, self.batch = int(n_users * scale) | 1, int(n_items * scale) | 1, dim, batch
        rng = np.random.default_rng(seed)

        def zipf_items(shape):
            r = np.exp(rng.random(shape) * np.log(self.n_items - 1)).astype(np.int64).clip(1, self.n_items - 1)
            return (r * 2654435761 % (self.n_items - 1)) + 1

        self.batches = [(rng.integers(1, self.n_users, batch), zipf_items(batch), rng.integers(1, self.n_items, batch))
                        for _ in range(n_batches)]
        self.uid_list = np.sort(rng.choice(np.arange(1, self.n_users), eval_users, replace=False))
        hist = np.sort(zipf_items((eval_users, per_user)), axis=1)
        test = np.sort(zipf_items((eval_users, n_test)), axis=1)
        # CSR rows must hold unique ids (a set in the reference)
        self.hist = csr_from_pairs(eval_users, np.repeat(np.arange(eval_users), per_user), hist.reshape(-1), self.n_items)
        self.pos = csr_from_pairs(eval_users, np.repeat(np.arange(eval_users), n_test), test.reshape(-1), self.n_items)
        self.per_user, self.n_test = per_user, n_test

    def describe(self):
        return dict(workload="cfg3", desc="BPR-MF synthetic 10M users x 2M items x d=128 (device-generated, %d train "
                    "items/user)" % self.per_user, n_users=self.n_users, n_items=self.n_items, dim=self.dim,
                    interactions=int((self.n_users - 1) * self.per_user), train_batch=self.batch, topk=10)
