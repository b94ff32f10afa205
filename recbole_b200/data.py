"""Device-resident index structures for the hot path (host logic; torch ops are plumbing).

``EvalIndex`` holds what ``GeneralFullDataLoader`` precomputes per user with Python loops
(recbole/data/dataloader/general_dataloader.py:294-328) -- evaluated users, their history
(= used ids of the phase minus its positives, sampler.py:206-227) and their positives -- as two
sorted CSRs in HBM, built once.
"""
import numpy as np
import torch


def _bits(n):
    b = 1
    while b < 32 and (1 << b) < n:
        b += 1
    return b


def _sorted_pairs(rows, cols, n_rows, n_cols):
    """(row, col) pairs in (row, col) order.  On the GPU the two sorts are the library's own bounded-key sort
    (rb2_sort_positions: least-significant key first, stable), the gathers are plumbing; on the CPU (host-side tests)
    a torch sort of the combined key."""
    if rows.is_cuda and rows.numel() < (1 << 31) and n_rows < (1 << 31) and n_cols < (1 << 31):
        from . import ops
        _, o1 = ops.sort_positions(cols.to(torch.int32).contiguous(), _bits(int(n_cols)))
        o1 = o1.to(torch.int64)
        _, o2 = ops.sort_positions(rows[o1].to(torch.int32).contiguous(), _bits(int(n_rows)))
        order = o1[o2.to(torch.int64)]
        return rows[order], cols[order]
    key = torch.sort(rows * int(n_cols) + cols).values
    r = torch.div(key, int(n_cols), rounding_mode="floor")
    return r, key - r * int(n_cols)


def build_csr(n_rows, rows, cols, n_cols, device):
    """Sorted, de-duplicated CSR of (row, col) pairs: (indptr int64[n_rows+1], indices int64)."""
    rows = torch.as_tensor(rows, dtype=torch.int64, device=device)
    cols = torch.as_tensor(cols, dtype=torch.int64, device=device)
    if rows.numel() == 0:
        return torch.zeros(n_rows + 1, dtype=torch.int64, device=device), torch.zeros(0, dtype=torch.int64,
                                                                                      device=device)
    r, c = _sorted_pairs(rows, cols, n_rows, n_cols)
    keep = torch.ones(r.numel(), dtype=torch.bool, device=device)
    keep[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
    r, c = r[keep], c[keep]
    counts = torch.bincount(r, minlength=n_rows)
    indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(counts, 0)
    return indptr, c.contiguous()


def csr_difference(a, b, n_cols):
    """Rows of CSR a minus rows of CSR b (same row count)."""
    device = a[0].device
    n_rows = a[0].numel() - 1
    ra = torch.repeat_interleave(torch.arange(n_rows, device=device), a[0][1:] - a[0][:-1])
    rb = torch.repeat_interleave(torch.arange(n_rows, device=device), b[0][1:] - b[0][:-1])
    ka, kb = ra * int(n_cols) + a[1], rb * int(n_cols) + b[1]
    keep = ~torch.isin(ka, kb)
    return build_csr(n_rows, ra[keep], a[1][keep], n_cols, device)


class EvalIndex:
    def __init__(self, uid_list, hist, pos, n_items):
        self.uid_list = uid_list        # int64[U] evaluated users (ascending), device
        self.hist_indptr, self.hist_indices = hist
        self.pos_indptr, self.pos_indices = pos
        self.n_items = int(n_items)

    @property
    def n_eval_users(self):
        return int(self.uid_list.numel())

    def pos_len(self):
        return self.pos_indptr[1:] - self.pos_indptr[:-1]

    @classmethod
    def from_phase_pairs(cls, n_users, n_items, phase_pairs, phase, device):
        """phase_pairs: [(users, items)] for train, valid, test; evaluate phase `phase`.
        used = union of phases 0..phase (sampler.py:213-218); positives = this phase's pairs;
        history = used - positives (general_dataloader.py:319-321)."""
        pu = torch.as_tensor(phase_pairs[phase][0], dtype=torch.int64, device=device)
        pi = torch.as_tensor(phase_pairs[phase][1], dtype=torch.int64, device=device)
        uid_list = torch.unique(pu)
        remap = torch.full((n_users,), -1, dtype=torch.int64, device=device)
        remap[uid_list] = torch.arange(uid_list.numel(), device=device)
        pos = build_csr(uid_list.numel(), remap[pu], pi, n_items, device)
        hu = torch.cat([torch.as_tensor(phase_pairs[p][0], dtype=torch.int64, device=device)
                        for p in range(phase + 1)])
        hi = torch.cat([torch.as_tensor(phase_pairs[p][1], dtype=torch.int64, device=device)
                        for p in range(phase + 1)])
        keep = remap[hu] >= 0
        used = build_csr(uid_list.numel(), remap[hu[keep]], hi[keep], n_items, device)
        hist = csr_difference(used, pos, n_items)
        return cls(uid_list, hist, pos, n_items)

    @classmethod
    def from_reference_dataloader(cls, eval_data, device):
        """From an unmodified reference GeneralFullDataLoader, WITHOUT its per-user Python structures: the loader's
        sampler keeps the datasets of all phases (sampler.py:171-184); used ids of the loader's phase = interactions of
        the phases up to it (sampler.py:206-218), positives = the loader's own dataset, history = used - positives
        (general_dataloader.py:319-321) -- three id-pair lists, sorted and de-duplicated on the device.  (The loops the
        reference runs to build uid2history_item / uid2swap_idx are what SURVEY.md 8f-2 replaces.)"""
        ds, smp = eval_data.dataset, eval_data.sampler
        if not (hasattr(smp, "datasets") and hasattr(smp, "phases") and getattr(smp, "phase", None) in smp.phases):
            return cls._from_reference_dataloader_loops(eval_data, device)
        upto = smp.phases.index(smp.phase)
        pairs = [(d.inter_feat[d.uid_field], d.inter_feat[d.iid_field]) for d in smp.datasets[:upto + 1]]
        pairs.append((ds.inter_feat[ds.uid_field], ds.inter_feat[ds.iid_field]))
        return cls.from_phase_pairs(ds.user_num, ds.item_num, pairs, len(pairs) - 1, device)

    @classmethod
    def _from_reference_dataloader_loops(cls, eval_data, device):
        """The same index from the loader's public per-user arrays uid_list / uid2history_item
        (general_dataloader.py:294-313); kept as the cross-check of from_reference_dataloader."""
        ds = eval_data.dataset
        n_items = ds.item_num
        uid_list = torch.as_tensor(np.asarray(eval_data.uid_list), dtype=torch.int64, device=device)
        n_users = ds.user_num
        remap = torch.full((n_users,), -1, dtype=torch.int64, device=device)
        remap[uid_list] = torch.arange(uid_list.numel(), device=device)
        pu = ds.inter_feat[ds.uid_field].to(device)
        pi = ds.inter_feat[ds.iid_field].to(device)
        pos = build_csr(uid_list.numel(), remap[pu], pi, n_items, device)
        rows, cols = [], []
        for r, u in enumerate(uid_list.tolist()):
            h = eval_data.uid2history_item[u]
            if h is not None and len(h):
                rows.append(torch.full((len(h),), r, dtype=torch.int64))
                cols.append(torch.as_tensor(h, dtype=torch.int64))
        if rows:
            hist = build_csr(uid_list.numel(), torch.cat(rows), torch.cat(cols), n_items, device)
        else:
            hist = build_csr(uid_list.numel(), [], [], n_items, device)
        return cls(uid_list, hist, pos, n_items)


# ---- synthetic workloads of BASELINE.json -------------------------------------------

def synth_interactions(n_users, n_items, n_inter, seed, device, zipf_alpha=1.0):
    """(user, item) pairs on the device: log-normal user activity, Zipf item popularity, no
    duplicate pairs.  ids start at 1 (0 = [PAD])."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    act = torch.exp(torch.randn(n_users - 1, generator=g, device=device) * 1.0)
    user = 1 + torch.multinomial(act / act.sum(), n_inter, replacement=True, generator=g) \
        if n_users - 1 <= (1 << 24) else 1 + torch.randint(0, n_users - 1, (n_inter,), generator=g, device=device)
    # Zipf via inverse CDF on ranks: rank ~ exp(U * log(N)) gives p(rank) ~ 1/rank for alpha = 1
    u = torch.rand(n_inter, generator=g, device=device, dtype=torch.float64)
    if abs(zipf_alpha - 1.0) < 1e-9:
        rank = torch.exp(u * np.log(n_items - 1))
    else:
        a = 1.0 - zipf_alpha
        rank = ((u * ((n_items - 1) ** a - 1.0)) + 1.0) ** (1.0 / a)
    item = torch.clamp(rank.to(torch.int64), 1, n_items - 1)
    perm = torch.randperm(n_items - 1, generator=g, device=device) + 1  # popularity not tied to id order
    item = perm[item - 1]
    key = torch.unique(user * n_items + item)
    key = key[torch.randperm(key.numel(), generator=g, device=device)]
    user = torch.div(key, n_items, rounding_mode="floor")
    return user.contiguous(), (key - user * n_items).contiguous()


def split_by_ratio(user, item, ratios=(0.8, 0.1, 0.1), seed=0):
    """Random 0.8/0.1/0.1 split of the pairs (reference: RO_RS, dataset.py:1281-1315 splits per
    user; a global random split has the same expected shape and is enough for synthetic data)."""
    n = user.numel()
    g = torch.Generator(device=user.device)
    g.manual_seed(int(seed))
    r = torch.rand(n, generator=g, device=user.device)
    a, b = ratios[0], ratios[0] + ratios[1]
    m0, m1, m2 = r < a, (r >= a) & (r < b), r >= b
    return [(user[m], item[m]) for m in (m0, m1, m2)]


class DeviceTrainLoader:
    """Device-resident replacement for the training side of ``GeneralNegSampleDataLoader``
    (recbole/data/dataloader/general_dataloader.py:212-241) + ``Dataset.shuffle``
    (recbole/data/interaction.py:272-276) + ``Interaction.to(device)`` (trainer.py:158-159): the
    interactions live in HBM, an epoch is a device-side permutation, negatives come from the sampler
    kernel, and every batch is already an on-device ``Interaction`` with the reference's field names
    (pair-wise format: user, item, neg_item).  SURVEY.md 8f rank 1."""

    def __init__(self, user, item, sampler, batch_size, uid_field="user_id", iid_field="item_id", neg_prefix="neg_",
                 shuffle=True, seed=2020):
        from .interaction import Interaction
        self._Interaction = Interaction
        self.user = torch.as_tensor(user, dtype=torch.int64, device=sampler.device).contiguous()
        self.item = torch.as_tensor(item, dtype=torch.int64, device=sampler.device).contiguous()
        self.sampler, self.batch_size, self.shuffle = sampler, int(batch_size), shuffle
        self.uid_field, self.iid_field, self.neg_field = uid_field, iid_field, neg_prefix + iid_field
        self.gen = torch.Generator(device=sampler.device)
        self.gen.manual_seed(int(seed))

    def __len__(self):
        return (self.user.numel() + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self.user.numel()
        perm = torch.randperm(n, device=self.user.device, generator=self.gen) if self.shuffle else None
        for lo in range(0, n, self.batch_size):
            idx = perm[lo:lo + self.batch_size] if perm is not None else slice(lo, lo + self.batch_size)
            u = self.user[idx].contiguous()
            yield self._Interaction({self.uid_field: u, self.iid_field: self.item[idx].contiguous(),
                                     self.neg_field: self.sampler.sample_by_user_ids(u, 1)})
