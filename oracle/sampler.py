"""Negative sampling restated in numpy (TEST INFRASTRUCTURE).

(1) ``RefSampler``: the reference's algorithm, bit for bit -- a once-shuffled candidate
    list walked by a moving pointer, re-drawing while the value is in the user's used set:
      ``AbstractSampler.random_num``         recbole/sampler/sampler.py:82-101
      ``AbstractSampler.sample_by_key_ids``  sampler.py:103-154 (both branches compute the same thing)
      ``Sampler.get_random_list``            sampler.py:191-204 (uniform: arange(1, n_items))
    Given the same ``random_list`` / ``random_pr`` it returns exactly what the reference returns.

(2) ``hash_sample``: the counter-based variant the CUDA path offers for device-resident
    training (no host permutation): candidate(slot, attempt) from a splitmix64 hash, rejected
    against the same used-set CSR.  Same contract (slot k*B+i belongs to user i, never a used
    item, never the pad id 0), different random stream; defined here so the kernel can be
    checked bit for bit.
"""
import numpy as np

U64 = np.uint64


class RefSampler:
    def __init__(self, random_list, used_indptr, used_indices, random_pr=0):
        self.random_list = np.asarray(random_list, dtype=np.int64)
        self.random_list_length = len(self.random_list)
        self.random_pr = int(random_pr)
        self.used_indptr = np.asarray(used_indptr, dtype=np.int64)
        self.used_indices = np.asarray(used_indices, dtype=np.int64)

    def random_num(self, num):  # sampler.py:82-101
        out = []
        self.random_pr %= self.random_list_length
        while True:
            if self.random_pr + num <= self.random_list_length:
                out.append(self.random_list[self.random_pr:self.random_pr + num])
                self.random_pr += num
                break
            out.append(self.random_list[self.random_pr:])
            num -= self.random_list_length - self.random_pr
            self.random_pr = 0
        return np.concatenate(out)

    def _is_used(self, keys, values):
        res = np.zeros(len(keys), dtype=bool)
        for j, (k, v) in enumerate(zip(keys, values)):
            row = self.used_indices[self.used_indptr[k]:self.used_indptr[k + 1]]
            i = np.searchsorted(row, v)
            res[j] = i < len(row) and row[i] == v
        return res

    def sample_by_key_ids(self, key_ids, num):  # sampler.py:103-154
        key_ids = np.tile(np.asarray(key_ids, dtype=np.int64), num)
        total = len(key_ids)
        value_ids = np.zeros(total, dtype=np.int64)
        check = np.arange(total)
        while len(check) > 0:
            value_ids[check] = self.random_num(len(check))
            check = check[self._is_used(key_ids[check], value_ids[check])]
        return value_ids


# ---- counter-based variant -------------------------------------------------------------------

_C1 = U64(0x9E3779B97F4A7C15)
_C2 = U64(0xBF58476D1CE4E5B9)
_C3 = U64(0x94D049BB133111EB)
MAX_ATTEMPTS = 64


def _mix(seed, step, slot, attempt):
    with np.errstate(over="ignore"):
        x = U64(seed) + _C1 * (slot.astype(U64) + U64(1))
        x = x ^ ((attempt.astype(U64) + U64(1)) * _C2)
        x = x + U64(step) * _C3
        x = (x ^ (x >> U64(30))) * _C2
        x = (x ^ (x >> U64(27))) * _C3
        x = x ^ (x >> U64(31))
    return x


def _mulhi(r, n):
    """floor(r * n / 2**64) for uint64 r and n < 2**32."""
    n = U64(n)
    hi, lo = r >> U64(32), r & U64(0xFFFFFFFF)
    return (hi * n + ((lo * n) >> U64(32))) >> U64(32)


def hash_candidates(seed, step, slot, attempt, n_items):
    """uniform id in 1..n_items-1 for (slot, attempt)."""
    return (U64(1) + _mulhi(_mix(seed, step, slot, attempt), n_items - 1)).astype(np.int64)


def hash_sample(user_ids, num, n_items, used_indptr, used_indices, seed, step):
    """int64[num*B]; slot k*B+i belongs to user i (sampler.py:112-115 layout).

    Attempt a = 0, 1, ... until the candidate is not in the user's used set.  After
    MAX_ATTEMPTS rejected draws the slot takes the first unused id found scanning upwards
    (cyclically over 1..n_items-1) from its last candidate -- deterministic, and only reachable
    for users who have used almost every item.
    """
    user_ids = np.asarray(user_ids, dtype=np.int64)
    B = len(user_ids)
    keys = np.tile(user_ids, num)
    slot = np.arange(B * num, dtype=np.int64)
    out = np.zeros(B * num, dtype=np.int64)
    pending = np.arange(B * num)

    def used(k, v):
        row = used_indices[used_indptr[k]:used_indptr[k + 1]]
        i = np.searchsorted(row, v)
        return i < len(row) and row[i] == v

    for a in range(MAX_ATTEMPTS):
        if len(pending) == 0:
            break
        cand = hash_candidates(seed, step, slot[pending], np.full(len(pending), a, dtype=np.int64), n_items)
        out[pending] = cand
        rej = np.array([used(k, v) for k, v in zip(keys[pending], cand)], dtype=bool)
        pending = pending[rej]
    for j in pending:
        v = int(out[j])
        for _ in range(n_items - 1):
            v = v + 1 if v + 1 < n_items else 1
            if not used(keys[j], v):
                break
        out[j] = v
    return out
