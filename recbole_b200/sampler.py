"""Device-resident negative sampler; mirrors ``recbole.sampler.Sampler`` (sampler.py:157-265).

``sample_by_user_ids(user_ids, num)`` keeps the reference's contract: the result has
``num * len(user_ids)`` ids, slot ``k*len + i`` belongs to user ``i``, no id is one of the user's
used ids of the phase, id 0 is never drawn.

mode='ref'  : the reference's own random stream.  ``random_list`` is the once-shuffled candidate
              list (np.random.shuffle of arange(1, n_items), sampler.py:54-57,199) and
              ``random_pr`` the moving pointer; given the same two, the output is bit-identical
              to the reference's and ``random_pr`` ends where the reference's ends (mod len).
mode='hash' : counter-based stream keyed by (seed, step, slot, attempt): one launch, no host state.
"""
import numpy as np
import torch

from . import ops


class DeviceSampler:
    def __init__(self, n_items, used_indptr, used_indices, mode="hash", seed=2020, random_list=None,
                 random_pr=0):
        self.n_items = int(n_items)
        self.used_indptr, self.used_indices = used_indptr, used_indices
        self.device = used_indptr.device
        self.mode = mode
        self.seed = int(seed)
        self.step = 0
        self.random_pr = int(random_pr)
        if mode == "ref":
            if random_list is None:  # what AbstractSampler.set_distribution does (sampler.py:54-57)
                random_list = np.arange(1, self.n_items)
                np.random.shuffle(random_list)
            self.random_list = torch.as_tensor(np.asarray(random_list), dtype=torch.int64, device=self.device)
        elif mode != "hash":
            raise ValueError("mode must be 'ref' or 'hash'")

    @classmethod
    def from_reference_sampler(cls, sampler, device, mode="ref"):
        """Build from an unmodified reference Sampler (after set_phase): same used ids, same
        random_list, same pointer."""
        indptr, indices = cls._used_csr(sampler, device)
        return cls(sampler.n_items, indptr, indices, mode=mode, random_list=np.asarray(sampler.random_list),
                   random_pr=sampler.random_pr)

    @staticmethod
    def _used_csr(sampler, device):
        from .data import build_csr
        if hasattr(sampler, "datasets") and getattr(sampler, "phase", None) in getattr(sampler, "phases", []):
            # used ids of the phase = the interactions of the phases up to it (sampler.py:206-218): id pairs straight
            # from the datasets the sampler keeps, no per-user sets
            upto = sampler.phases.index(sampler.phase)
            rows = torch.cat([d.inter_feat[sampler.uid_field] for d in sampler.datasets[:upto + 1]])
            cols = torch.cat([d.inter_feat[sampler.iid_field] for d in sampler.datasets[:upto + 1]])
        else:
            rows, cols = [], []
            for u, s in enumerate(sampler.used_ids):
                if len(s):
                    rows.append(np.full(len(s), u, dtype=np.int64))
                    cols.append(np.fromiter(s, dtype=np.int64, count=len(s)))
            rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
            cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
        return build_csr(sampler.n_users, rows, cols, sampler.n_items, device)

    def sample_by_user_ids(self, user_ids, num):
        user_ids = torch.as_tensor(user_ids, dtype=torch.int64, device=self.device).contiguous()
        if self.mode == "ref":
            out, self.random_pr = ops.neg_sample_ref(user_ids, int(num), self.random_list, self.random_pr,
                                                     self.used_indptr, self.used_indices)
            return out
        self.step += 1
        return ops.neg_sample_hash(user_ids, int(num), self.n_items, self.used_indptr, self.used_indices, self.seed,
                                   self.step)
