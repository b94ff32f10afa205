"""Debug: adam_lazy vs the oracle's dense Adam with small batches (gaps of several steps), single-GPU step and the
peer-memory step with a world of one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import bpr as obpr
from recbole_b200 import ops
from recbole_b200.dist import ShardedBPR

dev = torch.device("cuda:0")


class OneRank:
    rank, world, staged = 0, 1, False
    def barrier(self): pass
    def all_gather_object(self, o): return [o]
    def all_gather_equal(self, t): return t
    def all_reduce_sum(self, t): return t


for d in (64, 128):
    for exchange in ("local",):
        rng = np.random.default_rng(1)
        n_users, n_items, B, steps = 400, 301, 150, 5
        U0 = (rng.standard_normal((n_users, d)) * 0.3).astype(np.float32)
        V0 = (rng.standard_normal((n_items, d)) * 0.3).astype(np.float32)
        m = ShardedBPR(n_users, n_items, d, OneRank(), dev, U_full=U0, V_full=V0, exchange=exchange)
        m.build_optimizer("adam_lazy", lr=2e-3)
        st = obpr.new_state(U0, V0)
        for s in range(steps):
            u, p, n = rng.integers(1, n_users, B), rng.integers(1, n_items, B), rng.integers(1, n_items, B)
            t = lambda a: torch.from_numpy(a).to(dev)
            lo = m.train_step(t(u), t(p), t(n), global_batch=B).item()
            ro = obpr.bpr_train_step(st, u, p, n, s + 1, optimizer="adam", lr=2e-3, dense=True)
            m.flush()
            torch.cuda.synchronize()
            eu = np.abs(m.U[:n_users].cpu().numpy() - st["U"]).max() / np.abs(st["U"]).max()
            ev = np.abs(m.V[:n_items].cpu().numpy() - st["V"]).max() / np.abs(st["V"]).max()
            print(d, exchange, "step", s + 1, "loss err %.2e" % (abs(lo - ro) / abs(ro)), "U %.2e V %.2e" % (eu, ev))
            for k in ("mU", "vU", "mV", "vV"):
                a, b = m.state[k][:st[k].shape[0]].cpu().numpy(), st[k]
                dd = np.abs(a - b)
                i = np.unravel_index(dd.argmax(), dd.shape)
                print("    %s err %.2e (rel to max)  worst row %d col %d got %.6e want %.6e touched_now %s" % (
                    k, dd.max() / np.abs(b).max(), i[0], i[1], a[i], b[i], (i[0] in (u if k[1] == "U" else np.concatenate([p, n])))))
            dd = np.abs(m.U[:n_users].cpu().numpy() - st["U"]); i = np.unravel_index(dd.argmax(), dd.shape)
            print("    U worst row %d col %d diff %.3e touched_now %s  count_in_batch %d" % (i[0], i[1], dd[i], i[0] in u, (u == i[0]).sum()))
        m.check_flags()
