"""Development aid: run the tensor-core scorer on one case and print diagnostics vs the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recbole_b200 import ops
from recbole_b200._lib import lib
from oracle import fullsort as ofs

def case(dim, nq, N, K, hist=4, seed=0):
    rng = np.random.default_rng(seed)
    Q = rng.standard_normal((nq, dim)).astype(np.float32)
    V = rng.standard_normal((N, dim)).astype(np.float32)
    hp = np.arange(nq + 1, dtype=np.int64) * hist
    hi = np.sort(rng.integers(1, N, (nq, hist)), axis=1).reshape(-1).astype(np.int64)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    ids, sc = ops.fullsort_topk(t(Q), None, t(V), K, t(hp), t(hi), mode="tc")
    torch.cuda.synchronize()
    fb = lib.rb2_fullsort_tc_last_fallback_rows()
    o_ids, o_sc = ofs.full_sort_topk(Q, V, np.arange(nq), hp, hi, K)
    ids = ids.cpu().numpy(); sc = sc.cpu().numpy()
    bad = (ids != o_ids).any(axis=1)
    print("dim=%d nq=%d N=%d K=%d: fallback rows %d, mismatching rows %d, score mismatches %d" % (
        dim, nq, N, K, fb, bad.sum(), (sc != o_sc).sum()), flush=True)
    if bad.any():
        r = np.argmax(bad)
        print(" row", r, "\n  got ", ids[r], sc[r], "\n  want", o_ids[r], o_sc[r])

if __name__ == "__main__":
    case(64, 128, 256, 10)
    case(64, 300, 5000, 10)
    case(128, 300, 5000, 10)
    case(128, 1000, 70001, 10)
    case(64, 4096, 200000, 10, hist=0)
    lib.rb2_fullsort_tc_set_kprime(16)
    case(128, 1000, 70001, 10)
    case(64, 300, 5000, 10, hist=50)
    lib.rb2_fullsort_tc_set_kprime(0)
