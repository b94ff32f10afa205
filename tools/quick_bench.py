"""Quick device timings of the hot-path kernels (development aid; bench.py is the contract)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recbole_b200 import ops


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    for (n_users, n_items, dim, B) in [(138494, 26745, 64, 1 << 20), (138494, 26745, 64, 2048),
                                       (10_000_001, 2_000_001, 128, 1 << 20), (10_000_001, 2_000_001, 128, 1 << 22)]:
        U = torch.randn(n_users, dim, device=dev, generator=gen) * 0.05
        V = torch.randn(n_items, dim, device=dev, generator=gen) * 0.05
        st = dict(mU=torch.zeros_like(U), vU=torch.zeros_like(U), mV=torch.zeros_like(V), vV=torch.zeros_like(V))
        u = torch.randint(1, n_users, (B,), device=dev, generator=gen)
        p = torch.randint(1, n_items, (B,), device=dev, generator=gen)
        n = torch.randint(1, n_items, (B,), device=dev, generator=gen)
        loss = torch.zeros(1, device=dev)
        ws = ops.bpr_workspace(B, dim, dev)
        for kind in ("adam", "sgd"):
            opt = ops.Optim(kind, lr=1e-3)
            ms = timeit(lambda: ops.bpr_train_step(U, V, st, u, p, n, opt, loss, None, ws))
            byt = B * ((72 if kind == "adam" else 24) * dim + 24)
            print("train %s users=%d items=%d d=%d B=%d: %.3f ms  %.1f Msamples/s  alg %.0f GB/s" % (
                kind, n_users, n_items, dim, B, ms, B / ms / 1e3, byt / ms / 1e6), flush=True)
        del st, ws
        if n_users < 1_000_000:
            users = torch.arange(1, n_users, device=dev)
            nq = users.numel()
            hp = torch.arange(0, 20 * nq + 1, 20, device=dev, dtype=torch.int64)
            hi = torch.sort(torch.randint(1, n_items, (nq, 20), device=dev, generator=gen), dim=1).values.reshape(-1)
            ms = timeit(lambda: ops.fullsort_topk(U, users, V, 10, hp, hi), warm=1, it=3)
            print("fullsort fp32 users=%d items=%d d=%d: %.2f ms  %.2f Musers/s  %.1f TFLOP/s" % (
                nq, n_items, dim, ms, nq / ms / 1e3, 2.0 * nq * n_items * dim / ms / 1e9), flush=True)
        else:
            users = torch.randint(1, n_users, (65536,), device=dev, generator=gen)
            ms = timeit(lambda: ops.fullsort_topk(U, users, V, 10), warm=1, it=2)
            print("fullsort fp32 users=%d items=%d d=%d: %.2f ms  %.3f Musers/s  %.1f TFLOP/s" % (
                65536, n_items, dim, ms, 65536 / ms / 1e3, 2.0 * 65536 * n_items * dim / ms / 1e9), flush=True)
        del U, V
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
