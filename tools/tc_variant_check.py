"""Development aid: exactness of a tensor-core scorer variant against the fp32 scorer on awkward inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recbole_b200 import ops
from recbole_b200._lib import lib


def main():
    variant = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    assert lib.rb2_fullsort_tc_set_variant(variant) == 0
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    cases = []
    for (nq, N, d, h, k) in [(1000, 5000, 64, 5, 10), (4097, 70001, 128, 20, 10), (300, 1000003, 64, 0, 10),
                            (20000, 300001, 128, 30, 5), (777, 513, 128, 3, 16), (5, 300, 64, 0, 10)]:
        for kind in ("gauss", "scaled", "trained", "ties", "tiny"):
            cases.append((nq, N, d, h, k, kind))
    bad = 0
    for nq, N, d, h, k, kind in cases:
        Q = torch.randn(nq, d, device=dev, generator=gen)
        V = torch.randn(N, d, device=dev, generator=gen)
        if kind == "gauss":
            Q *= 0.1; V *= 0.1
        elif kind == "scaled":      # wildly different row norms
            Q *= torch.exp(torch.randn(nq, 1, device=dev, generator=gen) * 3)
            V *= torch.exp(torch.randn(N, 1, device=dev, generator=gen) * 2) * 1e-3
        elif kind == "trained":     # low-rank structure + popularity direction
            B = torch.randn(8, d, device=dev, generator=gen)
            Q = torch.randn(nq, 8, device=dev, generator=gen) @ B + 0.05 * Q
            V = torch.randn(N, 8, device=dev, generator=gen) @ B + 0.05 * V
        elif kind == "ties":        # few distinct values: many exact ties
            Q = torch.randint(-2, 3, (nq, d), device=dev, generator=gen).float() * 0.25
            V = torch.randint(-2, 3, (N, d), device=dev, generator=gen).float() * 0.5
        elif kind == "tiny":
            Q *= 1e-20; V *= 1e-12
        hp = hi = None
        if h:
            hp = torch.arange(0, h * nq + 1, h, device=dev, dtype=torch.int64)
            hi = torch.sort(torch.randint(1, N, (nq, h), device=dev, generator=gen), dim=1).values.reshape(-1).contiguous()
        ids_t, sc_t = ops.fullsort_topk(Q, None, V, k, hp, hi, mode="tc")
        fb = lib.rb2_fullsort_tc_last_fallback_rows()
        ids_f, sc_f = ops.fullsort_topk(Q, None, V, k, hp, hi, mode="fp32")
        ok = torch.equal(ids_t, ids_f) and torch.equal(sc_t, sc_f)
        bad += 0 if ok else 1
        print("variant %d nq=%-6d N=%-8d d=%-3d h=%-2d k=%-2d %-8s: %s, fallback rows %d (%.2f%%)" % (
            variant, nq, N, d, h, k, kind, "exact" if ok else "MISMATCH", fb, 100.0 * fb / nq), flush=True)
    print("mismatching cases:", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
