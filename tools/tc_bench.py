"""Development aid: time the scorer (fp32 vs tensor-core) on the BASELINE shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recbole_b200 import ops
from recbole_b200._lib import lib


def timeit(fn, warm=1, it=3):
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
    if len(sys.argv) > 1:
        lib.rb2_fullsort_tc_set_variant(int(sys.argv[1]))
        print("variant", sys.argv[1])
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    shapes = [("cfg2", 138493, 26745, 64, 20), ("cfg4", 4096, 1_000_001, 64, 0), ("cfg3-1gpu-64k", 65536, 2_000_001, 128, 20),
              ("cfg3-shard-256k", 262144, 250_001, 128, 20), ("cfg3-1gpu-512k", 524288, 2_000_001, 128, 0)]
    for name, nq, N, d, h in shapes:
        Q = torch.randn(nq, d, device=dev, generator=gen) * 0.1
        V = torch.randn(N, d, device=dev, generator=gen) * 0.1
        hp = hi = None
        if h:
            hp = torch.arange(0, h * nq + 1, h, device=dev, dtype=torch.int64)
            hi = torch.sort(torch.randint(1, N, (nq, h), device=dev, generator=gen), dim=1).values.reshape(-1).contiguous()
        for mode in ("tc", "fp32"):
            if mode == "fp32" and nq * N * d > 5e13:
                continue
            ops.profile_enable(True); ops.profile_read()
            ms = timeit(lambda: ops.fullsort_topk(Q, None, V, 10, hp, hi, mode=mode), warm=1, it=2)
            st = ops.profile_read(); ops.profile_enable(False)
            fl = 2.0 * nq * N * d
            extra = ""
            if mode == "tc":
                sc = st["tc_score"][0] / st["tc_score"][1]
                extra = " | tc_score %.2f ms = %.0f TFLOP/s, convert %.2f, refine %.2f, fallback rows %d" % (
                    sc, fl / sc / 1e9, st["tc_convert"][0] / st["tc_convert"][1], st["tc_refine"][0] / st["tc_refine"][1],
                    lib.rb2_fullsort_tc_last_fallback_rows())
            print("%-16s %-4s nq=%d N=%d d=%d: %.2f ms %.3f Mrows/s %.1f TFLOP/s%s" % (
                name, mode, nq, N, d, ms, nq / ms / 1e3, fl / ms / 1e9, extra), flush=True)
        del Q, V
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
