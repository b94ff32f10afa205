"""Development aid: where do the roles of k_fullsort_tc wait?  (rb2_fullsort_tc_set_trace)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from recbole_b200 import ops
from recbole_b200._lib import lib

NAMES = ["producer: wait empty slot", "producer: wait A free", "mma: wait full slot", "mma: wait accumulator drained",
         "mma: wait A tile", "mma: total", "epilogue set 0: wait accumulator", "epilogue set 0: drain",
         "epilogue set 1: wait accumulator", "epilogue set 1: drain", "tiles", "producer: total"]


def show_events(ev):
    base = int(ev[0, 0])
    print("  CTA 0, tiles 1000..: cycles relative to the first event (epilogue columns: warp set 0, lane quarter 0)")
    print("  tile  mma_start mma_issued | stage_full  released  examined")
    for i in range(12):
        e = [int(x) - base if int(x) else -1 for x in ev[i]]
        print("  %4d  %9d %10d | %10d %9d %9d" % (1000 + i, e[0], e[2], e[4], e[5], e[6]))


def main():
    variant = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    nq, N, d = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (131072, 2_000_001, 128)
    lib.rb2_fullsort_tc_set_variant(variant)
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    Q = torch.randn(nq, d, device=dev, generator=gen) * 0.1
    V = torch.randn(N, d, device=dev, generator=gen) * 0.1
    ops.fullsort_topk(Q, None, V, 10, mode="tc")
    buf = torch.zeros(148 * 16 + 64 * 16, dtype=torch.int64, device=dev)
    lib.rb2_fullsort_tc_set_trace(ctypes.c_void_p(buf.data_ptr()))
    ops.profile_enable(True); ops.profile_read()
    ops.fullsort_topk(Q, None, V, 10, mode="tc")
    st = ops.profile_read(); ops.profile_enable(False)
    lib.rb2_fullsort_tc_set_trace(None)
    torch.cuda.synchronize()
    t = buf[:148 * 16].view(148, 16).double().cpu()
    ev = buf[148 * 16:].view(64, 16).cpu()
    tiles = t[:, 10].clone()
    tiles[1::2] = torch.where(tiles[1::2] > 0, tiles[1::2], tiles[0::2])   # cta_group::2: only CTA 0 of a pair issues
    tiles = tiles.clamp(min=1)
    ms = st["tc_score"][0] / st["tc_score"][1]
    print("variant %d, %d x %d x %d: tc_score %.2f ms = %.0f TFLOP/s; tiles per CTA %.0f, cycles per tile (mma total) %.0f" % (
        variant, nq, N, d, ms, 2.0 * nq * N * d / ms / 1e9, tiles.mean().item(), (t[:, 5] / tiles).mean().item()))
    for i, n in enumerate(NAMES):
        if i == 10:
            continue
        per = t[:, i] / tiles
        print("  %-36s %8.0f cycles per tile (min %.0f, max %.0f over CTAs)" % (n, per.mean().item(), per.min().item(), per.max().item()))
    show_events(ev)


if __name__ == "__main__":
    main()
