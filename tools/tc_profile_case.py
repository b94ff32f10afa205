"""One tensor-core scorer call at a cfg3-like shape (for ncu): nq x N x d given on the command line."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recbole_b200 import ops
from recbole_b200._lib import lib

nq, N, d = (int(x) for x in sys.argv[1:4])
h = int(sys.argv[4]) if len(sys.argv) > 4 else 20
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev); gen.manual_seed(0)
Q = torch.randn(nq, d, device=dev, generator=gen) * 0.1
V = torch.randn(N, d, device=dev, generator=gen) * 0.1
hp = torch.arange(0, h * nq + 1, h, device=dev, dtype=torch.int64)
hi = torch.sort(torch.randint(1, N, (nq, h), device=dev, generator=gen), dim=1).values.reshape(-1).contiguous()
for _ in range(2):
    ids, sc = ops.fullsort_topk(Q, None, V, 10, hp, hi, mode="tc")
torch.cuda.synchronize()
print("ok fallback rows", lib.rb2_fullsort_tc_last_fallback_rows(), ids[0].tolist())
