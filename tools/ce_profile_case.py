"""One tensor-core CE head call at the BASELINE config 4 shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recbole_b200 import ops

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev); gen.manual_seed(4)
nq, N, d = 4096, 1_000_001, 64
X = torch.randn(nq, d, device=dev, generator=gen)
X = (X - X.mean(1, keepdim=True)) / X.std(1, keepdim=True)
E = torch.randn(N, d, device=dev, generator=gen) * 0.02
E[0] = 0
tgt = torch.randint(1, N, (nq,), device=dev, generator=gen)
for _ in range(2):
    out = ops.ce_head(X, E, tgt, 10)
torch.cuda.synchronize()
print("ok", float(out["loss"].item()))
