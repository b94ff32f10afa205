"""One rb2_ce_head_backward call at a cfg4-like shape for ncu (4096 x N x 64; N from argv, default 250 000)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recbole_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
nq = 4096
X = torch.nn.functional.layer_norm(torch.randn(nq, 64, device=dev, generator=g), (64,))
E = torch.randn(N, 64, device=dev, generator=g) * 0.02
tgt = torch.randint(1, N, (nq,), device=dev, generator=g)
out = ops.ce_head(X, E, tgt, k=1)
for _ in range(2):
    dx, de = ops.ce_head_backward(X, E, tgt, out["lse"])
torch.cuda.synchronize()
print("ok", float(dx.abs().max()), float(de.abs().max()))
