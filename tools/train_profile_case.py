"""A few fused BPR steps at the cfg3 shape (10M x 2M x d=128, B=2^20, uniform ids) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recbole_b200 import ops

dev = torch.device("cuda:0")
n_users, n_items, dim, B = 10_000_001, 2_000_001, 128, 1 << 20
gen = torch.Generator(device=dev); gen.manual_seed(0)
U = torch.randn(n_users, dim, device=dev, generator=gen) * 0.05
V = torch.randn(n_items, dim, device=dev, generator=gen) * 0.05
st = dict(mU=torch.zeros_like(U), vU=torch.zeros_like(U), mV=torch.zeros_like(V), vV=torch.zeros_like(V))
loss = torch.zeros(1, device=dev)
ws = ops.bpr_workspace(B, dim, dev)
opt = ops.Optim("adam", lr=1e-3)
for i in range(4):
    u = torch.randint(1, n_users, (B,), device=dev, generator=gen)
    p = torch.randint(1, n_items, (B,), device=dev, generator=gen)
    n = torch.randint(1, n_items, (B,), device=dev, generator=gen)
    ops.bpr_train_step(U, V, st, u, p, n, opt, loss, None, ws)
torch.cuda.synchronize()
print("ok", loss.item())
