"""A few fused FM steps at the Criteo shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_extra as be
from recbole_b200 import ops

dev = torch.device("cuda:0")
card, d, F, B = be.CRITEO_CARD, 16, 26, 1 << 18
rows = int(card.sum())
gen = torch.Generator(device=dev); gen.manual_seed(0)
E = torch.randn(rows, d, device=dev, generator=gen) * 0.01
W = torch.randn(rows, device=dev, generator=gen) * 0.01
bias3 = torch.zeros(3, device=dev)
st = dict(mE=torch.zeros_like(E), vE=torch.zeros_like(E), mW=torch.zeros_like(W), vW=torch.zeros_like(W))
import numpy as np
off = torch.from_numpy(np.concatenate([[0], np.cumsum(card)[:-1]]).astype(np.int64)).to(dev)
batches = [(torch.from_numpy(i).to(dev), torch.from_numpy(l).to(dev)) for i, l in be.fm_batches(card, B, 3)]
opt = ops.Optim("adam", lr=1e-3)
loss = torch.zeros(1, device=dev)
ws = ops.fm_workspace(B, F, d, dev)
for k in range(4):
    ops.fm_train_step(E, W, bias3, st, batches[k % 3][0], off, batches[k % 3][1], opt, loss, None, ws)
torch.cuda.synchronize()
print("ok", float(loss.item()))
