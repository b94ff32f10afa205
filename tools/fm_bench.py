"""Development aid: time the fused FM step at the Criteo shape of BASELINE config 5."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recbole_b200 import ops

dev = torch.device("cuda:0")
# Criteo-Kaggle categorical cardinalities (26 fields), scaled so that the sum is ~33M
card = np.array([1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992, 5461306,
                 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572], dtype=np.int64)
off = np.concatenate([[0], np.cumsum(card)[:-1]])
rows, d, F = int(card.sum()), 16, 26
print("rows", rows)
gen = torch.Generator(device=dev); gen.manual_seed(0)
E = torch.randn(rows, d, device=dev, generator=gen) * 0.01
W = torch.randn(rows, device=dev, generator=gen) * 0.01
bias3 = torch.zeros(3, device=dev)
st = dict(mE=torch.zeros_like(E), vE=torch.zeros_like(E), mW=torch.zeros_like(W), vW=torch.zeros_like(W))
for B in (2048, 1 << 18, 1 << 20):
    u = torch.rand(B, F, device=dev, generator=gen, dtype=torch.float64)
    ids = torch.minimum(torch.exp(u * torch.log(torch.tensor(card, device=dev, dtype=torch.float64))).long(),
                        torch.tensor(card - 1, device=dev))  # Zipf(1)-like per field
    lab = (torch.rand(B, device=dev, generator=gen) < 0.256).float()
    offs = torch.from_numpy(off).to(dev)
    opt = ops.Optim("adam", lr=1e-3)
    loss = torch.zeros(1, device=dev)
    ws = ops.fm_workspace(B, F, d, dev)
    for _ in range(3):
        ops.fm_train_step(E, W, bias3, st, ids, offs, lab, opt, loss, None, ws)
    torch.cuda.synchronize()
    ops.profile_enable(True); ops.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.fm_train_step(E, W, bias3, st, ids, offs, lab, opt, loss, None, ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    stg = ops.profile_read(); ops.profile_enable(False)
    alg = B * (F * (24 * d + 24) + 8 * F + 4)
    print("FM B=%d: %.3f ms  %.1f Msamples/s  alg %.0f GB/s  loss %.4f  stages %s" % (
        B, ms, B / ms / 1e3, alg / ms / 1e6, loss.item(), {k: round(v[0] / v[1], 3) for k, v in stg.items()}), flush=True)
