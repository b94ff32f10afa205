"""A few fused BPR steps on the real cfg3 workload (Cfg3Device: 10M x 2M x d=128, Zipf items, B=2^20) for ncu
and for a per-stage timing table.  usage: train_profile_cfg3.py [steps] [optimizer]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_workloads as bw
from recbole_b200 import ops

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
kind = sys.argv[2] if len(sys.argv) > 2 else "adam"
dev = torch.device("cuda:0")
w = bw.Cfg3Device(0, 1, dev, n_batches=4)
B, dim = w.batch, w.dim
gen = torch.Generator(device=dev); gen.manual_seed(0)
U = torch.randn(w.n_users, dim, device=dev, generator=gen) * 0.05
V = torch.randn(w.n_items, dim, device=dev, generator=gen) * 0.05
st = dict(mU=torch.zeros_like(U), vU=torch.zeros_like(U), mV=torch.zeros_like(V), vV=torch.zeros_like(V))
if kind == "adam_lazy":
    st["lastU"] = torch.zeros(w.n_users, dtype=torch.int32, device=dev)
    st["lastV"] = torch.zeros(w.n_items, dtype=torch.int32, device=dev)
loss = torch.zeros(1, device=dev)
ws = ops.bpr_workspace(B, dim, dev)
opt = ops.Optim(kind, lr=1e-3)
for i in range(2):
    ops.bpr_train_step(U, V, st, *w.batches[i % 4], opt, loss, None, ws)
torch.cuda.synchronize()
ops.profile_enable(True); ops.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    ops.bpr_train_step(U, V, st, *w.batches[(2 + i) % 4], opt, loss, None, ws)
e1.record(); torch.cuda.synchronize()
stg = ops.profile_read()
ms = e0.elapsed_time(e1) / steps
print("cfg3 %s: %.3f ms/step  %.1f M triples/s  frac %.3f" % (kind, ms, B / ms / 1e3, B * (72 * dim + 24) / ms / 1e6 / 6468.3))
print({k: round(v[0] / steps, 4) for k, v in stg.items()})
print("ok", loss.item())
