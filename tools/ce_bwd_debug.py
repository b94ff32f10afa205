"""development aid: error statistics of rb2_ce_head_backward against the float64 oracle"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ce_head as oce
from recbole_b200 import ops

def run(nq, N, seed=None, escale=0.4):
    rng = np.random.default_rng(nq + N if seed is None else seed)
    d = 64
    X = rng.standard_normal((nq, d)).astype(np.float32)
    X = (X - X.mean(1, keepdims=True)) / X.std(1, keepdims=True)
    E = (rng.standard_normal((N, d)) * escale).astype(np.float32)
    E[0] = 0
    tgt = rng.integers(1, N, nq) if N > 1 else np.zeros(nq, np.int64)
    dev = torch.device("cuda:0")
    Xd, Ed, td = torch.from_numpy(X).to(dev), torch.from_numpy(E).to(dev), torch.from_numpy(tgt).to(dev)
    out = ops.ce_head(Xd, Ed, td, k=1)
    dx, de = ops.ce_head_backward(Xd, Ed, td, out["lse"])
    dx, de = dx.cpu().numpy().astype(np.float64), de.cpu().numpy().astype(np.float64)
    o_dx, o_de = oce.ce_backward(X, E, tgt)
    for nm, a, b in (("dX", dx, o_dx.astype(np.float64)), ("dE", de, o_de.astype(np.float64))):
        dd = np.abs(a - b)
        rms = np.sqrt((b * b).mean())
        bad = dd > 1e-5 * np.abs(b) + 1e-5 * rms
        i = np.unravel_index(dd.argmax(), dd.shape)
        print("%5d x %7d %s: global %.2e  bad %d/%d  rms(b) %.3e max|b| %.3e  worst at %s: got %.6e want %.6e" % (
            nq, N, nm, dd.max() / np.abs(b).max(), bad.sum(), bad.size, rms, np.abs(b).max(), i, a[i], b[i]))
        if bad.sum():
            j = np.argwhere(bad)[:5]
            for jj in j:
                jj = tuple(jj)
                print("      bad", jj, a[jj], b[jj], "relerr %.2e" % (abs(a[jj] - b[jj]) / max(abs(b[jj]), 1e-30)))
    lse64 = None

for args in ((1, 3), (2, 5), (130, 127), (700, 40000), (512, 100000)):
    run(*args)

# timing at BASELINE config 4 (4096 x 1 000 001 x 64)
if "--time" in sys.argv:
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(0)
    nq, N = 4096, 1000001
    X = torch.nn.functional.layer_norm(torch.randn(nq, 64, device=dev, generator=g), (64,))
    E = torch.randn(N, 64, device=dev, generator=g) * 0.02
    tgt = torch.randint(1, N, (nq,), device=dev, generator=g)
    out = ops.ce_head(X, E, tgt, k=1)
    ws = ops.Workspace(ops.lib.rb2_ce_head_backward_workspace_bytes(nq, N, 64), dev)
    for _ in range(2):
        ops.ce_head_backward(X, E, tgt, out["lse"], ws=ws)
    torch.cuda.synchronize()
    ops.profile_enable(True); ops.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.ce_head_backward(X, E, tgt, out["lse"], ws=ws)
    e1.record(); torch.cuda.synchronize()
    st = ops.profile_read()
    ms = e0.elapsed_time(e1) / 5
    print("cfg4 backward: %.2f ms  (%.0f TFLOP/s algorithmic on 3 x 2BNH)" % (ms, 3 * 2.0 * nq * N * 64 / ms / 1e9))
    print({k: round(v[0] / 5, 3) for k, v in st.items() if v[1]})
