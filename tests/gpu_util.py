import numpy as np
import torch


def dev():
    return torch.device("cuda:0")


def t(a, dtype=None):
    x = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        x = x.to(dtype)
    return x.to(dev()).contiguous()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def adam_state(U, V, lazy=False):
    st = dict(mU=torch.zeros_like(U), vU=torch.zeros_like(U), mV=torch.zeros_like(V), vV=torch.zeros_like(V))
    if lazy:
        st["lastU"] = torch.zeros(U.shape[0], dtype=torch.int32, device=U.device)
        st["lastV"] = torch.zeros(V.shape[0], dtype=torch.int32, device=V.device)
    return st
