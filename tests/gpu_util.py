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


def elem_rel_err(a, b, floor=1e-3):
    """Element-wise relative error with an absolute floor: max_i |a_i - b_i| / max(|b_i|, floor * max |b|).  rel_err
    above divides every difference by the GLOBAL maximum, which says little about small elements; here an element
    100x smaller than the largest is still held to its own magnitude, and only elements below floor * max |b| (fp32
    sums of d products of O(1)-relative terms carry an absolute error of that order) are measured against the floor."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), floor * max(np.abs(b).max(), 1e-30))
    return float((np.abs(a - b) / den).max())


def adam_state(U, V, lazy=False):
    st = dict(mU=torch.zeros_like(U), vU=torch.zeros_like(U), mV=torch.zeros_like(V), vV=torch.zeros_like(V))
    if lazy:
        st["lastU"] = torch.zeros(U.shape[0], dtype=torch.int32, device=U.device)
        st["lastV"] = torch.zeros(V.shape[0], dtype=torch.int32, device=V.device)
    return st
