"""BPR-MF training step restated in numpy fp32 (TEST INFRASTRUCTURE).

Follows, in the reference:
  * ``BPR.calculate_loss``   recbole/model/general_recommender/bpr.py:74-83
  * ``BPRLoss.forward``      recbole/model/loss.py:43-49   (gamma = 1e-10)
  * autograd of the above into dense [rows, d] grads (triggered at trainer.py:170)
  * ``BPR.predict``          bpr.py:85-89
and the point-wise "dot" variant of the fork's MFSimple
(recbole/model/general_recommender/mfsimple.py:39-57, BCELoss on sigmoid(u.v + b_u + b_i + b)).
"""
import numpy as np

from . import optim

F32 = np.float32
GAMMA = 1e-10


def _sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def bpr_forward(U, V, user, pos, neg):
    """loss (fp32 scalar) and per-sample x = s+ - s- ; bpr.py:74-83 + loss.py:48."""
    u, p, n = U[user], V[pos], V[neg]
    ps = (u * p).sum(axis=1, dtype=F32)
    ns = (u * n).sum(axis=1, dtype=F32)
    x = ps - ns
    sig = _sigmoid(x)
    loss = -np.log(F32(GAMMA) + sig, dtype=F32).mean(dtype=F32)
    return F32(loss), x, sig


def bpr_grads(U, V, user, pos, neg):
    """Dense grads exactly as autograd accumulates them (duplicates summed)."""
    B = len(user)
    loss, x, sig = bpr_forward(U, V, user, pos, neg)
    # d/dx of -mean(log(gamma + sigmoid(x)))
    g = (-(F32(1) / F32(B)) * sig * (F32(1) - sig) / (F32(GAMMA) + sig)).astype(F32)
    u, p, n = U[user], V[pos], V[neg]
    dU = np.zeros_like(U)
    dV = np.zeros_like(V)
    np.add.at(dU, user, g[:, None] * p - g[:, None] * n)
    np.add.at(dV, pos, g[:, None] * u)
    np.add.at(dV, neg, -g[:, None] * u)
    return loss, dU, dV, g


def bpr_train_step(state, user, pos, neg, t, optimizer="adam", lr=1e-3, weight_decay=0.0,
                   beta1=0.9, beta2=0.999, eps=1e-8, dense=True):
    """One ``zero_grad -> calculate_loss -> backward -> step`` (trainer.py:160-173).

    ``state`` is a dict with U, V and (for Adam) mU, vU, mV, vV; updated in place.
    ``dense=True`` is the reference's semantics (every row steps); ``dense=False``
    restricts the update to rows that occur in the batch (the row-sparse mode).
    Returns the fp32 loss.
    """
    U, V = state["U"], state["V"]
    loss, dU, dV, _ = bpr_grads(U, V, user, pos, neg)
    if optimizer == "sgd":
        optim.sgd_step(U, dU, lr, weight_decay)
        optim.sgd_step(V, dV, lr, weight_decay)
        return loss
    kw = dict(lr=lr, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay)
    if dense:
        optim.adam_dense_step(U, state["mU"], state["vU"], dU, t, **kw)
        optim.adam_dense_step(V, state["mV"], state["vV"], dV, t, **kw)
    else:
        ur = np.unique(user)
        ir = np.unique(np.concatenate([pos, neg]))
        optim.adam_rowsparse_step(U, state["mU"], state["vU"], ur, dU[ur], t, **kw)
        optim.adam_rowsparse_step(V, state["mV"], state["vV"], ir, dV[ir], t, **kw)
    return loss


def new_state(U, V):
    U = np.array(U, dtype=F32, copy=True)
    V = np.array(V, dtype=F32, copy=True)
    return dict(U=U, V=V, mU=np.zeros_like(U), vU=np.zeros_like(U), mV=np.zeros_like(V), vV=np.zeros_like(V))


def predict(U, V, user, item):
    """bpr.py:85-89 -- torch.mul(u, i).sum(dim=1); fp32, summation order unspecified."""
    return (U[user] * V[item]).sum(axis=1, dtype=F32)


# ---- point-wise "dot" loss (fork's MFSimple, mfsimple.py:39-57) -------------------------------

def dot_bce_forward(U, V, bu, bi, b, user, item, label):
    z = (U[user] * V[item]).sum(axis=1, dtype=F32) + bu[user] + bi[item] + F32(b)
    y = _sigmoid(z)
    # nn.BCELoss clamps log at -100 (SURVEY.md 8c probe)
    ly = np.maximum(np.log(y, dtype=F32), F32(-100))
    l1y = np.maximum(np.log(F32(1) - y, dtype=F32), F32(-100))
    loss = -(label * ly + (F32(1) - label) * l1y).mean(dtype=F32)
    return F32(loss), y


def dot_bce_grads(U, V, bu, bi, b, user, item, label):
    B = len(user)
    loss, y = dot_bce_forward(U, V, bu, bi, b, user, item, label)
    gz = ((y - label) / F32(B)).astype(F32)  # d/dz of mean BCE(sigmoid(z)) away from the clamp
    dU, dV = np.zeros_like(U), np.zeros_like(V)
    dbu, dbi = np.zeros_like(bu), np.zeros_like(bi)
    np.add.at(dU, user, gz[:, None] * V[item])
    np.add.at(dV, item, gz[:, None] * U[user])
    np.add.at(dbu, user, gz)
    np.add.at(dbi, item, gz)
    return loss, dU, dV, dbu, dbi, F32(gz.sum(dtype=F32))
