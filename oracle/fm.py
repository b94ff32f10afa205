"""FM (token fields) CTR step restated in numpy fp32 (TEST INFRASTRUCTURE).

Reference:
  * ``ContextRecommender.__init__`` token table + offsets   recbole/model/abstract_recommender.py:220-224
  * ``FMEmbedding.forward``  (id + per-field offset -> one table)  recbole/model/layers.py:141-144
  * ``BaseFactorizationMachine.forward`` 0.5*sum_k[(sum_f v)^2 - sum_f v^2]   layers.py:164-171
  * ``FMFirstOrderLinear`` token part + bias              layers.py:1021-1061
  * ``FM.forward / calculate_loss``  sigmoid, nn.BCELoss (mean, log clamped at -100)   fm.py:47-56
Only TOKEN fields (config 5 names categorical fields only).
"""
import numpy as np

from . import optim

F32 = np.float32


def _sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def fm_forward(E, W, bias, rows):
    """E [V,d] second-order table, W [V] first-order table (d=1), rows int64[B,F] = id + offset."""
    v = E[rows]                                   # [B,F,d]
    s = v.sum(axis=1, dtype=F32)                  # [B,d]
    fm = F32(0.5) * ((s * s) - (v * v).sum(axis=1, dtype=F32)).sum(axis=1, dtype=F32)
    first = W[rows].sum(axis=1, dtype=F32) + F32(bias)
    z = first + fm
    return _sigmoid(z), s


def fm_loss(y, label):
    ly = np.maximum(np.log(y, dtype=F32), F32(-100))
    l1y = np.maximum(np.log(F32(1) - y, dtype=F32), F32(-100))
    return F32(-(label * ly + (F32(1) - label) * l1y).mean(dtype=F32))


def fm_grads(E, W, bias, rows, label):
    B = rows.shape[0]
    y, s = fm_forward(E, W, bias, rows)
    loss = fm_loss(y, label)
    gz = ((y - label) / F32(B)).astype(F32)       # d loss / d z  (away from the BCE clamp)
    v = E[rows]
    gv = gz[:, None, None] * (s[:, None, :] - v)  # d z / d v_f = S - v_f
    dE = np.zeros_like(E)
    dW = np.zeros_like(W)
    np.add.at(dE, rows.reshape(-1), gv.reshape(-1, E.shape[1]))
    np.add.at(dW, rows.reshape(-1), np.repeat(gz, rows.shape[1]))
    return loss, dE, dW, F32(gz.sum(dtype=F32)), y


def fm_train_step(state, rows, label, t, dense=True, **kw):
    """state: E, W, b (shape [1]) and Adam moments mE, vE, mW, vW, mb, vb."""
    loss, dE, dW, db, _ = fm_grads(state["E"], state["W"], state["b"][0], rows, label)
    if dense:
        optim.adam_dense_step(state["E"], state["mE"], state["vE"], dE, t, **kw)
        optim.adam_dense_step(state["W"], state["mW"], state["vW"], dW, t, **kw)
    else:
        r = np.unique(rows)
        optim.adam_rowsparse_step(state["E"], state["mE"], state["vE"], r, dE[r], t, **kw)
        optim.adam_rowsparse_step(state["W"], state["mW"], state["vW"], r, dW[r], t, **kw)
    optim.adam_dense_step(state["b"], state["mb"], state["vb"], np.array([db], dtype=F32), t, **kw)
    return loss


def new_state(E, W, b=0.0):
    E = np.array(E, dtype=F32, copy=True)
    W = np.array(W, dtype=F32, copy=True).reshape(-1)
    bb = np.array([b], dtype=F32)
    z = np.zeros_like
    return dict(E=E, W=W, b=bb, mE=z(E), vE=z(E), mW=z(W), vW=z(W), mb=z(bb), vb=z(bb))
