"""FM CTR step restated in numpy fp32 (TEST INFRASTRUCTURE).

Reference:
  * ``ContextRecommender.__init__`` token table + offsets   recbole/model/abstract_recommender.py:220-224
  * ``FMEmbedding.forward``  (id + per-field offset -> one table)  recbole/model/layers.py:141-144
  * ``BaseFactorizationMachine.forward`` 0.5*sum_k[(sum_f v)^2 - sum_f v^2]   layers.py:164-171
  * ``FMFirstOrderLinear`` token part + bias              layers.py:1021-1061
  * ``FM.forward / calculate_loss``  sigmoid, nn.BCELoss (mean, log clamped at -100)   fm.py:47-56
The functions up to new_state() cover TOKEN fields only (config 5 names categorical fields only);
fm_full_step() at the end restates the whole ContextRecommender field set (TOKEN + FLOAT + TOKEN_SEQ,
abstract_recommender.py:236-314, layers.py:947-1019) and is pinned on tests/golden/fm_float.npz / fm_seq.npz.
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


# ---- all three field kinds (abstract_recommender.py:236-314, layers.py:947-1019) ----------------------------------------
def fm_full_step(P, ids, fx, seqs, label, t, lr=1e-3, weight_decay=0.0):
    """One dense-Adam step of the reference's FM with TOKEN, FLOAT and TOKEN_SEQ fields, restated in numpy fp32.

    P: dict of parameter arrays under the reference's names (modified in place), with Adam moments under
    ``"m:" + name`` / ``"v:" + name`` (created on first use).  ids int64 [B, n_token] raw per-field ids (the token
    table's per-field offsets are derived from ``P["token_field_dims"]``), fx fp32 [B, n_float] or None,
    seqs: list of int64 [B, L_j] padded sequences (id 0 = padding) or [].  Returns the mean BCE loss (fp32).

    * token field f:  e = T[id + offset_f],  first order w = Tw[id + offset_f]          (layers.py:141-144, 966-987)
    * float field f:  e = x * F[f],          first order x * Fw[f]                      (:236-258, layers.py:947-966)
    * sequence j:     e = sum_{id != 0} S_j[id] / (count + 1e-8)   (mean pooling, :293-309),
                      first order sum_{id != 0} Sw_j[id]           (a plain masked SUM, layers.py:1000-1012)
    * y = sigmoid(sum first order + bias + 0.5 * sum_k[(sum_f e)^2 - sum_f e^2]),  nn.BCELoss (fm.py:47-56)
    """
    B = ids.shape[0]
    dims = [int(x) for x in P["token_field_dims"]]
    off = np.concatenate([[0], np.cumsum(dims)[:-1]]).astype(np.int64)
    T, Tw = P["token_embedding_table.embedding.weight"], P["first_order_linear.token_embedding_table.embedding.weight"]
    rows = ids + off[None, :]
    fields = [T[rows[:, f]] for f in range(rows.shape[1])]                     # list of [B, d]
    first = Tw[rows, 0].sum(axis=1, dtype=F32)
    n_float = 0 if fx is None else fx.shape[1]
    if n_float:
        Fe, Fw = P["float_embedding_table.weight"], P["first_order_linear.float_embedding_table.weight"]
        for f in range(n_float):
            fields.append((fx[:, f:f + 1] * Fe[f][None, :]).astype(F32))
        first = first + (fx * Fw[:, 0][None, :]).sum(axis=1, dtype=F32)
    coefs = []
    for j, q in enumerate(seqs):
        S = P["token_seq_embedding_table.%d.weight" % j]
        Sw = P["first_order_linear.token_seq_embedding_table.%d.weight" % j]
        mask = (q != 0).astype(F32)                                            # [B, L]
        cnt = mask.sum(axis=1, keepdims=True, dtype=F32)
        c = (F32(1) / (cnt + F32(1e-8))).astype(F32)                           # [B, 1]
        pooled = ((S[q] * mask[:, :, None]).sum(axis=1, dtype=F32) * c).astype(F32)
        fields.append(pooled)
        coefs.append((mask, c))
        first = first + (Sw[q, 0] * mask).sum(axis=1, dtype=F32)
    V = np.stack(fields, axis=1)                                               # [B, F_total, d]
    S_all = V.sum(axis=1, dtype=F32)
    second = F32(0.5) * ((S_all * S_all) - (V * V).sum(axis=1, dtype=F32)).sum(axis=1, dtype=F32)
    z = first + P["first_order_linear.bias"][0] + second
    y = _sigmoid(z.astype(F32))
    loss = fm_loss(y, label)
    gz = ((y - label) / F32(B)).astype(F32)
    G = {k: np.zeros_like(v) for k, v in P.items() if k != "token_field_dims" and not k[:2] in ("m:", "v:")}
    dV = gz[:, None, None] * (S_all[:, None, :] - V)                           # d loss / d field vector
    nt = rows.shape[1]
    for f in range(nt):
        np.add.at(G["token_embedding_table.embedding.weight"], rows[:, f], dV[:, f])
        np.add.at(G["first_order_linear.token_embedding_table.embedding.weight"][:, 0], rows[:, f], gz)
    for f in range(n_float):
        G["float_embedding_table.weight"][f] = (dV[:, nt + f] * fx[:, f:f + 1]).sum(axis=0, dtype=F32)
        G["first_order_linear.float_embedding_table.weight"][f, 0] = (gz * fx[:, f]).sum(dtype=F32)
    for j, q in enumerate(seqs):
        mask, c = coefs[j]
        g_entry = dV[:, nt + n_float + j][:, None, :] * (mask * c)[:, :, None]  # [B, L, d]
        np.add.at(G["token_seq_embedding_table.%d.weight" % j], q.reshape(-1), g_entry.reshape(-1, g_entry.shape[-1]))
        np.add.at(G["first_order_linear.token_seq_embedding_table.%d.weight" % j][:, 0], q.reshape(-1),
                  (gz[:, None] * mask).reshape(-1))
    G["first_order_linear.bias"][0] = gz.sum(dtype=F32)
    for k, g in G.items():
        m = P.setdefault("m:" + k, np.zeros_like(P[k]))
        v = P.setdefault("v:" + k, np.zeros_like(P[k]))
        optim.adam_dense_step(P[k], m, v, g.astype(F32), t, lr=lr, weight_decay=weight_decay)
    return loss
