"""Optimizer arithmetic of the reference path, restated in numpy fp32 (TEST INFRASTRUCTURE).

The reference builds ``torch.optim.Adam`` / ``torch.optim.SGD`` in
``recbole/trainer/trainer.py:109-130`` and calls ``optimizer.step()`` at
``trainer.py:173``.  The arithmetic itself lives in PyTorch (third-party, pinned
``torch>=1.7.0`` in the reference's requirements.txt:2; 2.11.0+cu128 installed here).
Restated from ``torch/optim/adam.py`` single-tensor path (2.11.0):

    grad  = grad + wd * p                       (adam.py:416-429, L2-in-gradient)
    m.lerp_(grad, 1 - beta1)                    (adam.py:457)   m += (1-b1) * (g - m)
    v.mul_(beta2).addcmul_(grad, grad, 1-beta2) (adam.py:476)
    bc1 = 1 - beta1**t ; bc2 = 1 - beta2**t     (python doubles, adam.py:531-532)
    step_size = lr / bc1 ; bc2_sqrt = bc2**0.5  (adam.py:534-536)
    denom = (v.sqrt() / bc2_sqrt) + eps         (adam.py:545)
    p.addcdiv_(m, denom, value=-step_size)      (adam.py:547)

``t`` starts at 1 and is one counter per parameter tensor.  SGD (sgd.py:343-375,
momentum 0): ``p.add_(grad + wd*p, alpha=-lr)``.
"""
import numpy as np

F32 = np.float32


def adam_hparams(t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """Host-side scalars exactly as torch computes them (python doubles)."""
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    return dict(step_size=lr / bc1, bc2_sqrt=bc2 ** 0.5, one_minus_beta1=1 - beta1, beta2=beta2,
                one_minus_beta2=1 - beta2, eps=eps)


def adam_dense_step(p, m, v, g, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """In-place dense Adam over the whole tensor: what the reference does every step."""
    h = adam_hparams(t, lr, beta1, beta2, eps)
    g = g.astype(F32, copy=True)
    if weight_decay != 0:
        g += F32(weight_decay) * p
    m += F32(h["one_minus_beta1"]) * (g - m)
    v *= F32(beta2)
    v += F32(h["one_minus_beta2"]) * g * g
    denom = np.sqrt(v) / F32(h["bc2_sqrt"]) + F32(eps)
    p += F32(-h["step_size"]) * (m / denom)
    return p, m, v


def adam_rowsparse_step(p, m, v, rows, g_rows, t, **kw):
    """Row-sparse Adam: the dense formula applied to the touched rows only.

    This is what the CUDA path runs in ``optimizer='adam_sparse'`` mode.  It equals
    the reference's dense Adam whenever every row with non-zero (m, v) is touched,
    in particular for the first step from zero state and for any single step
    restricted to the touched rows (SURVEY.md section 7 hard part A).
    """
    pr, mr, vr = p[rows].copy(), m[rows].copy(), v[rows].copy()
    adam_dense_step(pr, mr, vr, g_rows, t, **kw)
    p[rows], m[rows], v[rows] = pr, mr, vr
    return p, m, v


def sgd_step(p, g, lr=1e-3, weight_decay=0.0):
    g = g.astype(F32, copy=True)
    if weight_decay != 0:
        g += F32(weight_decay) * p
    p += F32(-lr) * g
    return p
