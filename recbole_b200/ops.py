"""Thin tensor-level wrappers over the C ABI: validate, pass data_ptr(), raise on error.

PyTorch is plumbing here (device memory, streams); every operation below is one call into
librecbole_b200.so.  Nothing in this module computes on the CPU or through ATen.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import RB2Optim, check, lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype=None, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError("required tensor is None")
    if not t.is_cuda:
        raise ValueError("recbole_b200: tensor must live on a CUDA device (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError("expected %s, got %s" % (dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


class Workspace:
    """Device scratch owned by the caller side (the library never allocates).  The first bytes
    hold sticky error flags (an id outside its table) which are checked whenever the host
    synchronises anyway."""

    def __init__(self, nbytes, device):
        self.buf = torch.zeros(int(nbytes) + 256, dtype=torch.uint8, device=device)
        self.nbytes = int(nbytes)

    def ptr(self):
        return ctypes.c_void_p(self.buf.data_ptr())

    def check_flags(self):
        flags = self.buf[:8].view(torch.int32).tolist()
        if flags[1]:
            self.buf[:8].zero_()
            raise RuntimeError("recbole_b200: a cross-GPU barrier of the peer-memory step timed out "
                               "(a rank died or the ranks fell out of step)")
        if flags[0]:
            self.buf[:8].zero_()
            raise IndexError("recbole_b200: an id was outside its embedding table "
                             "(the reference raises IndexError in F.embedding)")


def grow_workspace(cache, batch, make):
    """`cache` is a one-entry dict {capacity: Workspace}.  Returns a workspace that holds at least `batch`
    samples, keeping the largest one ever made (the C ABI only needs workspace_bytes >= need, so the short tail
    batch of an epoch reuses the full-batch buffer and its sticky error flags survive until they are checked)."""
    for cap, ws in cache.items():
        if cap >= batch:
            return ws
    for ws in cache.values():
        ws.check_flags()            # about to be dropped: do not lose a recorded error
    cache.clear()
    cache[int(batch)] = make(int(batch))
    return cache[int(batch)]


class Optim:
    """Optimizer hyper-parameters + step counter (replaces the torch.optim object of
    recbole/trainer/trainer.py:109-130).  Scalars are derived in Python doubles exactly as
    torch/optim/adam.py:531-536 derives them."""

    def __init__(self, kind="adam", lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
        kinds = {"sgd": _lib.OPT_SGD, "adam": _lib.OPT_ADAM, "adam_lazy": _lib.OPT_ADAM_LAZY}
        if kind not in kinds:
            raise ValueError("optimizer kind must be one of %s" % sorted(kinds))
        self.kind_name = kind
        self.kind = kinds[kind]
        self.lr, self.weight_decay, self.betas, self.eps = float(lr), float(weight_decay), tuple(betas), float(eps)
        self.step = 0
        self._lazy = None  # (step_size_tab, bc2_sqrt_tab) device tensors

    def set_hyper(self, lr=None, weight_decay=None, betas=None, eps=None):
        """Change hyper-parameters (checkpoint resume, trainer.py:230); the adam_lazy bias-correction tables are
        functions of (lr, betas) and are rebuilt on the next step."""
        if lr is not None:
            self.lr = float(lr)
        if weight_decay is not None:
            self.weight_decay = float(weight_decay)
        if betas is not None:
            self.betas = tuple(float(b) for b in betas)
        if eps is not None:
            self.eps = float(eps)
        self._lazy = None

    def _lazy_tables(self, device, upto):
        cap = 0 if self._lazy is None else self._lazy[0].numel()
        if upto + 1 > cap:
            n = max(1024, 2 * (upto + 1))
            j = np.arange(n, dtype=np.float64)
            b1, b2 = self.betas
            with np.errstate(divide="ignore"):
                ss = self.lr / (1.0 - np.power(b1, j))
                bs = np.sqrt(1.0 - np.power(b2, j))
            ss[0], bs[0] = 0.0, 1.0
            self._lazy = (torch.from_numpy(ss.astype(np.float32)).to(device),
                          torch.from_numpy(bs.astype(np.float32)).to(device))
        return self._lazy

    def c_struct(self, device=None, step=None):
        t = self.step if step is None else step
        b1, b2 = self.betas
        o = RB2Optim()
        o.kind, o.step = self.kind, int(t)
        o.lr, o.weight_decay = self.lr, self.weight_decay
        o.beta1, o.beta2 = b1, b2
        o.one_minus_beta1, o.one_minus_beta2 = 1 - b1, 1 - b2
        o.eps = self.eps
        if self.kind != _lib.OPT_SGD and t >= 1:
            o.step_size = self.lr / (1 - b1 ** t)
            o.bc2_sqrt = math.sqrt(1 - b2 ** t)
        else:
            o.step_size, o.bc2_sqrt = self.lr, 1.0
        if self.kind == _lib.OPT_ADAM_LAZY:
            ss, bs = self._lazy_tables(device, int(t))
            o.lazy_step_size, o.lazy_bc2_sqrt = ss.data_ptr(), bs.data_ptr()
        return o


def bpr_workspace(batch, dim, device):
    return Workspace(lib.rb2_bpr_workspace_bytes(int(batch), int(dim)), device)


def bpr_train_step(U, V, state, user, pos, neg, optim, loss_out, loss_accum, ws, step=None):
    """One fused step (include/recbole_b200.h: rb2_bpr_train_step).  `state` holds mU, vU, mV, vV
    (and lastU, lastV for adam_lazy).  optim.step is incremented here (t starts at 1) unless the
    caller passes the step it has already counted."""
    if step is None:
        optim.step += 1
        step = optim.step
    o = optim.c_struct(U.device, step)
    f32, i64 = torch.float32, torch.int64
    check(lib.rb2_bpr_train_step(
        _ptr(U, f32), _ptr(state.get("mU"), f32, True), _ptr(state.get("vU"), f32, True),
        _ptr(state.get("lastU"), torch.int32, True),
        _ptr(V, f32), _ptr(state.get("mV"), f32, True), _ptr(state.get("vV"), f32, True),
        _ptr(state.get("lastV"), torch.int32, True),
        U.shape[0], V.shape[0], U.shape[1], _ptr(user, i64), _ptr(pos, i64), _ptr(neg, i64), user.numel(),
        ctypes.byref(o), _ptr(loss_out, f32), _ptr(loss_accum, torch.float64, True), ws.ptr(), ws.nbytes, _stream()))


def bpr_train_step_sharded(U, state, item_rows, user, pos_c, neg_c, global_batch, optim, loss_out, loss_accum,
                           item_grad_out, ws, step=None, item_touched=None, item_plan=None, rows_ready=None):
    """User side + per-compact-row item gradient sums (rb2_bpr_train_step_sharded).  optim.step is NOT
    incremented here (the owner-side update of the same logical step shares it).  rows_ready: a
    torch.cuda.Event recorded after the collective that fills item_rows on another stream; the id-only part
    of the step (keys, sorts) runs before it is waited for."""
    o = optim.c_struct(U.device, step)
    f32, i64 = torch.float32, torch.int64
    check(lib.rb2_bpr_train_step_sharded_ev(
        _ptr(U, f32), _ptr(state.get("mU"), f32, True), _ptr(state.get("vU"), f32, True),
        _ptr(state.get("lastU"), torch.int32, True), _ptr(item_rows, f32), U.shape[0], item_rows.shape[0],
        U.shape[1], _ptr(user, i64), _ptr(pos_c, i64), _ptr(neg_c, i64), user.numel(), int(global_batch),
        ctypes.byref(o), _ptr(loss_out, f32), _ptr(loss_accum, torch.float64, True), _ptr(item_grad_out, f32),
        _ptr(item_touched, torch.int32, True), item_plan.ptr() if item_plan is not None else None,
        ws.ptr(), ws.nbytes, _stream(), ctypes.c_void_p(rows_ready.cuda_event) if rows_ready is not None else None))


class PeerArena:
    """The per-rank buffers of the peer-memory step (include/recbole_b200.h (1e)), carved from ONE device
    allocation so that a single cudaIpc handle per rank maps everything: the item shard, the gradient slots,
    their stamps, the barrier flags and the loss slots.  `comm` needs all_gather_object() and barrier()."""

    def __init__(self, comm, device, item_block, dim):
        world, me = comm.world, comm.rank
        if world > _lib.MAX_PEERS:
            raise ValueError("the peer-memory step supports up to %d ranks" % _lib.MAX_PEERS)
        self.world, self.me, self.item_block, self.dim = world, me, int(item_block), int(dim)
        al = lambda x: (int(x) + 255) // 256 * 256       # noqa: E731
        sizes = [("item_p", self.item_block * dim * 4), ("grad_slots", world * self.item_block * dim * 4),
                 ("stamps", world * self.item_block * 4), ("flags", 2 * _lib.MAX_PEERS * 4),
                 ("loss_slots", 2 * _lib.MAX_PEERS * 8)]
        self.offsets, off = {}, 0
        for name, nbytes in sizes:
            self.offsets[name] = off
            off += al(nbytes)
        self.buf = torch.zeros(off, dtype=torch.uint8, device=device)
        o = self.offsets
        self.item_p = self.buf[o["item_p"]: o["item_p"] + self.item_block * dim * 4].view(torch.float32).view(
            self.item_block, dim)
        base = [self.buf.data_ptr()] * world
        if world > 1:
            handle = (ctypes.c_char * 64)()
            offset = ctypes.c_int64(0)
            check(lib.rb2_ipc_export(ctypes.c_void_p(self.buf.data_ptr()), handle, ctypes.byref(offset)))
            torch.cuda.synchronize(device)               # the zero fill is complete before anybody maps the arena
            everyone = comm.all_gather_object((bytes(handle.raw), int(offset.value)))
            for r, (h, ofs) in enumerate(everyone):
                if r == me:
                    continue
                mapped = ctypes.c_void_p()
                check(lib.rb2_ipc_open(ctypes.c_char_p(h), ofs, ctypes.byref(mapped)))
                base[r] = mapped.value
            comm.barrier()
        self.peers = _lib.RB2Peers()
        self.peers.world, self.peers.me, self.peers.item_block = world, me, self.item_block
        for r in range(world):
            for name in ("item_p", "grad_slots", "stamps", "flags", "loss_slots"):
                getattr(self.peers, name)[r] = base[r] + o[name]
        self.cache = torch.empty((world * self.item_block, dim), dtype=torch.float32, device=device)
        self.seq = 0            # calls of rb2_bpr_train_step_p2p on these buffers (the barrier sequence)


def bpr_p2p_workspace(batch, dim, device):
    return Workspace(lib.rb2_bpr_p2p_workspace_bytes(int(batch), int(dim)), device)


def bpr_train_step_p2p(U, state, arena, user, user_base, pos, neg, n_items, global_batch, optim, loss_out, loss_accum,
                       ws, step=None, next_batch=None):
    """rb2_bpr_train_step_p2p: the whole sharded step (both barriers, the owner update and the global mean
    loss included).  optim.step is NOT incremented here; `step` must count 1, 2, 3, ... identically on every rank."""
    o = optim.c_struct(U.device, step)
    f32, i64 = torch.float32, torch.int64
    arena.seq += 1
    arena.peers.seq = arena.seq
    # the previous call sorted THIS batch's keys while it waited for its peers (same workspace, same id tensors)?
    key_now = (ws.buf.data_ptr(), user.data_ptr(), pos.data_ptr(), neg.data_ptr(), int(user.numel()))
    prepared = 1 if getattr(arena, "prepared_for", None) == key_now else 0
    nxt, key_next = (None, None, None), None
    if next_batch is not None and int(next_batch[0].numel()) == int(user.numel()):
        nu, np_, nn = next_batch
        nxt = (_ptr(nu, i64), _ptr(np_, i64), _ptr(nn, i64))
        key_next = (ws.buf.data_ptr(), nu.data_ptr(), np_.data_ptr(), nn.data_ptr(), int(nu.numel()))
    check(lib.rb2_bpr_train_step_p2p(
        _ptr(U, f32), _ptr(state.get("mU"), f32, True), _ptr(state.get("vU"), f32, True),
        _ptr(state.get("mV"), f32, True), _ptr(state.get("vV"), f32, True), U.shape[0], int(n_items), U.shape[1],
        _ptr(user, i64), int(user_base), _ptr(pos, i64), _ptr(neg, i64), user.numel(), int(global_batch),
        ctypes.byref(o), ctypes.byref(arena.peers), _ptr(arena.cache, f32), _ptr(loss_out, f32),
        _ptr(loss_accum, torch.float64, True), ws.ptr(), ws.nbytes, _stream(), prepared, *nxt,
        _ptr(state.get("lastU"), torch.int32, True)))
    arena.prepared_for = key_next


def item_plan(pos, neg, n_items, bounds_dev, world, plan_ws=None):
    """rb2_item_plan: returns dict(uniq [2B] (first n_uniq valid), pos_c, neg_c, cuts [world+2] on the
    device, ws = the plan workspace to hand to bpr_train_step_sharded)."""
    B, dev = int(pos.numel()), pos.device
    need = lib.rb2_item_plan_workspace_bytes(B)
    if plan_ws is None or plan_ws.nbytes < need:
        plan_ws = Workspace(need, dev)
    uniq = torch.empty(2 * B, dtype=torch.int64, device=dev)
    pos_c = torch.empty(B, dtype=torch.int64, device=dev)
    neg_c = torch.empty(B, dtype=torch.int64, device=dev)
    cuts = torch.empty(world + 2, dtype=torch.int64, device=dev)
    check(lib.rb2_item_plan(_ptr(pos, torch.int64), _ptr(neg, torch.int64), B, int(n_items),
                            _ptr(bounds_dev, torch.int64), int(world), _ptr(uniq), _ptr(pos_c), _ptr(neg_c),
                            _ptr(cuts), plan_ws.ptr(), plan_ws.nbytes, _stream()))
    return dict(uniq=uniq, pos_c=pos_c, neg_c=neg_c, cuts=cuts, ws=plan_ws)


def dense_rows_update(P, M, V, grads, touched, optim, step=None):
    o = optim.c_struct(P.device, step)
    f32 = torch.float32
    check(lib.rb2_dense_rows_update(_ptr(P, f32), _ptr(M, f32, True), _ptr(V, f32, True), P.shape[0], P.shape[1],
                                    _ptr(grads, f32), _ptr(touched, torch.int32), ctypes.byref(o), _stream()))


def sparse_rows_update(P, M, V, last, ids, grads, optim, ws=None, step=None):
    """Sum duplicate rows of `grads` by `ids`, then one optimizer step per touched row."""
    o = optim.c_struct(P.device, step)
    n = int(ids.numel())
    if n == 0:
        return ws
    need = lib.rb2_sparse_rows_update_workspace_bytes(n, P.shape[1])
    if ws is None or ws.nbytes < need:
        ws = Workspace(need, P.device)
    f32 = torch.float32
    check(lib.rb2_sparse_rows_update(_ptr(P, f32), _ptr(M, f32, True), _ptr(V, f32, True),
                                     _ptr(last, torch.int32, True), P.shape[0], P.shape[1], _ptr(ids, torch.int64),
                                     _ptr(grads, f32), n, ctypes.byref(o), ws.ptr(), ws.nbytes, _stream()))
    return ws


def bpr_loss(U, V, user, pos, neg, loss_out, ws):
    f32, i64 = torch.float32, torch.int64
    check(lib.rb2_bpr_loss(_ptr(U, f32), _ptr(V, f32), U.shape[0], V.shape[0], U.shape[1], _ptr(user, i64),
                           _ptr(pos, i64), _ptr(neg, i64), user.numel(), _ptr(loss_out, f32), ws.ptr(), ws.nbytes,
                           _stream()))


def sort_positions(keys, key_bits, ws=None):
    """Stable (key, position) sort of uint32-valued keys < 2^key_bits held in an int32 tensor (rb2_sort_positions).
    Returns (keys_sorted, positions) as int32 tensors (bit patterns of uint32)."""
    n = int(keys.numel())
    if ws is None:
        ws = Workspace(lib.rb2_sort_positions_workspace_bytes(n), keys.device)
    ks, pos = torch.empty_like(keys), torch.empty_like(keys)
    check(lib.rb2_sort_positions(_ptr(keys, torch.int32), n, int(key_bits), _ptr(ks), _ptr(pos), ws.ptr(), ws.nbytes,
                                 _stream()))
    return ks, pos


def adam_lazy_flush(P, M, V, last, optim):
    o = optim.c_struct(P.device)
    check(lib.rb2_adam_lazy_flush(_ptr(P, torch.float32), _ptr(M, torch.float32), _ptr(V, torch.float32),
                                  _ptr(last, torch.int32), P.shape[0], P.shape[1], ctypes.byref(o), _stream()))


def fm_workspace(batch, n_fields, dim, device):
    return Workspace(lib.rb2_fm_workspace_bytes(int(batch), int(n_fields), int(dim)), device)


def _fm_float(fl, state=None):
    """fl = (values [B, Ff], Ef [Ff, d], Wf [Ff]) or None -> ctypes pointer to an rb2_fm_float (or None)."""
    if fl is None:
        return None
    values, Ef, Wf = fl
    f32 = torch.float32
    c = _lib.RB2FmFloat()
    c.values, c.n_float = _ptr(values, f32).value, int(values.shape[1])
    c.Ef, c.Wf = _ptr(Ef, f32).value, _ptr(Wf, f32).value
    if state is not None and "mEf" in state:
        c.mEf, c.vEf = _ptr(state["mEf"], f32).value, _ptr(state["vEf"], f32).value
        c.mWf, c.vWf = _ptr(state["mWf"], f32).value, _ptr(state["vWf"], f32).value
    return ctypes.byref(c)


FM_MAX_SEQ = 16   # RB2_FM_MAX_SEQ


def _fm_seq(seq):
    """seq = dict(n_token_cols, seq_start int32[n_seq+1], col_seq int32[n_cols], seq_row_base, pooled, coef) or None."""
    if seq is None:
        return None
    c = _lib.RB2FmSeq()
    c.n_seq, c.n_token_cols = int(seq["seq_start"].numel()) - 1, int(seq["n_token_cols"])
    c.seq_start, c.col_seq = _ptr(seq["seq_start"], torch.int32).value, _ptr(seq["col_seq"], torch.int32).value
    c.seq_row_base = int(seq["seq_row_base"])
    if seq.get("pooled") is not None:
        c.pooled, c.coef = _ptr(seq["pooled"], torch.float32).value, _ptr(seq["coef"], torch.float32).value
    return ctypes.byref(c)


def fm_train_step(E, W, bias3, state, ids, offsets, label, optim, loss_out, loss_accum, ws, floats=None, seq=None):
    """One fused FM step (rb2_fm_train_step).  state: mE, vE, mW, vW for Adam (+ mEf, vEf, mWf, vWf with float fields:
    floats = (values [B, Ff], Ef [Ff, d], Wf [Ff]))."""
    optim.step += 1
    o = optim.c_struct(E.device)
    f32 = torch.float32
    check(lib.rb2_fm_train_step(_ptr(E, f32), _ptr(state.get("mE"), f32, True), _ptr(state.get("vE"), f32, True),
                                _ptr(W, f32), _ptr(state.get("mW"), f32, True), _ptr(state.get("vW"), f32, True),
                                _ptr(bias3, f32), _ptr(state.get("last"), torch.int32, True), E.shape[0], E.shape[1],
                                _ptr(ids, torch.int64),
                                _ptr(offsets, torch.int64), ids.shape[1], _ptr(label, f32), ids.shape[0],
                                ctypes.byref(o), _ptr(loss_out, f32), _ptr(loss_accum, torch.float64, True), ws.ptr(),
                                ws.nbytes, _stream(), _fm_float(floats, state), _fm_seq(seq)))


def fm_lazy_flush(E, W, state, optim):
    """adam_lazy: bring every row of E and W to the optimizer's current step (rb2_fm_lazy_flush)."""
    o = optim.c_struct(E.device)
    f32 = torch.float32
    check(lib.rb2_fm_lazy_flush(_ptr(E, f32), _ptr(state["mE"], f32), _ptr(state["vE"], f32), _ptr(W, f32),
                                _ptr(state["mW"], f32), _ptr(state["vW"], f32), _ptr(state["last"], torch.int32),
                                E.shape[0], E.shape[1], ctypes.byref(o), _stream()))


def fm_grad_step(rows_e, rows_w, bias3, ids, offsets, label, global_batch, loss2, ws):
    """Local part of a row-sharded FM step (rb2_fm_grad_step): the fetched rows are replaced by their gradients."""
    f32 = torch.float32
    check(lib.rb2_fm_grad_step(_ptr(rows_e, f32), _ptr(rows_w, f32), _ptr(bias3, f32), rows_e.shape[0], rows_e.shape[1],
                               _ptr(ids, torch.int64), _ptr(offsets, torch.int64), ids.shape[1], _ptr(label, f32),
                               ids.shape[0], int(global_batch), _ptr(loss2, f32), ws.ptr(), ws.nbytes, _stream()))


def scalar_rows_update(P, M, V, ids, grads, optim, ws=None, step=None):
    """d = 1 table: sum duplicate ids' gradients (fixed order), one optimizer step per touched row."""
    o = optim.c_struct(P.device, step)
    n = int(ids.numel())
    if n == 0:
        return ws
    need = lib.rb2_scalar_rows_update_workspace_bytes(n)
    if ws is None or ws.nbytes < need:
        ws = Workspace(need, P.device)
    f32 = torch.float32
    check(lib.rb2_scalar_rows_update(_ptr(P, f32), _ptr(M, f32, True), _ptr(V, f32, True), P.shape[0],
                                     _ptr(ids, torch.int64), _ptr(grads, f32), n, ctypes.byref(o), ws.ptr(), ws.nbytes,
                                     _stream()))
    return ws


def scalar_step(p3, grad, optim, step=None):
    o = optim.c_struct(p3.device, step)
    check(lib.rb2_scalar_step(_ptr(p3, torch.float32), _ptr(grad, torch.float32), ctypes.byref(o), _stream()))


def fm_predict(E, W, bias3, ids, offsets, ws=None, floats=None, seq=None):
    B, F = ids.shape
    if ws is None:
        ws = fm_workspace(B, F, E.shape[1], E.device)
    y = torch.empty(B, dtype=torch.float32, device=E.device)
    f32 = torch.float32
    check(lib.rb2_fm_predict(_ptr(E, f32), _ptr(W, f32), _ptr(bias3, f32), E.shape[0], E.shape[1],
                             _ptr(ids, torch.int64), _ptr(offsets, torch.int64), F, B, _ptr(y), ws.ptr(), ws.nbytes,
                             _stream(), _fm_float(floats), _fm_seq(seq)))
    return y


def fm_loss(E, W, bias3, ids, offsets, label, loss_out, ws, floats=None, seq=None):
    """Forward + mean BCE only (rb2_fm_loss); loss_out[0] = the batch's loss."""
    f32 = torch.float32
    check(lib.rb2_fm_loss(_ptr(E, f32), _ptr(W, f32), _ptr(bias3, f32), E.shape[0], E.shape[1], _ptr(ids, torch.int64),
                          _ptr(offsets, torch.int64), ids.shape[1], _ptr(label, f32), ids.shape[0], _ptr(loss_out, f32),
                          ws.ptr(), ws.nbytes, _stream(), _fm_float(floats), _fm_seq(seq)))


def gather_dot(U, V, user, item):
    out = torch.empty(user.numel(), dtype=torch.float32, device=U.device)
    check(lib.rb2_gather_dot(_ptr(U, torch.float32), _ptr(V, torch.float32), U.shape[0], V.shape[0], U.shape[1],
                             _ptr(user, torch.int64), _ptr(item, torch.int64), user.numel(), _ptr(out), _stream()))
    return out


class ScorerState:
    """Caller-owned scorer state (include/recbole_b200.h rb2_scorer_state): knobs, the adaptive statistics that pick
    the tensor-core scorer's first pass, and what the last call did.  One per model / device / stream; calls that
    pass none share the calling thread's default state inside the library."""

    def __init__(self, variant=0, kprime=0, ce_scorer=0):
        self.c = _lib.RB2ScorerState()
        self.c.variant, self.c.kprime, self.c.ce_scorer = int(variant), int(kprime), int(ce_scorer)

    @property
    def last_fallback_rows(self):
        return int(self.c.last_fallback_rows)

    @property
    def last_pass2_rows(self):
        return int(self.c.last_pass2_rows)


def fullsort_topk(Q, query_ids, V, k, hist_indptr=None, hist_indices=None, item_base=0, mode="fp32", ws=None,
                  out=None, state=None):
    """Top-k item ids / scores per query row (rb2_fullsort_topk).  Returns (ids int64[nq,k],
    scores fp32[nq,k])."""
    nq = int(query_ids.numel()) if query_ids is not None else int(Q.shape[0])
    dev = V.device
    m = {"fp32": _lib.SCORER_FP32, "tc": _lib.SCORER_TC}[mode]
    need = lib.rb2_fullsort_workspace_bytes(nq, V.shape[0], V.shape[1], int(k), m)
    if ws is None or ws.nbytes < need:
        ws = Workspace(need, dev)
    if out is None:
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    else:
        ids, sc = out
    args = (_ptr(Q, torch.float32), _ptr(query_ids, torch.int64, True), nq, _ptr(V, torch.float32), V.shape[0],
            int(item_base), V.shape[1], _ptr(hist_indptr, torch.int64, True), _ptr(hist_indices, torch.int64, True),
            int(k), m, _ptr(ids), _ptr(sc), ws.ptr(), ws.nbytes, _stream())
    if state is None:
        check(lib.rb2_fullsort_topk(*args))
    else:
        check(lib.rb2_fullsort_topk_s(*args, ctypes.byref(state.c)))
    return ids, sc


def fullsort_scores(Q, query_ids, V):
    """The [nq, n_items] score matrix (rb2_fullsort_scores): compatibility path for an unmodified reference Trainer."""
    nq = int(query_ids.numel()) if query_ids is not None else Q.shape[0]
    out = torch.empty((nq, V.shape[0]), dtype=torch.float32, device=Q.device)
    check(lib.rb2_fullsort_scores(_ptr(Q, torch.float32), _ptr(query_ids, torch.int64, True), nq, Q.shape[0],
                                  _ptr(V, torch.float32), V.shape[0], V.shape[1], _ptr(out), _stream()))
    return out


def ce_head(X, E, target, k=10, scorer="auto"):
    """Fused full-sort CE head (rb2_ce_head): returns dict(loss 0-dim, lse [nq], ids [nq,k], scores [nq,k]).
    scorer: "auto" = tensor cores where covered (dim 64, k <= 16), "fp32" = the CUDA-core kernel."""
    nq, dev = X.shape[0], X.device
    state = ScorerState(ce_scorer={"auto": 0, "tc": 0, "fp32": 1}[scorer])
    ws = Workspace(lib.rb2_ce_head_workspace_bytes(nq, E.shape[0], E.shape[1], int(k)), dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    lse = torch.empty(nq, dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    check(lib.rb2_ce_head_s(_ptr(X, torch.float32), nq, _ptr(E, torch.float32), E.shape[0], E.shape[1],
                            _ptr(target, torch.int64, True), int(k), _ptr(loss), _ptr(lse), _ptr(ids), _ptr(sc),
                            ws.ptr(), ws.nbytes, _stream(), ctypes.byref(state.c)))
    return dict(loss=loss[0], lse=lse, ids=ids, scores=sc)


def ce_head_backward(X, E, target, lse, grad_scale=None, want_dx=True, want_de=True, ws=None):
    """Backward of the full-sort CE head (rb2_ce_head_backward): returns (dX [nq, d] or None, dE [N, d] or None) for
    loss = mean_b(logsumexp_b - logit[b, target_b]) * upstream; grad_scale defaults to 1 / nq."""
    nq, dev = X.shape[0], X.device
    if grad_scale is None:
        grad_scale = 1.0 / nq
    need = lib.rb2_ce_head_backward_workspace_bytes(nq, E.shape[0], E.shape[1])
    if need == 0:
        raise ValueError("ce_head_backward: hidden size %d is not served (64)" % E.shape[1])
    if ws is None or ws.nbytes < need:
        ws = Workspace(need, dev)
    dx = torch.empty_like(X) if want_dx else None
    de = torch.empty_like(E) if want_de else None
    f32 = torch.float32
    check(lib.rb2_ce_head_backward(_ptr(X, f32), nq, _ptr(E, f32), E.shape[0], E.shape[1], _ptr(target, torch.int64),
                                   _ptr(lse, f32), float(grad_scale), _ptr(dx, f32, True), _ptr(de, f32, True),
                                   ws.ptr(), ws.nbytes, _stream()))
    return dx, de


def dense_step(P, M, V, grad, optim, step=None):
    """One dense optimizer step over a whole tensor (rb2_dense_step; optim.step is NOT incremented here)."""
    o = optim.c_struct(P.device, step)
    f32 = torch.float32
    check(lib.rb2_dense_step(_ptr(P, f32), _ptr(M, f32, True), _ptr(V, f32, True), _ptr(grad, f32), P.numel(),
                             ctypes.byref(o), _stream()))


class CEHeadFunction(torch.autograd.Function):
    """loss = CrossEntropy(seq_output @ item_emb.T, pos) (sasrec.py:137-141) as ONE autograd node: forward = rb2_ce_head
    (tensor cores, logits never written), backward = rb2_ce_head_backward.  Gradients flow to seq_output (so the
    transformer below it trains with ordinary autograd) and to the item table."""

    @staticmethod
    def forward(ctx, seq_output, item_weight, pos_items):
        x, e = seq_output.detach().contiguous(), item_weight.detach().contiguous()
        out = ce_head(x, e, pos_items.contiguous(), k=1)
        ctx.save_for_backward(x, e, pos_items.contiguous(), out["lse"])
        return out["loss"].clone()

    @staticmethod
    def backward(ctx, grad_out):
        x, e, pos, lse = ctx.saved_tensors
        scale = float(grad_out.item()) / x.shape[0]        # the upstream of a scalar loss is a scalar (usually 1)
        dx, de = ce_head_backward(x, e, pos, lse, scale, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dx, de, None


def ce_head_loss(seq_output, item_weight, pos_items):
    return CEHeadFunction.apply(seq_output, item_weight, pos_items)


def topk_merge(ids, scores):
    """[parts, nq, k] sorted lists -> global [nq, k]."""
    parts, nq, k = ids.shape
    out_ids = torch.empty((nq, k), dtype=torch.int64, device=ids.device)
    out_sc = torch.empty((nq, k), dtype=torch.float32, device=ids.device)
    check(lib.rb2_topk_merge(_ptr(ids, torch.int64), _ptr(scores, torch.float32), parts, nq, k, _ptr(out_ids),
                             _ptr(out_sc), _stream()))
    return out_ids, out_sc


_DISCOUNT_CACHE = {}


def _discount_tables(k, device):
    key = (k, str(device))
    if key not in _DISCOUNT_CACHE:
        # metrics.py:131-141, float64 exactly as numpy computes them
        iranks = np.arange(1, k + 1, dtype=np.float64)
        disc = 1.0 / np.log2(iranks + 1)
        _DISCOUNT_CACHE[key] = (torch.from_numpy(disc).to(device), torch.from_numpy(np.cumsum(disc)).to(device))
    return _DISCOUNT_CACHE[key]


def topk_metrics(topk_ids, pos_indptr, pos_indices, n_items, want_hit=False, want_ref_idx=False):
    """Sums over rows of the six metrics at every rank (rb2_topk_metrics).
    Returns dict(sums=float64[6,k] (device), hit=uint8[nq,k]|None, ref_idx=int64[nq,k+1]|None)."""
    nq, k = topk_ids.shape
    dev = topk_ids.device
    disc, idcg = _discount_tables(k, dev)
    sums = torch.empty((_lib.NUM_METRICS, k), dtype=torch.float64, device=dev)
    hit = torch.empty((nq, k), dtype=torch.uint8, device=dev) if want_hit else None
    ref = torch.empty((nq, k + 1), dtype=torch.int64, device=dev) if want_ref_idx else None
    ws = Workspace(lib.rb2_topk_metrics_workspace_bytes(nq, k), dev)
    check(lib.rb2_topk_metrics(_ptr(topk_ids, torch.int64), nq, k, int(n_items), _ptr(pos_indptr, torch.int64),
                               _ptr(pos_indices, torch.int64), _ptr(disc), _ptr(idcg), _ptr(sums),
                               _ptr(hit, None, True), _ptr(ref, None, True), ws.ptr(), ws.nbytes, _stream()))
    return dict(sums=sums, hit=hit, ref_idx=ref)


def neg_sample_ref(key_ids, num, random_list, random_pr, used_indptr, used_indices):
    """Reference-stream sampler (rb2_neg_sample_ref).  Returns (ids int64[num*n], new random_pr)."""
    n = key_ids.numel()
    dev = key_ids.device
    out = torch.empty(n * num, dtype=torch.int64, device=dev)
    ws = Workspace(lib.rb2_neg_sample_workspace_bytes(n, num), dev)
    pr = ctypes.c_int64(int(random_pr))
    check(lib.rb2_neg_sample_ref(_ptr(key_ids, torch.int64), n, num, _ptr(random_list, torch.int64),
                                 random_list.numel(), ctypes.byref(pr), _ptr(used_indptr, torch.int64),
                                 _ptr(used_indices, torch.int64), used_indptr.numel() - 1, _ptr(out), ws.ptr(),
                                 ws.nbytes, _stream()))
    ws.check_flags()
    return out, pr.value


def neg_sample_hash(key_ids, num, n_items, used_indptr, used_indices, seed, step, out=None):
    n = key_ids.numel()
    if out is None:
        out = torch.empty(n * num, dtype=torch.int64, device=key_ids.device)
    check(lib.rb2_neg_sample_hash(_ptr(key_ids, torch.int64), n, num, int(n_items), _ptr(used_indptr, torch.int64),
                                  _ptr(used_indices, torch.int64), used_indptr.numel() - 1, int(seed), int(step),
                                  _ptr(out), _stream()))
    return out


def profile_enable(on=True):
    check(lib.rb2_profile_enable(1 if on else 0))


def profile_read():
    """{stage: (ms, calls, launches)} since the last read (synchronises the device)."""
    n = len(_lib.STAGES)
    ms = (ctypes.c_float * n)()
    calls = (ctypes.c_int64 * n)()
    launches = (ctypes.c_int64 * n)()
    check(lib.rb2_profile_read(ms, calls, launches))
    return {name: (float(ms[i]), int(calls[i]), int(launches[i])) for i, name in enumerate(_lib.STAGES)
            if calls[i]}
