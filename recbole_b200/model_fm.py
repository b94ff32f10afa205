"""Host-side mirror of the reference's FM context-aware recommender for the fused path.

``FusedFM`` exposes the plugin API of ``recbole.model.context_aware_recommender.fm.FM``
(fm.py:26-59) on top of ``ContextRecommender`` (recbole/model/abstract_recommender.py:151-412):
constructor ``(config, dataset)``, ``calculate_loss / predict``, and parameters under the reference's
names (``token_embedding_table.embedding.weight``, ``first_order_linear.token_embedding_table.embedding.weight``,
``first_order_linear.bias``, ``float_embedding_table.weight``, ``token_seq_embedding_table.<i>.weight`` ...) so state
dicts interchange.  TOKEN, FLOAT and TOKEN_SEQ (mean-pooled) fields.
"""
import numpy as np
import torch
from torch import nn

from . import ops
from .enums import InputType, ModelType
from .model import FusedOptimizer, fused_learner, xavier_normal_

FEATURE_TOKEN = "token"


def _is_token(ftype):
    return getattr(ftype, "value", ftype) == FEATURE_TOKEN


class _RecordPointwise(torch.autograd.Function):
    """Forward: the forward + BCE kernel (rb2_fm_loss).  Backward: remember the batch for FusedOptimizer.step() --
    or, when no fused optimizer was handed out (unmodified reference Trainer), take the fused step right here."""

    @staticmethod
    def forward(ctx, anchor, model, interaction):
        ctx.model, ctx.batch = model, interaction
        return model._loss_value(interaction)

    @staticmethod
    def backward(ctx, grad_out):
        m = ctx.model
        m._pending = ctx.batch
        if m._autostep:
            m._apply_pending()
        return None, None, None


class _PointwiseMixin:
    """What FusedOptimizer and the unmodified-Trainer mode need from a point-wise fused model."""

    def _init_fused(self, config):
        self._optim, self._state, self._ws = None, {}, {}
        self._offsets = self._bias3 = self._loss_out = self._loss_accum = None
        self._pending = None
        self._hyper = (config["learner"], config["learning_rate"], config["weight_decay"])
        self._autostep = True

    def flush(self):
        """adam_lazy only: bring every row up to the current step before the tables are read."""
        if self._optim is not None and self._optim.kind_name == "adam_lazy" and self._optim.step > 0:
            E, W = self._tables()
            ops.fm_lazy_flush(E, W, self._state, self._optim)

    def state_dict(self, *args, **kwargs):
        self.flush()
        return self._state_dict_impl(*args, **kwargs)

    def _new_state(self, kind, E, W):
        st = {}
        if kind != "sgd":
            st = dict(mE=torch.zeros_like(E), vE=torch.zeros_like(E), mW=torch.zeros_like(W), vW=torch.zeros_like(W))
        if kind == "adam_lazy":
            st["last"] = torch.zeros(E.shape[0], dtype=torch.int32, device=E.device)
        return st

    def _after_state_load(self):
        if "last" in self._state:
            self._state["last"].fill_(self._optim.step)

    def _apply_pending(self):
        if self._optim is None:     # unmodified Trainer: hyper-parameters from the config it reads itself
            learner, lr, wd = self._hyper
            self._make_optimizer(fused_learner(learner), lr if lr is not None else 1e-3, wd or 0.0)
        inter = self._pending
        self._pending = None
        self._fused_step(inter, None)

    def build_optimizer(self, learner="adam", learning_rate=1e-3, weight_decay=0.0):
        """Trainer._build_optimizer (trainer.py:109-130) for the fused path: 'adam' (row-sparse), 'adam_lazy' (row-
        sparse work, the trajectory of the reference's dense Adam incl. weight decay) or 'sgd'.  Returns a
        FusedOptimizer (zero_grad / step / state_dict / load_state_dict); train_step() needs nothing else."""
        self._make_optimizer(learner.lower(), learning_rate, weight_decay or 0.0)
        self._autostep = False
        return FusedOptimizer(self)

    def train_step(self, interaction):
        if self._optim is None:
            raise RuntimeError("call build_optimizer() first")
        return self._fused_step(interaction, self._loss_accum)

    def calculate_loss(self, interaction):
        """0-dim loss tensor (fm.py:52-56 / mfsimple.py:48-57) computed by the fused forward kernel; its backward()
        records the batch (see model.py)."""
        anchor = next(self.parameters())
        return _RecordPointwise.apply(anchor, self, interaction)


class _Table(nn.Module):
    def __init__(self, rows, dim):
        super().__init__()
        self.embedding = nn.Embedding(rows, dim)


class _FirstOrder(nn.Module):
    def __init__(self, rows, n_float=0, seq_dims=()):
        super().__init__()
        self.token_embedding_table = _Table(rows, 1)
        if n_float:
            self.float_embedding_table = nn.Embedding(n_float, 1)       # layers.py:939-940
        if seq_dims:
            self.token_seq_embedding_table = nn.ModuleList(nn.Embedding(n, 1) for n in seq_dims)   # layers.py:941-944
        self.bias = nn.Parameter(torch.zeros((1,)), requires_grad=True)  # layers.py:945


FEATURE_FLOAT = "float"
FEATURE_TOKEN_SEQ = "token_seq"
_KERNEL_DIMS = (16, 32, 64, 128)


class FusedFM(_PointwiseMixin, nn.Module):
    """TOKEN, FLOAT and TOKEN_SEQ fields (abstract_recommender.py:205-314; sequences are mean-pooled over their non-zero
    ids, the reference's default mode).  Any embedding_size up to 128 (FM.yaml's default is 10): the kernels work on
    rows of 16 / 32 / 64 / 128 floats, other sizes live in zero-padded rows whose padding never moves (its gradient is
    exactly zero) and the parameters are views of the first embedding_size columns.  With TOKEN_SEQ fields the token
    table and every sequence table are views of ONE row range (token rows first), which is what the kernels index."""
    input_type = InputType.POINTWISE   # abstract_recommender.py:157
    type = ModelType.CONTEXT           # abstract_recommender.py:156

    def __init__(self, config, dataset):
        super().__init__()
        self.LABEL = config["LABEL_FIELD"]
        self.embedding_size = int(config["embedding_size"] or 10)          # properties/model/FM.yaml:1
        if self.embedding_size > _KERNEL_DIMS[-1]:
            raise ValueError("FusedFM: embedding_size %d > %d" % (self.embedding_size, _KERNEL_DIMS[-1]))
        self._dpad = next(k for k in _KERNEL_DIMS if k >= self.embedding_size)
        self.device = config["device"]
        self.token_field_names, self.token_field_dims, self.float_field_names = [], [], []
        self.token_seq_field_names, self.token_seq_field_dims = [], []
        for name in dataset.fields():                       # abstract_recommender.py:205-219
            if name == self.LABEL:
                continue
            ftype = getattr(dataset.field2type[name], "value", dataset.field2type[name])
            if ftype == FEATURE_TOKEN:
                self.token_field_names.append(name)
                self.token_field_dims.append(int(dataset.num(name)))
            elif ftype == FEATURE_FLOAT:
                if int(dataset.num(name)) != 1:
                    raise NotImplementedError("FusedFM: float field %r has %d columns (1 supported)" % (name, dataset.num(name)))
                self.float_field_names.append(name)
            elif ftype == FEATURE_TOKEN_SEQ:
                self.token_seq_field_names.append(name)
                self.token_seq_field_dims.append(int(dataset.num(name)))
            else:
                raise NotImplementedError("FusedFM handles TOKEN, FLOAT and TOKEN_SEQ fields; field %r is %r" % (name, ftype))
        if len(self.token_seq_field_names) > ops.FM_MAX_SEQ:
            raise NotImplementedError("FusedFM: at most %d TOKEN_SEQ fields" % ops.FM_MAX_SEQ)
        if not self.token_field_names:
            raise NotImplementedError("FusedFM needs at least one TOKEN field")
        self.n_float = len(self.float_field_names)
        self.n_seq = len(self.token_seq_field_names)
        self.num_feature_field = len(self.token_field_names)      # token fields (the kernels' F without sequences)
        # abstract_recommender.py:220-224: one table, per-field offsets
        self.token_field_offsets = np.array((0, *np.cumsum(self.token_field_dims)[:-1]), dtype=np.int64)
        rows = int(sum(self.token_field_dims))
        self._tok_rows = rows
        self._seq_bases = [rows + int(x) for x in np.cumsum([0] + self.token_seq_field_dims[:-1])] if self.n_seq else []
        self._all_rows = rows + int(sum(self.token_seq_field_dims))
        # registration order = ContextRecommender.__init__'s (the optimizer state is indexed by parameter position)
        self.token_embedding_table = _Table(rows, self.embedding_size)
        if self.n_float:
            self.float_embedding_table = nn.Embedding(self.n_float, self.embedding_size)     # :225-228
        if self.n_seq:                                                                       # :229-232
            self.token_seq_embedding_table = nn.ModuleList(nn.Embedding(n, self.embedding_size)
                                                           for n in self.token_seq_field_dims)
        self.first_order_linear = _FirstOrder(rows, self.n_float, self.token_seq_field_dims)
        # fm.py:41-45: xavier_normal_ on every nn.Embedding
        for m in self.modules():
            if isinstance(m, nn.Embedding):
                xavier_normal_(m.weight.data)
        self._pads = {}          # parameter name -> zero-padded [rows, dpad] storage the parameter is a view of
        self._seq_meta = {}      # (device, total id columns) -> offsets / column maps of a batch layout
        self._init_fused(config)

    # ---- plumbing -----------------------------------------------------------------------------------
    def _padded(self, key, param):
        """The [rows, dpad] tensor the kernels work on.  embedding_size == dpad: the parameter itself.  Otherwise the
        parameter's data is (re)made a view of the first embedding_size columns of a zero-padded buffer -- after
        construction and after every Module.to(), which replaces the data by a compact copy."""
        d, dp = self.embedding_size, self._dpad
        if d == dp:
            return param.data
        pad = self._pads.get(key)
        w = param.data
        if pad is None or pad.device != w.device or w.data_ptr() != pad.data_ptr() or w.stride(0) != dp:
            pad = torch.zeros((w.shape[0], dp), dtype=torch.float32, device=w.device)
            pad[:, :d] = w
            param.data = pad[:, :d]
            self._pads[key] = pad
        return pad

    def _joined(self, key, params, width):
        """One [all rows, width] tensor holding the token table followed by every TOKEN_SEQ table; the parameters are
        (re)made views of their row ranges (first embedding_size columns), like _padded does for a single table."""
        dev = params[0].device
        buf = self._pads.get(key)
        d = params[0].shape[1]
        bases = [0] + self._seq_bases
        ok = buf is not None and buf.device == dev
        if ok:
            for p, b in zip(params, bases):
                w = p.data
                if w.data_ptr() != buf.data_ptr() + 4 * b * width or w.stride(0) != width or w.device != dev:
                    ok = False
        if not ok:
            buf = torch.zeros((self._all_rows, width), dtype=torch.float32, device=dev)
            for p, b in zip(params, bases):
                buf[b:b + p.shape[0], :d] = p.data
                p.data = buf[b:b + p.shape[0], :d]
            self._pads[key] = buf
        return buf

    def _tables(self):
        if self.n_seq:
            fo = self.first_order_linear
            E = self._joined("E", [self.token_embedding_table.embedding.weight] +
                             [t.weight for t in self.token_seq_embedding_table], self._dpad)
            W = self._joined("W", [fo.token_embedding_table.embedding.weight] +
                             [t.weight for t in fo.token_seq_embedding_table], 1)
            return E, W.view(-1)
        E = self._padded("E", self.token_embedding_table.embedding.weight)
        W = self.first_order_linear.token_embedding_table.embedding.weight.data.view(-1)
        return E, W

    def _float_tables(self):
        Ef = self._padded("Ef", self.float_embedding_table.weight)
        Wf = self.first_order_linear.float_embedding_table.weight.data.view(-1)
        return Ef, Wf

    def _floats(self, interaction):
        """(values [B, Ff], Ef, Wf) for the kernels, or None (abstract_recommender.py:361-372)."""
        if not self.n_float:
            return None
        vals = torch.stack([interaction[n].reshape(-1).to(torch.float32) for n in self.float_field_names], dim=1)
        return (vals.contiguous(),) + self._float_tables()

    def _ids(self, interaction):
        # abstract_recommender.py:381-395: the per-field id columns -> [B, F] (+ the padded sequences' columns)
        cols = [interaction[n].reshape(-1, 1) for n in self.token_field_names]
        cols += [interaction[n].reshape(cols[0].shape[0], -1) for n in self.token_seq_field_names]
        return torch.cat(cols, dim=1).contiguous()

    def _make_layout(self, dev, lens):
        n_tok = len(self.token_field_names)
        starts = np.concatenate([[n_tok], n_tok + np.cumsum(lens)]).astype(np.int32)
        offs = list(self.token_field_offsets) + [b for b, n in zip(self._seq_bases, lens) for _ in range(n)]
        col_seq = [-1] * n_tok + [j for j, n in enumerate(lens) for _ in range(n)]
        return dict(n_token_cols=n_tok, seq_row_base=self._tok_rows,
                    offsets=torch.tensor(offs, dtype=torch.int64, device=dev),
                    seq_start=torch.from_numpy(starts).to(dev),
                    col_seq=torch.tensor(col_seq, dtype=torch.int32, device=dev))

    def _batch(self, interaction, train):
        """ids, offsets and the TOKEN_SEQ description of one batch (pooled / coef buffers when training)."""
        ids = self._ids(interaction)
        if not self.n_seq:
            self._ensure_device_state()
            return ids, self._offsets, None
        lens = [int(interaction[n].shape[1]) if interaction[n].dim() > 1 else 1 for n in self.token_seq_field_names]
        key = (str(ids.device),) + tuple(lens)
        meta = self._seq_meta.get(key)
        if meta is None:
            meta = self._seq_meta[key] = self._make_layout(ids.device, lens)
        seq = dict(meta)
        if train:
            B = ids.shape[0]
            buf = self._seq_meta.get("pooled")
            if buf is None or buf[0].shape[0] < B or buf[0].device != ids.device:
                buf = (torch.empty((B, self.n_seq, self._dpad), dtype=torch.float32, device=ids.device),
                       torch.empty((B, self.n_seq), dtype=torch.float32, device=ids.device))
                self._seq_meta["pooled"] = buf
            seq["pooled"], seq["coef"] = buf
        return ids, meta["offsets"], seq

    def _workspace(self, batch, n_cols=None):
        dev = self.token_embedding_table.embedding.weight.device
        F = int(n_cols or self.num_feature_field)
        if getattr(self, "_ws_dev", None) != (str(dev), F):
            self._ws, self._ws_dev = {}, (str(dev), F)
        return ops.grow_workspace(self._ws, batch, lambda b: ops.fm_workspace(b, F, self._dpad, dev))

    def _ensure_device_state(self):
        dev = self.token_embedding_table.embedding.weight.device
        if self._offsets is None or self._offsets.device != dev:
            self._offsets = torch.from_numpy(self.token_field_offsets).to(dev)
            self._bias3 = torch.zeros(3, dtype=torch.float32, device=dev)
            self._loss_out = torch.zeros(1, dtype=torch.float32, device=dev)
            self._loss_accum = torch.zeros(1, dtype=torch.float64, device=dev)
        self._bias3[0:1].copy_(self.first_order_linear.bias.data)

    def _make_optimizer(self, kind, learning_rate, weight_decay):
        if kind not in ("adam", "adam_lazy", "sgd"):
            raise ValueError("FusedFM implements the fused kinds {adam (row-sparse), adam_lazy, sgd}")
        self._optim = ops.Optim(kind, learning_rate, weight_decay)
        self._state = self._new_state(kind, *self._tables())
        if self.n_float and kind != "sgd":
            Ef, Wf = self._float_tables()
            z = torch.zeros_like
            self._state.update(mEf=z(Ef), vEf=z(Ef), mWf=z(Wf), vWf=z(Wf))
        self._ensure_device_state()

    def _opt_entries(self):
        # model.parameters() order of the reference's FM: token table, float table, first_order_linear.bias,
        # first-order token table, first-order float table (a module's own parameters precede its children's)
        st, d = self._state, self.embedding_size
        # (with TOKEN_SEQ fields: ... float table, sequence tables, bias, first-order token / float / sequence tables)
        tr = self._tok_rows
        seq = [(b, b + n) for b, n in zip(self._seq_bases, self.token_seq_field_dims)]
        out = [(st["mE"][:tr, :d], st["vE"][:tr, :d])]
        if self.n_float:
            out.append((st["mEf"][:, :d], st["vEf"][:, :d]))
        out += [(st["mE"][a:b, :d], st["vE"][a:b, :d]) for a, b in seq]
        out.append((self._bias3[1:2], self._bias3[2:3]))
        out.append((st["mW"][:tr], st["vW"][:tr]))
        if self.n_float:
            out.append((st["mWf"], st["vWf"]))
        out += [(st["mW"][a:b], st["vW"][a:b]) for a, b in seq]
        return out

    def _state_dict_impl(self, *args, **kwargs):
        sd = nn.Module.state_dict(self, *args, **kwargs)
        if self.embedding_size != self._dpad:          # compact copies: a checkpoint should not drag the padding along
            for k, v in list(sd.items()):
                if v.dim() == 2 and v.stride(0) == self._dpad and v.shape[1] == self.embedding_size:
                    sd[k] = v.contiguous()
        return sd

    # ---- fused step ----------------------------------------------------------------------------------
    def _fused_step(self, interaction, loss_accum):
        ids, offsets, seq = self._batch(interaction, True)
        E, W = self._tables()
        ops.fm_train_step(E, W, self._bias3, self._state, ids, offsets, interaction[self.LABEL].contiguous(),
                          self._optim, self._loss_out, loss_accum, self._workspace(ids.shape[0], ids.shape[1]),
                          floats=self._floats(interaction), seq=seq)
        self.first_order_linear.bias.data.copy_(self._bias3[0:1])
        return self._loss_out

    def _loss_value(self, interaction):
        self.flush()
        ids, offsets, seq = self._batch(interaction, False)
        E, W = self._tables()
        out = torch.empty(1, dtype=torch.float32, device=E.device)
        ops.fm_loss(E, W, self._bias3, ids, offsets, interaction[self.LABEL].contiguous(), out,
                    self._workspace(ids.shape[0], ids.shape[1]), floats=self._floats(interaction), seq=seq)
        return out[0]

    # ---- the reference's plugin API ---------------------------------------------------------------------
    def predict(self, interaction):  # fm.py:58-59
        self.flush()
        ids, offsets, seq = self._batch(interaction, False)
        E, W = self._tables()
        return ops.fm_predict(E, W, self._bias3, ids, offsets, self._workspace(ids.shape[0], ids.shape[1]),
                              floats=self._floats(interaction), seq=seq)


class FusedMFSimple(_PointwiseMixin, nn.Module):
    """The fork's point-wise "dot" model (recbole/model/general_recommender/mfsimple.py:8-62):
    ``sigmoid(<u, v> + b_u + b_i + b)`` with ``nn.BCELoss``.  It is exactly a two-field FM
    (field 0 = user id, field 1 = item id: 0.5[(u+v)^2 - u^2 - v^2] = <u, v>, first-order terms = the
    biases), so it runs on the fused FM kernels with the two tables stored back to back; the state
    dict keeps the reference's names (user_embedding.weight, item_embedding.weight, user_bias,
    item_bias, bias)."""
    input_type = InputType.POINTWISE   # mfsimple.py:10
    type = ModelType.GENERAL           # abstract_recommender.py:82

    def __init__(self, config, dataset):
        super().__init__()
        self.USER_ID, self.ITEM_ID = config["USER_ID_FIELD"], config["ITEM_ID_FIELD"]
        self.LABEL = config["LABEL_FIELD"]
        self.n_users, self.n_items = dataset.num(self.USER_ID), dataset.num(self.ITEM_ID)
        self.embedding_dim = config["embedding_dimension"] or 128      # MFSimple.yaml:1
        self.device = config["device"]
        rows = self.n_users + self.n_items
        self.table = nn.Parameter(torch.empty(rows, self.embedding_dim).normal_(0.0, 0.01))   # mfsimple.py:35-37
        self.biases = nn.Parameter(torch.zeros(rows))
        self.bias = nn.Parameter(torch.zeros(1))
        self._init_fused(config)

    def _tables(self):
        return self.table.data, self.biases.data

    # reference-compatible state dict -------------------------------------------------------------------
    def _state_dict_impl(self, *a, **k):
        nu = self.n_users
        return {"user_embedding.weight": self.table.data[:nu], "item_embedding.weight": self.table.data[nu:],
                "user_bias": self.biases.data[:nu], "item_bias": self.biases.data[nu:], "bias": self.bias.data}

    def load_state_dict(self, sd, strict=True):
        nu = self.n_users
        with torch.no_grad():
            self.table[:nu].copy_(sd["user_embedding.weight"])
            self.table[nu:].copy_(sd["item_embedding.weight"])
            self.biases[:nu].copy_(sd["user_bias"])
            self.biases[nu:].copy_(sd["item_bias"])
            self.bias.copy_(sd["bias"])

    def _prep(self):
        dev = self.table.device
        if self._offsets is None or self._offsets.device != dev:
            self._offsets = torch.tensor([0, self.n_users], dtype=torch.int64, device=dev)
            self._bias3 = torch.zeros(3, dtype=torch.float32, device=dev)
            self._loss_out = torch.zeros(1, dtype=torch.float32, device=dev)
            self._loss_accum = torch.zeros(1, dtype=torch.float64, device=dev)
        self._bias3[0:1].copy_(self.bias.data)

    def _workspace(self, batch):
        return ops.grow_workspace(self._ws, batch, lambda b: ops.fm_workspace(b, 2, self.embedding_dim, self.table.device))

    def _make_optimizer(self, kind, learning_rate, weight_decay):
        if kind not in ("adam", "adam_lazy", "sgd"):
            raise ValueError("FusedMFSimple implements the fused kinds {adam (row-sparse), adam_lazy, sgd}")
        self._optim = ops.Optim(kind, learning_rate, weight_decay)
        self._state = self._new_state(kind, *self._tables())
        self._prep()

    def _opt_entries(self):
        # the reference's MFSimple.parameters() order: user_bias, item_bias, bias, user_embedding, item_embedding
        # (a module's own parameters come before its children's, mfsimple.py:23-27)
        st, nu = self._state, self.n_users
        return [(st["mW"][:nu], st["vW"][:nu]), (st["mW"][nu:], st["vW"][nu:]), (self._bias3[1:2], self._bias3[2:3]),
                (st["mE"][:nu], st["vE"][:nu]), (st["mE"][nu:], st["vE"][nu:])]

    def _ids(self, interaction):
        return torch.stack([interaction[self.USER_ID], interaction[self.ITEM_ID]], dim=1).contiguous()

    def _fused_step(self, interaction, loss_accum):
        self._prep()
        ids = self._ids(interaction)
        ops.fm_train_step(self.table.data, self.biases.data, self._bias3, self._state, ids, self._offsets,
                          interaction[self.LABEL].contiguous(), self._optim, self._loss_out, loss_accum,
                          self._workspace(ids.shape[0]))
        self.bias.data.copy_(self._bias3[0:1])
        return self._loss_out

    def _loss_value(self, interaction):
        self.flush()
        self._prep()
        ids = self._ids(interaction)
        out = torch.empty(1, dtype=torch.float32, device=self.table.device)
        ops.fm_loss(self.table.data, self.biases.data, self._bias3, ids, self._offsets,
                    interaction[self.LABEL].contiguous(), out, self._workspace(ids.shape[0]))
        return out[0]

    def predict(self, interaction):  # mfsimple.py:59-62
        self.flush()
        self._prep()
        ids = self._ids(interaction)
        return ops.fm_predict(self.table.data, self.biases.data, self._bias3, ids, self._offsets,
                              self._workspace(ids.shape[0]))
