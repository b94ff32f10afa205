"""Host-side mirror of the reference's FM context-aware recommender for the fused path.

``FusedFM`` exposes the plugin API of ``recbole.model.context_aware_recommender.fm.FM``
(fm.py:26-59) on top of ``ContextRecommender`` (recbole/model/abstract_recommender.py:151-412):
constructor ``(config, dataset)``, ``calculate_loss / predict``, and parameters under the reference's
names (``token_embedding_table.embedding.weight``, ``first_order_linear.token_embedding_table.embedding.weight``,
``first_order_linear.bias``) so state dicts interchange.  TOKEN fields only (BASELINE config 5 names
categorical fields; float / token_seq fields stay on the reference path and raise here).
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
    def __init__(self, rows):
        super().__init__()
        self.token_embedding_table = _Table(rows, 1)
        self.bias = nn.Parameter(torch.zeros((1,)), requires_grad=True)  # layers.py:945


class FusedFM(_PointwiseMixin, nn.Module):
    input_type = InputType.POINTWISE   # abstract_recommender.py:157
    type = ModelType.CONTEXT           # abstract_recommender.py:156

    def __init__(self, config, dataset):
        super().__init__()
        self.LABEL = config["LABEL_FIELD"]
        self.embedding_size = config["embedding_size"]
        self.device = config["device"]
        self.token_field_names, self.token_field_dims = [], []
        for name in dataset.fields():                       # abstract_recommender.py:205-219
            if name == self.LABEL:
                continue
            if not _is_token(dataset.field2type[name]):
                raise NotImplementedError("FusedFM handles TOKEN fields only; field %r is %r"
                                          % (name, dataset.field2type[name]))
            self.token_field_names.append(name)
            self.token_field_dims.append(int(dataset.num(name)))
        self.num_feature_field = len(self.token_field_names)
        # abstract_recommender.py:220-224: one table, per-field offsets
        self.token_field_offsets = np.array((0, *np.cumsum(self.token_field_dims)[:-1]), dtype=np.int64)
        rows = int(sum(self.token_field_dims))
        self.token_embedding_table = _Table(rows, self.embedding_size)
        self.first_order_linear = _FirstOrder(rows)
        # fm.py:41-45: xavier_normal_ on every nn.Embedding
        xavier_normal_(self.token_embedding_table.embedding.weight.data)
        xavier_normal_(self.first_order_linear.token_embedding_table.embedding.weight.data)
        self._init_fused(config)

    # ---- plumbing -----------------------------------------------------------------------------------
    def _tables(self):
        E = self.token_embedding_table.embedding.weight.data
        W = self.first_order_linear.token_embedding_table.embedding.weight.data.view(-1)
        return E, W

    def _ids(self, interaction):
        # abstract_recommender.py:381-388: stack the per-field id columns -> [B, F]
        return torch.stack([interaction[n] for n in self.token_field_names], dim=1).contiguous()

    def _workspace(self, batch):
        dev = self.token_embedding_table.embedding.weight.device
        if getattr(self, "_ws_dev", None) != str(dev):
            self._ws, self._ws_dev = {}, str(dev)
        return ops.grow_workspace(self._ws, batch,
                                  lambda b: ops.fm_workspace(b, self.num_feature_field, self.embedding_size, dev))

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
        self._ensure_device_state()

    def _opt_entries(self):
        # model.parameters() order: E, first_order_linear.bias, first_order_linear...weight (same as the reference's FM)
        st = self._state
        return [(st["mE"], st["vE"]), (self._bias3[1:2], self._bias3[2:3]), (st["mW"], st["vW"])]

    def _state_dict_impl(self, *args, **kwargs):
        return nn.Module.state_dict(self, *args, **kwargs)

    # ---- fused step ----------------------------------------------------------------------------------
    def _fused_step(self, interaction, loss_accum):
        self._ensure_device_state()
        ids = self._ids(interaction)
        E, W = self._tables()
        ops.fm_train_step(E, W, self._bias3, self._state, ids, self._offsets, interaction[self.LABEL].contiguous(),
                          self._optim, self._loss_out, loss_accum, self._workspace(ids.shape[0]))
        self.first_order_linear.bias.data.copy_(self._bias3[0:1])
        return self._loss_out

    def _loss_value(self, interaction):
        self.flush()
        self._ensure_device_state()
        ids = self._ids(interaction)
        E, W = self._tables()
        out = torch.empty(1, dtype=torch.float32, device=E.device)
        ops.fm_loss(E, W, self._bias3, ids, self._offsets, interaction[self.LABEL].contiguous(), out,
                    self._workspace(ids.shape[0]))
        return out[0]

    # ---- the reference's plugin API ---------------------------------------------------------------------
    def predict(self, interaction):  # fm.py:58-59
        self.flush()
        self._ensure_device_state()
        ids = self._ids(interaction)
        E, W = self._tables()
        return ops.fm_predict(E, W, self._bias3, ids, self._offsets, self._workspace(ids.shape[0]))


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
