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
from .model import xavier_normal_

FEATURE_TOKEN = "token"


def _is_token(ftype):
    return getattr(ftype, "value", ftype) == FEATURE_TOKEN


class _Table(nn.Module):
    def __init__(self, rows, dim):
        super().__init__()
        self.embedding = nn.Embedding(rows, dim)


class _FirstOrder(nn.Module):
    def __init__(self, rows):
        super().__init__()
        self.token_embedding_table = _Table(rows, 1)
        self.bias = nn.Parameter(torch.zeros((1,)), requires_grad=True)  # layers.py:945


class FusedFM(nn.Module):
    input_type = "pointwise"   # abstract_recommender.py:157
    type = "context"           # ModelType.CONTEXT, abstract_recommender.py:156

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
        self._optim, self._state, self._ws, self._bias3 = None, None, {}, None
        self._loss_out = self._loss_accum = self._offsets = None

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
        key = (int(batch), str(dev))
        if key not in self._ws:
            self._ws = {key: ops.fm_workspace(batch, self.num_feature_field, self.embedding_size, dev)}
        return self._ws[key]

    def _ensure_device_state(self):
        dev = self.token_embedding_table.embedding.weight.device
        if self._offsets is None or self._offsets.device != dev:
            self._offsets = torch.from_numpy(self.token_field_offsets).to(dev)
            self._bias3 = torch.zeros(3, dtype=torch.float32, device=dev)
            self._loss_out = torch.zeros(1, dtype=torch.float32, device=dev)
            self._loss_accum = torch.zeros(1, dtype=torch.float64, device=dev)
        self._bias3[0:1].copy_(self.first_order_linear.bias.data)

    def build_optimizer(self, learner="adam", learning_rate=1e-3, weight_decay=0.0):
        if learner.lower() not in ("adam", "sgd"):
            raise ValueError("FusedFM implements learner in {adam, sgd}")
        self._optim = ops.Optim(learner.lower(), learning_rate, weight_decay)
        E, W = self._tables()
        self._state = {}
        if learner.lower() == "adam":
            self._state = dict(mE=torch.zeros_like(E), vE=torch.zeros_like(E), mW=torch.zeros_like(W),
                               vW=torch.zeros_like(W))
        self._ensure_device_state()
        return self

    # ---- fused step ----------------------------------------------------------------------------------
    def train_step(self, interaction):
        if self._optim is None:
            raise RuntimeError("call build_optimizer() first")
        self._ensure_device_state()
        ids = self._ids(interaction)
        E, W = self._tables()
        ops.fm_train_step(E, W, self._bias3, self._state, ids, self._offsets, interaction[self.LABEL].contiguous(),
                          self._optim, self._loss_out, self._loss_accum, self._workspace(ids.shape[0]))
        self.first_order_linear.bias.data.copy_(self._bias3[0:1])
        return self._loss_out

    # ---- the reference's plugin API ---------------------------------------------------------------------
    def predict(self, interaction):  # fm.py:58-59
        self._ensure_device_state()
        ids = self._ids(interaction)
        E, W = self._tables()
        return ops.fm_predict(E, W, self._bias3, ids, self._offsets, self._workspace(ids.shape[0]))

    def calculate_loss(self, interaction):  # fm.py:52-56 (forward only; training goes through train_step)
        y = self.predict(interaction)
        return nn.functional.binary_cross_entropy(y, interaction[self.LABEL])
