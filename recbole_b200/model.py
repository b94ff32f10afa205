"""Host-side mirror of the reference's model interface for the BPR hot path.

``FusedBPR`` exposes exactly the plugin API of ``recbole.model.general_recommender.bpr.BPR``
(bpr.py:27-96) on top of ``GeneralRecommender`` (recbole/model/abstract_recommender.py:78-95):
constructor ``(config, dataset)``, attributes ``USER_ID / ITEM_ID / NEG_ITEM_ID / n_users / n_items /
device``, parameters named ``user_embedding.weight`` / ``item_embedding.weight`` (state-dict
compatible with the reference's checkpoints, trainer.py:191-206) and the methods
``calculate_loss / predict / full_sort_predict``.  All arithmetic runs in librecbole_b200.so;
there is no ATen fallback (a CPU tensor raises).

Two ways to train:
  * ``FusedTrainer`` (recbole_b200/trainer.py) calls ``model.train_step(interaction)``: one fused
    launch sequence per batch, loss accumulated on the device.
  * an UNMODIFIED reference ``Trainer`` works too: ``calculate_loss`` returns a 0-dim tensor whose
    ``backward()`` only records the batch, and ``FusedOptimizer.step()`` (returned by
    ``build_optimizer``) runs the fused step for it (SURVEY.md 8b "Optimizer interface" (ii)).
"""
import math

import torch
from torch import nn

from . import ops


def xavier_normal_(weight, generator=None):
    """recbole/model/init.py:15-31 -> torch.nn.init.xavier_normal_: std = sqrt(2 / (rows + d))."""
    rows, d = weight.shape
    std = math.sqrt(2.0 / float(rows + d))
    with torch.no_grad():
        return weight.normal_(0.0, std, generator=generator)


class _RecordBatch(torch.autograd.Function):
    """Forward: the loss kernel.  Backward: remember the batch for FusedOptimizer.step()."""

    @staticmethod
    def forward(ctx, anchor, model, user, pos, neg):
        ctx.model, ctx.batch = model, (user, pos, neg)
        out = torch.empty(1, dtype=torch.float32, device=anchor.device)
        ops.bpr_loss(model.user_embedding.weight.data, model.item_embedding.weight.data, user, pos, neg, out,
                     model._workspace(user.numel()))
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        ctx.model._pending = ctx.batch
        return None, None, None, None, None


class FusedBPR(nn.Module):
    input_type = "pairwise"   # InputType.PAIRWISE, bpr.py:31
    type = "general"          # ModelType.GENERAL, abstract_recommender.py:82

    def __init__(self, config, dataset):
        super().__init__()
        # abstract_recommender.py:84-95
        self.USER_ID = config["USER_ID_FIELD"]
        self.ITEM_ID = config["ITEM_ID_FIELD"]
        self.NEG_ITEM_ID = config["NEG_PREFIX"] + self.ITEM_ID
        self.n_users = dataset.num(self.USER_ID)
        self.n_items = dataset.num(self.ITEM_ID)
        self.device = config["device"]
        # bpr.py:36-45
        self.embedding_size = config["embedding_size"]
        self.user_embedding = nn.Embedding(self.n_users, self.embedding_size)
        self.item_embedding = nn.Embedding(self.n_items, self.embedding_size)
        xavier_normal_(self.user_embedding.weight.data)
        xavier_normal_(self.item_embedding.weight.data)
        self._ws, self._ws_dev = {}, None
        self._pending = None
        self._optim = None      # ops.Optim
        self._opt_state = None  # dict of moment tensors
        self._loss_out = None
        self._loss_accum = None

    # ---- plumbing -----------------------------------------------------------------------------
    def _workspace(self, batch):
        dev = self.user_embedding.weight.device
        if self._ws_dev != str(dev):
            self._ws, self._ws_dev = {}, str(dev)
        return ops.grow_workspace(self._ws, batch, lambda b: ops.bpr_workspace(b, self.embedding_size, dev))

    def build_optimizer(self, learner="adam", learning_rate=1e-3, weight_decay=0.0):
        """Trainer._build_optimizer (trainer.py:109-130) for the fused path.  ``learner``:
        'adam' (row-sparse), 'adam_lazy' (row-sparse, trajectory identical to the reference's dense
        Adam), 'sgd'."""
        self._optim = ops.Optim(learner.lower(), learning_rate, weight_decay)
        dev = self.user_embedding.weight.device
        U, V = self.user_embedding.weight.data, self.item_embedding.weight.data
        st = {}
        if learner.lower() != "sgd":
            st = dict(mU=torch.zeros_like(U), vU=torch.zeros_like(U), mV=torch.zeros_like(V), vV=torch.zeros_like(V))
        if learner.lower() == "adam_lazy":
            st["lastU"] = torch.zeros(U.shape[0], dtype=torch.int32, device=dev)
            st["lastV"] = torch.zeros(V.shape[0], dtype=torch.int32, device=dev)
        self._opt_state = st
        self._loss_out = torch.zeros(1, dtype=torch.float32, device=dev)
        self._loss_accum = torch.zeros(1, dtype=torch.float64, device=dev)
        return FusedOptimizer(self)

    def _ids(self, interaction):
        return interaction[self.USER_ID], interaction[self.ITEM_ID], interaction[self.NEG_ITEM_ID]

    # ---- fused training step --------------------------------------------------------------------
    def train_step(self, interaction):
        """calculate_loss + backward + optimizer.step in one fused launch sequence.  Returns the
        device scalar holding this batch's loss (no host sync)."""
        if self._optim is None:
            raise RuntimeError("call build_optimizer() first")
        user, pos, neg = self._ids(interaction)
        ops.bpr_train_step(self.user_embedding.weight.data, self.item_embedding.weight.data, self._opt_state,
                           user, pos, neg, self._optim, self._loss_out, self._loss_accum,
                           self._workspace(user.numel()))
        return self._loss_out

    def flush(self):
        """adam_lazy only: bring every row up to the current step before the tables are read."""
        if self._optim is not None and self._optim.kind_name == "adam_lazy" and self._optim.step > 0:
            st = self._opt_state
            ops.adam_lazy_flush(self.user_embedding.weight.data, st["mU"], st["vU"], st["lastU"], self._optim)
            ops.adam_lazy_flush(self.item_embedding.weight.data, st["mV"], st["vV"], st["lastV"], self._optim)

    # ---- the reference's plugin API -------------------------------------------------------------
    def calculate_loss(self, interaction):  # bpr.py:74-83
        # adam_lazy: the forward-only kernel reads the raw tables, so bring every row up to date first
        # (O(table), what the reference's dense Adam costs anyway; train_step() does not need this)
        self.flush()
        user, pos, neg = self._ids(interaction)
        return _RecordBatch.apply(self.user_embedding.weight, self, user, pos, neg)

    def predict(self, interaction):  # bpr.py:85-89
        self.flush()
        return ops.gather_dot(self.user_embedding.weight.data, self.item_embedding.weight.data,
                              interaction[self.USER_ID], interaction[self.ITEM_ID])

    def full_sort_predict(self, interaction):
        """bpr.py:91-96 materialises a [users, n_items] matrix; the fused path never does.  Raising
        NotImplementedError is the reference's own protocol for "no full-sort matrix" and makes an
        unmodified Trainer fall back to predict() (trainer.py:333-340); FusedTrainer.evaluate uses
        full_sort_topk() below instead."""
        raise NotImplementedError("FusedBPR does not materialise score matrices; use full_sort_topk()")

    def full_sort_topk(self, user_ids, k, hist_indptr=None, hist_indices=None, mode="fp32"):
        self.flush()
        return ops.fullsort_topk(self.user_embedding.weight.data, user_ids, self.item_embedding.weight.data, k,
                                 hist_indptr, hist_indices, mode=mode)


class FusedOptimizer:
    """What Trainer.__init__ stores as self.optimizer (trainer.py:103); supports the four calls the
    reference makes on it: zero_grad (trainer.py:160), step (:173), state_dict (:204),
    load_state_dict (:230).  state_dict() has the layout of torch.optim.Adam's."""

    def __init__(self, model):
        self.model = model

    def zero_grad(self, set_to_none=True):
        self.model._pending = None

    def step(self):
        m = self.model
        if m._pending is None:
            raise RuntimeError("FusedOptimizer.step(): no batch recorded (call loss.backward() first)")
        user, pos, neg = m._pending
        m._pending = None
        ops.bpr_train_step(m.user_embedding.weight.data, m.item_embedding.weight.data, m._opt_state, user, pos, neg,
                           m._optim, m._loss_out, None, m._workspace(user.numel()))

    def state_dict(self):
        m = self.model
        m.flush()
        o, st = m._optim, m._opt_state
        state = {}
        if o.kind_name != "sgd":
            step = torch.tensor(float(o.step))
            state = {0: dict(step=step, exp_avg=st["mU"], exp_avg_sq=st["vU"]),
                     1: dict(step=step.clone(), exp_avg=st["mV"], exp_avg_sq=st["vV"])}
        group = dict(lr=o.lr, betas=o.betas, eps=o.eps, weight_decay=o.weight_decay, params=[0, 1])
        return dict(state=state, param_groups=[group], fused_kind=o.kind_name)

    def load_state_dict(self, sd):
        m = self.model
        o, st = m._optim, m._opt_state
        g = sd["param_groups"][0]
        o.lr, o.weight_decay = g["lr"], g["weight_decay"]
        if sd["state"]:
            o.step = int(sd["state"][0]["step"])
            st["mU"].copy_(sd["state"][0]["exp_avg"])
            st["vU"].copy_(sd["state"][0]["exp_avg_sq"])
            st["mV"].copy_(sd["state"][1]["exp_avg"])
            st["vV"].copy_(sd["state"][1]["exp_avg_sq"])
            if "lastU" in st:
                st["lastU"].fill_(o.step)
                st["lastV"].fill_(o.step)
