"""Host-side mirror of the reference's model interface for the BPR hot path.

``FusedBPR`` exposes exactly the plugin API of ``recbole.model.general_recommender.bpr.BPR``
(bpr.py:27-96) on top of ``GeneralRecommender`` (recbole/model/abstract_recommender.py:78-95):
constructor ``(config, dataset)``, attributes ``USER_ID / ITEM_ID / NEG_ITEM_ID / n_users / n_items /
device``, parameters named ``user_embedding.weight`` / ``item_embedding.weight`` (state-dict
compatible with the reference's checkpoints, trainer.py:191-206) and the methods
``calculate_loss / predict / full_sort_predict``.  All arithmetic runs in librecbole_b200.so;
there is no ATen fallback (a CPU tensor raises).

Three ways to train:
  * ``FusedTrainer`` (recbole_b200/trainer.py) calls ``model.train_step(interaction)``: one fused
    launch sequence per batch, loss accumulated on the device.
  * a reference ``Trainer`` whose ``_build_optimizer`` returns ``model.build_optimizer(...)``:
    ``calculate_loss`` returns a 0-dim tensor whose ``backward()`` only records the batch, and
    ``FusedOptimizer.step()`` runs the fused step for it (SURVEY.md 8b "Optimizer interface" (ii)).
  * a completely UNMODIFIED reference ``Trainer(config, model)``: it builds a ``torch.optim`` optimizer over
    ``model.parameters()`` (trainer.py:103,109-130) whose ``step()`` finds no gradients and does nothing; the model
    reads ``learner / learning_rate / weight_decay`` from the same config (trainer.py:80-81,93) and takes the fused
    step inside ``loss.backward()`` (trainer.py:170).  ``full_sort_predict`` returns the score matrix the
    Trainer masks and top-k's itself (trainer.py:328-352).
"""
import math

import torch
from torch import nn

from . import ops
from .enums import InputType, ModelType


def xavier_normal_(weight, generator=None):
    """recbole/model/init.py:15-31 -> torch.nn.init.xavier_normal_: std = sqrt(2 / (rows + d))."""
    rows, d = weight.shape
    std = math.sqrt(2.0 / float(rows + d))
    with torch.no_grad():
        return weight.normal_(0.0, std, generator=generator)


class _RecordBatch(torch.autograd.Function):
    """Forward: the loss kernel.  Backward: remember the batch for FusedOptimizer.step() -- or, when no fused
    optimizer was handed out (unmodified reference Trainer), take the fused step right here."""

    @staticmethod
    def forward(ctx, anchor, model, user, pos, neg):
        ctx.model, ctx.batch = model, (user, pos, neg)
        out = torch.empty(1, dtype=torch.float32, device=anchor.device)
        ops.bpr_loss(model.user_embedding.weight.data, model.item_embedding.weight.data, user, pos, neg, out,
                     model._workspace(user.numel()))
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        m = ctx.model
        m._pending = ctx.batch
        if m._autostep:
            m._apply_pending()
        return None, None, None, None, None


def fused_learner(config_learner, explicit=None):
    """Trainer._build_optimizer's `learner` (trainer.py:109-130) -> the fused optimizer kind.
    'adam' is the reference's DENSE torch.optim.Adam: served by 'adam_lazy', which reproduces its trajectory
    (rows without gradient keep moving on their momentum).  The row-sparse kernel (moments and parameters of
    untouched rows stay put -- torch.optim.SparseAdam's contract, the reference's learner 'sparse_adam') is a
    different algorithm and must be asked for by that name (or 'adam_sparse')."""
    name = (explicit or config_learner or "adam").lower()
    table = {"adam": "adam_lazy", "adam_lazy": "adam_lazy", "sparse_adam": "adam", "adam_sparse": "adam",
             "sgd": "sgd"}
    if name not in table:
        raise ValueError("the fused path implements learner in {adam, sparse_adam, sgd}; got %r" % name)
    return table[name]


class FusedBPR(nn.Module):
    input_type = InputType.PAIRWISE   # bpr.py:31
    type = ModelType.GENERAL          # abstract_recommender.py:82

    def __init__(self, config, dataset):
        super().__init__()
        # abstract_recommender.py:84-95
        self.USER_ID = config["USER_ID_FIELD"]
        self.ITEM_ID = config["ITEM_ID_FIELD"]
        self.NEG_ITEM_ID = config["NEG_PREFIX"] + self.ITEM_ID
        self.n_users = dataset.num(self.USER_ID)
        self.n_items = dataset.num(self.ITEM_ID)
        self.device = config["device"]
        # bpr.py:36-45.  Config(model=FusedBPR) finds no properties/model/FusedBPR.yaml (configurator.py:219-228):
        # BPR.yaml's value is the default
        self.embedding_size = config["embedding_size"] or 64
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
        # hyper-parameters for the unmodified-Trainer mode (the Trainer reads the same keys, trainer.py:80-81,93)
        self._hyper = (config["learner"], config["learning_rate"], config["weight_decay"])
        self._autostep = True   # cleared by build_optimizer(): the caller then drives FusedOptimizer.step()

    # ---- plumbing -----------------------------------------------------------------------------
    def _workspace(self, batch):
        dev = self.user_embedding.weight.device
        if self._ws_dev != str(dev):
            self._ws, self._ws_dev = {}, str(dev)
        return ops.grow_workspace(self._ws, batch, lambda b: ops.bpr_workspace(b, self.embedding_size, dev))

    def _make_optimizer(self, kind, learning_rate, weight_decay):
        self._optim = ops.Optim(kind, learning_rate, weight_decay)
        dev = self.user_embedding.weight.device
        U, V = self.user_embedding.weight.data, self.item_embedding.weight.data
        st = {}
        if kind != "sgd":
            st = dict(mU=torch.zeros_like(U), vU=torch.zeros_like(U), mV=torch.zeros_like(V), vV=torch.zeros_like(V))
        if kind == "adam_lazy":
            st["lastU"] = torch.zeros(U.shape[0], dtype=torch.int32, device=dev)
            st["lastV"] = torch.zeros(V.shape[0], dtype=torch.int32, device=dev)
        self._opt_state = st
        self._loss_out = torch.zeros(1, dtype=torch.float32, device=dev)
        self._loss_accum = torch.zeros(1, dtype=torch.float64, device=dev)

    def build_optimizer(self, learner="adam", learning_rate=1e-3, weight_decay=0.0):
        """Trainer._build_optimizer (trainer.py:109-130) for the fused path.  ``learner`` here names the fused
        kernel directly: 'adam' (row-sparse), 'adam_lazy' (row-sparse work, trajectory identical to the
        reference's dense Adam), 'sgd'.  (FusedTrainer maps the CONFIG's learner through fused_learner().)"""
        kind = learner.lower()
        if kind not in ("adam", "adam_lazy", "sgd"):
            raise ValueError("fused optimizer kinds: adam (row-sparse), adam_lazy (dense-Adam trajectory), sgd")
        self._make_optimizer(kind, learning_rate, weight_decay or 0.0)
        self._autostep = False
        return FusedOptimizer(self)

    def _ids(self, interaction):
        return interaction[self.USER_ID], interaction[self.ITEM_ID], interaction[self.NEG_ITEM_ID]

    def _opt_entries(self):
        """(exp_avg, exp_avg_sq) per parameter, in model.parameters() order (torch.optim.Adam's state layout)."""
        st = self._opt_state
        return [(st["mU"], st["vU"]), (st["mV"], st["vV"])]

    def _after_state_load(self):
        st = self._opt_state
        if "lastU" in st:
            st["lastU"].fill_(self._optim.step)
            st["lastV"].fill_(self._optim.step)

    def _apply_pending(self):
        if self._optim is None:     # unmodified Trainer: hyper-parameters from the config it reads itself
            learner, lr, wd = self._hyper
            self._make_optimizer(fused_learner(learner), lr if lr is not None else 1e-3, wd or 0.0)
        user, pos, neg = self._pending
        self._pending = None
        ops.bpr_train_step(self.user_embedding.weight.data, self.item_embedding.weight.data, self._opt_state, user, pos,
                           neg, self._optim, self._loss_out, None, self._workspace(user.numel()))

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

    def state_dict(self, *args, **kwargs):
        self.flush()            # a checkpoint holds the tables as dense Adam would have them (trainer.py:203)
        return super().state_dict(*args, **kwargs)

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
        """bpr.py:91-96: the flat [users * n_items] score vector.  COMPATIBILITY path for an unmodified reference
        Trainer, which masks it and runs torch.topk itself (trainer.py:328-352, evaluators.py:68-72); every score is
        the canonical fp32 chain of the fused scorer.  FusedTrainer.evaluate never builds this matrix: it uses
        full_sort_topk() (scores, history mask and top-K fused)."""
        self.flush()
        return ops.fullsort_scores(self.user_embedding.weight.data, interaction[self.USER_ID].contiguous(),
                                   self.item_embedding.weight.data).view(-1)

    def full_sort_topk(self, user_ids, k, hist_indptr=None, hist_indices=None, mode="fp32"):
        self.flush()
        return ops.fullsort_topk(self.user_embedding.weight.data, user_ids, self.item_embedding.weight.data, k,
                                 hist_indptr, hist_indices, mode=mode)


class FusedOptimizer:
    """What Trainer.__init__ stores as self.optimizer (trainer.py:103); supports the four calls the
    reference makes on it: zero_grad (trainer.py:160), step (:173), state_dict (:204),
    load_state_dict (:230).  state_dict() has the layout of torch.optim.Adam's (state index = position of the
    parameter in model.parameters()).  Works for every fused model that provides `_pending`, `_apply_pending()`,
    `_opt_entries()`, `_after_state_load()`, `flush()` and `_optim`."""

    def __init__(self, model):
        self.model = model

    @property
    def param_groups(self):
        o = self.model._optim
        return [dict(lr=o.lr, betas=o.betas, eps=o.eps, weight_decay=o.weight_decay)]

    def zero_grad(self, set_to_none=True):
        self.model._pending = None

    def step(self):
        m = self.model
        if m._pending is None:
            raise RuntimeError("FusedOptimizer.step(): no batch recorded (call loss.backward() first)")
        m._apply_pending()

    def state_dict(self):
        m = self.model
        m.flush()
        o = m._optim
        state = {}
        if o.kind_name != "sgd":
            for i, (ea, es) in enumerate(m._opt_entries()):
                state[i] = dict(step=torch.tensor(float(o.step)), exp_avg=ea, exp_avg_sq=es)
        n_params = len(state) or len(list(m.parameters()))
        group = dict(lr=o.lr, betas=o.betas, eps=o.eps, weight_decay=o.weight_decay, params=list(range(n_params)))
        return dict(state=state, param_groups=[group], fused_kind=o.kind_name)

    def load_state_dict(self, sd):
        m = self.model
        o = m._optim
        g = sd["param_groups"][0]
        o.set_hyper(lr=g["lr"], weight_decay=g["weight_decay"], betas=tuple(g.get("betas", o.betas)),
                    eps=g.get("eps", o.eps))
        if sd["state"]:
            o.step = int(sd["state"][0]["step"])
            for i, (ea, es) in enumerate(m._opt_entries()):
                ea.copy_(sd["state"][i]["exp_avg"].reshape(ea.shape))
                es.copy_(sd["state"][i]["exp_avg_sq"].reshape(es.shape))
            m._after_state_load()
