"""``FusedTrainer``: mirror of the reference ``Trainer`` (recbole/trainer/trainer.py:77-428) for
the fused BPR path.  Same constructor ``(config, model)``, same ``fit`` / ``evaluate`` /
``_train_epoch`` entry points and return values; what changes is below the line:

* ``_train_epoch`` (trainer.py:132-174) issues one fused step per batch and reads the loss back
  ONCE per epoch (the reference calls ``.item()`` per batch, trainer.py:168); the NaN check
  (trainer.py:234-236) runs on the epoch total.
* ``evaluate`` (trainer.py:354-412) for full-sort data never builds ``item_tensor``
  (trainer.py:384-387) nor a score matrix: users are scored in large tiles with the history mask
  and top-K fused, and the metrics are reduced on the device.
* ``learner`` (trainer.py:109-130): the config's 'adam' means the reference's DENSE ``torch.optim.Adam`` and is
  served by the fused kind 'adam_lazy' (same trajectory); 'sparse_adam' selects the row-sparse kernel
  (``torch.optim.SparseAdam``'s contract: untouched rows keep parameters and moments); 'sgd' is the same thing
  dense or sparse when ``weight_decay`` is 0.
* checkpoints (trainer.py:191-232,372-380): same file layout (``config, epoch, cur_step, best_valid_score,
  state_dict, optimizer``), parameter names and ``torch.optim.Adam`` state layout as the reference's, so a file
  written by either trainer resumes in the other.
"""
import os
from time import strftime, time

import numpy as np
import torch

from .data import EvalIndex
from .evaluator import FusedTopKEvaluator
from .interaction import Interaction  # noqa: F401
from .model import fused_learner


class FusedTrainer:
    def __init__(self, config, model):
        self.config = config
        self.model = model
        self.learner = config["learner"] or "adam"
        self.learning_rate = config["learning_rate"] if config["learning_rate"] is not None else 1e-3
        self.epochs = config["epochs"] or 1
        self.eval_step = min(config["eval_step"] or 1, self.epochs)
        self.stopping_step = config["stopping_step"] or 10
        self.valid_metric = (config["valid_metric"] or "MRR@10").lower()
        self.valid_metric_bigger = config["valid_metric_bigger"] if config["valid_metric_bigger"] is not None else True
        self.device = config["device"]
        self.weight_decay = config["weight_decay"] or 0.0
        self.scorer_mode = config["scorer_mode"] or "fp32"
        self.eval_user_tile = config["eval_user_tile"] or 1 << 16
        self.clip_grad_norm = config["clip_grad_norm"]
        if self.clip_grad_norm:
            # trainer.py:171-172 clips the norm of the DENSE gradient of all parameters; the fused step never forms it
            raise NotImplementedError("clip_grad_norm is not available on the fused path (no dense gradient exists)")
        self.checkpoint_dir = config["checkpoint_dir"] or "saved"
        self.saved_model_file = os.path.join(self.checkpoint_dir, "{}-{}.pth".format(
            config["model"] or type(model).__name__, strftime("%b-%d-%Y_%H-%M-%S")))          # trainer.py:91-92
        self.start_epoch = 0
        self.cur_step = 0
        self.best_valid_score = -np.inf if self.valid_metric_bigger else np.inf
        self.best_valid_result = None
        self.train_loss_dict = dict()
        self.optimizer = self._build_optimizer(self.model.parameters())
        self.evaluator = FusedTopKEvaluator(config)
        self._index_cache = {}

    def _build_optimizer(self, params):  # trainer.py:109-130
        kind = fused_learner(self.learner, self.config["fused_learner"])
        return self.model.build_optimizer(kind, self.learning_rate, self.weight_decay)

    # ---- checkpoints (trainer.py:191-232) ---------------------------------------------------------------------
    def _save_checkpoint(self, epoch):
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        cfg = self.config
        state = {
            "config": cfg if not isinstance(cfg, dict) or type(cfg) is dict else dict(cfg),
            "epoch": epoch,
            "cur_step": self.cur_step,
            "best_valid_score": self.best_valid_score,
            "state_dict": self.model.state_dict(),
            "optimizer": self.optimizer.state_dict(),
        }
        torch.save(state, self.saved_model_file)

    def resume_checkpoint(self, resume_file):
        checkpoint = torch.load(str(resume_file), weights_only=False, map_location=self.device)
        self.start_epoch = checkpoint["epoch"] + 1
        self.cur_step = checkpoint["cur_step"]
        self.best_valid_score = checkpoint["best_valid_score"]
        self.model.load_state_dict(checkpoint["state_dict"])
        self.optimizer.load_state_dict(checkpoint["optimizer"])

    def _check_nan(self, loss):  # trainer.py:234-236
        if np.isnan(loss):
            raise ValueError("Training loss is nan")

    def _train_epoch(self, train_data, epoch_idx, loss_func=None, show_progress=False):
        self.model.train()
        self.model._loss_accum.zero_()
        ws = None
        for interaction in train_data:
            interaction = interaction.to(self.device)
            self.model.train_step(interaction)
        total = float(self.model._loss_accum.item())  # the one host sync of the epoch
        for ws in self.model._ws.values():
            ws.check_flags()
        self._check_nan(total)
        return total

    def _valid_epoch(self, valid_data, show_progress=False):  # trainer.py:176-189
        result = self.evaluate(valid_data, load_best_model=False)
        return result[self.valid_metric], result

    def fit(self, train_data, valid_data=None, verbose=True, saved=True, show_progress=False, callback_fn=None):
        """trainer.py:250-326, same control flow and return value."""
        if saved and self.start_epoch >= self.epochs:
            self._save_checkpoint(-1)
        for epoch_idx in range(self.start_epoch, self.epochs):
            t0 = time()
            train_loss = self._train_epoch(train_data, epoch_idx)
            self.train_loss_dict[epoch_idx] = train_loss
            if verbose:
                print("epoch %d training [time: %.2fs, train loss: %.4f]" % (epoch_idx, time() - t0, train_loss))
            if self.eval_step <= 0 or not valid_data:
                if saved:
                    self._save_checkpoint(epoch_idx)
                continue
            if (epoch_idx + 1) % self.eval_step == 0:
                score, result = self._valid_epoch(valid_data)
                # utils.py:99-140 early_stopping
                better = score > self.best_valid_score if self.valid_metric_bigger else score < self.best_valid_score
                if better:
                    self.best_valid_score, self.cur_step = score, 0
                    if saved:
                        self._save_checkpoint(epoch_idx)
                    self.best_valid_result = result
                else:
                    self.cur_step += 1
                if callback_fn:
                    callback_fn(epoch_idx, score)
                if self.cur_step > self.stopping_step:
                    break
        return self.best_valid_score, self.best_valid_result

    def _eval_index(self, eval_data):
        if isinstance(eval_data, EvalIndex):
            return eval_data
        key = id(eval_data)
        if key not in self._index_cache:
            self._index_cache[key] = EvalIndex.from_reference_dataloader(eval_data, self.device)
        return self._index_cache[key]

    @torch.no_grad()
    def full_sort_topk(self, index):
        """Top-max(topk) item ids for every evaluated user, in user tiles."""
        K = self.evaluator.max_k
        n = index.n_eval_users
        ids = torch.empty((n, K), dtype=torch.int64, device=index.uid_list.device)
        for lo in range(0, n, self.eval_user_tile):
            hi = min(lo + self.eval_user_tile, n)
            ptr = index.hist_indptr[lo:hi + 1].contiguous()
            t_ids, _ = self.model.full_sort_topk(index.uid_list[lo:hi].contiguous(), K, ptr, index.hist_indices,
                                                 mode=self.scorer_mode)
            ids[lo:hi] = t_ids
        return ids

    @torch.no_grad()
    def evaluate(self, eval_data, load_best_model=False, model_file=None, show_progress=False):
        if eval_data is None:
            return
        if load_best_model:  # trainer.py:372-380
            checkpoint = torch.load(model_file or self.saved_model_file, weights_only=False, map_location=self.device)
            self.model.load_state_dict(checkpoint["state_dict"])
        self.model.eval()
        index = self._eval_index(eval_data)
        ids = self.full_sort_topk(index)
        return self.evaluator.evaluate(ids, index)
