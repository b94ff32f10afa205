"""The reference's CPU path restated with the same PyTorch calls it makes (TEST INFRASTRUCTURE).

The reference is pure Python over ATen: its "CPU implementation" of the hot path IS this
sequence of torch calls.  This module restates it, op for op, so that it can be timed on the
GPU box's host cores with every thread torch can use (bench.py: ``cpu_baseline`` and
``--impl reference``); /root/reference itself cannot travel to the box.  It is checked against
the reference's own outputs in tests/test_oracle_golden.py.

  RefBPR                 recbole/model/general_recommender/bpr.py:33-96, loss.py:43-49, init.py:15-31
  train_epoch            recbole/trainer/trainer.py:109-130 (optimizer), :157-173 (loop body)
  full_sort_eval         general_dataloader.py:330-364 (batching: step = max(4096 // n_items, 1) users),
                         trainer.py:328-352 (mask + swap), evaluators.py:53-105,122-141 (flip, topk, metrics)
  sampler_epoch          sampler.py:103-154 general branch (the Python list comprehension with set lookups)
"""
import numpy as np
import torch
from torch import nn

from . import fullsort as _fs
from . import metrics as _metrics


class RefBPR(nn.Module):
    def __init__(self, n_users, n_items, dim):
        super().__init__()
        self.user_embedding = nn.Embedding(n_users, dim)
        self.item_embedding = nn.Embedding(n_items, dim)
        for m in (self.user_embedding, self.item_embedding):
            nn.init.xavier_normal_(m.weight.data)  # init.py:27
        self.gamma = 1e-10

    def calculate_loss(self, user, pos_item, neg_item):  # bpr.py:74-83
        user_e = self.user_embedding(user)
        pos_e = self.item_embedding(pos_item)
        neg_e = self.item_embedding(neg_item)
        pos_s, neg_s = torch.mul(user_e, pos_e).sum(dim=1), torch.mul(user_e, neg_e).sum(dim=1)
        return -torch.log(self.gamma + torch.sigmoid(pos_s - neg_s)).mean()  # loss.py:48

    def predict(self, user, item):  # bpr.py:85-89
        return torch.mul(self.user_embedding(user), self.item_embedding(item)).sum(dim=1)

    def full_sort_predict(self, user):  # bpr.py:91-96
        return torch.matmul(self.user_embedding(user), self.item_embedding.weight.transpose(0, 1)).view(-1)


def build_optimizer(model, learner="adam", lr=1e-3, weight_decay=0.0):  # trainer.py:109-130
    if learner == "sgd":
        return torch.optim.SGD(model.parameters(), lr=lr, weight_decay=weight_decay)
    return torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)


def train_steps(model, optimizer, batches):
    """trainer.py:157-173 for pre-built (user, pos, neg) int64 CPU batches.  Returns the summed loss."""
    model.train()
    total = 0.0
    for user, pos, neg in batches:
        optimizer.zero_grad()
        loss = model.calculate_loss(user, pos, neg)
        total += loss.item()          # trainer.py:168, the per-step host sync
        if torch.isnan(loss):
            raise ValueError("Training loss is nan")
        loss.backward()
        optimizer.step()
    return total


@torch.no_grad()
def full_sort_eval(model, uid_list, hist, pos, n_items, topk=(10,), metrics=("recall", "mrr", "ndcg", "hit",
                                                                             "precision"),
                   eval_batch_size=4096):
    """Reference evaluation loop.  hist / pos are (indptr, indices) numpy CSRs over uid_list rows.
    Returns (result dict, int64 [users, K+1] topk matrix as TopKEvaluator.collect builds it)."""
    model.eval()
    step = max(eval_batch_size // n_items, 1)  # general_dataloader.py:330-334
    K = max(topk)
    mats = []
    for lo in range(0, len(uid_list), step):
        rows = range(lo, min(lo + step, len(uid_list)))
        users = torch.as_tensor(uid_list[lo:lo + step], dtype=torch.int64)
        scores = model.full_sort_predict(users).view(-1, n_items)
        scores[:, 0] = -np.inf                                   # trainer.py:343
        h_row, h_col, s_row, s_after, s_before = [], [], [], [], []
        for i, r in enumerate(rows):                              # general_dataloader.py:349-364
            h = hist[1][hist[0][r]:hist[0][r + 1]]
            h_row.append(np.full(len(h), i, dtype=np.int64))
            h_col.append(h)
            a, b = _fs.reference_swap_index(pos[1][pos[0][r]:pos[0][r + 1]])
            s_row.append(np.full(len(a), i, dtype=np.int64))
            s_after.append(a)
            s_before.append(b)
        h_row, h_col = torch.from_numpy(np.concatenate(h_row)), torch.from_numpy(np.concatenate(h_col))
        scores[(h_row, h_col)] = -np.inf                          # trainer.py:344-345
        s_row = torch.from_numpy(np.concatenate(s_row))
        s_after, s_before = torch.from_numpy(np.concatenate(s_after)), torch.from_numpy(np.concatenate(s_before))
        scores[s_row, s_after] = scores[s_row, s_before]          # trainer.py:347-350
        flipped = torch.flip(scores, dims=[-1])                   # evaluators.py:68
        _, idx = torch.topk(flipped, K, dim=-1)                   # evaluators.py:72
        shape = torch.full((len(users), 1), n_items)
        mats.append(torch.cat((idx, shape), dim=1))               # evaluators.py:75
    mat = torch.cat(mats, dim=0).numpy()                          # evaluators.py:90
    pos_len = np.diff(pos[0])
    pos_idx = mat[:, :-1] >= (mat[:, -1] - pos_len).reshape(-1, 1)  # evaluators.py:134
    return _metrics.evaluate(pos_idx, pos_len, list(metrics), list(topk)), mat


def sampler_draw(random_list, random_pr, used_sets, user_ids, num):
    """sampler.py:103-154, general branch, verbatim semantics incl. the Python-level loop."""
    L = len(random_list)

    def random_num(n, pr):
        out = []
        pr %= L
        while True:
            if pr + n <= L:
                out.append(random_list[pr:pr + n])
                pr += n
                break
            out.append(random_list[pr:])
            n -= L - pr
            pr = 0
        return np.concatenate(out), pr

    key_ids = np.tile(np.asarray(user_ids), num)
    total = len(key_ids)
    value_ids = np.zeros(total, dtype=np.int64)
    check = np.arange(total)
    pr = random_pr
    while len(check) > 0:
        value_ids[check], pr = random_num(len(check), pr)
        check = np.array([i for i, used, v in zip(check, used_sets[key_ids[check]], value_ids[check]) if v in used],
                         dtype=np.int64)
    return value_ids, pr


class RefFM(nn.Module):
    """FM on TOKEN fields restated with the reference's own torch calls:
    FMEmbedding (layers.py:121-144), BaseFactorizationMachine (layers.py:147-171), FMFirstOrderLinear token
    part + bias (layers.py:1021-1061), FM.forward / calculate_loss (fm.py:47-56)."""

    def __init__(self, field_dims, dim):
        super().__init__()
        rows = int(sum(field_dims))
        self.offsets = torch.as_tensor(np.array((0, *np.cumsum(field_dims)[:-1]), dtype=np.int64))
        self.embedding = nn.Embedding(rows, dim)
        self.first_order = nn.Embedding(rows, 1)
        self.bias = nn.Parameter(torch.zeros((1,)))
        nn.init.xavier_normal_(self.embedding.weight.data)
        nn.init.xavier_normal_(self.first_order.weight.data)
        self.loss = nn.BCELoss()

    def forward(self, ids):  # ids int64 [B, F]
        x = ids + self.offsets.unsqueeze(0)                       # layers.py:142
        e = self.embedding(x)                                     # [B, F, d]
        square_of_sum = torch.sum(e, dim=1) ** 2                  # layers.py:165-166
        sum_of_square = torch.sum(e ** 2, dim=1)
        fm = 0.5 * torch.sum(square_of_sum - sum_of_square, dim=1, keepdim=True)
        first = torch.sum(self.first_order(x), dim=1) + self.bias  # layers.py:1021-1061
        return torch.sigmoid(first + fm).squeeze()                # fm.py:49-50

    def calculate_loss(self, ids, label):
        return self.loss(self.forward(ids), label)                # fm.py:52-56


def fm_train_steps(model, optimizer, batches):
    model.train()
    total = 0.0
    for ids, label in batches:
        optimizer.zero_grad()
        loss = model.calculate_loss(ids, label)
        total += loss.item()
        loss.backward()
        optimizer.step()
    return total


@torch.no_grad()
def ce_head(X, E, target, k=10):
    """sasrec.py:137-141 (CE branch) + 152-158 / trainer.py:343 / evaluators.py:72 on one chunk."""
    logits = torch.matmul(X, E.transpose(0, 1))
    loss = nn.functional.cross_entropy(logits, target)
    scores = logits.clone()
    scores[:, 0] = -np.inf
    _, idx = torch.topk(scores, k, dim=-1)
    return loss, idx
