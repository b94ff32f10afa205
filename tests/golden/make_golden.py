"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

The import shims under tests/golden/_shims are the ones listed in SURVEY.md section 8c
(stub colorlog / colorama / gensim / torch_sparse, numpy.float alias); no reference file is
modified or copied.  Every array saved here is an INPUT or an OUTPUT of reference code:

  metrics.npz    recbole.evaluator.metrics.* on the known-answer case of
                 tests/metrics/test_topk_metrics.py:15-79 and on a random case
  bpr_steps.npz  recbole BPR + BPRLoss + torch.optim.Adam / SGD, 3 steps, duplicate ids in batch
  bpr_gaps.npz   recbole BPR under dense Adam (with and without weight decay), 16 steps of small batches: most rows are
                 untouched in most steps
  dot_steps.npz  the fork's MFSimple forward / BCELoss gradients (mfsimple.py:39-57)
  fullsort_small.npz / fullsort_ml100k.npz
                 Config -> create_dataset -> data_preparation -> GeneralFullDataLoader ->
                 BPR.full_sort_predict -> Trainer._full_sort_batch_eval -> TopKEvaluator
  sampler.npz    recbole.sampler.Sampler.sample_by_user_ids with its random_list / random_pr state
  fm_steps.npz   recbole FM (token fields) + BCELoss + Adam, 2 steps
  dense_adam_pointwise.npz  FM (wd 0 / 1e-3) and MFSimple (its yaml: wd 1e-8, lr 2e-3) under dense torch.optim.Adam for
                 4-5 steps on batches that leave most rows untouched (pins the fused 'adam_lazy' kind)
  fm_float.npz   recbole FM with TOKEN + FLOAT fields, embedding_size 10 (FM.yaml), dense Adam, 4 steps
  fm_seq.npz     recbole FM with TOKEN + FLOAT + two TOKEN_SEQ fields (mean pooling over the non-zero ids, some samples
                 with an empty sequence), embedding_size 10, dense Adam, 4 steps
  ce_backward.npz  torch autograd of that head w.r.t. seq_output and the item table + one dense Adam step
  ce_head.npz    the SASRec head expressions of sasrec.py:137-141,152-158
  cfg1_train.npz BASELINE config 1: the reference pipeline on ml-100k, 2 epochs of Trainer._train_epoch
                 (every batch recorded) + Trainer.evaluate
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)
sys.argv = sys.argv[:1]  # the reference's Config parses argv

import numpy as np  # noqa: E402

np.float = float  # removed alias used at metrics.py:58,84,87,135,141

import torch  # noqa: E402

torch.set_num_threads(4)


class StubConfig(dict):
    def __getitem__(self, k):
        return self.get(k, None)


class StubDataset:
    """Just enough of the dataloader 'dlapi' surface the model constructors read
    (abstract_recommender.py:78-95,159-234)."""

    def __init__(self, nums, field2type=None):
        self._nums = nums
        self.field2type = field2type or {}

    def num(self, field):
        return self._nums[field]

    def fields(self):
        return list(self._nums.keys())


def g_metrics(out):
    from recbole.evaluator.metrics import metrics_dict
    names = ["recall", "mrr", "ndcg", "hit", "precision", "map"]
    pos_idx = np.array([[0, 0, 0], [1, 1, 1], [1, 0, 1], [0, 0, 1]])
    pos_len = np.array([1, 3, 4, 2])
    rng = np.random.default_rng(7)
    big_len = rng.integers(1, 30, size=200)
    big_idx = rng.random((200, 10)) < 0.2
    d = dict(names=np.array(names), pos_idx=pos_idx, pos_len=pos_len, big_idx=big_idx, big_len=big_len)
    for n in names:
        d["ka_" + n] = np.asarray(metrics_dict[n](pos_idx, pos_len), dtype=np.float64)
        d["big_" + n] = np.asarray(metrics_dict[n](big_idx, big_len), dtype=np.float64)
    np.savez_compressed(out, **d)


def g_bpr(out):
    from recbole.model.general_recommender.bpr import BPR
    d = {}
    for dim in (16, 64, 128):
        torch.manual_seed(2020 + dim)
        n_users, n_items, B = 37, 53, 96
        cfg = StubConfig(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device="cpu",
                         embedding_size=dim)
        ds = StubDataset({"user_id": n_users, "item_id": n_items})
        rng = np.random.default_rng(dim)
        ids = [dict(user_id=torch.from_numpy(rng.integers(1, n_users, B)),
                    item_id=torch.from_numpy(rng.integers(1, n_items, B)),
                    neg_item_id=torch.from_numpy(rng.integers(1, n_items, B))) for _ in range(3)]
        for opt_name in ("adam", "sgd", "adam_wd"):
            torch.manual_seed(2020 + dim)
            model = BPR(cfg, ds)
            # make gradients large enough that 1e-5 relative on the update is a real test
            with torch.no_grad():
                model.user_embedding.weight.mul_(4.0)
                model.item_embedding.weight.mul_(4.0)
            key = "d%d_%s_" % (dim, opt_name)
            d[key + "U0"] = model.user_embedding.weight.detach().numpy().copy()
            d[key + "V0"] = model.item_embedding.weight.detach().numpy().copy()
            # trainer.py:115-118
            if opt_name == "adam":
                opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=0.0)
            elif opt_name == "adam_wd":
                opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=1e-3)
            else:
                opt = torch.optim.SGD(model.parameters(), lr=0.5, weight_decay=0.0)
            for s, inter in enumerate(ids):
                opt.zero_grad()
                loss = model.calculate_loss(inter)  # bpr.py:74-83
                loss.backward()
                if s == 0 and opt_name == "adam":
                    d[key + "gU"] = model.user_embedding.weight.grad.numpy().copy()
                    d[key + "gV"] = model.item_embedding.weight.grad.numpy().copy()
                opt.step()
                d[key + "loss%d" % s] = np.float32(loss.item())
                d[key + "U%d" % (s + 1)] = model.user_embedding.weight.detach().numpy().copy()
                d[key + "V%d" % (s + 1)] = model.item_embedding.weight.detach().numpy().copy()
            if opt_name.startswith("adam"):
                st = opt.state[model.user_embedding.weight]
                d[key + "mU"] = st["exp_avg"].numpy().copy()
                d[key + "vU"] = st["exp_avg_sq"].numpy().copy()
                st = opt.state[model.item_embedding.weight]
                d[key + "mV"] = st["exp_avg"].numpy().copy()
                d[key + "vV"] = st["exp_avg_sq"].numpy().copy()
            with torch.no_grad():
                d[key + "pred"] = model.predict(ids[0]).numpy().copy()  # bpr.py:85-89
        for s, inter in enumerate(ids):
            for k, v in inter.items():
                d["d%d_%s%d" % (dim, k, s)] = v.numpy()
    np.savez_compressed(out, **d)


def g_bpr_gaps(out):
    """The reference's BPR under dense torch.optim.Adam for 16 steps of SMALL batches: most rows are untouched in most
    steps and keep moving on their momentum (and, with weight decay, on wd * p) -- the trajectory the fused 'adam_lazy'
    kind replays (pins oracle.bpr's dense mode over long gaps)."""
    from recbole.model.general_recommender.bpr import BPR
    d = {}
    n_users, n_items, B, dim, steps = 300, 200, 24, 64, 16
    cfg = StubConfig(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device="cpu", embedding_size=dim)
    ds = StubDataset({"user_id": n_users, "item_id": n_items})
    rng = np.random.default_rng(99)
    ids = [dict(user_id=torch.from_numpy(rng.integers(1, n_users, B)), item_id=torch.from_numpy(rng.integers(1, n_items, B)),
                neg_item_id=torch.from_numpy(rng.integers(1, n_items, B))) for _ in range(steps)]
    for s, inter in enumerate(ids):
        for k, v in inter.items():
            d["%s%d" % (k, s)] = v.numpy()
    for tag, wd in (("wd0", 0.0), ("wd", 1e-4)):
        torch.manual_seed(77)
        model = BPR(cfg, ds)
        with torch.no_grad():
            model.user_embedding.weight.mul_(4.0)
            model.item_embedding.weight.mul_(4.0)
        d[tag + "_U0"] = model.user_embedding.weight.detach().numpy().copy()
        d[tag + "_V0"] = model.item_embedding.weight.detach().numpy().copy()
        opt = torch.optim.Adam(model.parameters(), lr=5e-3, weight_decay=wd)      # trainer.py:115-116
        for s, inter in enumerate(ids):
            opt.zero_grad()
            loss = model.calculate_loss(inter)
            loss.backward()
            opt.step()
            d[tag + "_loss%d" % s] = np.float32(loss.item())
        d[tag + "_UN"] = model.user_embedding.weight.detach().numpy().copy()
        d[tag + "_VN"] = model.item_embedding.weight.detach().numpy().copy()
        for nm, prm in (("U", model.user_embedding.weight), ("V", model.item_embedding.weight)):
            d[tag + "_m" + nm] = opt.state[prm]["exp_avg"].numpy().copy()
            d[tag + "_v" + nm] = opt.state[prm]["exp_avg_sq"].numpy().copy()
    d["steps"] = np.int64(steps)
    np.savez_compressed(out, **d)


def g_dot(out):
    from recbole.model.general_recommender.mfsimple import MFSimple
    torch.manual_seed(11)
    n_users, n_items, B, dim = 29, 41, 80, 32
    cfg = StubConfig(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device="cpu",
                     embedding_size=dim, embedding_dimension=dim, LABEL_FIELD="label")
    ds = StubDataset({"user_id": n_users, "item_id": n_items})
    model = MFSimple(cfg, ds)
    with torch.no_grad():
        for p in model.parameters():
            if p.ndim == 2:
                p.mul_(30.0)  # init std is 0.01 (mfsimple.py:37)
            else:
                p.add_(torch.randn_like(p) * 0.1)
    rng = np.random.default_rng(5)
    inter = dict(user_id=torch.from_numpy(rng.integers(1, n_users, B)),
                 item_id=torch.from_numpy(rng.integers(1, n_items, B)),
                 label=torch.from_numpy((rng.random(B) < 0.3).astype(np.float32)))
    d = {k: v.numpy() for k, v in inter.items()}
    for n, p in model.named_parameters():
        d["p_" + n] = p.detach().numpy().copy()
    loss = model.calculate_loss(inter)
    loss.backward()
    d["loss"] = np.float32(loss.item())
    for n, p in model.named_parameters():
        d["g_" + n] = p.grad.numpy().copy()
    with torch.no_grad():
        d["pred"] = model.predict(inter).numpy().copy()
    np.savez_compressed(out, **d)


def _pipeline(config_dict):
    import logging
    from recbole.config import Config
    from recbole.data import create_dataset, data_preparation
    from recbole.utils import init_seed
    config = Config(config_dict=config_dict)
    init_seed(config["seed"], config["reproducibility"])
    logging.basicConfig(level=logging.ERROR)
    dataset = create_dataset(config)
    return (config,) + tuple(data_preparation(config, dataset))


def _fullsort(config, train_data, test_data, out, seed, scale=1.0, extra=None, train_epochs=0):
    from recbole.model.general_recommender.bpr import BPR
    from recbole.trainer import Trainer
    torch.manual_seed(seed)
    model = BPR(config, train_data).to(config["device"])
    if scale != 1.0:
        with torch.no_grad():
            model.user_embedding.weight.mul_(scale)
            model.item_embedding.weight.mul_(scale)
    cwd = os.getcwd()
    os.makedirs("/tmp/rb2_golden_scratch", exist_ok=True)
    os.chdir("/tmp/rb2_golden_scratch")  # Trainer makes ./saved
    try:
        trainer = Trainer(config, model)
    finally:
        os.chdir(cwd)
    for ep in range(train_epochs):  # reference training loop, trainer.py:132-174
        trainer._train_epoch(train_data, ep)
    model.eval()
    trainer.tot_item_num = test_data.dataset.item_num
    n_items = test_data.dataset.item_num
    uids, raw, masked, mats, hist_rows, hist_cols, pos_items = [], [], [], [], [], [], []
    swaps = []
    with torch.no_grad():
        row0 = 0
        for batched in test_data:
            inter, hist, swap_row, swap_after, swap_before = batched
            u = inter["user_id"].numpy()
            uids.append(u)
            raw.append(model.full_sort_predict(inter).view(-1, n_items).numpy().copy())  # bpr.py:91-96
            inter2, scores = trainer._full_sort_batch_eval(batched)  # trainer.py:328-352
            masked.append(scores.numpy().copy())
            mats.append(trainer.evaluator.collect(inter2, scores)[0].numpy().copy())  # evaluators.py:53-76
            hist_rows.append(hist[0].numpy() + row0)
            hist_cols.append(hist[1].numpy())
            swaps.append((swap_row.numpy() + row0, swap_after.numpy(), swap_before.numpy()))
            row0 += len(u)
        result = trainer.evaluate(test_data, load_best_model=False)  # trainer.py:354-412
    uids = np.concatenate(uids)
    d = dict(
        n_items=n_items, uid_list=uids,
        U=model.user_embedding.weight.detach().numpy(), V=model.item_embedding.weight.detach().numpy(),
        raw_scores=np.concatenate(raw), topk_matrix=np.concatenate(mats),
        hist_row=np.concatenate(hist_rows), hist_col=np.concatenate(hist_cols),
        swap_row=np.concatenate([s[0] for s in swaps]), swap_after=np.concatenate([s[1] for s in swaps]),
        swap_before=np.concatenate([s[2] for s in swaps]),
        pos_len=np.asarray(test_data.get_pos_len_list()),
        result_keys=np.array(list(result.keys())), result_vals=np.array(list(result.values()), dtype=np.float64),
        topk=np.asarray(config["topk"]), metrics=np.array([m.lower() for m in config["metrics"]]),
    )
    # positives of the evaluated phase, straight from the dataloader's dataset (general_dataloader.py:304-313)
    ds = test_data.dataset
    d["pos_user"] = ds.inter_feat[ds.uid_field].numpy()
    d["pos_item"] = ds.inter_feat[ds.iid_field].numpy()
    # used ids of the phase (sampler.py:206-227) as (user, item) pairs
    used = test_data.sampler.used_ids
    uu, ii = [], []
    for u in range(len(used)):
        for i in sorted(used[u]):
            uu.append(u)
            ii.append(i)
    d["used_user"], d["used_item"] = np.array(uu, dtype=np.int64), np.array(ii, dtype=np.int64)
    if extra:
        d.update(extra)
    if d["raw_scores"].size > 400000:
        d.pop("raw_scores")
    np.savez_compressed(out, **d)
    return result


def g_fullsort_small(out):
    cfg = {
        "model": "BPR", "dataset": "general_full_dataloader", "data_path": os.path.join(REF, "tests", "data"),
        "load_col": None, "eval_setting": "TO_RS,full", "training_neg_sample_num": 1,
        "split_ratio": [0.8, 0.1, 0.1], "train_batch_size": 6, "eval_batch_size": 100, "use_gpu": False,
        "topk": [1, 5, 10], "metrics": ["Recall", "MRR", "NDCG", "Hit", "Precision", "MAP"],
    }
    config, train, valid, test = _pipeline(cfg)
    print("small:", _fullsort(config, train, test, out, seed=3, scale=8.0))


def g_fullsort_ml100k(out, out_sampler):
    cfg = {
        "model": "BPR", "dataset": "ml-100k", "data_path": os.path.join(REF, "dataset"),
        "load_col": {"inter": ["user_id", "item_id"]}, "use_gpu": False,
        "topk": [10], "metrics": ["Recall", "MRR", "NDCG", "Hit", "Precision"],
    }
    config, train, valid, test = _pipeline(cfg)
    # ---- sampler golden (sampler.py:103-154,246-265) on the real train sampler ----
    smp = train.sampler
    used = smp.used_ids
    uu, ii = [], []
    for u in range(len(used)):
        for i in sorted(used[u]):
            uu.append(u)
            ii.append(i)
    sd = dict(random_list=np.asarray(smp.random_list), n_items=smp.n_items, n_users=smp.n_users,
              used_user=np.array(uu, dtype=np.int64), used_item=np.array(ii, dtype=np.int64))
    rng = np.random.default_rng(99)
    # start near the end of the list so the wrap-around branch (sampler.py:95-99) is exercised
    smp.random_pr = len(smp.random_list) - 700
    for c, (B, num) in enumerate([(2048, 1), (500, 3), (64, 1), (1, 5)]):
        users = rng.integers(1, smp.n_users, B)
        if c == 2:
            users[:] = users[0]  # single-key branch (sampler.py:117-143)
        sd["pr_before%d" % c] = smp.random_pr
        sd["users%d" % c] = users
        sd["num%d" % c] = num
        sd["out%d" % c] = smp.sample_by_user_ids(users, num).numpy()
        sd["pr_after%d" % c] = smp.random_pr
    np.savez_compressed(out_sampler, **sd)
    # one training batch as the dataloader builds it (general_dataloader.py:212-241)
    train.shuffle = False
    batch = next(iter(train))
    extra = dict(train_user=batch["user_id"].numpy(), train_pos=batch["item_id"].numpy(),
                 train_neg=batch["neg_item_id"].numpy())
    print("ml-100k:", _fullsort(config, train, test, out, seed=2020, scale=1.0, extra=extra, train_epochs=8))


def g_cfg1_train(out):
    """BASELINE config 1 trajectory: the reference's own pipeline on ml-100k (BPR, d=64, Adam 1e-3,
    B=2048), 2 epochs through Trainer._train_epoch, then Trainer.evaluate on the test split.  Saves
    the initial tables, every batch the reference's dataloader + sampler produced, the final tables
    and the result dict."""
    from recbole.model.general_recommender.bpr import BPR
    from recbole.trainer import Trainer
    cfg = {
        "model": "BPR", "dataset": "ml-100k", "data_path": os.path.join(REF, "dataset"),
        "load_col": {"inter": ["user_id", "item_id"]}, "use_gpu": False,
        "topk": [10], "metrics": ["Recall", "MRR", "NDCG", "Hit", "Precision"],
    }
    config, train, valid, test = _pipeline(cfg)
    torch.manual_seed(2020)
    model = BPR(config, train).to(config["device"])
    cwd = os.getcwd()
    os.makedirs("/tmp/rb2_golden_scratch", exist_ok=True)
    os.chdir("/tmp/rb2_golden_scratch")
    try:
        trainer = Trainer(config, model)
    finally:
        os.chdir(cwd)
    d = dict(U0=model.user_embedding.weight.detach().numpy().copy(),
             V0=model.item_embedding.weight.detach().numpy().copy())
    # record the batches by wrapping the model's loss function (trainer.py:151: loss_func argument)
    rec = []

    def loss_func(interaction):
        rec.append(np.stack([interaction["user_id"].numpy(), interaction["item_id"].numpy(),
                             interaction["neg_item_id"].numpy()]).astype(np.int32))
        return model.calculate_loss(interaction)

    losses = []
    for ep in range(2):
        losses.append(trainer._train_epoch(train, ep, loss_func=loss_func))
    d["epoch_loss"] = np.array(losses, dtype=np.float64)
    d["batch_sizes"] = np.array([b.shape[1] for b in rec], dtype=np.int64)
    d["batches"] = np.concatenate(rec, axis=1)
    d["U"] = model.user_embedding.weight.detach().numpy().copy()
    d["V"] = model.item_embedding.weight.detach().numpy().copy()
    with torch.no_grad():
        result = trainer.evaluate(test, load_best_model=False)
    d["result_keys"] = np.array(list(result.keys()))
    d["result_vals"] = np.array(list(result.values()), dtype=np.float64)
    ds = test.dataset
    d["pos_user"] = ds.inter_feat[ds.uid_field].numpy()
    d["pos_item"] = ds.inter_feat[ds.iid_field].numpy()
    used = test.sampler.used_ids
    uu, ii = [], []
    for u in range(len(used)):
        for i in sorted(used[u]):
            uu.append(u)
            ii.append(i)
    d["used_user"], d["used_item"] = np.array(uu, dtype=np.int32), np.array(ii, dtype=np.int32)
    d["uid_list"] = np.asarray(test.uid_list)
    d["n_items"] = ds.item_num
    print("cfg1 train:", losses, result)
    np.savez_compressed(out, **d)


def g_fm(out):
    from recbole.model.context_aware_recommender.fm import FM
    from recbole.utils import FeatureType
    torch.manual_seed(5)
    field_dims = {"f0": 7, "f1": 30, "f2": 3, "f3": 19, "f4": 11}
    dim, B = 16, 64
    f2t = {k: FeatureType.TOKEN for k in field_dims}
    f2t["label"] = FeatureType.FLOAT
    nums = dict(field_dims)
    nums["label"] = 1
    cfg = StubConfig(LABEL_FIELD="label", embedding_size=dim, device="cpu", double_tower=None)
    model = FM(cfg, StubDataset(nums, f2t))
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(2.0)
    rng = np.random.default_rng(3)
    d = dict(field_dims=np.array(list(field_dims.values())), offsets=np.asarray(model.token_field_offsets))
    for n, p in model.named_parameters():
        d["p0_" + n] = p.detach().numpy().copy()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=0.0)
    for s in range(2):
        inter = {k: torch.from_numpy(rng.integers(0, v, B)) for k, v in field_dims.items()}
        inter["label"] = torch.from_numpy((rng.random(B) < 0.3).astype(np.float32))
        d["ids%d" % s] = np.stack([inter[k].numpy() for k in field_dims], axis=1)
        d["label%d" % s] = inter["label"].numpy()
        with torch.no_grad():
            d["pred%d" % s] = model.predict(inter).numpy().copy()
        opt.zero_grad()
        loss = model.calculate_loss(inter)  # fm.py:52-56
        loss.backward()
        if s == 0:
            for n, p in model.named_parameters():
                d["g0_" + n] = p.grad.numpy().copy()
        opt.step()
        d["loss%d" % s] = np.float32(loss.item())
        for n, p in model.named_parameters():
            d["p%d_" % (s + 1) + n] = p.detach().numpy().copy()
    np.savez_compressed(out, **d)


def g_dense_adam_pointwise(out):
    """Dense torch.optim.Adam trajectories of the two point-wise models, on batches that leave most rows UNTOUCHED
    (so row-sparse Adam and the reference's dense Adam differ from step 2 on): recbole FM, 4 steps, weight_decay 0
    and 1e-3; the fork's MFSimple with ITS config (MFSimple.yaml: weight_decay 1e-08, learning_rate 0.002), 5 steps."""
    from recbole.model.context_aware_recommender.fm import FM
    from recbole.model.general_recommender.mfsimple import MFSimple
    from recbole.utils import FeatureType
    d = {}
    field_dims = {"f0": 40, "f1": 300, "f2": 3, "f3": 90, "f4": 11, "f5": 700}
    dim, B, steps = 16, 48, 4
    f2t = {k: FeatureType.TOKEN for k in field_dims}
    f2t["label"] = FeatureType.FLOAT
    nums = dict(field_dims)
    nums["label"] = 1
    for tag, wd in (("fm_wd0", 0.0), ("fm_wd", 1e-3)):
        torch.manual_seed(9)
        model = FM(StubConfig(LABEL_FIELD="label", embedding_size=dim, device="cpu", double_tower=None),
                   StubDataset(nums, f2t))
        with torch.no_grad():
            for p in model.parameters():
                p.mul_(2.0)
        rng = np.random.default_rng(13)
        d[tag + "_field_dims"] = np.array(list(field_dims.values()))
        d[tag + "_offsets"] = np.asarray(model.token_field_offsets)
        for n, p in model.named_parameters():
            d[tag + "_p0_" + n] = p.detach().numpy().copy()
        opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=wd)
        for s in range(steps):
            inter = {k: torch.from_numpy(rng.integers(0, v, B)) for k, v in field_dims.items()}
            inter["label"] = torch.from_numpy((rng.random(B) < 0.3).astype(np.float32))
            d[tag + "_ids%d" % s] = np.stack([inter[k].numpy() for k in field_dims], axis=1)
            d[tag + "_label%d" % s] = inter["label"].numpy()
            opt.zero_grad()
            loss = model.calculate_loss(inter)
            loss.backward()
            opt.step()
            d[tag + "_loss%d" % s] = np.float32(loss.item())
        for n, p in model.named_parameters():
            d[tag + "_pN_" + n] = p.detach().numpy().copy()
        with torch.no_grad():
            d[tag + "_predN"] = model.predict(inter).numpy().copy()
    # MFSimple, its own hyper-parameters (properties/model/MFSimple.yaml)
    torch.manual_seed(21)
    n_users, n_items, B, dim, steps = 200, 300, 64, 32, 5
    cfg = StubConfig(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device="cpu",
                     embedding_dimension=dim, LABEL_FIELD="label")
    model = MFSimple(cfg, StubDataset({"user_id": n_users, "item_id": n_items}))
    with torch.no_grad():
        for p in model.parameters():
            if p.ndim == 2:
                p.mul_(30.0)
    rng = np.random.default_rng(17)
    for n, p in model.named_parameters():
        d["mf_p0_" + n] = p.detach().numpy().copy()
    opt = torch.optim.Adam(model.parameters(), lr=0.002, weight_decay=1e-08)
    for s in range(steps):
        inter = dict(user_id=torch.from_numpy(rng.integers(1, n_users, B)),
                     item_id=torch.from_numpy(rng.integers(1, n_items, B)),
                     label=torch.from_numpy((rng.random(B) < 0.3).astype(np.float32)))
        for k, v in inter.items():
            d["mf_%s%d" % (k, s)] = v.numpy()
        opt.zero_grad()
        loss = model.calculate_loss(inter)
        loss.backward()
        opt.step()
        d["mf_loss%d" % s] = np.float32(loss.item())
    for n, p in model.named_parameters():
        d["mf_pN_" + n] = p.detach().numpy().copy()
    np.savez_compressed(out, **d)


def g_fm_float(out):
    """recbole FM with TOKEN and FLOAT fields at FM.yaml's default embedding_size (10), dense torch.optim.Adam, 4 steps
    (abstract_recommender.py:236-258 float embeddings scaled by the value; layers.py:947-966 first order)."""
    from recbole.model.context_aware_recommender.fm import FM
    from recbole.utils import FeatureType
    torch.manual_seed(31)
    token_dims = {"t0": 40, "t1": 300, "t2": 3}
    float_names = ["x0", "x1"]
    dim, B, steps = 10, 48, 4
    f2t = {k: FeatureType.TOKEN for k in token_dims}
    f2t.update({k: FeatureType.FLOAT for k in float_names})
    f2t["label"] = FeatureType.FLOAT
    nums = dict(token_dims)
    nums.update({k: 1 for k in float_names})
    nums["label"] = 1
    model = FM(StubConfig(LABEL_FIELD="label", embedding_size=dim, device="cpu", double_tower=None),
               StubDataset(nums, f2t))
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(2.0)
    rng = np.random.default_rng(41)
    d = dict(token_dims=np.array(list(token_dims.values())), offsets=np.asarray(model.token_field_offsets),
             n_float=np.int64(len(float_names)))
    d["param_order"] = np.array([n for n, _ in model.named_parameters()])
    for n, p in model.named_parameters():
        d["p0_" + n] = p.detach().numpy().copy()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    for s in range(steps):
        inter = {k: torch.from_numpy(rng.integers(0, v, B)) for k, v in token_dims.items()}
        for k in float_names:
            inter[k] = torch.from_numpy(rng.random(B).astype(np.float32))        # normalised floats in [0, 1]
        inter["label"] = torch.from_numpy((rng.random(B) < 0.3).astype(np.float32))
        d["ids%d" % s] = np.stack([inter[k].numpy() for k in token_dims], axis=1)
        d["fx%d" % s] = np.stack([inter[k].numpy() for k in float_names], axis=1)
        d["label%d" % s] = inter["label"].numpy()
        opt.zero_grad()
        loss = model.calculate_loss(inter)
        loss.backward()
        opt.step()
        d["loss%d" % s] = np.float32(loss.item())
    for n, p in model.named_parameters():
        d["pN_" + n] = p.detach().numpy().copy()
    with torch.no_grad():
        d["predN"] = model.predict(inter).numpy().copy()
    np.savez_compressed(out, **d)


def g_fm_seq(out):
    """recbole FM with TOKEN, FLOAT and TOKEN_SEQ fields (abstract_recommender.py:277-314 masked mean pooling; first order
    = masked sum, layers.py:989-1019) at embedding_size 10, dense torch.optim.Adam, 4 steps."""
    from recbole.model.context_aware_recommender.fm import FM
    from recbole.utils import FeatureType
    torch.manual_seed(37)
    token_dims = {"t0": 30, "t1": 200}
    float_names = ["x0"]
    seq_dims = {"q0": 25, "q1": 7}
    seq_lens = {"q0": 6, "q1": 3}
    dim, B, steps = 10, 64, 4
    f2t = {k: FeatureType.TOKEN for k in token_dims}
    f2t.update({k: FeatureType.FLOAT for k in float_names})
    f2t.update({k: FeatureType.TOKEN_SEQ for k in seq_dims})
    f2t["label"] = FeatureType.FLOAT
    nums = dict(token_dims)
    nums.update({k: 1 for k in float_names})
    nums.update(seq_dims)
    nums["label"] = 1
    model = FM(StubConfig(LABEL_FIELD="label", embedding_size=dim, device="cpu", double_tower=None),
               StubDataset(nums, f2t))
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(2.0)
    rng = np.random.default_rng(43)
    d = dict(token_dims=np.array(list(token_dims.values())), n_float=np.int64(len(float_names)),
             seq_dims=np.array(list(seq_dims.values())), seq_lens=np.array(list(seq_lens.values())))
    d["param_order"] = np.array([n for n, _ in model.named_parameters()])
    for n, p in model.named_parameters():
        d["p0_" + n] = p.detach().numpy().copy()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    for s in range(steps):
        inter = {k: torch.from_numpy(rng.integers(0, v, B)) for k, v in token_dims.items()}
        for k in float_names:
            inter[k] = torch.from_numpy(rng.random(B).astype(np.float32))
        for k, v in seq_dims.items():
            # padded sequences: a random number (0 .. L) of non-zero ids first, zeros behind (dataset.py's padding)
            L = seq_lens[k]
            q = rng.integers(1, v, (B, L))
            n = rng.integers(0, L + 1, B)
            q[np.arange(L)[None, :] >= n[:, None]] = 0
            inter[k] = torch.from_numpy(q)
            d["%s_%d" % (k, s)] = q
        inter["label"] = torch.from_numpy((rng.random(B) < 0.3).astype(np.float32))
        d["ids%d" % s] = np.stack([inter[k].numpy() for k in token_dims], axis=1)
        d["fx%d" % s] = np.stack([inter[k].numpy() for k in float_names], axis=1)
        d["label%d" % s] = inter["label"].numpy()
        opt.zero_grad()
        loss = model.calculate_loss(inter)
        loss.backward()
        opt.step()
        d["loss%d" % s] = np.float32(loss.item())
    for n, p in model.named_parameters():
        d["pN_" + n] = p.detach().numpy().copy()
    with torch.no_grad():
        d["predN"] = model.predict(inter).numpy().copy()
    np.savez_compressed(out, **d)


def g_ce(out):
    torch.manual_seed(9)
    B, N, H, K = 48, 700, 64, 10
    X = torch.nn.functional.layer_norm(torch.randn(B, H), (H,))
    E = torch.randn(N, H) * 0.02 * 20
    E[0] = 0  # padding_idx=0 (sasrec.py:53)
    pos = torch.randint(1, N, (B,))
    logits = torch.matmul(X, E.transpose(0, 1))  # sasrec.py:139
    loss = torch.nn.CrossEntropyLoss()(logits, pos)  # sasrec.py:140
    scores = logits.clone()
    scores[:, 0] = -np.inf  # trainer.py:343
    _, idx = torch.topk(scores, K, dim=-1)
    np.savez_compressed(out, X=X.numpy(), E=E.numpy(), pos=pos.numpy(), loss=np.float32(loss.item()),
                        logits=logits.numpy(), topk_ids=idx.numpy(),
                        lse=torch.logsumexp(logits, dim=1).numpy())


def g_ce_backward(out):
    """autograd of the CE branch (sasrec.py:137-141) w.r.t. seq_output and item_embedding.weight, and one dense
    torch.optim.Adam step on the item table with that gradient (trainer.py:170-173)."""
    torch.manual_seed(10)
    B, N, H = 200, 1500, 64
    X = torch.nn.functional.layer_norm(torch.randn(B, H), (H,)).requires_grad_(True)
    E = (torch.randn(N, H) * 0.02 * 20)
    with torch.no_grad():
        E[0] = 0
    E.requires_grad_(True)
    pos = torch.randint(1, N, (B,))
    opt = torch.optim.Adam([E], lr=1e-3)
    logits = torch.matmul(X, E.transpose(0, 1))
    loss = torch.nn.CrossEntropyLoss()(logits, pos)
    loss.backward()
    d = dict(X=X.detach().numpy().copy(), E=E.detach().numpy().copy(), pos=pos.numpy(), loss=np.float32(loss.item()),
             dX=X.grad.numpy().copy(), dE=E.grad.numpy().copy())
    opt.step()
    d["E1"] = E.detach().numpy().copy()
    np.savez_compressed(out, **d)


if __name__ == "__main__":
    o = lambda n: os.path.join(HERE, n)  # noqa: E731
    g_metrics(o("metrics.npz"))
    g_bpr(o("bpr_steps.npz"))
    g_bpr_gaps(o("bpr_gaps.npz"))
    g_dot(o("dot_steps.npz"))
    g_fm(o("fm_steps.npz"))
    g_dense_adam_pointwise(o("dense_adam_pointwise.npz"))
    g_fm_float(o("fm_float.npz"))
    g_fm_seq(o("fm_seq.npz"))
    g_ce(o("ce_head.npz"))
    g_ce_backward(o("ce_backward.npz"))
    g_fullsort_small(o("fullsort_small.npz"))
    g_fullsort_ml100k(o("fullsort_ml100k.npz"), o("sampler.npz"))
    g_cfg1_train(o("cfg1_train.npz"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
