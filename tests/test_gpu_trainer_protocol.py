"""GPU: the drop-in boundary as the reference's UNMODIFIED Trainer drives it
(recbole/trainer/trainer.py:157-173): optimizer.zero_grad() -> loss = model.calculate_loss(interaction)
-> loss.item() / isnan -> loss.backward() -> optimizer.step().  calculate_loss returns the loss kernel's
output, backward() only records the batch, FusedOptimizer.step() runs the fused step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Cfg(dict):
    def __getitem__(self, k):
        return self.get(k)


def _make(g, key, dim, learner, lr, wd=0.0):
    from recbole_b200 import FusedBPR

    class DS:
        def num(self, f):
            return {"user_id": g[key + "U0"].shape[0], "item_id": g[key + "V0"].shape[0]}[f]

    m = FusedBPR(Cfg(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device="cuda",
                     embedding_size=dim), DS()).to("cuda")
    m.load_state_dict({"user_embedding.weight": torch.from_numpy(g[key + "U0"]),
                       "item_embedding.weight": torch.from_numpy(g[key + "V0"])})
    return m, m.build_optimizer(learner, lr, wd)


@pytest.mark.parametrize("dim", [16, 64])
@pytest.mark.parametrize("opt_name,learner,lr,wd", [("adam", "adam_lazy", 1e-2, 0.0), ("sgd", "sgd", 0.5, 0.0),
                                                    ("adam_wd", "adam_lazy", 1e-2, 1e-3)])
def test_reference_trainer_loop_protocol(golden, dim, opt_name, learner, lr, wd):
    from recbole_b200 import Interaction
    from gpu_util import rel_err
    g = golden("bpr_steps.npz")
    key = "d%d_%s_" % (dim, opt_name)
    model, optimizer = _make(g, key, dim, learner, lr, wd)
    model.train()
    total = 0.0
    for s in range(3):
        inter = Interaction({f: torch.from_numpy(g["d%d_%s%d" % (dim, f, s)]) for f in
                             ("user_id", "item_id", "neg_item_id")}).to("cuda")
        optimizer.zero_grad()                       # trainer.py:160
        loss = model.calculate_loss(inter)          # trainer.py:161
        assert loss.dim() == 0 and loss.requires_grad
        total += loss.item()                        # trainer.py:168
        assert not torch.isnan(loss)                # trainer.py:234-236
        loss.backward()                             # trainer.py:170
        optimizer.step()                            # trainer.py:173
        ref = float(g[key + "loss%d" % s])
        assert abs(loss.item() - ref) <= 1e-5 * abs(ref)
    sd = optimizer.state_dict()                     # trainer.py:204 (flushes the lazy rows)
    assert rel_err(model.state_dict()["user_embedding.weight"].cpu().numpy(), g[key + "U3"]) < 1e-5
    assert rel_err(model.state_dict()["item_embedding.weight"].cpu().numpy(), g[key + "V3"]) < 1e-5
    if learner != "sgd":
        # the optimizer state has torch.optim.Adam's layout and, after the flush, its values
        assert int(sd["state"][0]["step"]) == 3
        assert rel_err(sd["state"][0]["exp_avg"].cpu().numpy(), g[key + "mU"]) < 1e-4
        assert rel_err(sd["state"][1]["exp_avg_sq"].cpu().numpy(), g[key + "vV"]) < 1e-4
        # round trip (trainer.py:230 resume_checkpoint)
        model2, opt2 = _make(g, key, dim, learner, lr, wd)
        opt2.load_state_dict(sd)
        assert opt2.model._optim.step == 3
        assert torch.equal(opt2.model._opt_state["mU"], sd["state"][0]["exp_avg"])
    # step() without a recorded batch is an error, like stepping without gradients would be a no-op there
    with pytest.raises(RuntimeError):
        optimizer.step()


def test_predict_and_full_sort_protocol(golden):
    """predict(): scores of explicit (user, item) pairs == the reference's predict() (bpr.py:85-89)."""
    from recbole_b200 import Interaction
    from gpu_util import rel_err
    g = golden("bpr_steps.npz")
    key = "d64_adam_"
    model, _ = _make(g, key, 64, "adam", 1e-2)
    model.load_state_dict({"user_embedding.weight": torch.from_numpy(g[key + "U3"]),
                           "item_embedding.weight": torch.from_numpy(g[key + "V3"])})
    inter = Interaction({"user_id": torch.from_numpy(g["d64_user_id0"]), "item_id": torch.from_numpy(g["d64_item_id0"])}).to("cuda")
    assert rel_err(model.predict(inter).cpu().numpy(), g[key + "pred"]) < 1e-5
    ids, sc = model.full_sort_topk(inter["user_id"][:5].contiguous(), 3)
    assert ids.shape == (5, 3) and (sc[:, 0] >= sc[:, 1]).all()


def _unmodified_trainer_epoch(model, optimizer, batches):
    """recbole/trainer/trainer.py:157-173 verbatim in structure: what an UNMODIFIED Trainer._train_epoch does."""
    total = None
    for inter in batches:
        optimizer.zero_grad()
        losses = model.calculate_loss(inter)
        total = losses.item() if total is None else total + losses.item()
        assert not torch.isnan(losses)
        losses.backward()
        optimizer.step()
    return total


@pytest.mark.parametrize("opt_name,learner,lr,wd", [("adam", "adam", 1e-2, 0.0), ("sgd", "sgd", 0.5, 0.0),
                                                    ("adam_wd", "adam", 1e-2, 1e-3)])
def test_completely_unmodified_trainer_autostep(golden, opt_name, learner, lr, wd):
    """Trainer(config, model) builds torch.optim over model.parameters() (trainer.py:103,109-130); its step() finds
    no gradients.  The model reads learner / learning_rate / weight_decay from the same config and takes the fused
    step inside loss.backward(); 'adam' means the reference's dense Adam (fused kind adam_lazy)."""
    from recbole_b200 import FusedBPR, Interaction
    from gpu_util import rel_err
    g = golden("bpr_steps.npz")
    dim = 64
    key = "d%d_%s_" % (dim, opt_name)

    class DS:
        def num(self, f):
            return {"user_id": g[key + "U0"].shape[0], "item_id": g[key + "V0"].shape[0]}[f]

    cfg = Cfg(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device="cuda", embedding_size=dim,
              learner=learner, learning_rate=lr, weight_decay=wd)
    model = FusedBPR(cfg, DS()).to("cuda")
    model.load_state_dict({"user_embedding.weight": torch.from_numpy(g[key + "U0"]),
                           "item_embedding.weight": torch.from_numpy(g[key + "V0"])})
    torch_opt = {"adam": torch.optim.Adam, "sgd": torch.optim.SGD}[learner](model.parameters(), lr=lr, weight_decay=wd)
    batches = [Interaction({f: torch.from_numpy(g["d%d_%s%d" % (dim, f, s)]) for f in
                            ("user_id", "item_id", "neg_item_id")}).to("cuda") for s in range(3)]
    total = _unmodified_trainer_epoch(model, torch_opt, batches)
    ref_total = sum(float(g[key + "loss%d" % s]) for s in range(3))
    assert abs(total - ref_total) <= 1e-5 * abs(ref_total)
    sd = model.state_dict()                       # trainer.py:203 (flushes the lazy rows)
    assert rel_err(sd["user_embedding.weight"].cpu().numpy(), g[key + "U3"]) < 1e-5
    assert rel_err(sd["item_embedding.weight"].cpu().numpy(), g[key + "V3"]) < 1e-5
    assert model._optim.kind_name == ("adam_lazy" if learner == "adam" else "sgd")


def test_full_sort_predict_is_the_reference_score_matrix(golden):
    """bpr.py:91-96 for an unmodified Trainer._full_sort_batch_eval (trainer.py:328-352): flat [users * n_items]
    scores, every one the canonical fp32 chain (bit-identical to the oracle), == the reference's matmul to 1e-5."""
    from oracle import _clib
    from recbole_b200 import FusedBPR, Interaction
    from gpu_util import rel_err
    g = golden("bpr_steps.npz")
    key = "d64_adam_"

    class DS:
        def num(self, f):
            return {"user_id": g[key + "U3"].shape[0], "item_id": g[key + "V3"].shape[0]}[f]

    model = FusedBPR(Cfg(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device="cuda",
                         embedding_size=64), DS()).to("cuda")
    U, V = g[key + "U3"], g[key + "V3"]
    model.load_state_dict({"user_embedding.weight": torch.from_numpy(U), "item_embedding.weight": torch.from_numpy(V)})
    users = np.array([3, 1, 1, U.shape[0] - 1, 2, 7, 5, 4, 6, 8, 9], dtype=np.int64)     # > one pass of 8 query rows
    out = model.full_sort_predict(Interaction({"user_id": torch.from_numpy(users)}).to("cuda"))
    assert out.shape == (len(users) * V.shape[0],)
    got = out.view(len(users), -1).cpu().numpy()
    assert rel_err(got, U[users] @ V.T) < 1e-5
    assert np.array_equal(got, _clib.lib.scores_fma(U[users], V))       # bit-exact vs the oracle's canonical chain
    # the Trainer's own masking + topk on that matrix agrees with the fused top-K (no history here)
    s = out.view(len(users), -1).clone()
    s[:, 0] = -np.inf                                       # trainer.py:343
    ids, sc = model.full_sort_topk(torch.from_numpy(users).cuda(), 5)
    top = torch.topk(s, 5, dim=1)
    assert torch.equal(top.values, sc)


def test_fm_and_mfsimple_calculate_loss_backward_step(golden):
    """FM / MFSimple under the reference Trainer's loop: calculate_loss is the fused forward + BCE kernel (no ATen on
    the path), backward() records the batch, FusedOptimizer.step() takes the fused step; the optimizer state has
    torch.optim.Adam's layout in the reference's parameter order."""
    from recbole_b200 import FusedFM, Interaction
    from gpu_util import rel_err
    g = golden("fm_steps.npz")
    names = [str(n) for n in g["field_names"]] if "field_names" in g.files else ["f%d" % i for i in range(g["ids0"].shape[1])]
    dims = np.diff(np.concatenate([g["offsets"], [g["p0_token_embedding_table.embedding.weight"].shape[0]]])).astype(int)

    class DS:
        field2type = {n: "token" for n in names}

        def fields(self):
            return names + ["label"]

        def num(self, f):
            return int(dims[names.index(f)])

    DS.field2type["label"] = "float"
    d = g["p0_token_embedding_table.embedding.weight"].shape[1]
    cfg = Cfg(LABEL_FIELD="label", embedding_size=d, device="cuda", learner="adam", learning_rate=1e-2)
    model = FusedFM(cfg, DS()).to("cuda")
    model.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p0_")})
    optimizer = model.build_optimizer("adam_lazy", 1e-2)      # dense-Adam trajectory (some rows miss a batch)
    for s in range(2):
        inter = Interaction(dict({n: torch.from_numpy(g["ids%d" % s][:, i].astype(np.int64)) for i, n in enumerate(names)},
                                 label=torch.from_numpy(g["label%d" % s]))).to("cuda")
        optimizer.zero_grad()
        loss = model.calculate_loss(inter)
        assert loss.dim() == 0 and loss.requires_grad
        ref = float(g["loss%d" % s])
        assert abs(loss.item() - ref) <= 1e-5 * abs(ref)
        loss.backward()
        optimizer.step()
    sd = model.state_dict()
    assert rel_err(sd["token_embedding_table.embedding.weight"].cpu().numpy(),
                   g["p2_token_embedding_table.embedding.weight"]) < 1e-5
    osd = optimizer.state_dict()
    assert sorted(osd["state"]) == [0, 1, 2] and int(osd["state"][0]["step"]) == 2
    assert osd["state"][0]["exp_avg"].shape == sd["token_embedding_table.embedding.weight"].shape      # E
    assert osd["state"][1]["exp_avg"].numel() == 1                                                      # bias
    assert osd["state"][2]["exp_avg"].numel() == sd["token_embedding_table.embedding.weight"].shape[0]  # W


def test_fused_trainer_checkpoint_save_resume(tmp_path, golden):
    """trainer.py:191-232,372-380: fit(saved=True) writes the reference's checkpoint layout; resume_checkpoint in a
    new trainer continues the SAME trajectory (tables, Adam moments, step counter, lazy rows)."""
    from recbole_b200 import FusedBPR, FusedTrainer, Interaction
    g = golden("cfg1_train.npz")

    class DS:
        def num(self, f):
            return {"user_id": g["U0"].shape[0], "item_id": g["V0"].shape[0]}[f]

    offs = np.concatenate([[0], np.cumsum(g["batch_sizes"])])
    batches = []
    for b in range(12):
        blk = g["batches"][:, offs[b]:offs[b + 1]].astype(np.int64)
        batches.append(Interaction({"user_id": torch.from_numpy(blk[0]), "item_id": torch.from_numpy(blk[1]),
                                    "neg_item_id": torch.from_numpy(blk[2])}))

    def make(epochs):
        cfg = Cfg(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device=torch.device("cuda:0"),
                  embedding_size=64, learner="adam", learning_rate=1e-3, epochs=epochs, metrics=["Recall"], topk=[10],
                  checkpoint_dir=str(tmp_path), model="FusedBPR")
        m = FusedBPR(cfg, DS()).to("cuda:0")
        m.load_state_dict({"user_embedding.weight": torch.from_numpy(g["U0"]),
                           "item_embedding.weight": torch.from_numpy(g["V0"])})
        return m, FusedTrainer(cfg, m)

    # straight through: 3 epochs over the same 4 batches each
    m_ref, t_ref = make(3)
    for ep in range(3):
        t_ref._train_epoch(batches[4 * ep:4 * ep + 4], ep)
    # one epoch, checkpoint, resume in a fresh trainer, two more
    m1, t1 = make(1)
    t1.fit(batches[0:4], valid_data=None, verbose=False, saved=True)
    ck = torch.load(t1.saved_model_file, weights_only=False)
    assert sorted(ck) == ["best_valid_score", "config", "cur_step", "epoch", "optimizer", "state_dict"]
    assert ck["epoch"] == 0 and int(ck["optimizer"]["state"][0]["step"]) == 4
    m2, t2 = make(3)
    t2.resume_checkpoint(t1.saved_model_file)
    assert t2.start_epoch == 1
    for ep in range(1, 3):
        t2._train_epoch(batches[4 * ep:4 * ep + 4], ep)
    a, b = m_ref.state_dict(), m2.state_dict()
    for k in a:
        assert torch.allclose(a[k], b[k], rtol=1e-6, atol=1e-9), k
    # evaluate(load_best_model=True) reads the trainer's own file (trainer.py:372-380)
    t2._save_checkpoint(2)
    assert t2.evaluate(None, load_best_model=True) is None
