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
    """predict() serves the Trainer's fallback when full_sort_predict raises NotImplementedError
    (trainer.py:333-340): scores of explicit (user, item) pairs == the reference's predict()."""
    from recbole_b200 import Interaction
    from gpu_util import rel_err
    g = golden("bpr_steps.npz")
    key = "d64_adam_"
    model, _ = _make(g, key, 64, "adam", 1e-2)
    model.load_state_dict({"user_embedding.weight": torch.from_numpy(g[key + "U3"]),
                           "item_embedding.weight": torch.from_numpy(g[key + "V3"])})
    inter = Interaction({"user_id": torch.from_numpy(g["d64_user_id0"]), "item_id": torch.from_numpy(g["d64_item_id0"])}).to("cuda")
    assert rel_err(model.predict(inter).cpu().numpy(), g[key + "pred"]) < 1e-5
    with pytest.raises(NotImplementedError):
        model.full_sort_predict(inter)
    ids, sc = model.full_sort_topk(inter["user_id"][:5].contiguous(), 3)
    assert ids.shape == (5, 3) and (sc[:, 0] >= sc[:, 1]).all()
