"""GPU: device-resident training pipeline (DeviceTrainLoader + DeviceSampler + FusedTrainer.fit) on a
small synthetic dataset: an epoch visits every interaction once, negatives respect the sampler
contract, the loss goes down and the evaluation improves -- everything without a host round trip per
batch."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Cfg(dict):
    def __getitem__(self, k):
        return self.get(k)


def test_device_pipeline_fit_and_evaluate():
    import bench_workloads as bw
    from recbole_b200 import DeviceSampler, DeviceTrainLoader, EvalIndex, FusedBPR, FusedTrainer
    from recbole_b200.data import build_csr
    dev = torch.device("cuda:0")
    n_users, n_items = 3000, 800
    user, item = bw.interactions(n_users, n_items, 60000, seed=1)
    phases = bw.split(user, item, seed=1)
    tu, ti = phases[0]
    used = build_csr(n_users, tu, ti, n_items, dev)
    sampler = DeviceSampler(n_items, used[0], used[1], mode="hash", seed=7)
    loader = DeviceTrainLoader(tu, ti, sampler, batch_size=4096, seed=3)
    seen = []
    for b in loader:
        assert b["user_id"].is_cuda and len(b) <= 4096
        key = b["user_id"] * n_items + b["neg_item_id"]
        assert not torch.isin(key, torch.as_tensor(tu * n_items + ti, device=dev)).any()   # never a train positive
        assert (b["neg_item_id"] >= 1).all() and (b["neg_item_id"] < n_items).all()
        seen.append((b["user_id"] * n_items + b["item_id"]).cpu().numpy())
    seen = np.sort(np.concatenate(seen))
    np.testing.assert_array_equal(seen, np.sort(tu * n_items + ti))                          # one full pass

    class DS:
        def num(self, f):
            return {"user_id": n_users, "item_id": n_items}[f]

    cfg = Cfg(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device=dev, embedding_size=64,
              learner="adam", learning_rate=5e-3, epochs=6, metrics=["Recall", "NDCG"], topk=[10],
              metric_decimal_place=4, valid_metric="recall@10", eval_step=6, scorer_mode="tc")
    model = FusedBPR(cfg, DS()).to(dev)
    trainer = FusedTrainer(cfg, model)
    index = EvalIndex.from_phase_pairs(n_users, n_items, phases, 2, dev)
    before = trainer.evaluate(index)
    trainer.fit(loader, valid_data=None, verbose=False, saved=False)
    losses = [trainer.train_loss_dict[e] for e in range(6)]
    assert losses[-1] < losses[0]
    after = trainer.evaluate(index)
    assert after["recall@10"] > before["recall@10"] and after["ndcg@10"] > before["ndcg@10"]
    # the two scorers agree exactly on the trained tables
    trainer.scorer_mode = "fp32"
    assert trainer.evaluate(index) == after
