"""Round-2 golden vectors from the IMPORTED REFERENCE (torch CPU).  Run in the build container only:
    python tests/golden/make_golden_round2.py
  content_branch.npz : ItemTower(use_content_embedding=True) — the content_projection branch
                       (src/models/two_tower.py:184-191,264-266) forward (train-mode BN statistics, dropout 0, and
                       eval mode) and the gradient of sum(emb * w) w.r.t. every parameter.
  trajectory_120.npz : 120 steps of the reference's OWN TwoTowerTrainer.train_epoch (src/training/trainers/
                       two_tower.py:84-156: mixed 0.7 explicit + 0.3 in-batch loss, clip 1.0, Adam lr 1e-3 wd 1e-5) from a
                       stored initial state on batches that the test regenerates from the stored seed (torch CPU
                       generator): per-step losses (">= 100 steps from the same init").
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)
from src.models.two_tower import ItemTower  # noqa: E402
from src.training.utils import create_two_tower_model_for_training  # noqa: E402
from src.training.trainers.two_tower import TwoTowerTrainer  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
from trajectory_batches import make_batches  # noqa: E402


def content_branch(seed=41):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    it = ItemTower(input_dim=12, embedding_dim=32, hidden_layers=[64, 48], dropout_rate=0.0, activation="relu",
                   categorical_features={"genre": 20}, use_content_embedding=True, content_embedding_dim=64)
    for m in it.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            with torch.no_grad():
                m.weight.copy_(1.0 + 0.2 * torch.randn(m.weight.shape, generator=gen))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=gen))
    B = 96
    num = torch.randn(B, 12, generator=gen)
    cat = torch.randint(0, 21, (B,), generator=gen)
    content = torch.randn(B, 64, generator=gen)
    w = torch.randn(B, 32, generator=gen)
    out = {"numerical": num.numpy(), "genre": cat.numpy(), "content": content.numpy(), "w": w.numpy()}
    out.update({f"sd.{k}": v.detach().numpy().copy() for k, v in it.state_dict().items()})
    it.train()
    emb = it(num, {"genre": cat}, content)
    (emb * w).sum().backward()
    out["emb_train"] = emb.detach().numpy()
    out.update({f"grad.{k}": p.grad.detach().numpy().copy() for k, p in it.named_parameters() if p.grad is not None})
    it.eval()
    with torch.no_grad():
        out["emb_eval"] = it(num, {"genre": cat}, content).numpy()
    np.savez(os.path.join(OUT, "content_branch.npz"), **out)
    print("content_branch", out["emb_train"].shape, sorted(k for k in out if k.startswith("grad."))[:4])


def trajectory(steps=120, seed=77, B=256, R=4, user_dim=3, item_dim=20):
    cfg = {"embedding_dim": 64, "hidden_layers": [128, 64], "dropout_rate": 0.0, "temperature": 0.05}
    torch.manual_seed(seed)
    model = create_two_tower_model_for_training(user_dim, item_dim, cfg)
    init = {f"sd.{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
    batches = make_batches(seed, steps, B, R, user_dim, item_dim)
    losses = []

    class Loader(list):
        pass

    tr = TwoTowerTrainer(model, Loader(batches), [], {"learning_rate": 1e-3, "weight_decay": 1e-5,
                                                      "checkpoint_dir": "/tmp/b200rec_golden_ckpt"}, device="cpu")
    # train_epoch only returns the epoch mean: record every step's loss through the model's own loss methods
    orig_c, orig_i = model.contrastive_loss, model.in_batch_negative_loss
    last = {}

    def c(u, p, n):
        last["c"] = orig_c(u, p, n)
        return last["c"]

    def i(u, p):
        last["i"] = orig_i(u, p)
        losses.append(float((0.7 * last["c"] + 0.3 * last["i"]).item()))
        return last["i"]

    model.contrastive_loss, model.in_batch_negative_loss = c, i
    mean = tr.train_epoch(1)
    out = dict(init)
    out.update({"batch_checksum": np.asarray([float(b["user_features"].double().sum() + b["pos_item_features"].double().sum()
                                                    + b["neg_item_features"].double().sum()) for b in batches]),
                "losses": np.asarray(losses, dtype=np.float64), "epoch_mean": np.float64(mean), "seed": np.int64(seed),
                "B": np.int64(B), "R": np.int64(R), "steps": np.int64(steps), "user_dim": np.int64(user_dim),
                "item_dim": np.int64(item_dim)})
    out.update({f"final.{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()})
    np.savez_compressed(os.path.join(OUT, "trajectory_120.npz"), **out)
    print("trajectory", len(losses), losses[0], losses[9], losses[-1], "mean", mean)


if __name__ == "__main__":
    content_branch()
    trajectory()
