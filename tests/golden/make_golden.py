"""Generate golden vectors by RUNNING THE IMPORTED REFERENCE (torch CPU) on seeded inputs.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The GPU box never reads /root/reference; it reads these files.

Reference entry points exercised (unmodified):
  src/models/two_tower.py  UserTower/ItemTower/TwoTowerModel (:12,:137,:284)
  src/training/utils.py    create_two_tower_model_for_training (:14)
  the trainer step body    src/training/trainers/two_tower.py:98-146 (restated inline: same calls)
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)
from src.models.two_tower import UserTower, ItemTower, TwoTowerModel  # noqa: E402
from src.training.utils import create_two_tower_model_for_training  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sd_np(prefix, module):
    return {f"{prefix}.{k}": v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def grads_np(prefix, module):
    return {f"grad.{prefix}.{k}": p.grad.detach().cpu().numpy().copy()
            for k, p in module.named_parameters() if p.grad is not None}


def randomize_bn(model, gen):
    # non-trivial gamma/beta so BN parity is actually exercised
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            with torch.no_grad():
                m.weight.copy_(1.0 + 0.2 * torch.randn(m.weight.shape, generator=gen))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=gen))
        if isinstance(m, torch.nn.Linear):
            with torch.no_grad():
                m.bias.copy_(0.05 * torch.randn(m.bias.shape, generator=gen))


def case_trainer_step(name, user_dim, item_dim, cfg, B, R, seed):
    """One mixed-loss step exactly as trainers/two_tower.py:98-141 (dropout 0 so it is deterministic)."""
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    model = create_two_tower_model_for_training(user_dim, item_dim, cfg)
    randomize_bn(model, gen)
    with torch.no_grad():
        model.user_bias.fill_(0.3)
        model.item_bias.fill_(-0.2)
    model.train()
    uf = torch.randn(B, user_dim, generator=gen)
    pf = torch.randn(B, item_dim, generator=gen)
    nf = torch.randn(B, R, item_dim, generator=gen)
    out = {"user_features": uf.numpy(), "pos_item_features": pf.numpy(), "neg_item_features": nf.numpy(),
           "temperature": np.float32(model.temperature)}
    out.update(sd_np("user", model.user_tower))
    out.update(sd_np("item", model.item_tower))
    out["user_bias"] = model.user_bias.detach().numpy().copy()
    out["item_bias"] = model.item_bias.detach().numpy().copy()

    u = model.get_user_embeddings({"numerical": uf, "categorical": {}})
    p = model.get_item_embeddings({"numerical": pf, "categorical": {}})
    n = model.get_item_embeddings({"numerical": nf.view(-1, item_dim), "categorical": {}})
    u.retain_grad(); p.retain_grad(); n.retain_grad()
    explicit = model.contrastive_loss(u, p, n)
    inbatch = model.in_batch_negative_loss(u, p)
    loss = 0.7 * explicit + 0.3 * inbatch
    loss.backward()
    out.update({"user_emb": u.detach().numpy(), "pos_emb": p.detach().numpy(), "neg_emb": n.detach().numpy(),
                "explicit_loss": np.float32(explicit.item()), "inbatch_loss": np.float32(inbatch.item()),
                "loss": np.float32(loss.item()),
                "grad.user_emb": u.grad.numpy(), "grad.pos_emb": p.grad.numpy(), "grad.neg_emb": n.grad.numpy(),
                "grad.user_bias": model.user_bias.grad.numpy(), "grad.item_bias": model.item_bias.grad.numpy()})
    out.update(grads_np("user", model.user_tower))
    out.update(grads_np("item", model.item_tower))
    # post-step running statistics (item tower saw two batches: two_tower.py BN momentum 0.1 twice)
    out.update({f"after.{k}": v for k, v in sd_np("user", model.user_tower).items() if "running" in k or "num_batches" in k})
    out.update({f"after.{k}": v for k, v in sd_np("item", model.item_tower).items() if "running" in k or "num_batches" in k})
    # eval-mode forward with those running stats
    model.eval()
    with torch.no_grad():
        out["user_emb_eval"] = model.get_user_embeddings({"numerical": uf, "categorical": {}}).numpy()
        out["pos_emb_eval"] = model.get_item_embeddings({"numerical": pf, "categorical": {}}).numpy()
        out["inbatch_loss_eval"] = np.float32(model.in_batch_negative_loss(
            torch.from_numpy(out["user_emb_eval"]), torch.from_numpy(out["pos_emb_eval"])).item())
    np.savez(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", loss.item(), "explicit", explicit.item(), "inbatch", inbatch.item())


def case_categorical(name, activation, seed):
    """Towers with embedding tables (two_tower.py:44-51,113-126,174-181,254-273) + in-batch loss backward."""
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    ucat = {"category": 10, "subcategory": 5}
    icat = {"genre": 30, "studio": 200}
    ut = UserTower(input_dim=10, embedding_dim=32, hidden_layers=[64, 32], dropout_rate=0.0,
                   activation=activation, categorical_features=ucat)
    it = ItemTower(input_dim=15, embedding_dim=32, hidden_layers=[64, 32], dropout_rate=0.0,
                   activation=activation, categorical_features=icat, use_content_embedding=False)
    model = TwoTowerModel(ut, it, temperature=0.1)
    randomize_bn(model, gen)
    model.train()
    B = 24
    un = torch.randn(B, 10, generator=gen)
    inn = torch.randn(B, 15, generator=gen)
    uc = {"category": torch.randint(0, 11, (B,), generator=gen), "subcategory": torch.randint(0, 6, (B,), generator=gen)}
    ic = {"genre": torch.randint(0, 31, (B,), generator=gen), "studio": torch.randint(0, 201, (B,), generator=gen)}
    uc["category"][0] = 0  # padding index present
    ic["studio"][1] = 0
    ic["studio"][2] = ic["studio"][3]  # duplicate row
    res = model({"numerical": un, "categorical": uc}, {"numerical": inn, "categorical": ic}, compute_loss=True)
    res["loss"].backward()
    out = {"user_num": un.numpy(), "item_num": inn.numpy(), "temperature": np.float32(0.1),
           "activation": np.array(activation)}
    for k, v in uc.items():
        out[f"user_cat.{k}"] = v.numpy()
    for k, v in ic.items():
        out[f"item_cat.{k}"] = v.numpy()
    out.update(sd_np("user", ut)); out.update(sd_np("item", it))
    out.update(grads_np("user", ut)); out.update(grads_np("item", it))
    out.update({"user_emb": res["user_embedding"].detach().numpy(), "item_emb": res["item_embedding"].detach().numpy(),
                "similarity": res["similarity"].detach().numpy(), "loss": np.float32(res["loss"].item())})
    np.savez(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", res["loss"].item())


def case_kat():
    """Closed-form known answers evaluated by the reference code."""
    ut = UserTower(input_dim=4, embedding_dim=8, hidden_layers=[8])
    it = ItemTower(input_dim=4, embedding_dim=8, hidden_layers=[8], use_content_embedding=False)
    m = TwoTowerModel(ut, it, temperature=0.1)
    U = torch.eye(4, 8)
    out = {"inbatch": np.float32(m.in_batch_negative_loss(U, U).item())}
    neg = torch.zeros(8, 8)
    for b in range(4):
        neg[2 * b, (b + 1) % 8 + 0] = 1.0 if (b + 1) % 8 != b else 0.0
        neg[2 * b + 1, (b + 5) % 8] = 1.0
    out["explicit"] = np.float32(m.contrastive_loss(U, U, neg).item())
    with torch.no_grad():
        m.user_bias.fill_(0.5); m.item_bias.fill_(0.25)
    out["explicit_bias"] = np.float32(m.contrastive_loss(U, U, neg).item())
    out["neg"] = neg.numpy()
    np.savez(os.path.join(OUT, "kat_losses.npz"), **out)
    print("kat", out["inbatch"], out["explicit"], out["explicit_bias"])


if __name__ == "__main__":
    torch.set_num_threads(1)
    case_trainer_step("step_ml1m", 3, 20, {"embedding_dim": 128, "hidden_layers": [256, 128], "dropout_rate": 0.0,
                                           "temperature": 0.05}, B=48, R=4, seed=11)
    case_trainer_step("step_small", 6, 9, {"embedding_dim": 64, "hidden_layers": [128, 64], "dropout_rate": 0.0,
                                           "temperature": 0.1}, B=40, R=16, seed=12)
    for act in ["relu", "gelu", "leaky_relu", "tanh", "sigmoid"]:
        case_categorical(f"cat_{act}", act, seed=20)
    case_kat()
