"""The seeded batches of the 120-step trajectory golden: shared by the generator (make_golden_round2.py, reference on
torch-CPU) and the GPU test, so both sides see the same tensors (torch's CPU generator is deterministic)."""
import torch


def make_batches(seed: int, steps: int, B: int, R: int, user_dim: int, item_dim: int):
    gen = torch.Generator().manual_seed(seed + 1)
    # a fixed random "teacher" makes the batches learnable, so the loss trajectory moves instead of sitting at ln B
    tu, ti = torch.randn(user_dim, 8, generator=gen), torch.randn(item_dim, 8, generator=gen)
    out = []
    for _ in range(steps):
        uf = torch.randn(B, user_dim, generator=gen)
        cand = torch.randn(B, 1 + R + 3, item_dim, generator=gen)
        score = torch.einsum("bk,bck->bc", uf @ tu, cand @ ti)
        order = score.argsort(dim=1, descending=True)
        rows = torch.arange(B)
        pos = cand[rows, order[:, 0]]
        neg = torch.stack([cand[rows, order[:, -1 - j]] for j in range(R)], dim=1)
        out.append({"user_features": uf, "pos_item_features": pos, "neg_item_features": neg})
    return out
