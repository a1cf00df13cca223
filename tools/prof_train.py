"""A few training steps for ncu captures: SHAPE=cfg2 (1M users x 100K items, dim 64, batch 8192, in-batch negatives),
SHAPE=cfg4 (50M x 5M-row tables, dim 128, row-sparse Adam) or
SHAPE=ml1m (batch 1024 x 16 explicit negatives, E 128, hidden [256,128], mixed loss)."""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from b200rec.trainer import TwoTowerTrainer
from b200rec.training_utils import create_two_tower_model_for_training
torch.manual_seed(1234)
rng = np.random.default_rng(1234)
if os.environ.get("SHAPE", "cfg2") in ("cfg2", "cfg4"):
    big = os.environ.get("SHAPE") == "cfg4"
    B, NU, NI, FD = (8192, 50_000_000, 5_000_000, 16) if big else (8192, 1_000_000, 100_000, 16)
    ed = 128 if big else 64
    cfg = {"embedding_dim": ed, "hidden_layers": [256, 128] if big else [128, 64], "dropout_rate": 0.2, "temperature": 0.05,
           "user_categorical_features": {"user_id": NU}, "item_categorical_features": {"item_id": NI},
           "embedding_dims": {"user_id": ed, "item_id": ed}, "sparse_tables": big}
    with torch.device("cuda"):
        model = create_two_tower_model_for_training(FD, FD, cfg)
    z = lambda n, hi: torch.from_numpy(np.clip(rng.zipf(1.05, size=n), 1, hi).astype(np.int64)).cuda()
    args = (torch.randn(B, FD).cuda(), torch.randn(B, FD).cuda(), None, {"user_id": z(B, NU)}, {"item_id": z(B, NI)})
else:
    B, R, FU, FI = 1024, 16, 3, 20
    model = create_two_tower_model_for_training(FU, FI, {"embedding_dim": 128, "hidden_layers": [256, 128], "dropout_rate": 0.2})
    args = (torch.randn(B, FU).cuda(), torch.randn(B, FI).cuda(), torch.randn(B, R, FI).cuda())
tr = TwoTowerTrainer(model, [], [], {"checkpoint_dir": "/tmp/b200rec_prof"}, device="cuda")
model.train()
for _ in range(3):
    loss = tr.train_step(*args)
torch.cuda.synchronize()
print("ok", float(loss))
