"""A few config-2 training steps (1M users x 100K items, dim 64, batch 8192, in-batch negatives) for ncu captures."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from b200rec.trainer import TwoTowerTrainer
from b200rec.training_utils import create_two_tower_model_for_training
B, NU, NI, FD = 8192, 1_000_000, 100_000, 16
torch.manual_seed(1234)
cfg = {"embedding_dim": 64, "hidden_layers": [128, 64], "dropout_rate": 0.2, "temperature": 0.05,
       "user_categorical_features": {"user_id": NU}, "item_categorical_features": {"item_id": NI},
       "embedding_dims": {"user_id": 64, "item_id": 64}}
model = create_two_tower_model_for_training(FD, FD, cfg)
tr = TwoTowerTrainer(model, [], [], {"checkpoint_dir": "/tmp/b200rec_prof"}, device="cuda")
model.train()
rng = np.random.default_rng(1234)
z = lambda n, hi: torch.from_numpy(np.clip(rng.zipf(1.05, size=n), 1, hi).astype(np.int64)).cuda()
uf, pf, uid, iid = torch.randn(B, FD).cuda(), torch.randn(B, FD).cuda(), z(B, NU), z(B, NI)
for _ in range(3):
    loss = tr.train_step(uf, pf, None, {"user_id": uid}, {"item_id": iid})
torch.cuda.synchronize()
print("ok", float(loss))
