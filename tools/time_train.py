"""ms per training step (CUDA-graph replay, CUDA events) for SHAPE=cfg2 | ml1m | cfg4 — quick A/B of step-level changes."""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from b200rec.trainer import TwoTowerTrainer
from b200rec.training_utils import create_two_tower_model_for_training
torch.manual_seed(1234)
rng = np.random.default_rng(1234)
for shape in os.environ.get("SHAPES", "cfg2,ml1m").split(","):
    if shape == "cfg2":
        B, NU, NI, FD = 8192, 1_000_000, 100_000, 16
        cfg = {"embedding_dim": 64, "hidden_layers": [128, 64], "dropout_rate": 0.2, "temperature": 0.05,
               "user_categorical_features": {"user_id": NU}, "item_categorical_features": {"item_id": NI},
               "embedding_dims": {"user_id": 64, "item_id": 64}}
        model = create_two_tower_model_for_training(FD, FD, cfg)
        z = lambda n, hi: torch.from_numpy(np.clip(rng.zipf(1.05, size=n), 1, hi).astype(np.int64)).cuda()
        args = (torch.randn(B, FD).cuda(), torch.randn(B, FD).cuda(), None, {"user_id": z(B, NU)}, {"item_id": z(B, NI)})
    else:
        B, R, FU, FI = 1024, 16, 3, 20
        model = create_two_tower_model_for_training(FU, FI, {"embedding_dim": 128, "hidden_layers": [256, 128], "dropout_rate": 0.2})
        args = (torch.randn(B, FU).cuda(), torch.randn(B, FI).cuda(), torch.randn(B, R, FI).cuda())
    tr = TwoTowerTrainer(model, [], [], {"checkpoint_dir": "/tmp/b200rec_prof"}, device="cuda")
    model.train()
    tr.enable_cuda_graph()
    for _ in range(6):
        loss = tr.train_step(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        loss = tr.train_step(*args)
    e1.record()
    torch.cuda.synchronize()
    print(f"{shape}: {e0.elapsed_time(e1) / n:.4f} ms/step  loss {float(loss):.4f}", flush=True)
    del tr, model
