"""torchrun --nproc-per-node G tools/check_dp_training.py : G-replica data-parallel training (synced BatchNorm statistics,
all-gathered in-batch negatives, summed gradients) == one process on the concatenated global batch.
Dropout is 0 (masks are per replica by design).  Exits non-zero on mismatch."""
import copy, os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, ".")
from b200rec.dist import DataParallel
from b200rec.trainer import TwoTowerTrainer
from b200rec.training_utils import create_two_tower_model_for_training
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B, FD, NU, NI, R = int(os.environ.get("BLOCAL", 1024)), 16, 50_000, 5_000, 4
ok = True
for sparse, mixed in ((False, False), (True, False), (False, True)):
    if True:
        torch.manual_seed(1234)
        cfg = {"embedding_dim": 64, "hidden_layers": [128, 64], "dropout_rate": 0.0, "temperature": 0.05}
        if not mixed:  # the mixed explicit + in-batch step feeds numerical negatives only (trainers/two_tower.py:116-117)
            cfg.update({"user_categorical_features": {"user_id": NU}, "item_categorical_features": {"item_id": NI},
                        "embedding_dims": {"user_id": 64, "item_id": 64}})
        def make():
            torch.manual_seed(1234)
            m = create_two_tower_model_for_training(FD, FD, cfg)
            if sparse:
                for t in (m.user_tower, m.item_tower):
                    for e in t.embeddings.values():
                        e.weight._b200_sparse = True
            return m
        tcfg = {"learning_rate": 1e-3, "weight_decay": 0.0 if sparse else 1e-5, "checkpoint_dir": f"/tmp/b200rec_dp_{rank}"}
        dp_model = make(); trainer = TwoTowerTrainer(dp_model, [], [], tcfg, device=str(dev)); DataParallel(dp_model); dp_model.train()
        ref_model = make(); ref_trainer = TwoTowerTrainer(ref_model, [], [], tcfg, device=str(dev)); ref_model.train()
        trainer.optimizer.clear_grad = ref_trainer.optimizer.clear_grad = False   # the check below reads the gradients
        g = torch.Generator(device=dev).manual_seed(99)
        losses, ref_losses = [], []
        grad_rel = 0.0
        for step in range(4):
            uf = torch.randn(world * B, FD, device=dev, generator=g); pf = torch.randn(world * B, FD, device=dev, generator=g)
            nf = torch.randn(world * B, R, FD, device=dev, generator=g) if mixed else None
            uid = torch.randint(1, NU + 1, (world * B,), device=dev, generator=g); iid = torch.randint(1, NI + 1, (world * B,), device=dev, generator=g)
            sl = slice(rank * B, (rank + 1) * B)
            cu, ci = (None, None) if mixed else ({"user_id": uid[sl]}, {"item_id": iid[sl]})
            losses.append(trainer.train_step(uf[sl], pf[sl], nf[sl] if mixed else None, cu, ci))
            cu, ci = (None, None) if mixed else ({"user_id": uid}, {"item_id": iid})
            ref_losses.append(ref_trainer.train_step(uf, pf, nf, cu, ci))
            if step == 0:   # the summed DP gradient vs the single-process gradient, before Adam turns noise into steps
                gd, gr = trainer.optimizer.grad, ref_trainer.optimizer.grad
                grad_rel = ((gd - gr).norm() / gr.norm()).item()
        l = torch.stack([x.reshape(()) for x in losses]).cpu(); r = torch.stack([x.reshape(()) for x in ref_losses]).cpu()
        rel = ((l - r).abs() / r.abs()).max().item()
        pmax = 0.0
        for (n1, p1), (_, p2) in zip(dp_model.named_parameters(), ref_model.named_parameters()):
            pmax = max(pmax, (p1.detach() - p2.detach()).abs().max().item())
        bn = max((b1 - b2).abs().max().item() for (k1, b1), (_, b2) in zip(dp_model.named_buffers(), ref_model.named_buffers()) if b1.dtype.is_floating_point)
        # losses and first-step gradients agree to rounding (the two runs sum in different orders: chunking, split-K);
        # parameters are only bounded loosely: Adam turns 1e-9 gradient noise on near-zero gradients into lr-sized steps
        # (lr = 1e-3, 4 steps; weight gradients are summed with fp32 atomics in a run-dependent order)
        good = rel <= 2e-6 and grad_rel <= 1e-5 and pmax <= 2e-3 and bn <= 1e-5
        ok = ok and good
        if rank == 0:
            print(f"sparse_tables={sparse} mixed_loss={mixed}: losses DP {l.tolist()} vs single {r.tolist()} max rel diff {rel:.2e}; "
                  f"first-step gradient rel diff {grad_rel:.2e}; max |param diff| {pmax:.2e}; max |BN running-stat diff| {bn:.2e} -> {'OK' if good else 'MISMATCH'}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
# throughput of the DP step at the config-2 shape (per-GPU batch 8192)
B2, NU2, NI2 = 8192, 1_000_000, 100_000
torch.manual_seed(1234)
cfg2 = {"embedding_dim": 64, "hidden_layers": [128, 64], "dropout_rate": 0.2, "temperature": 0.05,
        "user_categorical_features": {"user_id": NU2}, "item_categorical_features": {"item_id": NI2},
        "embedding_dims": {"user_id": 64, "item_id": 64}}
m2 = create_two_tower_model_for_training(FD, FD, cfg2)
t2 = TwoTowerTrainer(m2, [], [], {"checkpoint_dir": f"/tmp/b200rec_dp_{rank}"}, device=str(dev)); DataParallel(m2); m2.train()
g = torch.Generator(device=dev).manual_seed(7 + rank)
uf = torch.randn(B2, FD, device=dev, generator=g); pf = torch.randn(B2, FD, device=dev, generator=g)
uid = torch.randint(1, NU2 + 1, (B2,), device=dev, generator=g); iid = torch.randint(1, NI2 + 1, (B2,), device=dev, generator=g)
for graphed in (False, True):
    if graphed: t2.enable_cuda_graph(warm_steps=1)
    for _ in range(5): t2.train_step(uf, pf, None, {"user_id": uid}, {"item_id": iid})
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(20): loss = t2.train_step(uf, pf, None, {"user_id": uid}, {"item_id": iid})
    torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / 20
    if rank == 0: print(f"DP config-2 step ({'CUDA graph' if graphed else 'eager'}), {world} GPUs x batch {B2} (global in-batch negatives {world*B2}): {dt*1e3:.2f} ms/step, {world*B2/dt:.0f} samples/s, loss {float(loss):.4f}", flush=True)
code = 0 if int(flag.item()) == 1 else 1
t2.release_graphs(); del t2, m2
import gc, threading; gc.collect(); torch.cuda.synchronize(); dist.barrier()
sys.stdout.flush()
threading.Timer(30.0, lambda: os._exit(code)).start()   # a wedged communicator teardown must not turn a pass into a hang
dist.destroy_process_group()
os._exit(code)
