"""torchrun --nproc-per-node G tools/dp_breakdown.py : per-phase device time of the exact data-parallel training step
(SHAPE=cfg2 dense tables | cfg4 row-sparse tables), eager (every collective bracketed by CUDA events on the training
stream), next to the CUDA-graph replay time of the same step."""
import os, sys, time, collections, torch, torch.distributed as dist
sys.path.insert(0, ".")
from b200rec.dist import DataParallel
from b200rec.trainer import TwoTowerTrainer
from b200rec.training_utils import create_two_tower_model_for_training
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
shape = os.environ.get("SHAPE", "cfg2")
B, FD = 8192, 16
NU, NI, ed, hid, E, sparse = (1_000_000, 100_000, 64, [128, 64], 64, False) if shape == "cfg2" else (50_000_000, 5_000_000, 128, [256, 128], 128, True)
cfg = {"embedding_dim": E, "hidden_layers": hid, "dropout_rate": 0.2, "temperature": 0.05, "sparse_tables": sparse,
       "user_categorical_features": {"user_id": NU}, "item_categorical_features": {"item_id": NI}, "embedding_dims": {"user_id": ed, "item_id": ed}}
torch.manual_seed(1234)
with torch.device(dev):
    model = create_two_tower_model_for_training(FD, FD, cfg)
tr = TwoTowerTrainer(model, [], [], {"checkpoint_dir": f"/tmp/b200rec_dpb_{rank}"}, device=str(dev)); dp = DataParallel(model); model.train()
g = torch.Generator(device=dev).manual_seed(7 + rank)
uf = torch.randn(B, FD, device=dev, generator=g); pf = torch.randn(B, FD, device=dev, generator=g)
uid = torch.randint(1, NU + 1, (B,), device=dev, generator=g); iid = torch.randint(1, NI + 1, (B,), device=dev, generator=g)
step = lambda: tr.train_step(uf, pf, None, {"user_id": uid}, {"item_id": iid})
for _ in range(4): step()
spans = collections.defaultdict(list)
def timed(name, fn):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record(); spans[name].append((e0, e1)); return r
    return w
import b200rec.dist as D
dp.reduce_sums = timed("BatchNorm statistic all-reduces (fp64 sums)", dp.reduce_sums)
dp.gather_sparse = timed("touched-row all-gathers (ids + gradient rows)", dp.gather_sparse)
dp.reduce_dense_grad_ = timed("flat dense-gradient all-reduce", dp.reduce_dense_grad_)
dp.global_loss = timed("loss all-reduce", dp.global_loss)
fwd0, bwd0 = D._AllGatherRows.forward, D._AllGatherRows.backward
D._AllGatherRows.forward = staticmethod(timed("item-embedding all-gather", fwd0))
D._AllGatherRows.backward = staticmethod(timed("item-gradient reduce-scatter", bwd0))
n = 10
tot = []
for _ in range(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); e0.record(); step(); e1.record(); tot.append((e0, e1))
torch.cuda.synchronize()
total = sum(a.elapsed_time(b) for a, b in tot) / n
rows = {k: sum(a.elapsed_time(b) for a, b in v) / n for k, v in spans.items()}
counts = {k: len(v) / n for k, v in spans.items()}
t = torch.tensor([total] + [rows[k] for k in sorted(rows)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"{shape} exact data-parallel step, {world} GPUs x batch {B}, eager: {t[0].item():.3f} ms per step (max over ranks)")
    comm = 0.0
    for i, k in enumerate(sorted(rows)):
        comm += t[1 + i].item()
        print(f"  {k:48s} {t[1 + i].item() * 1e3:8.1f} us  ({counts[k]:.0f} per step)")
    print(f"  {'compute (everything else, incl. launch gaps)':48s} {(t[0].item() - comm) * 1e3:8.1f} us")
# CUDA-graph replay of the same step
D._AllGatherRows.forward, D._AllGatherRows.backward = staticmethod(fwd0), staticmethod(bwd0)
tr.enable_cuda_graph(warm_steps=1)
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"  as one CUDA graph: {ms.item():.3f} ms per step, {world * B / ms.item() * 1e3 / 1e6:.2f} M samples/s", flush=True)
tr.release_graphs(); del tr, model
import gc, threading; gc.collect(); torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush()
threading.Timer(30.0, lambda: os._exit(0)).start()
dist.destroy_process_group()
os._exit(0)
