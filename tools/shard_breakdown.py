"""torchrun --nproc-per-node G tools/shard_breakdown.py : per-phase device time of one sharded search (development tool)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
from b200rec import kernels as K
from b200rec.dist import ShardedFlatIndex, shard_bounds
from b200rec.retrieval import FlatIPDeviceIndex
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N, Q, D, k = int(os.environ.get("NROWS", 10_000_000)), int(os.environ.get("NQ", 4096)), 128, 100
lo, hi = shard_bounds(N, world, rank)
g = torch.Generator(device=dev).manual_seed(7 + rank)
cat = torch.nn.functional.normalize(torch.randn(hi - lo, D, device=dev, generator=g), dim=1).to(torch.bfloat16)
g2 = torch.Generator(device=dev).manual_seed(99)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device=dev, generator=g2), dim=1)
ix = FlatIPDeviceIndex(D, storage="bf16", device=dev, row_offset=lo); ix.add_bf16_rows(cat)
sh = ShardedFlatIndex.from_device_index(ix)
qo = ix.prepare_queries(qry, normalize=False)
for _ in range(5): sh.search(qo, k)
torch.cuda.synchronize(); dist.barrier()
names = ["sample", "allgather-tau", "merge-tau", "slice", "local-search", "allgather", "merge"]
acc = [0.0] * len(names)
ids = torch.arange(world * Q * k, dtype=torch.int64, device=dev).view(world, Q, k)
mine = torch.empty((12 * Q * k,), dtype=torch.uint8, device=dev); everyone = torch.empty((world * 12 * Q * k,), dtype=torch.uint8, device=dev)
iters = 20
for it in range(iters):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record()
    vals = ix.sample_device(qo, k); ev[1].record()
    allv = torch.empty((world * Q, k), dtype=vals.dtype, device=dev); dist.all_gather_into_tensor(allv, vals); ev[2].record()
    top, _ = K.topk_merge(allv.view(world, Q, k), ids, k); ev[3].record()
    tau = top[:, k - 1].contiguous(); ev[4].record()
    s_view = mine[: 4 * Q * k].view(torch.float32).view(Q, k); i_view = mine[4 * Q * k:].view(torch.int64).view(Q, k)
    ix.search_device(qo, k, tau_init=tau, out=(s_view, i_view)); ev[5].record()
    dist.all_gather_into_tensor(everyone, mine); ev[6].record()
    ch = everyone.view(world, 12 * Q * k)
    K.topk_merge(ch[:, : 4 * Q * k].view(torch.float32).view(world, Q, k), ch[:, 4 * Q * k:].view(torch.int64).view(world, Q, k), k); ev[7].record()
    torch.cuda.synchronize()
    for j in range(len(names)): acc[j] += ev[j].elapsed_time(ev[j + 1])
if rank == 0:
    print(f"world {world}, N {N} ({hi-lo} rows/GPU): " + ", ".join(f"{n} {a/iters*1e3:.0f}us" for n, a in zip(names, acc)) + f" | total {sum(acc)/iters:.3f} ms", flush=True)
dist.destroy_process_group()
