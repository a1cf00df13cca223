"""torchrun --nproc-per-node G tools/check_sharded.py : sharded top-k (NCCL) == single-GPU top-k on the same rows."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
from b200rec.dist import ShardedFlatIndex, shard_bounds
from b200rec.retrieval import FlatIPDeviceIndex
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N, Q, D, k = int(os.environ.get("NROWS", 2_000_000)), int(os.environ.get("NQ", 4096)), 128, 100
g = torch.Generator(device=dev).manual_seed(7)
cat = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device=dev, generator=g), dim=1)
lo, hi = shard_bounds(N, world, rank)
ix = FlatIPDeviceIndex(D, storage="bf16", device=dev, row_offset=lo); ix.add_bf16_rows(cat[lo:hi].contiguous())
sh = ShardedFlatIndex.from_device_index(ix, peer_exchange=os.environ.get("PEER", "1") == "1")
qo = ix.prepare_queries(qry, normalize=False)
s, i = sh.search(qo, k); torch.cuda.synchronize()
full = FlatIPDeviceIndex(D, storage="bf16", device=dev); full.add_bf16_rows(cat)
s1, i1 = full.search_device(full.prepare_queries(qry, normalize=False), k); torch.cuda.synchronize()
ok = bool(torch.equal(i, i1) and torch.equal(s, s1))
mism = (i != i1).float().mean().item()
print(f"rank {rank}/{world}: sharded == single: {ok} (id mismatch fraction {mism:.2e})", flush=True)
import time
for _ in range(5): sh.search(qo, k)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for _ in range(20): sh.search(qo, k)
torch.cuda.synchronize(); dist.barrier()
if rank == 0: print(f"peer exchange active: {any(k_[0] == 'peer' and v is not None for k_, v in sh._ids_cache.items() if isinstance(k_, tuple))}")
if rank == 0: print(f"sharded search N={N}: {(time.perf_counter()-t0)/20*1e3:.3f} ms/batch", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
