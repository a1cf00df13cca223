"""One GPU emulating one of G row shards of the 10M catalogue: per-phase device time of the sharded search minus the
two all-gathers (development tool)."""
import os, sys, torch
sys.path.insert(0, ".")
from b200rec import kernels as K
from b200rec.dist import ShardedFlatIndex, shard_bounds
from b200rec.retrieval import FlatIPDeviceIndex
dev = torch.device("cuda", 0)
G, N, Q, D, k = int(os.environ.get("SHARDS", 8)), 10_000_000, 4096, 128, 100
g = torch.Generator(device=dev).manual_seed(7)
cat = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device=dev, generator=g), dim=1)
shards = []
for r in range(G):
    lo, hi = shard_bounds(N, G, r)
    ix = FlatIPDeviceIndex(D, storage="bf16", device=dev, row_offset=lo); ix.add_bf16_rows(cat[lo:hi]); shards.append(ix)
qo = shards[0].prepare_queries(qry, normalize=False)
kx = ShardedFlatIndex.exchange_width(k, G)
vals = torch.stack([ix.sample_device(qo, k, kx, G) for ix in shards]).contiguous()
ids = torch.arange(vals.numel(), device=dev).view_as(vals)
names = ["sample", "merge-tau", "slice", "local-search", "merge"]
acc = [0.0] * len(names)
outs = [ix.search_device(qo, k) for ix in shards]
gs = torch.stack([o[0] for o in outs]).contiguous(); gi = torch.stack([o[1] for o in outs]).contiguous()
ix = shards[0]
iters = 30
for it in range(iters + 5):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record()
    v = ix.sample_device(qo, k, kx, G); ev[1].record()
    top, _ = K.topk_merge(vals, ids, k); ev[2].record()
    tau = top[:, k - 1].contiguous(); ev[3].record()
    ix.search_device(qo, k, tau_init=tau); ev[4].record()
    K.topk_merge(gs, gi, k); ev[5].record()
    torch.cuda.synchronize()
    if it >= 5:
        for j in range(len(names)): acc[j] += ev[j].elapsed_time(ev[j + 1])
print(f"{G}-way shard ({shards[0].ntotal} rows), exchange width {kx}: " + ", ".join(f"{n} {a/iters*1e3:.0f}us" for n, a in zip(names, acc)) + f" | total {sum(acc)/iters:.3f} ms (+2 all-gathers)", flush=True)
import ctypes
from b200rec import _native as NV
lib = NV.lib()
os.environ["B200REC_TOPK_DEBUG"] = "2"
buf = (ctypes.c_ulonglong * 24)()
ix.search_device(qo, k, tau_init=tau); torch.cuda.synchronize()
lib.b200rec_debug_topk_stats16(buf, 1)
ix.search_device(qo, k, tau_init=tau); torch.cuda.synchronize()
lib.b200rec_debug_topk_stats16(buf, 0)
ends = (ctypes.c_ulonglong * 512)(); lib.b200rec_debug_topk_cta_end(ends)
start = (~buf[5]) & 0xFFFFFFFFFFFFFFFF
span = (buf[7] - start) / 1e3
tiles = max(buf[20], 1)
d = sorted((ends[i] - start) / 1e3 for i in range(0, 144, 2))
print(f"main kernel with pooled tau: span {span:.0f} us @ {buf[6]} MHz; unit end times us: min {d[0]:.0f} median {d[len(d)//2]:.0f} max {d[-1]:.0f}")
print(f"  appends/query {buf[0]/Q:.0f}, requests/query {buf[1]/Q:.1f}, entries/warp-tile {buf[9]/tiles:.3f}, cycles/entry {buf[8]/max(buf[9],1):.0f}, "
      f"MMA wait-acc {buf[16]/72/(tiles/(144*8)):.0f} cyc/tile, epilogue consume {buf[19]/tiles:.0f} cyc/tile, wait {buf[18]/tiles:.0f}")
