"""Interleaved A/B timing of flat_ip_topk variants on one box (development tool).
CONFIGS="name:ENV=V,ENV=V;name2:..."  ROUNDS=3 ITERS=20"""
import os, sys, ctypes, torch, statistics
sys.path.insert(0, ".")
from b200rec import kernels as KR, _native as NV
lib = NV.lib()
N, Q, D, k = int(os.environ.get("NROWS", 10_000_000)), int(os.environ.get("NQ", 4096)), int(os.environ.get("DIM", 128)), int(os.environ.get("TOPK", 100))
g = torch.Generator(device="cuda").manual_seed(1234)
cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
ws = torch.empty(KR.topk_workspace_bytes(N, D, Q, k) * 2, dtype=torch.uint8, device="cuda")
cfgs = []
for c in os.environ.get("CONFIGS", "base:").split(";"):
    name, _, envs = c.partition(":")
    cfgs.append((name, dict(e.split("=") for e in envs.split(",") if e)))
keys = sorted({k_ for _, e in cfgs for k_ in e})
iters, rounds = int(os.environ.get("ITERS", 20)), int(os.environ.get("ROUNDS", 3))
res = {n: [] for n, _ in cfgs}
for r in range(rounds):
    for name, env in cfgs:
        for k_ in keys:
            os.environ.pop(k_, None)
        os.environ.update(env)
        for _ in range(3):
            KR.flat_ip_topk(cat, qry, k, workspace=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            KR.flat_ip_topk(cat, qry, k, workspace=ws)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        buf = (ctypes.c_ulonglong * 24)()
        lib.b200rec_debug_topk_stats16(buf, 1)
        KR.flat_ip_topk(cat, qry, k, workspace=ws); torch.cuda.synchronize()
        lib.b200rec_debug_topk_stats16(buf, 1)
        span = (buf[7] - ((~buf[5]) & 0xFFFFFFFFFFFFFFFF)) / 1e6
        res[name].append((ms, span, buf[6]))
for name, _ in cfgs:
    v = res[name]
    print(f"{name:28s} call ms " + " ".join(f"{x[0]:7.3f}" for x in v) + f" | median {statistics.median(x[0] for x in v):7.3f} | main-kernel ms (1 call) "
          + " ".join(f"{x[1]:6.2f}@{x[2]}" for x in v), flush=True)
