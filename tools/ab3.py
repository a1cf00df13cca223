"""Interleaved A/B timing of library variants / env knobs, one subprocess per measurement (development tool).
CONFIGS="name:ENV=V,ENV=V;..." (B200REC_LIB=variants/x.so selects a variant build)  ROUNDS=3 ITERS=30"""
import os, subprocess, sys, statistics
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ".")
    from b200rec import kernels as KR
    N, Q, D, k = int(os.environ.get("NROWS", 10_000_000)), int(os.environ.get("NQ", 4096)), int(os.environ.get("DIM", 128)), int(os.environ.get("TOPK", 100))
    g = torch.Generator(device="cuda").manual_seed(1234)
    cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    ws = torch.empty(KR.topk_workspace_bytes(N, D, Q, k), dtype=torch.uint8, device="cuda")
    iters = int(os.environ.get("ITERS", 30))
    for _ in range(5):
        KR.flat_ip_topk(cat, qry, k, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        KR.flat_ip_topk(cat, qry, k, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    print("MS", e0.elapsed_time(e1) / iters)
    sys.exit(0)
cfgs = []
for c in os.environ.get("CONFIGS", "base:").split(";"):
    name, _, envs = c.partition(":")
    cfgs.append((name, dict(e.split("=", 1) for e in envs.split(",") if e)))
res = {n: [] for n, _ in cfgs}
for r in range(int(os.environ.get("ROUNDS", 3))):
    for name, env in cfgs:
        e = dict(os.environ); e.update(env)
        if "B200REC_LIB" in e:
            e["B200REC_LIB"] = os.path.abspath(e["B200REC_LIB"])
        out = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
        ms = [float(l.split()[1]) for l in out.stdout.splitlines() if l.startswith("MS")]
        res[name].append(ms[0] if ms else float("nan"))
        if not ms:
            print(name, "FAILED", out.stderr[-400:], flush=True)
for name, _ in cfgs:
    v = res[name]
    print(f"{name:26s} " + " ".join(f"{x:7.3f}" for x in v) + f" | median {statistics.median(v):7.3f} ms  {4096/statistics.median(v)*1e3:8.0f} QPS", flush=True)
