"""k = 1000 at small / large query batches: list hand-over threshold sweep (B200REC_TOPK_FLUSH).  NROWS, DIM via env."""
import os, sys, torch
sys.path.insert(0, ".")
from b200rec import kernels as KR
N, D, k = int(os.environ.get("NROWS", 50_000_000)), int(os.environ.get("DIM", 64)), int(os.environ.get("TOPK", 1000))
g = torch.Generator(device="cuda").manual_seed(1234)
cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
for Q in [int(x) for x in os.environ.get("QS", "1,16,128,1024").split(",")]:
    qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    ws = torch.empty(KR.topk_workspace_bytes(N, D, Q, k), dtype=torch.uint8, device="cuda")
    ref = None
    for fl in os.environ.get("FLUSHES", "-1,224,512,1024,2000").split(","):
        os.environ["B200REC_TOPK_FLUSH"] = fl
        KR.N.lib().b200rec_debug_reload_env()
        for _ in range(2): s, i = KR.flat_ip_topk(cat, qry, k, workspace=ws)
        torch.cuda.synchronize()
        if ref is None: ref = i.clone()
        same = bool((i == ref).all())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): KR.flat_ip_topk(cat, qry, k, workspace=ws)
        e1.record(); torch.cuda.synchronize()
        print(f"Q={Q:5d} k={k} flush={fl:>5s}: {e0.elapsed_time(e1)/5:8.3f} ms  same_ids={same}", flush=True)
