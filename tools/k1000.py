import os, sys, torch
sys.path.insert(0, ".")
from b200rec import kernels as KR
N, D, Q, k = 10_000_000, int(os.environ.get("DIM", 128)), 1024, 1000
g = torch.Generator(device="cuda").manual_seed(1234)
cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
ws = torch.empty(KR.topk_workspace_bytes(N, D, Q, k), dtype=torch.uint8, device="cuda")
for fl in os.environ.get("FLUSHES", "32,64,128,256,512").split(","):
    os.environ["B200REC_TOPK_FLUSH"] = fl
    KR.N.lib().b200rec_debug_reload_env()   # knobs are read once per process otherwise
    for _ in range(2): KR.flat_ip_topk(cat, qry, k, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): KR.flat_ip_topk(cat, qry, k, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    print(f"k={k} flush={fl}: {e0.elapsed_time(e1)/5:.3f} ms", flush=True)
