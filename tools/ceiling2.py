import os, sys, torch
sys.path.insert(0, ".")
from b200rec import kernels as KR
os.environ["B200REC_TOPK_NOSAMPLE"] = "1"
os.environ["B200REC_TOPK_DEBUG"] = "3"
def run(N, Q, D, k, iters=10, tag=""):
    g = torch.Generator(device="cuda").manual_seed(1234)
    cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    ws = torch.empty(KR.topk_workspace_bytes(N, D, Q, k), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        KR.flat_ip_topk(cat, qry, k, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        KR.flat_ip_topk(cat, qry, k, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{tag} N={N} Q={Q} D={D} k={k}: {ms:.3f} ms {2.0*Q*N*D/ms/1e9:.1f} TFLOP/s  L2->SM {N*D*2*((Q+255)//256)/ms/1e6:.0f} GB/s", flush=True)
for N in (100_000, 200_000, 400_000, 1_000_000, 5_000_000):
    run(N, 4096, 128, 100, tag="pure-mma")
run(5_000_000, 8192, 128, 100, tag="pure-mma")
run(5_000_000, 1024, 128, 100, tag="pure-mma")
