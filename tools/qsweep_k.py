import os, sys, torch
sys.path.insert(0, ".")
from b200rec import kernels as KR
N, D = int(os.environ.get("NROWS", 10_000_000)), int(os.environ.get("DIM", 128))
g = torch.Generator(device="cuda").manual_seed(1234)
cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
for Q, k in [(1, 1000), (1, 100), (16, 1000), (128, 1000), (128, 100), (1024, 1000), (1024, 100)]:
    qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    ws = torch.empty(KR.topk_workspace_bytes(N, D, Q, k), dtype=torch.uint8, device="cuda")
    for _ in range(3): KR.flat_ip_topk(cat, qry, k, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): KR.flat_ip_topk(cat, qry, k, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"Q={Q:5d} k={k:4d}: {ms:8.3f} ms  {Q/ms*1e3:10.0f} QPS  catalogue stream {N*D*2/ms/1e6:7.0f} GB/s  {2.0*Q*N*D/ms/1e9:7.1f} TFLOP/s", flush=True)
