import os, sys, torch
sys.path.insert(0, ".")
from b200rec import kernels as KR
N, Q, D, k = int(os.environ.get("NROWS", 10_000_000)), 4096, 128, 100
g = torch.Generator(device="cuda").manual_seed(1234)
cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
def run(tag, iters=10):
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
    print(f"{tag}: {ms:.3f} ms {Q/ms*1e3:.0f} QPS {2.0*Q*N*D/ms/1e9:.1f} TFLOP/s", flush=True)
cfgs = [("v1", {"B200REC_TOPK_V2": "0"}), ("v2 ks1", {"B200REC_TOPK_V2": "1", "B200REC_KS": "1"}),
        ("v2 ks2", {"B200REC_TOPK_V2": "1", "B200REC_KS": "2"}), ("v2s2 ks2", {"B200REC_TOPK_V2": "2", "B200REC_KS": "2"})]
for rep in range(2):
    for tag, env in cfgs:
        os.environ.update(env)
        run(tag)
