import os, sys, torch
sys.path.insert(0, ".")
from b200rec import kernels as KR
N, Q, D, k = int(os.environ.get("NROWS", 1_250_000)), 4096, 128, 100
g = torch.Generator(device="cuda").manual_seed(1234)
cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
ws = torch.empty(KR.topk_workspace_bytes(N, D, Q, k), dtype=torch.uint8, device="cuda")
for _ in range(4): KR.flat_ip_topk(cat, qry, k, workspace=ws)
torch.cuda.synchronize()
