import os, sys, torch
sys.path.insert(0, ".")
from b200rec import kernels as KR
N, Q, D, k = int(os.environ.get("NROWS", 2_000_000)), int(os.environ.get("NQRY", 4096)), int(os.environ.get("DIM", 128)), int(os.environ.get("TOPK", 100))
g = torch.Generator(device="cuda").manual_seed(1234)
cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
ws = torch.empty(KR.topk_workspace_bytes(N, D, Q, k), dtype=torch.uint8, device="cuda")
for _ in range(3):
    s, i = KR.flat_ip_topk(cat, qry, k, workspace=ws)
torch.cuda.synchronize()
print("ok", s[0, :3].tolist())
