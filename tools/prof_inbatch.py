"""One forward + backward of the in-batch loss for ncu captures:  B=8192 G=1 E=64 python tools/prof_inbatch.py"""
import os, sys, torch
sys.path.insert(0, ".")
from b200rec import ops
B, G, E = int(os.environ.get("B", 8192)), int(os.environ.get("G", 1)), int(os.environ.get("E", 64))
u = torch.nn.functional.normalize(torch.randn(B, E, device="cuda"), dim=1).requires_grad_()
v = torch.nn.functional.normalize(torch.randn(G * B, E, device="cuda"), dim=1).requires_grad_()
for it in range(3):
    u.grad = v.grad = None
    loss = ops.InBatchCEFn.apply(u, v, 20.0, 6, 0, G * B)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
