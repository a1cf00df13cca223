"""Device time of the in-batch loss forward (fused LSE) and backward (fused gradient kernel) at the config-2 / ML-1M shapes."""
import sys, torch
sys.path.insert(0, ".")
from b200rec import ops
for B, G, E in ((8192, 1, 64), (8192, 1, 128), (1024, 1, 128), (8192, 8, 64)):
    u = torch.nn.functional.normalize(torch.randn(B, E, device="cuda"), dim=1).requires_grad_()
    v = torch.nn.functional.normalize(torch.randn(G * B, E, device="cuda"), dim=1).requires_grad_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for it in range(8):
        u.grad = v.grad = None
        ev[0].record()
        loss = ops.InBatchCEFn.apply(u, v, 20.0, 6, 0, G * B)
        ev[1].record()
        loss.backward()
        ev[2].record()
        torch.cuda.synchronize()
        if it >= 3:
            tf += ev[0].elapsed_time(ev[1]) / 5
            tb += ev[1].elapsed_time(ev[2]) / 5
    print(f"B {B} NI {G * B} E {E}: forward {tf * 1e3:.1f} us  backward {tb * 1e3:.1f} us", flush=True)
