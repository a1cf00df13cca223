"""Warm (L2-resident, CUDA-graph replayed) time of one tower forward / forward+backward, fused vs unfused kernels."""
import os, sys, torch
sys.path.insert(0, ".")
from b200rec.two_tower import UserTower

def timeit(fn, n=30):
    g = torch.cuda.CUDAGraph()
    fn(); fn(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for name, B, K0, hidden, E in (("cfg2 tower", 8192, 80, [128, 64], 64), ("ml1m user", 1024, 3, [256, 128], 128),
                               ("ml1m negatives", 16384, 20, [256, 128], 128)):
    for mode in ("fused", "unfused"):
        os.environ["B200REC_MLP"] = mode
        torch.manual_seed(0)
        t = UserTower(K0, embedding_dim=E, hidden_layers=hidden, dropout_rate=0.2).cuda()
        t.train()
        x = torch.randn(B, K0, device="cuda", requires_grad=True)
        R = torch.randn(B, E, device="cuda")
        def fwd():
            with torch.no_grad():
                t(x)
        def both():
            for p in t.parameters():
                p.grad = None
            x.grad = None
            (t(x) * R).sum().backward()
        print(f"{name:16s} {mode:8s} fwd {timeit(fwd):7.1f} us   fwd+bwd {timeit(both):7.1f} us", flush=True)
