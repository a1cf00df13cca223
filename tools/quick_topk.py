"""Quick device timing of the fused top-K kernel (development aid; bench.py is the judged entry point)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from b200rec import kernels as KR  # noqa: E402


def run(N, Q, D, k, iters=5):
    g = torch.Generator(device="cuda").manual_seed(1234)
    cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    ws = torch.empty(KR.topk_workspace_bytes(N, D, Q, k), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        s, i = KR.flat_ip_topk(cat, qry, k, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        s, i = KR.flat_ip_topk(cat, qry, k, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * Q * N * D
    print(f"N={N} Q={Q} D={D} k={k}: {ms:.3f} ms  {Q / ms * 1e3:.0f} QPS  {flops / ms / 1e9:.1f} TFLOP/s  "
          f"cat-read {N * D * 2 / ms / 1e6:.1f} GB/s  ws={ws.numel() / 1e6:.0f} MB", flush=True)
    # spot check vs torch on a few queries
    ref = (qry[:4].float() @ cat.float().T).topk(k, dim=1)
    ok = torch.equal(ref.indices, i[:4]) or (ref.values - s[:4]).abs().max().item() < 1e-5
    print("   spot-check:", "ok" if ok else "MISMATCH", (ref.values - s[:4]).abs().max().item(), flush=True)


if __name__ == "__main__":
    t = time.time()
    run(1_000_000, 4096, 128, 100)
    run(10_000_000, 4096, 128, 100)
    run(10_000_000, 1, 128, 100)
    run(10_000_000, 128, 128, 100)
    run(10_000_000, 1024, 64, 1000, iters=2)
    print("total", time.time() - t)
