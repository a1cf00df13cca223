import os, sys, ctypes, torch
sys.path.insert(0, ".")
from b200rec import kernels as KR, _native as NV
lib = NV.lib()
Q, D, k = 4096, 128, 100
g = torch.Generator(device="cuda").manual_seed(1234)
big = torch.nn.functional.normalize(torch.randn(5_000_000, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
for N in (156_250, 312_500, 625_000, 1_250_000, 2_500_000, 5_000_000):
    cat = big[:N]
    ws = torch.empty(KR.topk_workspace_bytes(N, D, Q, k), dtype=torch.uint8, device="cuda")
    for dbg in ("0", "1"):
        os.environ["B200REC_TOPK_DEBUG"] = dbg
        for _ in range(5): KR.flat_ip_topk(cat, qry, k, workspace=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): KR.flat_ip_topk(cat, qry, k, workspace=ws)
        e1.record(); torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 24)()
        lib.b200rec_debug_topk_stats16(buf, 1)
        KR.flat_ip_topk(cat, qry, k, workspace=ws); torch.cuda.synchronize()
        lib.b200rec_debug_topk_stats16(buf, 1)
        span = (buf[7] - ((~buf[5]) & 0xFFFFFFFFFFFFFFFF)) / 1e3
        print(f"N={N:8d} dbg={dbg}: call {e0.elapsed_time(e1)/20*1e3:7.0f} us, main kernel {span:7.0f} us @ {buf[6]} MHz, ideal at 1063 TF/s {2.0*Q*N*D/1063e12*1e6:6.0f} us", flush=True)
