import os, sys, ctypes, torch
sys.path.insert(0, ".")
os.environ["B200REC_TOPK_DEBUG"] = "2"
from b200rec import kernels as KR, _native as N
lib = N.lib()
def run(Nr, Q, D, k):
    g = torch.Generator(device="cuda").manual_seed(1234)
    cat = torch.nn.functional.normalize(torch.randn(Nr, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
    ws = torch.empty(KR.topk_workspace_bytes(Nr, D, Q, k), dtype=torch.uint8, device="cuda")
    KR.flat_ip_topk(cat, qry, k, workspace=ws); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 8)()
    lib.b200rec_debug_topk_stats(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); KR.flat_ip_topk(cat, qry, k, workspace=ws); e1.record(); torch.cuda.synchronize()
    lib.b200rec_debug_topk_stats(buf, 1)
    ms = e0.elapsed_time(e1)
    print(f"N={Nr} Q={Q} k={k}: {ms:.2f} ms appends/query={buf[0]/Q:.0f} requests/query={buf[1]/Q:.1f} "
          f"epi-wait Mcyc/CTA={buf[2]/148/1e6:.2f} helper-busy Mcyc/CTA={buf[3]/148/1e6:.2f} lock-miss={buf[4]} (kernel ~{ms*1.9:.1f} Mcyc)", flush=True)
run(250_000, 4096, 128, 100)
run(1_000_000, 4096, 128, 100)
run(10_000_000, 4096, 128, 100)
run(10_000_000, 64, 128, 100)
