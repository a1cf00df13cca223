import os, sys, torch
sys.path.insert(0, ".")
from b200rec import kernels as KR, _native as NV
import ctypes
lib = NV.lib()
N, Q, D, k = int(os.environ.get("NROWS", 10_000_000)), int(os.environ.get("NQ", 4096)), 128, 100
g = torch.Generator(device="cuda").manual_seed(1234)
cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
def run(tag, iters=int(os.environ.get('ITERS', 8))):
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
    buf = (ctypes.c_ulonglong * 8)()
    lib.b200rec_debug_topk_stats(buf, 1)
    KR.flat_ip_topk(cat, qry, k, workspace=ws); torch.cuda.synchronize()
    lib.b200rec_debug_topk_stats(buf, 1)
    span = (buf[7] - ((~buf[5]) & 0xFFFFFFFFFFFFFFFF)) / 1e6
    print(f"{tag}: {ms:.3f} ms {Q/ms*1e3:.0f} QPS {2.0*Q*N*D/ms/1e9:.1f} TFLOP/s  main kernel {span:.3f} ms @ {buf[6]} MHz", flush=True)
if os.environ.get("CEIL", "1") == "1":
    os.environ["B200REC_TOPK_NOSAMPLE"] = "1"
    for v2 in ("2",):
        os.environ["B200REC_TOPK_V2"] = v2
        for dbg, name in (("3", "fed,noepi"), ("5", "fed,tmem-read only"), ("1", "fed,fastpath reject-all")):
            os.environ["B200REC_TOPK_DEBUG"] = dbg
            run(f"v2={v2} dbg={dbg} ({name})")
    os.environ.pop("B200REC_TOPK_DEBUG"); os.environ.pop("B200REC_TOPK_NOSAMPLE")
os.environ["B200REC_TOPK_V2"] = "2"
for mo in os.environ.get("ORDERS", "0").split(","):
  os.environ["B200REC_MMA_ORDER"] = mo
  os.environ["B200REC_TOPK_NOSAMPLE"] = "1"
  for dbg in ("3", "1"):
    os.environ["B200REC_TOPK_DEBUG"] = dbg
    run(f"order={mo} dbg={dbg}")
  os.environ.pop("B200REC_TOPK_DEBUG"); os.environ.pop("B200REC_TOPK_NOSAMPLE")
  run(f"order={mo} full")
for hs in os.environ.get("HSLEEPS", "64,400,1500").split(","):
    os.environ["B200REC_TOPK_HSLEEP"] = hs
    run(f"v2=2 full hsleep={hs}")
