"""Pipeline wait-cycle breakdown of the fused scoring + top-K kernel (development tool; B200REC_TOPK_DEBUG=2)."""
import os, sys, ctypes, torch
sys.path.insert(0, ".")
from b200rec import kernels as KR, _native as NV
lib = NV.lib()
Nr, Q, D, k = int(os.environ.get("NROWS", 10_000_000)), int(os.environ.get("NQ", 4096)), 128, 100
g = torch.Generator(device="cuda").manual_seed(1234)
cat = torch.nn.functional.normalize(torch.randn(Nr, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
ws = torch.empty(KR.topk_workspace_bytes(Nr, D, Q, k), dtype=torch.uint8, device="cuda")
def run(tag):
    buf = (ctypes.c_ulonglong * 24)()
    for _ in range(3):
        KR.flat_ip_topk(cat, qry, k, workspace=ws)
    torch.cuda.synchronize()
    lib.b200rec_debug_topk_stats16(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); KR.flat_ip_topk(cat, qry, k, workspace=ws); e1.record(); torch.cuda.synchronize()
    lib.b200rec_debug_topk_stats16(buf, 1)
    span = (buf[7] - ((~buf[5]) & 0xFFFFFFFFFFFFFFFF)) / 1e6
    ctas, tiles = 148, max(buf[20], 1)
    per_tile = lambda x, warps: x / warps / (tiles / (148 * 8))
    tiles_cta = tiles / (148 * 8)
    print(f"{tag}: call {e0.elapsed_time(e1):.2f} ms, main {span:.2f} ms @ {buf[6]} MHz, {tiles_cta:.0f} tiles/CTA, "
          f"cycles/tile {span*1e-3*buf[6]*1e6/tiles_cta:.0f}", flush=True)
    print(f"   MMA warp (74 leaders): wait-acc {buf[16]/74/tiles_cta:.0f} cyc/tile, wait-operands {buf[17]/74/tiles_cta:.0f} cyc/tile", flush=True)
    print(f"   epilogue warp: wait-acc {buf[18]/tiles:.0f} cyc/tile, consume {buf[19]/tiles:.0f} cyc/tile", flush=True)
    print(f"   appends/query {buf[0]/Q:.0f}, requests/query {buf[1]/Q:.1f}, slow-path entries/warp-tile {buf[9]/tiles:.3f}, "
          f"cycles/entry {buf[8]/max(buf[9],1):.0f}, group re-reads/entry {buf[10]/max(buf[9],1):.2f}, helper busy {buf[3]/148/4/1e6:.3f} Mcyc/warp, lock-miss {buf[4]}, epi mailbox wait {buf[2]/148/8/1e6:.3f} Mcyc/warp", flush=True)
    print(f"   rare path per entry: find groups {buf[11]/max(buf[9],1):.0f} cyc, TMEM re-read {buf[12]/max(buf[9],1):.0f} cyc, tests+append {buf[13]/max(buf[9],1):.0f} cyc", flush=True)
os.environ["B200REC_TOPK_DEBUG"] = "2"
lib.b200rec_debug_reload_env()   # knobs are read once per process otherwise
run("full+stats")
os.environ["B200REC_TOPK_DEBUG"] = "1"; os.environ["B200REC_STREAM_STATS"] = "1"
lib.b200rec_debug_reload_env()
run("reject-all+stats")
