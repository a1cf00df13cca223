import os, sys, ctypes, torch
os.environ["B200REC_TOPK_DEBUG"] = "2"
sys.path.insert(0, ".")
from b200rec import kernels as KR, _native as NV
lib = NV.lib()
Nr, Q, D, k = int(os.environ.get("NROWS", 1_250_000)), 4096, 128, 100
g = torch.Generator(device="cuda").manual_seed(1234)
cat = torch.nn.functional.normalize(torch.randn(Nr, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
ws = torch.empty(KR.topk_workspace_bytes(Nr, D, Q, k), dtype=torch.uint8, device="cuda")
for _ in range(3): KR.flat_ip_topk(cat, qry, k, workspace=ws)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 24)(); lib.b200rec_debug_topk_stats16(buf, 1)
KR.flat_ip_topk(cat, qry, k, workspace=ws); torch.cuda.synchronize()
lib.b200rec_debug_topk_stats16(buf, 0)
ends = (ctypes.c_ulonglong * 512)(); lib.b200rec_debug_topk_cta_end(ends)
start = (~buf[5]) & 0xFFFFFFFFFFFFFFFF
d = [(ends[i] - start) / 1e3 for i in range(0, 148, 2)]
print("unit 72 segments (begin,end) us:", " ".join(f"({(ends[256+2*i]-start)/1e3:.0f},{(ends[257+2*i]-start)/1e3:.0f})" for i in range(6) if ends[256+2*i] > start))
import struct
print("unit 72 warp4 cumulative per segment (entries, lane0 appends, lane0 tau at end, slow cycles):", " ".join(f"({ends[400+4*i]},{ends[401+4*i]},{struct.unpack('f', struct.pack('I', ends[402+4*i] & 0xFFFFFFFF))[0]:.4f},{ends[403+4*i]})" for i in range(4)))
print("unit durations us:", " ".join(f"{x:.0f}" for x in d))

print("per-warp (entries, appends, slow kcycles): CTA144:", " ".join(f"({ends[300+3*w]},{ends[301+3*w]},{ends[302+3*w]//1000})" for w in range(8)), " CTA145:", " ".join(f"({ends[300+3*w]},{ends[301+3*w]},{ends[302+3*w]//1000})" for w in range(8, 16)), " CTA0:", " ".join(f"({ends[300+3*w]},{ends[301+3*w]},{ends[302+3*w]//1000})" for w in range(16, 24)))
for u in (0, 36, 72):
    os.environ["B200REC_STATS_UNIT"] = str(u)
    lib.b200rec_debug_reload_env()
    lib.b200rec_debug_topk_stats16(buf, 1)
    KR.flat_ip_topk(cat, qry, k, workspace=ws); torch.cuda.synchronize()
    lib.b200rec_debug_topk_stats16(buf, 0)
    tiles = max(buf[20], 1) / 16
    print(f"unit {u}: tiles {tiles:.0f}; MMA wait-acc {buf[16]/tiles:.0f} cyc/tile, wait-operands {buf[17]/tiles:.0f}; epilogue wait-acc {buf[18]/16/tiles:.0f}, consume {buf[19]/16/tiles:.0f} cyc/tile (MHz {buf[6]})")
