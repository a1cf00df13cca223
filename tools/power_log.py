"""Power / clock trace of the fused top-K kernel next to cuBLAS bf16 on the SAME box (is the kernel
power-limited?).  Samples nvidia-smi every 10 ms while each workload loops for >= 3 s; writes gpurun_out/power_*.csv and
prints medians.   python tools/power_log.py"""
import os, subprocess, sys, threading, time, torch
sys.path.insert(0, ".")
from b200rec import kernels as KR

def sample(tag, fn, seconds=3.5):
    rows = []
    p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=power.draw,clocks.sm,clocks.mem,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                          "--format=csv,noheader,nounits", "-lms", "10"], stdout=subprocess.PIPE, text=True)
    th = threading.Thread(target=lambda: [rows.append((time.perf_counter(), l.strip())) for l in p.stdout], daemon=True)
    th.start()
    for _ in range(3): fn()
    torch.cuda.synchronize(); time.sleep(0.3)
    t0 = time.perf_counter(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(5): fn()
        n += 5
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    p.terminate()
    ms = e0.elapsed_time(e1) / n
    keep = [l for t, l in rows if t0 + 0.5 <= t <= t1]
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/power_{tag}.csv", "w") as fh:
        fh.write("power_w,sm_mhz,mem_mhz,temp_c,sw_power_cap,hw_slowdown,sw_thermal\n" + "\n".join(keep) + "\n")
    pw = sorted(float(l.split(",")[0]) for l in keep); sm = sorted(float(l.split(",")[1]) for l in keep)
    cap = sum("Active" in l.split(",")[4] for l in keep)
    print(f"{tag}: {ms:.3f} ms/iter, {len(keep)} samples, power median {pw[len(pw)//2]:.0f} W (min {pw[0]:.0f}, max {pw[-1]:.0f}), "
          f"SM clock median {sm[len(sm)//2]:.0f} MHz (min {sm[0]:.0f}, max {sm[-1]:.0f}), sw_power_cap active in {cap} samples", flush=True)
    return ms

N, Q, D, k = 10_000_000, 4096, 128, 100
g = torch.Generator(device="cuda").manual_seed(1234)
cat = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
qry = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda", generator=g), dim=1).to(torch.bfloat16)
ws = torch.empty(KR.topk_workspace_bytes(N, D, Q, k), dtype=torch.uint8, device="cuda")
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
print(os.popen("nvidia-smi --query-gpu=name,power.limit,power.max_limit,clocks.max.sm --format=csv,noheader").read().strip(), flush=True)
for rnd in range(2):
    ms = sample(f"topk_r{rnd}", lambda: KR.flat_ip_topk(cat, qry, k, workspace=ws))
    print(f"   top-K: {2*Q*N*D/ms/1e9:.0f} TFLOP/s algorithmic", flush=True)
    ms = sample(f"cublas_r{rnd}", lambda: torch.matmul(a, b))
    print(f"   cuBLAS bf16 8192^3: {2*8192**3/ms/1e9:.0f} TFLOP/s", flush=True)
