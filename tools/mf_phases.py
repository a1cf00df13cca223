"""Phase time stamps of one fused MLP layer launch (variant build: python tools/build_variant.py mf_timing -DMF_TIMING;
run with B200REC_LIB=variants/mf_timing.so)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from b200rec import kernels as K, _native as N
from b200rec.two_tower import UserTower
names = {0: "start", 1: "prologue done", 14: "main loop done", 15: "accumulators ready", 16: "epilogue done", 17: "dealloc", 18: "first tmem ld", 19: "groups done", 20: "staged+synced"}
for ci in range(4):
    names[2 + 3 * ci], names[3 + 3 * ci], names[4 + 3 * ci] = f"chunk{ci} loaded", f"chunk{ci} synced", f"chunk{ci} issued"
def stamps():
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    assert N.lib().b200rec_debug_mf_stamps(buf) == 0
    return list(buf)
for name, B, K0, hidden, E in (("cfg2 tower", 8192, 80, [128, 64], 64), ("ml1m user", 1024, 3, [256, 128], 128)):
    t = UserTower(K0, embedding_dim=E, hidden_layers=hidden, dropout_rate=0.2).cuda().train()
    x = torch.randn(B, K0, device="cuda", requires_grad=True)
    for _ in range(3):
        e = t(x)
    # re-run the layers one by one through the kernel API to read the stamps of each launch
    import b200rec.kernels as KK
    orig = {n: getattr(KK, n) for n in ("mlp_forward", "mlp_dgrad", "mlp_wgrad")}
    log = []
    def wrap(n):
        def f(*a, **k):
            r = orig[n](*a, **k)
            log.append((n, stamps()))
            return r
        return f
    for n in orig:
        setattr(KK, n, wrap(n))
    e = t(x)
    (e * torch.randn_like(e)).sum().backward()
    for n in orig:
        setattr(KK, n, orig[n])
    print("==", name)
    for n, s in log:
        t0 = s[0]
        seq = sorted((v - t0, i) for i, v in enumerate(s[:21]) if v >= t0 and i in names)
        print(n, " | ".join(f"{names[i]} {d / 1000:.1f}" for d, i in seq))
