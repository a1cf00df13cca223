"""Relative error of the in-batch loss (forward only) against fp64 torch with 6 / 3 split-bf16 piece products."""
import sys, torch
sys.path.insert(0, ".")
from b200rec import ops
torch.manual_seed(0)
for B, NI, E, corr in ((8192, 8192, 64, 0.5), (8192, 8192, 64, 0.0), (8192, 8192, 64, 2.0), (1024, 1024, 128, 0.5), (8192, 8192, 128, 0.5),
                       (4096, 32768, 64, 0.5), (1024, 1024, 128, 0.0), (968, 968, 128, 1.0)):
    errs = {3: [], 6: []}
    for trial in range(4):
        u = torch.nn.functional.normalize(torch.randn(B, E), dim=1)
        v = torch.nn.functional.normalize(torch.randn(NI, E), dim=1)
        v[:B] = torch.nn.functional.normalize(v[:B] + corr * u, dim=1)
        lg = u.double().cuda() @ v.double().cuda().T * 20.0
        ref = (torch.logsumexp(lg, 1) - lg[torch.arange(B), torch.arange(B)]).sum().item() / NI * (NI / B)
        for t in (3, 6):
            loss = ops.InBatchCEFn.apply(u.cuda(), v.cuda(), 20.0, t, 0, B).item()
            errs[t].append(abs(loss - ref) / abs(ref))
    print(f"B {B} NI {NI} E {E} corr {corr}: loss {ref:.4f}  rel err 3 products max {max(errs[3]):.2e}  6 products max {max(errs[6]):.2e}", flush=True)
