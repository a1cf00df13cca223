"""GPU parity of the training-side kernels against plain fp32 torch / the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_gather_concat_and_sparse_grad_match_oracle():
    from b200rec import kernels as K
    from oracle.two_tower import embedding_touched_rows
    rng = np.random.default_rng(0)
    B = 1000
    tabs = [torch.randn(501, 50), torch.randn(20001, 64), torch.randn(7, 3)]
    idx = [torch.from_numpy(rng.integers(0, t.shape[0], size=B)) for t in tabs]
    idx[1][:300] = 17  # hot row
    num = torch.randn(B, 5)
    widths = [50, 64, 3]
    offs = [5, 55, 119]
    err = torch.zeros((1,), dtype=torch.int32, device=DEV)
    out = K.gather_concat(num.to(DEV), [t.to(DEV) for t in tabs], [i.to(DEV) for i in idx], widths, offs, 122, B, DEV, err)
    ref = torch.cat([num] + [t[i] for t, i in zip(tabs, idx)], dim=1)
    assert torch.equal(out.cpu(), ref)  # gathers are bit-exact
    assert int(err.item()) == 0
    dout = torch.randn(B, 122)
    for f, (t, i) in enumerate(zip(tabs, idx)):
        rows, vals, n = K.embedding_sparse_grad(i.to(DEV), dout.to(DEV)[:, offs[f]:], widths[f], t.shape[0], 0)
        n = int(n.item())
        want_rows = embedding_touched_rows(i.numpy())
        assert np.array_equal(rows[:n].cpu().numpy(), want_rows)  # gradient index sets are bit-exact
        dense = torch.zeros_like(t).index_add_(0, i, dout[:, offs[f]:offs[f] + widths[f]])
        dense[0] = 0
        got = torch.zeros_like(t)
        got[rows[:n].cpu()] = vals[:n].cpu()
        assert torch.allclose(got, dense, atol=1e-4, rtol=1e-5)


def test_out_of_range_ids_raise_and_never_touch_other_memory():
    """nn.Embedding raises IndexError for an id outside the table (reference two_tower.py:46-50,116).  Our gather flags
    it (ops.check_index_errors raises), and the backward treats such a sample like the padding row: no row >= the table
    size ever reaches scatter_rows / sparse Adam."""
    from b200rec import kernels as K, ops
    from b200rec.two_tower import UserTower
    torch.manual_seed(0)
    t = UserTower(4, embedding_dim=16, hidden_layers=[32], dropout_rate=0.0, categorical_features={"f": 10}).to(DEV)
    ids = torch.tensor([1, 2, 11, 3, -5, 2, 10, 7], device=DEV)           # table has 11 rows: 11 and -5 are invalid
    out = t(torch.randn(8, 4, device=DEV), {"f": ids})
    with pytest.raises(IndexError):
        ops.check_index_errors()
    ops.check_index_errors()                                               # the flag was consumed
    guard = torch.zeros(64, 6, device=DEV)                                 # table view inside a larger buffer
    table = guard[16:27]
    dY = torch.ones(8, 6, device=DEV)
    rows, vals, nt = K.embedding_sparse_grad(ids, dY, 6, 11, 0)
    n = int(nt.item())
    assert rows[:n].cpu().tolist() == [1, 2, 3, 7, 10]
    assert vals[:n, 0].cpu().tolist() == [1.0, 2.0, 1.0, 1.0, 1.0]
    K.scatter_rows(rows, vals, nt, table, accumulate=True)
    assert guard[:16].abs().sum().item() == 0 and guard[27:].abs().sum().item() == 0
    out.sum().backward()                                                   # full autograd path stays in bounds too
    assert torch.isfinite(t.embeddings["f"].weight.grad).all()
    with pytest.raises(RuntimeError):                                      # one id per row is required
        t(torch.randn(8, 4, device=DEV), {"f": ids[:5]})


def test_explicit_ce_rejects_ragged_negatives():
    from b200rec import ops
    u, p = torch.randn(4, 8, device=DEV), torch.randn(4, 8, device=DEV)
    with pytest.raises(RuntimeError):
        ops.ExplicitCEFn.apply(u, p, torch.randn(10, 8, device=DEV), None, None, 20.0)


def test_flat_adam_matches_torch_adam_with_clipping():
    from b200rec.trainer import FlatAdam
    torch.manual_seed(1)
    shapes = [(64, 20), (64,), (128, 64), (1,)]
    ref_p = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref_p]
    ref_opt = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=1e-5)
    ours = FlatAdam(our_p, lr=1e-3, weight_decay=1e-5, max_grad_norm=1.0)
    ours.touch_all()           # gradients are copied straight into .grad below (no autograd hooks fire)
    for step in range(5):
        grads = [torch.randn(s) * (3.0 if step % 2 == 0 else 0.01) for s in shapes]
        ref_opt.zero_grad()
        ours.zero_grad()
        for p, g in zip(ref_p, grads):
            p.grad = g.clone()
        for p, g in zip(our_p, grads):
            p.grad.copy_(g.to(DEV))
        norm = torch.nn.utils.clip_grad_norm_(ref_p, max_norm=1.0)
        ref_opt.step()
        ours.step()
        assert abs(ours.grad_norm.item() - norm.item()) <= 1e-5 * norm.item()
        for a, b in zip(our_p, ref_p):
            assert torch.allclose(a.detach().cpu(), b.detach(), atol=1e-6, rtol=1e-5)


def test_flat_adam_skips_parameters_without_gradient_and_round_trips_torch_state():
    """torch.optim.Adam leaves a parameter whose grad is None untouched (no update, no weight decay): FlatAdam tracks
    which flat segments autograd reached.  Its state_dict has torch.optim.Adam's layout and loads one written by torch."""
    from b200rec.trainer import FlatAdam
    torch.manual_seed(4)
    shapes = [(1,), (8, 4), (8,), (3, 8)]
    ref_p = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref_p]
    ref_opt = torch.optim.Adam(ref_p, lr=1e-2, weight_decay=1e-2)
    ours = FlatAdam(our_p, lr=1e-2, weight_decay=1e-2, max_grad_norm=None)
    x = torch.randn(5, 4)
    for _ in range(3):
        ref_opt.zero_grad()
        ours.zero_grad()
        (x @ ref_p[1].T + ref_p[2]).pow(2).sum().backward()               # parameters 0 and 3 are never reached
        (x.to(DEV) @ our_p[1].T + our_p[2]).pow(2).sum().backward()
        ref_opt.step()
        ours.step()
    for a, b in zip(our_p, ref_p):
        assert torch.allclose(a.detach().cpu(), b.detach(), atol=1e-6, rtol=1e-5)
    sd, rsd = ours.state_dict(), ref_opt.state_dict()
    assert sorted(sd.keys()) == sorted(rsd.keys()) and sd["param_groups"][0]["params"] == rsd["param_groups"][0]["params"]
    for i in rsd["state"]:
        # the gradients come from the 3-product split-bf16 GEMMs (~2^-17 per product): moments agree to ~1e-5 relative
        assert torch.allclose(sd["state"][i]["exp_avg"].cpu(), rsd["state"][i]["exp_avg"], rtol=1e-4, atol=1e-6)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"].cpu(), rsd["state"][i]["exp_avg_sq"], rtol=1e-4, atol=1e-8)
    # resume from the torch optimiser's state: the next step matches torch's next step
    fresh_p = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref_p]
    fresh = FlatAdam(fresh_p, lr=1.0, weight_decay=0.0, max_grad_norm=None)
    fresh.load_state_dict(rsd)
    ref_opt.zero_grad()
    fresh.zero_grad()
    (x @ ref_p[1].T + ref_p[2]).pow(2).sum().backward()
    (x.to(DEV) @ fresh_p[1].T + fresh_p[2]).pow(2).sum().backward()
    ref_opt.step()
    fresh.step()
    for a, b in zip(fresh_p, ref_p):
        assert torch.allclose(a.detach().cpu(), b.detach(), atol=1e-6, rtol=1e-5)


def test_inbatch_loss_large_batch_and_backward_vs_torch():
    from b200rec import ops
    torch.manual_seed(2)
    for B, E, terms, tol in ((8192, 64, 6, 1e-5), (3000, 128, 6, 1e-5), (4096, 64, 1, 2e-2)):
        u = torch.nn.functional.normalize(torch.randn(B, E), dim=1)
        i = torch.nn.functional.normalize(torch.randn(B, E) + 0.5 * u, dim=1)
        uc, ic = u.clone().to(DEV).requires_grad_(), i.clone().to(DEV).requires_grad_()
        loss = ops.InBatchCEFn.apply(uc, ic, 20.0, terms, 0, B)
        loss.backward()
        ur, ir = u.clone().double().requires_grad_(), i.clone().double().requires_grad_()
        ref = torch.nn.functional.cross_entropy(ur @ ir.T * 20.0, torch.arange(B))
        ref.backward()
        assert abs(loss.item() - ref.item()) <= tol * abs(ref.item()), (B, E, terms, loss.item(), ref.item())
        gtol = 1e-4 if terms == 6 else 5e-2
        assert (uc.grad.cpu().double() - ur.grad).abs().max() <= gtol * ur.grad.abs().max()
        assert (ic.grad.cpu().double() - ir.grad).abs().max() <= gtol * ir.grad.abs().max()


@pytest.mark.parametrize("B,G,E,rank", [(1024, 4, 64, 2), (2048, 8, 64, 7), (700, 3, 128, 1), (8192, 2, 128, 0)])
def test_fused_inbatch_backward_with_gathered_negatives(B, G, E, rank):
    """The data-parallel shape of the in-batch loss (two_tower.py:453-479 on the all-gathered item batch): B local users
    against NI = G*B items, positives at rows [rank*B, rank*B + B).  Column ranges longer than one work unit are split
    and combined with atomics (csrc/inbatch_grad.cu); fused path == chunked path == fp64 torch."""
    import os
    from b200rec import ops
    torch.manual_seed(B + G)
    NI = G * B + (37 if B == 700 else 0)
    u = torch.nn.functional.normalize(torch.randn(B, E), dim=1)
    v = torch.nn.functional.normalize(torch.randn(NI, E), dim=1)
    v[rank * B:rank * B + B] = torch.nn.functional.normalize(v[rank * B:rank * B + B] + 0.7 * u, dim=1)
    ur, vr = u.clone().double().requires_grad_(), v.clone().double().requires_grad_()
    logits = ur @ vr.T * 20.0
    ref = (torch.logsumexp(logits, 1) - logits[torch.arange(B), rank * B + torch.arange(B)]).sum() / (G * B)
    (ref * 3.0).backward()
    # the loss is a DIFFERENCE of O(10) terms (strong positives: lse ~ pos): 1e-5 relative to the terms that are summed
    scale = max(abs(ref.item()), torch.logsumexp(logits, 1).abs().mean().item() / G)
    res = {}
    for mode in ("fused", "chunked"):
        os.environ["B200REC_INBATCH_BWD"] = mode
        try:
            uc, vc = u.clone().to(DEV).requires_grad_(), v.clone().to(DEV).requires_grad_()
            loss = ops.InBatchCEFn.apply(uc, vc, 20.0, 6, rank * B, G * B)
            (loss * 3.0).backward()
        finally:
            os.environ.pop("B200REC_INBATCH_BWD", None)
        assert abs(loss.item() - ref.item()) <= 1e-5 * scale, (loss.item(), ref.item())
        assert (uc.grad.cpu().double() - ur.grad).abs().max() <= 1e-4 * ur.grad.abs().max(), mode
        assert (vc.grad.cpu().double() - vr.grad).abs().max() <= 1e-4 * vr.grad.abs().max(), mode
        res[mode] = (uc.grad.cpu(), vc.grad.cpu())
    assert (res["fused"][0] - res["chunked"][0]).abs().max() <= 1e-4 * ur.grad.abs().max()


@pytest.mark.parametrize("B,E", [(2048, 64), (1500, 128)])
def test_fused_inbatch_backward_six_product_recompute(B, E, monkeypatch):
    """B200REC_BWD_TERMS=6: the logits recompute of the fused gradient kernel uses the 6 fp32-grade piece products
    (inbatch_grad_kernel<E, 6, 3>; E = 128 fits since the gradient GEMM reads the row operands as MN-major tiles and G
    lives in tensor memory).  The gradient GEMMs keep 3 products (~2^-17 each): gradients within 5e-5 of fp64 torch."""
    from b200rec import kernels as K, ops
    monkeypatch.setenv("B200REC_BWD_TERMS", "6")
    assert K.inbatch_grad_supported(B, B, E, 6, 3)
    torch.manual_seed(B)
    u = torch.nn.functional.normalize(torch.randn(B, E), dim=1)
    v = torch.nn.functional.normalize(torch.randn(B, E) + 0.5 * u, dim=1)
    uc, vc = u.clone().to(DEV).requires_grad_(), v.clone().to(DEV).requires_grad_()
    ops.InBatchCEFn.apply(uc, vc, 20.0, 6, 0, B).backward()
    ur, vr = u.clone().double().requires_grad_(), v.clone().double().requires_grad_()
    torch.nn.functional.cross_entropy(ur @ vr.T * 20.0, torch.arange(B)).backward()
    assert (uc.grad.cpu().double() - ur.grad).abs().max() <= 5e-5 * ur.grad.abs().max()
    assert (vc.grad.cpu().double() - vr.grad).abs().max() <= 5e-5 * vr.grad.abs().max()


def test_row_flagged_table_adam_matches_flat_adam(monkeypatch):
    """Dense id tables under the reference's dense Adam + weight decay (trainers/two_tower.py:60-64): the row-flagged
    update (b200rec_scatter_add_rows_flagged / _table_sumsq / _adam_table: gradient of touched rows only) moves EVERY row
    exactly as the flat kernels do (B200REC_TABLE_ADAM=0), clears its flags, and reports the same gradient norm."""
    from b200rec.trainer import TwoTowerTrainer
    from b200rec.training_utils import create_two_tower_model_for_training
    cfg = {"embedding_dim": 64, "hidden_layers": [128, 64], "dropout_rate": 0.0, "temperature": 0.05,
           "user_categorical_features": {"user_id": 5000}, "item_categorical_features": {"item_id": 700},
           "embedding_dims": {"user_id": 64, "item_id": 50}}
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(512, 8, generator=g), torch.randn(512, 8, generator=g),
                torch.randint(0, 5001, (512,), generator=g), torch.randint(0, 701, (512,), generator=g)) for _ in range(3)]
    runs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("B200REC_TABLE_ADAM", mode)
        torch.manual_seed(11)
        model = create_two_tower_model_for_training(8, 8, cfg)
        tr = TwoTowerTrainer(model, [], [], {"checkpoint_dir": "/tmp/b200rec_rowflag", "weight_decay": 1e-2}, device=DEV)
        model.train()
        opt = tr.optimizer
        assert any(opt._table) == (mode == "1")
        norms = []
        for uf, pf, uid, iid in batches:
            tr.train_step(uf.to(DEV), pf.to(DEV), None, {"user_id": uid.to(DEV)}, {"item_id": iid.to(DEV)})
            norms.append(opt.grad_norm.item())
        if mode == "1":
            for p in opt.dense:
                if getattr(p, "_b200_row_flags", None) is not None:
                    assert int(p._b200_row_flags.sum().item()) == 0        # consumed and cleared by the update
            assert float(opt.grad.abs().max().item()) == 0.0               # Adam reset every gradient it consumed
        runs[mode] = ({k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, norms,
                      opt.m.detach().cpu().clone(), opt.v.detach().cpu().clone())
    a, b = runs["1"], runs["0"]
    assert a[1][0] == pytest.approx(b[1][0], rel=1e-6)          # step 1 starts from identical parameters
    # later steps: both runs add gradient rows with atomics (unordered), and Adam turns 1e-9 of noise on a near-zero
    # gradient into an lr-sized step, so parameters agree to a few lr only; exactness is checked kernel against kernel below
    for k in a[0]:
        assert torch.allclose(a[0][k].float(), b[0][k].float(), atol=3.5e-3, rtol=0), k
    # untouched rows moved too (weight decay), i.e. the row-flagged path did not skip them
    w = a[0]["user_tower.embeddings.user_id.weight"]
    torch.manual_seed(11)
    w0 = create_two_tower_model_for_training(8, 8, cfg).state_dict()["user_tower.embeddings.user_id.weight"]
    untouched = torch.ones(5001, dtype=torch.bool)
    for _, _, uid, _ in batches:
        untouched[uid] = False
    assert (w[untouched] != w0[untouched]).any()


@pytest.mark.parametrize("rows,e", [(4099, 64), (1000, 50), (257, 128), (33, 7)])
def test_adam_table_kernel_matches_adam_dense(rows, e):
    """b200rec_adam_table (gradient of flagged rows only) == b200rec_adam_dense over the same buffers when the gradient
    is zero outside the flagged rows; b200rec_table_sumsq == b200rec_sumsq; flags and gradients are reset."""
    from b200rec import kernels as K
    g = torch.Generator().manual_seed(rows + e)
    p = torch.randn(rows, e, generator=g)
    m = torch.randn(rows, e, generator=g) * 1e-2
    v = torch.rand(rows, e, generator=g) * 1e-3
    flags = (torch.rand(rows, generator=g) < 0.2).to(torch.int32)
    grad = torch.randn(rows, e, generator=g) * flags[:, None]
    coef = torch.tensor([0.37])
    outs = []
    for table in (True, False):
        P, G, M, V, F, C = (t.clone().to(DEV) for t in (p, grad, m, v, flags, coef))
        acc = torch.zeros(1, dtype=torch.float64, device=DEV)
        if table:
            K.table_sumsq_(G, F, acc)
            K.adam_table_(P, G, M, V, F, 1e-3, 0.9, 0.999, 1e-8, 1e-2, 7, C, None, clear_grad=True)
            assert int(F.sum().item()) == 0
        else:
            K.sumsq_(G.reshape(-1), acc)
            K.adam_dense_(P.reshape(-1), G.reshape(-1), M.reshape(-1), V.reshape(-1), 1e-3, 0.9, 0.999, 1e-8, 1e-2, 7, C,
                          clear_grad=True)
        assert float(G.abs().max().item()) == 0.0
        outs.append((P.cpu(), M.cpu(), V.cpu(), acc.item()))
    for name, x, y in zip("pmv", outs[0][:3], outs[1][:3]):
        # same arithmetic in both kernels; the compiler may contract a multiply-add differently in the two loop bodies
        assert torch.allclose(x, y, atol=1e-7, rtol=1e-6), (name, (x - y).abs().max().item())   # last-ulp differences
    assert outs[0][3] == pytest.approx(outs[1][3], rel=1e-12)
    assert outs[0][3] == pytest.approx(float((grad.double() ** 2).sum()), rel=1e-12)


def test_peer_allreduce_single_rank_and_graph_replay():
    """csrc/peer_reduce.cu with world = 1 (the buffer is its own peer): data is unchanged, the sequence number in the
    buffer advances once per call — also when the calls are replayed from a CUDA graph, which is how the data-parallel
    training step runs them.  (World >= 2 needs one GPU per rank: tests/test_gpu_model.py DP test, tools/check_dp_training.py.)"""
    import ctypes
    from b200rec import _native as N
    max_n = 512
    nbytes = int(N.lib().b200rec_peer_allreduce_bytes(max_n))
    assert nbytes == 256 + 2 * 8 * max_n * 8
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
    ptrs = (ctypes.c_uint64 * 1)(buf.data_ptr())
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    x = torch.arange(300, dtype=torch.float64, device=DEV) * 0.25
    ref = x.clone()

    def call():
        N.check(N.lib().b200rec_peer_allreduce_f64(N.ptr(x), x.numel(), 0, 1, ptrs, max_n, N.ptr(status), N.stream()), "peer")

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    assert torch.equal(x, ref) and int(buf[:8].view(torch.int64).item()) == 3 and int(status.item()) == 0
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            call()
            call()
    for _ in range(4):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(x, ref) and int(buf[:8].view(torch.int64).item()) == 3 + 2 * 4 and int(status.item()) == 0
    assert N.lib().b200rec_peer_allreduce_f64(N.ptr(x), max_n + 1, 0, 1, ptrs, max_n, N.ptr(status), N.stream()) != 0


def test_sparse_claim_accumulate_equals_sort_and_segment_sum():
    """Row-sparse table gradients without a sort (b200rec_sparse_claim_accumulate): every distinct valid row appears
    once in rows_out (-1 for duplicates, the padding id and out-of-range ids), its summed gradient sits in the same
    position of the compact buffer, the claim map is all zero again; scattered into a dense table the result equals
    the sort + segment-sum path (b200rec_embedding_sparse_grad) up to the order of the fp32 additions."""
    from b200rec import kernels as K
    g = torch.Generator().manual_seed(9)
    rows_t, B, e = 300, 4096, 40
    ids = torch.randint(0, rows_t, (B,), generator=g)
    ids[::97] = 0            # padding id
    ids[5::211] = rows_t + 3  # out of range
    ids[7::223] = -2
    dY = torch.randn(B, e + 8, generator=g)          # a column block of a wider gradient
    slot = torch.zeros(rows_t, dtype=torch.int32, device=DEV)
    rows, acc, n = K.sparse_claim_accumulate(ids.to(DEV), dY.to(DEV)[:, 4:], e, rows_t, slot, 0)
    torch.cuda.synchronize()
    assert int(n.item()) == B and int(slot.abs().sum().item()) == 0
    r = rows.cpu()
    valid = (ids > 0) & (ids < rows_t)
    lead = r >= 0
    assert torch.equal(torch.sort(r[lead]).values, torch.unique(ids[valid]))      # every distinct valid row exactly once
    assert torch.equal(r[lead], ids[lead]) and not lead[~valid].any()
    assert float(acc.cpu()[~lead].abs().max()) == 0.0
    ref = torch.zeros(rows_t, e, dtype=torch.float64)
    ref.index_add_(0, ids[valid], dY[valid][:, 4:4 + e].double())
    dense = torch.zeros(rows_t, e, device=DEV)
    K.scatter_rows(rows.clamp(min=0), acc * lead.to(DEV)[:, None], n, dense, accumulate=True)
    assert torch.allclose(dense.cpu().double(), ref, atol=1e-5, rtol=1e-5)
    rows2, vals2, n2 = K.embedding_sparse_grad(ids.clamp(min=-1).to(DEV), dY.to(DEV)[:, 4:], e, rows_t, 0)
    dense2 = torch.zeros(rows_t, e, device=DEV)
    K.scatter_rows(rows2, vals2, n2, dense2)
    assert torch.allclose(dense.cpu(), dense2.cpu(), atol=1e-5, rtol=1e-5)


def test_bf16_mode_within_budget():
    from b200rec.training_utils import create_two_tower_model_for_training
    torch.manual_seed(3)
    m = create_two_tower_model_for_training(16, 16, {"embedding_dim": 64, "hidden_layers": [128, 64],
                                                     "dropout_rate": 0.0}).to(DEV)
    m.eval()
    x = torch.randn(512, 16, device=DEV)
    a = m.get_user_embeddings({"numerical": x})
    m.precision = "bf16"
    b = m.get_user_embeddings({"numerical": x})
    assert (a - b).abs().max().item() <= 2e-2 * a.abs().max().item()


def _torch_tower(t, dtype=torch.float64):
    """The reference module chain (two_tower.py:56-72) rebuilt from a b200rec tower's parameters, on the CPU in fp64."""
    import copy
    mlp = copy.deepcopy(t.mlp).cpu().to(dtype)
    for m in mlp:
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return mlp


@pytest.mark.parametrize("B,K0,hidden,E,act", [(300, 20, [64, 48], 64, "relu"), (1000, 3, [256, 128], 128, "gelu"),
                                               (257, 130, [100], 32, "tanh"), (64, 80, [512, 256, 128], 128, "leaky_relu"),
                                               (129, 16, [], 64, "relu"), (2048, 80, [128, 64], 64, "sigmoid")])
def test_fused_mlp_layers_match_torch_fp64(B, K0, hidden, E, act):
    """csrc/mlp_fused.cuh (one launch per Linear and direction) against the torch module chain in fp64: embeddings,
    input gradient, every parameter gradient, BatchNorm running statistics and batch counter; train and eval mode."""
    from b200rec.two_tower import UserTower
    torch.manual_seed(B + K0)
    t = UserTower(K0, embedding_dim=E, hidden_layers=hidden, dropout_rate=0.0, activation=act).to(DEV)
    for m in t.modules():
        if isinstance(m, torch.nn.Linear):
            m.bias.data.normal_(0, 0.1)
        if isinstance(m, torch.nn.BatchNorm1d):
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    assert t._fused_ok()
    ref = _torch_tower(t)
    x = torch.randn(B, K0)
    R = torch.randn(B, E)
    for training in (True, False):
        t.train(training)
        ref.train(training)
        xg = x.clone().to(DEV).requires_grad_()
        e = t(xg)
        (e * R.to(DEV)).sum().backward()
        xr = x.clone().double().requires_grad_()
        er = torch.nn.functional.normalize(ref(xr), p=2, dim=-1)
        (er * R.double()).sum().backward()
        assert (e.detach().cpu().double() - er.detach()).abs().max() <= 1e-5
        scale = lambda g: max(g.abs().max().item(), 1e-6)
        assert (xg.grad.cpu().double() - xr.grad).abs().max() <= 2e-4 * scale(xr.grad)
        ours = dict(t.mlp.named_parameters())
        for name, p in ref.named_parameters():
            assert (ours[name].grad.cpu().double() - p.grad).abs().max() <= 2e-4 * scale(p.grad), (training, name)
            ours[name].grad = None
            p.grad = None
        for (n1, b1), (n2, b2) in zip(t.mlp.named_buffers(), ref.named_buffers()):
            assert n1 == n2
            assert torch.allclose(b1.cpu().double(), b2.double(), rtol=1e-5, atol=1e-6), (training, n1)


def test_fused_mlp_equals_unfused_chain_with_dropout(monkeypatch):
    """Same dropout streams in both paths: the fused kernels reproduce the unfused kernel chain (prep + GEMM + BN
    kernels) on a dropout > 0 training step, forward and backward, and the item tower's two passes accumulate."""
    from b200rec.two_tower import ItemTower
    torch.manual_seed(5)
    t = ItemTower(20, embedding_dim=128, hidden_layers=[256, 128], dropout_rate=0.3, use_content_embedding=False).to(DEV)
    t.train()
    x1, x2 = torch.randn(700, 20, device=DEV), torch.randn(1300, 20, device=DEV)
    R1, R2 = torch.randn(700, 128, device=DEV), torch.randn(1300, 128, device=DEV)
    out = {}
    state = {k: v.clone() for k, v in t.state_dict().items()}
    for mode in ("fused", "unfused"):
        monkeypatch.setenv("B200REC_MLP", mode)
        t.load_state_dict(state)
        t._seed_counter = 0
        for p in t.parameters():
            p.grad = None
        a, b = t(x1), t(x2)                     # two passes of one tower (positives, negatives): gradients add up
        ((a * R1).sum() + (b * R2).sum()).backward()
        out[mode] = (a.detach().clone(), b.detach().clone(), {n: p.grad.clone() for n, p in t.named_parameters()},
                     {n: v.clone() for n, v in t.named_buffers()})
    f, u = out["fused"], out["unfused"]
    assert (f[0] - u[0]).abs().max().item() <= 2e-6 and (f[1] - u[1]).abs().max().item() <= 2e-6
    for n in f[2]:
        assert (f[2][n] - u[2][n]).abs().max().item() <= 2e-4 * max(u[2][n].abs().max().item(), 1e-6), n
    for n in f[3]:
        assert torch.allclose(f[3][n].double(), u[3][n].double(), rtol=1e-5, atol=1e-6), n


def test_scatter_add_rows_equals_index_add_with_hot_rows_and_guards():
    """The one-launch dense embedding gradient (atomics) equals index_add_ on valid ids; the padding row and ids outside
    the table contribute nothing and nothing outside the table view is written."""
    from b200rec import kernels as K
    rng = np.random.default_rng(1)
    B, W, rows = 5000, 48, 301
    idx = torch.from_numpy(np.clip(rng.zipf(1.1, size=B), 1, rows - 1).astype(np.int64))
    idx[:50] = 0                                       # padding row
    idx[50:60] = rows + 5                              # out of range
    idx[60:70] = -3
    dY = torch.randn(B, W + 7)
    guard = torch.zeros(rows + 64, W, device=DEV)
    table_grad = guard[32:32 + rows]
    K.scatter_add_rows(idx.to(DEV), dY.to(DEV), W, table_grad, 0)
    valid = (idx > 0) & (idx < rows)
    ref = torch.zeros(rows, W, dtype=torch.float64).index_add_(0, idx[valid], dY[valid][:, :W].double())
    assert (table_grad.cpu().double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()   # fp32 sums of ~2000 terms
    assert guard[:32].abs().sum().item() == 0 and guard[32 + rows:].abs().sum().item() == 0
    assert table_grad[0].abs().sum().item() == 0
