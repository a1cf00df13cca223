"""torch.autograd glue around the b200rec kernels: each Function's forward AND backward run hand-written sm_100a
kernels through the C-ABI; torch only owns the tensors and the autograd graph.  Shapes follow the reference modules
(src/models/two_tower.py)."""
from __future__ import annotations

import math
import os
from typing import List, Optional

import torch
from torch.autograd import Function

from . import kernels as K

NUM_SMS = 148


def _c32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("b200rec modules run on CUDA (sm_100a) tensors only; there is no CPU path — "
                               "move the model and its inputs to a B200 device")


# ---------------------------------------------------------------------------------------------- index-range errors
# nn.Embedding raises IndexError for an id outside [0, cardinality]; a kernel cannot raise, so every gather ORs into one
# device flag per GPU (field number + 1 of the offending table) and the host reads it at a point where it synchronises
# anyway: TwoTowerTrainer checks it once per epoch, `check_index_errors()` on demand, B200REC_SYNC_CHECKS=1 after every
# gather (debugging).  Offending samples read row 0 in the forward and contribute NO gradient (csrc/embedding.cu).
_ERR_FLAGS = {}


def _err_flag(device: torch.device) -> torch.Tensor:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    t = _ERR_FLAGS.get(key)
    if t is None:
        t = _ERR_FLAGS[key] = torch.zeros((1,), dtype=torch.int32, device=device)
    return t


def check_index_errors(device=None) -> None:
    """Raise IndexError (as nn.Embedding does) if any embedding gather since the last check saw an out-of-range id."""
    for key, t in list(_ERR_FLAGS.items()):
        if device is not None and torch.device(device).index not in (None, key[1]):
            continue
        v = int(t.item())
        if v:
            t.zero_()
            raise IndexError(f"index out of range in self (embedding field #{v - 1} of a b200rec gather received an id "
                             "outside [0, num_embeddings))")


def _bwd_terms(terms: int) -> int:
    """Split-bf16 terms of the GRADIENT GEMMs (dgrad, wgrad, the two in-batch gradient products) when the forward runs
    fp32-grade (6 terms).  3 terms ([h m h] x [m h h]: products accurate to ~2^-17, fp32 accumulation) halve their
    FLOPs and operand traffic; the golden parity tests (per-step losses, embeddings, gradients and parameters against the
    reference at 1e-5) hold with them; the fused in-batch backward recomputes its logits with the same 3 products.
    B200REC_BWD_TERMS=6 restores 6-product dgrad / wgrad GEMMs and a 6-product logits recompute."""
    if terms == 6:
        t = int(os.environ.get("B200REC_BWD_TERMS", "3"))
        return t if t in (3, 6) else 3
    return terms


class LinearFn(Function):
    """z = x W^T + b on tcgen05 (nn.Linear, two_tower.py:62,70).  terms: 6 = fp32-grade split-bf16, 1 = bf16."""

    @staticmethod
    def forward(ctx, x, w, b, terms: int):
        require_cuda(x, w)
        x, w = _c32(x), _c32(w)
        if x.dim() != 2 or x.shape[1] != w.shape[1]:   # nn.Linear's own error for a wrong feature width
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({x.shape[0]}x{x.shape[-1]} and "
                               f"{w.shape[1]}x{w.shape[0]})")
        xo = K.split_bf16(x, terms, 0)
        wo = K.split_bf16(w, terms, 1)
        z = K.gemm_tn(xo, wo, x.shape[0], w.shape[0], xo.shape[1], None if b is None else _c32(b))
        ctx.save_for_backward(x, w)
        ctx.terms = terms
        ctx.has_bias = b is not None
        return z

    @staticmethod
    def backward(ctx, dz):
        x, w = ctx.saved_tensors
        terms = _bwd_terms(ctx.terms)
        dz = _c32(dz)
        B, out_f, in_f = x.shape[0], w.shape[0], w.shape[1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dzo = K.split_bf16(dz, terms, 0)
            wt = K.split_bf16(w, terms, 1, transpose=True)          # operand of W^T: [in, terms*kpad(out)]
            dx = K.gemm_tn(dzo, wt, B, in_f, dzo.shape[1])
        if ctx.needs_input_grad[1]:
            dzt = K.split_bf16(dz, terms, 0, transpose=True)        # [out, terms*kpad(B)]
            xt = K.split_bf16(x, terms, 1, transpose=True)          # [in,  terms*kpad(B)]
            tiles = math.ceil(out_f / 128) * math.ceil(in_f / (64 if in_f <= 64 else 128))
            ks = max(1, min(64, NUM_SMS // tiles, dzt.shape[1] // 64))
            dw = K.gemm_tn(dzt, xt, out_f, in_f, dzt.shape[1], k_splits=ks)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = K.colsum(dz)
        return dx, dw, db, None


class ActBNDropFn(Function):
    """y = Dropout(BatchNorm1d(act(z)))  — the hidden block after each Linear (two_tower.py:60-66).
    `dp` (a dist.DataParallel context, training only): batch statistics over the rows of all replicas."""

    @staticmethod
    def forward(ctx, z, gamma, beta, running_mean, running_var, act: int, training: bool, eps: float, momentum: float,
                drop_p: float, seed: int, dp=None):
        z = _c32(z)
        if dp is not None and training:
            y, mean, invstd = K.bn_forward_dp(z, act, eps, momentum, gamma, beta, running_mean, running_var, drop_p,
                                              seed, dp.reduce_sums, z.shape[0] * dp.world)
        else:
            dp = None
            y, mean, invstd = K.bn_forward(z, act, training, eps, momentum, gamma, beta, running_mean, running_var,
                                           drop_p, seed)
        ctx.save_for_backward(z, mean, invstd, gamma)
        ctx.cfg = (act, training, drop_p, seed, dp)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, mean, invstd, gamma = ctx.saved_tensors
        act, training, drop_p, seed, dp = ctx.cfg
        if dp is not None:
            dz, dgamma, dbeta = K.bn_backward_dp(_c32(dy), z, act, mean, invstd, gamma, drop_p, seed, dp.reduce_sums,
                                                 z.shape[0] * dp.world)
        else:
            dz, dgamma, dbeta = K.bn_backward(_c32(dy), z, act, training, mean, invstd, gamma, drop_p, seed)
        return dz, dgamma, dbeta, None, None, None, None, None, None, None, None, None


def _grad_target(p: torch.Tensor, needed: bool):
    """Where a fused backward kernel accumulates the gradient of parameter p: its pre-allocated .grad (FlatAdam's flat
    buffer, zeroed by zero_grad — the kernel adds with atomics, so a tower that runs twice per step simply adds twice)
    or a fresh zero tensor handed back to autograd.  Returns (buffer, value_to_return_from_backward)."""
    if not needed:
        return None, None
    g = p.grad
    if (isinstance(g, torch.Tensor) and g.shape == p.shape and g.dtype == torch.float32 and g.is_contiguous()
            and getattr(p, "_b200_touch", None) is not None):
        touch = p._b200_touch
        touch[0]._mark(touch[1])
        return g, None
    buf = torch.zeros_like(p, dtype=torch.float32)
    return buf, buf


_WGRAD_STREAMS = {}


def _wgrad_stream(device: torch.device) -> torch.cuda.Stream:
    """One side stream per (device, ambient stream): two towers whose backward passes already run on different streams
    each get their own."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    st = _WGRAD_STREAMS.get(key)
    if st is None:
        st = _WGRAD_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


class MLPSpec:
    """Non-tensor description of one tower MLP call for TowerMLPFn (activation, mode, dropout streams, BN buffers)."""

    def __init__(self, act: int, training: bool, drop_p: float, seeds, bns, np_fwd: int, np_bwd: int, dp=None):
        self.act, self.training, self.drop_p, self.seeds, self.bns = act, training, drop_p, seeds, bns
        self.np_fwd, self.np_bwd, self.dp = np_fwd, np_bwd, dp


class TowerMLPFn(Function):
    """[Linear -> act -> BatchNorm1d -> Dropout] x L -> Linear -> F.normalize (two_tower.py:56-72,128-132) with ONE
    kernel launch per Linear in the forward and two (weight gradient, data gradient) in the backward
    (csrc/mlp_fused.cuh).  Inputs: x, spec, then per hidden layer (W, b, gamma, beta) and (W, b) of the last Linear.
    Parameter gradients are accumulated in place when the optimiser pre-allocated them."""

    @staticmethod
    def forward(ctx, x, spec: MLPSpec, *params):
        require_cuda(x, *params)
        x = _c32(x)
        L = len(spec.bns)
        B, dev = x.shape[0], x.device
        training = spec.training
        dp = spec.dp if training else None
        b_stat = B * (dp.world if dp is not None else 1)
        widths = [params[4 * l].shape[0] for l in range(L)]
        # forward statistics and backward sums of every block in one zero-filled buffer: [fwd_0 | bwd_0 | fwd_1 | ...]
        # (eval mode needs no forward statistics; its backward sums still give dgamma / dbeta)
        scratch = torch.zeros((max(4 * sum(widths), 1),), dtype=torch.float64, device=dev) \
            if (training or any(ctx.needs_input_grad)) else None
        off, fsum, bsum = 0, [], []
        for h in widths:
            fsum.append(scratch[off:off + 2 * h] if training else None)
            bsum.append(scratch[off + 2 * h:off + 4 * h] if scratch is not None else None)
            off += 4 * h
        blocks, lower, inp = [], None, x
        for l in range(L):
            w, b, gamma, beta = params[4 * l:4 * l + 4]
            bn = spec.bns[l]
            z = torch.empty((B, widths[l]), dtype=torch.float32, device=dev)
            K.mlp_forward(inp, lower, w, b, z, spec.np_fwd, spec.act, fsum[l], update_running=training)
            if dp is not None:
                dp.reduce_sums(fsum[l])
            lower = K.BnBlock(z, spec.act, training, fsum[l], gamma, beta, bn.running_mean, bn.running_var,
                              bn.num_batches_tracked, bn.eps, bn.momentum if bn.momentum is not None else 0.1,
                              spec.drop_p, spec.seeds[l], b_stat)
            blocks.append(lower)
            inp = None
        w, b = params[4 * L], params[4 * L + 1]
        e = torch.empty((B, w.shape[0]), dtype=torch.float32, device=dev)
        norms = torch.empty((B,), dtype=torch.float32, device=dev)
        K.mlp_forward(inp, lower, w, b, e, spec.np_fwd, 5, None, normalize=True, norms=norms, update_running=training)
        ctx.save_for_backward(x, e, norms)
        ctx.spec, ctx.params, ctx.blocks, ctx.bsum, ctx.scratch = spec, params, blocks, bsum, scratch
        return e

    @staticmethod
    def backward(ctx, de):
        x, e, norms = ctx.saved_tensors
        spec, params, blocks, bsum = ctx.spec, ctx.params, ctx.blocks, ctx.bsum
        L = len(blocks)
        B, dev = x.shape[0], x.device
        npb = spec.np_bwd
        dp = spec.dp if spec.training else None
        need = ctx.needs_input_grad
        grads = [None] * len(params)
        dy = K.normalize_bwd(_c32(de), e, norms)                      # gradient wrt the last Linear's output
        local = [None] * L
        dx0 = None
        # The weight gradient of a layer and its data gradient both start from dy; only the data gradients form a chain.
        # Weight-gradient launches go to a side stream (fork after dy is ready, join before returning) so the chain
        # dgrad(L) -> dgrad(L-1) -> ... is the critical path and each wgrad fills SMs the 64-CTA dgrad leaves idle.
        main = torch.cuda.current_stream()
        side = _wgrad_stream(dev) if os.environ.get("B200REC_OVERLAP", "1") != "0" else None
        keep = [dy]
        for l in range(L, -1, -1):
            own = blocks[l] if l < L else None
            own_b = bsum[l] if l < L else None
            lower = blocks[l - 1] if l > 0 else None
            base = 4 * l
            w = params[base]
            has_b = params[base + 1] is not None
            dw, grads[base] = _grad_target(w, need[2 + base])
            if dw is None:
                dw = torch.zeros_like(w)
            db, grads[base + 1] = _grad_target(params[base + 1], has_b and need[3 + base])
            dgamma = dbeta = None
            if l < L:
                dgamma, grads[base + 2] = _grad_target(params[base + 2], need[4 + base])
                dbeta, grads[base + 3] = _grad_target(params[base + 3], need[5 + base])
                if (dgamma is None) != (dbeta is None):               # the kernel writes both or none
                    dgamma = dgamma if dgamma is not None else torch.zeros_like(params[base + 2])
                    dbeta = dbeta if dbeta is not None else torch.zeros_like(params[base + 3])
            if side is not None and (dp is None or getattr(dp, "streams_safe", False)):
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    K.mlp_wgrad(dy, own, own_b, (local[l] if l < L else None) if dp is not None else None,
                                x if l == 0 else None, lower, npb, dw, db, dgamma, dbeta)
            else:
                K.mlp_wgrad(dy, own, own_b, local[l] if l < L else None, x if l == 0 else None, lower, npb, dw, db, dgamma,
                            dbeta)
            if l > 0 or need[0]:
                k_in = w.shape[1]
                dx = torch.empty((B, k_in), dtype=torch.float32, device=dev)
                want_sums = l > 0
                K.mlp_dgrad(dy, own, own_b, w, dx, npb, lower if want_sums else None, bsum[l - 1] if want_sums else None)
                if want_sums and dp is not None:
                    local[l - 1] = bsum[l - 1].clone()               # dgamma / dbeta add this replica's part only
                    dp.reduce_sums(bsum[l - 1])
                dy = dx
                keep.append(dx)
                if l == 0:
                    dx0 = dx
        if side is not None and (dp is None or getattr(dp, "streams_safe", False)):
            main.wait_stream(side)
        return (dx0, None, *grads)


class ActDropFn(Function):
    """y = Dropout(act(z)) (content_projection, two_tower.py:184-191)."""

    @staticmethod
    def forward(ctx, z, act: int, drop_p: float, seed: int):
        z = _c32(z)
        ctx.save_for_backward(z)
        ctx.cfg = (act, drop_p, seed)
        return K.act_dropout(z, act, drop_p, seed)

    @staticmethod
    def backward(ctx, dy):
        (z,) = ctx.saved_tensors
        act, drop_p, seed = ctx.cfg
        return K.act_dropout_bwd(_c32(dy), z, act, drop_p, seed), None, None, None


class NormalizeFn(Function):
    """F.normalize(p=2, dim=-1, eps=1e-12) (two_tower.py:132,279)."""

    @staticmethod
    def forward(ctx, o):
        o = _c32(o)
        e, norms, _ = K.normalize_rows(o, normalize=True, faiss_rule=False, want_f32=True, want_norms=True)
        ctx.save_for_backward(e, norms)
        return e

    @staticmethod
    def backward(ctx, de):
        e, norms = ctx.saved_tensors
        return K.normalize_bwd(de, e, norms)


def sparse_slot_map(table: torch.Tensor) -> torch.Tensor:
    """int32 [rows] claim map of a row-sparse table (all zero between launches; 4 B per row next to the row's 4*e B)."""
    slot = getattr(table, "_b200_slot", None)
    if slot is None or slot.device != table.device or slot.shape[0] != table.shape[0]:
        slot = table._b200_slot = torch.zeros((table.shape[0],), dtype=torch.int32, device=table.device)
    return slot


class GatherConcatFn(Function):
    """cat([numerical, emb_f(idx_f) ...], -1) in one kernel (two_tower.py:113-126); backward = coalesced sparse row
    gradients per table, delivered as the dense gradient nn.Embedding(sparse=False) produces (padding row 0 gets none),
    or kept row-sparse on `weight._b200_sparse_grad` when the table opted in (large tables + sparse Adam)."""

    @staticmethod
    def forward(ctx, numerical, n_fields: int, *rest):
        indices = list(rest[:n_fields])
        tables = list(rest[n_fields:])
        require_cuda(numerical, *tables)
        numerical = _c32(numerical)
        B = numerical.shape[0]
        widths = [t.shape[1] for t in tables]
        offs, c = [], numerical.shape[1]
        for w in widths:
            offs.append(c)
            c += w
        idx64 = [i.to(torch.int64).contiguous() for i in indices]
        for f, i in enumerate(idx64):
            require_cuda(i)
            if i.numel() != B:
                raise RuntimeError(f"categorical field #{f}: expected {B} ids (one per row of the numerical block), "
                                   f"got {i.numel()}")
        err = _err_flag(numerical.device)
        out = K.gather_concat(numerical if numerical.shape[1] > 0 else None, tables, idx64, widths, offs, c, B,
                              numerical.device, err)
        if os.environ.get("B200REC_SYNC_CHECKS") == "1":
            check_index_errors(numerical.device)
        ctx.save_for_backward(*idx64)
        ctx.tables = tables
        ctx.meta = (numerical.shape[1], widths, offs)
        return out

    @staticmethod
    def backward(ctx, dout):
        idx64 = ctx.saved_tensors
        num_cols, widths, offs = ctx.meta
        dout = _c32(dout)
        dnum = dout[:, :num_cols] if ctx.needs_input_grad[0] else None
        grads: List[Optional[torch.Tensor]] = []
        for f, table in enumerate(ctx.tables):
            if not ctx.needs_input_grad[2 + len(idx64) + f]:
                grads.append(None)
                continue
            inplace = (isinstance(table.grad, torch.Tensor) and table.grad.shape == table.shape
                       and table.grad.dtype == torch.float32 and table.grad.stride(1) == 1)
            if (inplace and not getattr(table, "_b200_sparse", False)
                    and os.environ.get("B200REC_EMB_BWD", "atomic") != "sorted"):
                # the optimiser pre-allocated (and zeroed) the dense gradient: add the sample rows straight into it,
                # one launch, instead of sort + segment sum + scatter (B200REC_EMB_BWD=sorted keeps the deterministic path)
                ex = getattr(table, "_b200_row_exchange", None)
                if ex is not None and getattr(table, "_b200_touch", None) is not None:
                    # data parallel: every replica adds the sample rows of ALL replicas (instead of all-reducing the
                    # whole dense table gradient afterwards)
                    ids_all, rows_all = ex.gather_sparse(idx64[f], dout[:, offs[f]:offs[f] + widths[f]].contiguous())
                    K.scatter_add_rows(ids_all, rows_all, widths[f], table.grad, 0, getattr(table, "_b200_row_flags", None))
                else:
                    K.scatter_add_rows(idx64[f], dout[:, offs[f]:], widths[f], table.grad, 0,
                                       getattr(table, "_b200_row_flags", None))
                touch = getattr(table, "_b200_touch", None)
                if touch is not None:
                    touch[0]._mark(touch[1], flagged=getattr(table, "_b200_row_flags", None) is not None)
                grads.append(None)
                continue
            if getattr(table, "_b200_sparse", False) and os.environ.get("B200REC_EMB_BWD", "atomic") != "sorted":
                # row-sparse table: claim a leader per distinct row and add every sample's gradient to it (one launch
                # instead of sort + scan + segment sum); duplicates and invalid ids come back as row -1
                rows, vals, n = K.sparse_claim_accumulate(idx64[f], dout[:, offs[f]:], widths[f], table.shape[0],
                                                          sparse_slot_map(table), 0)
                prev = getattr(table, "_b200_sparse_grad", None)
                table._b200_sparse_grad = (prev or []) + [(rows, vals, n)]
                grads.append(None)
                continue
            rows, vals, n = K.embedding_sparse_grad(idx64[f], dout[:, offs[f]:], widths[f], table.shape[0], 0)
            if getattr(table, "_b200_sparse", False):
                prev = getattr(table, "_b200_sparse_grad", None)
                table._b200_sparse_grad = (prev or []) + [(rows, vals, n)]
                grads.append(None)
            elif (isinstance(table.grad, torch.Tensor) and table.grad.shape == table.shape
                  and table.grad.dtype == torch.float32 and table.grad.stride(1) == 1):
                # the optimiser pre-allocated the gradient (FlatAdam's flat buffer, zeroed by zero_grad): add the touched
                # rows in place instead of materialising a dense [rows, e] tensor for autograd to add (for the 1 M-row
                # table of config 2 that was a 256 MB fill + a 770 MB read-modify-write per step)
                K.scatter_rows(rows, vals, n, table.grad, accumulate=True)
                touch = getattr(table, "_b200_touch", None)   # FlatAdam: this table received a gradient this step
                if touch is not None:
                    touch[0]._mark(touch[1])
                grads.append(None)
            else:
                dense = torch.zeros_like(table)
                K.scatter_rows(rows, vals, n, dense)
                grads.append(dense)
        return (dnum, None, *([None] * len(idx64)), *grads)


class RowDotFn(Function):
    """compute_similarity: sum(u*i, -1)/T + user_bias + item_bias (two_tower.py:395-402)."""

    @staticmethod
    def forward(ctx, u, i, user_bias, item_bias, inv_t: float):
        require_cuda(u, i)
        u, i = _c32(u), _c32(i)
        ctx.save_for_backward(u, i)
        ctx.inv_t = inv_t
        ctx.has_bias = user_bias is not None
        return K.rowdot(u, i, inv_t, user_bias, item_bias)

    @staticmethod
    def backward(ctx, g):
        u, i = ctx.saved_tensors
        g = _c32(g)
        du, di = K.rowdot_bwd(g, u, i, ctx.inv_t)
        gb = None
        if ctx.has_bias:
            acc = torch.zeros((1,), dtype=torch.float32, device=g.device)
            K.ce_sum(g, None, acc)
            gb = acc
        return du, di, gb, gb, None


class ExplicitCEFn(Function):
    """contrastive_loss with explicit negatives (two_tower.py:406-451): mean CE over [pos | R negatives], label 0."""

    @staticmethod
    def forward(ctx, u, p, n, user_bias, item_bias, inv_t: float):
        require_cuda(u, p, n)
        u, p, n = _c32(u), _c32(p), _c32(n)
        B = u.shape[0]
        if n.shape[0] % B != 0:   # the reference's neg.view(B, R, -1) raises for this shape (two_tower.py:428)
            raise RuntimeError(f"shape '[{B}, -1, {n.shape[1]}]' is invalid for input of size {n.numel()}")
        R = n.shape[0] // B
        row_loss, *_ = K.explicit_ce(u, p, n, R, inv_t, user_bias, item_bias)
        acc = torch.zeros((1,), dtype=torch.float32, device=u.device)
        K.ce_sum(row_loss, None, acc)
        ctx.save_for_backward(u, p, n, user_bias, item_bias)
        ctx.cfg = (R, inv_t)
        return (acc / B).reshape(())

    @staticmethod
    def backward(ctx, g):
        u, p, n, ub, ib = ctx.saved_tensors
        R, inv_t = ctx.cfg
        B = u.shape[0]
        gdev = _c32(g).reshape(1)
        _, du, dp, dn, dbias = K.explicit_ce(u, p, n, R, inv_t, ub, ib, grad_scale=1.0 / B, grad_scale_dev=gdev,
                                             want_grad=True)
        gb = None
        if ub is not None:
            gb = torch.zeros((1,), dtype=torch.float32, device=u.device)
            K.ce_sum(dbias, None, gb)
        return du, dp, dn, gb, gb, None


def _lse_terms(terms: int) -> int:
    """Piece products of the in-batch LOGITS (forward log-sum-exp and the operands the backward reuses) when the model
    runs fp32-grade (6 terms).  3 products put ~2^-17 * |u||v| / T (<= 1.5e-4 for unit vectors at T = 0.05) of random
    error on a logit; the log-sum-exp is a softmax-weighted average of those errors and the loss a mean over rows:
    measured against fp64 (tools/lse_terms.py: 8 shapes from ML-1M to the data-parallel 4096 x 32768) the loss error
    is the same 7e-8 .. 1.4e-6 relative with 3 and with 6 products (what remains is the fp32 evaluation of lse - pos).
    Half the tensor work of the forward.  B200REC_LSE_TERMS=6, or a 6-product backward, keeps the 6-product operands."""
    if terms == 6 and _bwd_terms(6) == 3 and os.environ.get("B200REC_LSE_TERMS", "3") != "6":
        return 3
    return terms


class InBatchCEFn(Function):
    """in_batch_negative_loss (two_tower.py:453-479): mean_b( logsumexp_j(<u_b,i_j>/T) - <u_b,i_b>/T ).

    Forward: logits GEMM fused with the online log-sum-exp (never materialised); when the split operand is too wide
    for the resident tile the row chunks go through the plain GEMM + row LSE.  Backward: recomputes logits chunk by
    chunk (chunk x NI fp32, L2 sized), turns them into softmax - onehot in place and feeds two GEMMs.
    `diag_offset`: first row of this rank's positives inside `i` (data parallel: i is the all-gathered item batch)."""

    CHUNK = 2048

    @staticmethod
    def forward(ctx, u, i, inv_t: float, terms: int, diag_offset: int, total_rows: int):
        require_cuda(u, i)
        u, i = _c32(u), _c32(i)
        B, NI = u.shape[0], i.shape[0]
        terms = _lse_terms(terms)
        uo = K.split_bf16(u, terms, 0)
        io = K.split_bf16(i, terms, 1)
        lse = K.inbatch_lse(uo, io, B, NI, inv_t)
        if lse is None:
            parts = []
            for r0 in range(0, B, InBatchCEFn.CHUNK):
                r1 = min(B, r0 + InBatchCEFn.CHUNK)
                S = K.gemm_tn(uo[r0:r1], io, r1 - r0, NI, uo.shape[1])
                parts.append(K.lse_rows(S, inv_t, 0, False)[0])
            lse = parts[0] if len(parts) == 1 else torch.cat(parts)
        ipos = i[diag_offset:diag_offset + B] if (diag_offset != 0 or NI != B) else i
        pos = K.rowdot(u, _c32(ipos), inv_t, None, None)
        acc = torch.zeros((1,), dtype=torch.float32, device=u.device)
        K.ce_sum(lse, pos, acc)
        ctx.save_for_backward(u, i, uo, io, lse)
        ctx.cfg = (inv_t, terms, diag_offset, total_rows)
        return (acc / total_rows).reshape(())

    @staticmethod
    def backward(ctx, g):
        u, i, uo, io, lse = ctx.saved_tensors
        inv_t, terms, diag_offset, total_rows = ctx.cfg
        B, NI, E = u.shape[0], i.shape[0], u.shape[1]
        gdev = _c32(g).reshape(1)
        gterms = _bwd_terms(terms)                                    # gradient GEMMs: 3 piece products (1 in bf16 mode)
        # fused flash-style backward (csrc/inbatch_grad.cu): the logits never reach HBM.  The recompute uses the same
        # number of piece products as the gradient GEMMs (3: error ~2^-17 per product before exp(), far inside the
        # gradient tolerance; B200REC_BWD_TERMS=6 recomputes with the 6 fp32-grade products when the tile fits).
        nps = 1 if terms == 1 else (6 if gterms == 6 else 3)
        npg = 1 if terms == 1 else 3
        if os.environ.get("B200REC_INBATCH_BWD") != "chunked" and K.inbatch_grad_supported(B, NI, E, nps, npg):
            du = torch.empty_like(u)
            di = torch.empty_like(i)
            K.inbatch_grad(uo, K.PIECE_BLOCKS[(terms, 0)], io, K.PIECE_BLOCKS[(terms, 1)], B, NI, E, nps, npg, inv_t, lse,
                           diag_offset, inv_t / total_rows, gdev, du, di)
            return du, di, None, None, None, None
        du = torch.empty_like(u)
        di = torch.zeros_like(i)
        it = K.split_bf16(i, gterms, 1, transpose=True)              # operand of I^T: [E, terms*kpad(NI)]
        for r0 in range(0, B, InBatchCEFn.CHUNK):
            r1 = min(B, r0 + InBatchCEFn.CHUNK)
            rows = r1 - r0
            S = K.gemm_tn(uo[r0:r1], io, rows, NI, uo.shape[1])
            K.softmax_grad_(S, inv_t, lse[r0:r1], diag_offset + r0, inv_t / total_rows, gdev)
            go = K.split_bf16(S, gterms, 0)                           # [rows, terms*kpad(NI)]
            # [rows x E] output with K = terms * NI: a handful of tiles and a very long reduction -> split K over the SMs
            # (16 CTAs on 148 SMs ran this GEMM at 115-200 us per chunk)
            tiles = math.ceil(rows / 128) * math.ceil(E / (64 if E <= 64 else 128))
            ks = max(1, min(16, NUM_SMS // tiles, go.shape[1] // 512))
            K.gemm_tn(go, it, rows, E, go.shape[1], k_splits=ks, out=du[r0:r1])
            gt = K.split_bf16(S, gterms, 0, transpose=True)           # [NI, terms*kpad(rows)]
            ut = K.split_bf16(u[r0:r1], gterms, 1, transpose=True)    # [E,  terms*kpad(rows)]
            K.gemm_tn(gt, ut, NI, E, gt.shape[1], k_splits=2, out=di, accumulate=True)
        return du, di, None, None, None, None
