"""Tensor-level wrappers over the C-ABI (include/b200rec.h).  Every function launches hand-written sm_100a kernels
asynchronously on torch's current CUDA stream; torch is used only to own device memory."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _native as N

BF16 = torch.bfloat16


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------ operand preparation
def split_bf16(x: torch.Tensor, terms: int = 1, side: int = 0, transpose: bool = False,
               kpad: Optional[int] = None) -> torch.Tensor:
    """fp32 [R, C] -> bf16 operand [R, terms*kpad] (or of x^T when transpose)."""
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    r, c = x.shape
    rows, cols = (c, r) if transpose else (r, c)
    kp = N.pad64(cols) if kpad is None else kpad
    out = torch.empty((rows, terms * kp), dtype=BF16, device=x.device)
    N.check(N.lib().b200rec_split_bf16(N.ptr(x), rows, cols, x.stride(0), int(transpose), N.ptr(out), kp, terms, side,
                                       N.stream()), "split_bf16")
    return out


def normalize_rows(x: torch.Tensor, normalize: bool = True, faiss_rule: bool = False, want_f32: bool = True,
                   want_norms: bool = False, terms: int = 0, side: int = 0):
    """Row L2 normalisation fused with the operand cast.  Returns (y_f32 | None, norms | None, y_bf16 | None)."""
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    r, c = x.shape
    y = torch.empty((r, c), dtype=torch.float32, device=x.device) if want_f32 else None
    nr = torch.empty((r,), dtype=torch.float32, device=x.device) if want_norms else None
    kp = N.pad64(c)
    yb = torch.empty((r, terms * kp), dtype=BF16, device=x.device) if terms else None
    N.check(N.lib().b200rec_normalize_rows(N.ptr(x), r, c, x.stride(0), int(normalize), int(faiss_rule), N.ptr(y),
                                           c, N.ptr(nr), N.ptr(yb), kp, max(terms, 1), side, N.stream()),
            "normalize_rows")
    return y, nr, yb


# ------------------------------------------------------------------------------------------------ GEMM
def gemm_tn(a_op: torch.Tensor, b_op: torch.Tensor, M: int, Nn: int, K: int, bias: Optional[torch.Tensor] = None,
            alpha: float = 1.0, k_splits: int = 1, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C[M,N] = alpha * A[M,K] . B[N,K]^T + bias, bf16 operands, fp32 result."""
    assert a_op.dtype == BF16 and b_op.dtype == BF16
    if out is None:
        out = (torch.zeros if k_splits > 1 else torch.empty)((M, Nn), dtype=torch.float32, device=a_op.device)
    elif k_splits > 1:
        out.zero_()
    N.check(N.lib().b200rec_gemm_bf16_tn(N.ptr(a_op), a_op.stride(0), M, N.ptr(b_op), b_op.stride(0), Nn, K,
                                         N.ptr(out), out.stride(0), N.ptr(bias), float(alpha), k_splits, N.stream()),
            "gemm_bf16_tn")
    return out


def matmul_nt(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, terms: int = 6,
              k_splits: int = 1) -> torch.Tensor:
    """x[M,K] . w[N,K]^T (+bias) through split-bf16 operands (terms=6: fp32-grade; 1: plain bf16)."""
    xo = split_bf16(_f32c(x), terms, 0)
    wo = split_bf16(_f32c(w), terms, 1)
    return gemm_tn(xo, wo, x.shape[0], w.shape[0], xo.shape[1], bias, 1.0, k_splits)


# ------------------------------------------------------------------------------------------------ exact IP top-K
def topk_workspace_bytes(n: int, ld: int, q: int, k: int) -> int:
    return int(N.lib().b200rec_topk_workspace_bytes(n, ld, q, k))


def flat_ip_topk(catalogue: torch.Tensor, queries: torch.Tensor, k: int, row_offset: int = 0,
                 exclude_indptr: Optional[torch.Tensor] = None, exclude_rows: Optional[torch.Tensor] = None,
                 workspace: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """catalogue bf16 [N, ld], queries bf16 [Q, ld] (ld multiple of 64) -> (scores fp32 [Q,k] desc, ids int64 [Q,k])."""
    assert catalogue.dtype == BF16 and queries.dtype == BF16
    assert catalogue.stride(1) == 1 and queries.stride(1) == 1 and catalogue.stride(0) == queries.stride(0)
    n, ld = catalogue.shape[0], catalogue.stride(0)
    q = queries.shape[0]
    need = topk_workspace_bytes(n, ld, q, k)
    if need == 0:
        raise RuntimeError(f"b200rec flat_ip_topk: {N.last_error()}")
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty((need,), dtype=torch.uint8, device=catalogue.device)
    scores = torch.empty((q, k), dtype=torch.float32, device=catalogue.device)
    ids = torch.empty((q, k), dtype=torch.int64, device=catalogue.device)
    N.check(N.lib().b200rec_flat_ip_topk(N.ptr(catalogue), n, ld, N.ptr(queries), q, k, row_offset,
                                         N.ptr(exclude_indptr), N.ptr(exclude_rows), N.ptr(scores), N.ptr(ids),
                                         N.ptr(workspace), workspace.numel(), N.stream()), "flat_ip_topk")
    return scores, ids


def topk_merge(scores: torch.Tensor, ids: torch.Tensor, k_out: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """scores/ids [parts, Q, k_in] -> global top k_out per query (score desc, id asc; id < 0 = padding)."""
    assert scores.dim() == 3 and scores.shape == ids.shape and scores.is_contiguous() and ids.is_contiguous()
    parts, q, k_in = scores.shape
    out_s = torch.empty((q, k_out), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((q, k_out), dtype=torch.int64, device=scores.device)
    N.check(N.lib().b200rec_topk_merge(N.ptr(scores), N.ptr(ids), parts, q, k_in, k_out, N.ptr(out_s), N.ptr(out_i),
                                       N.stream()), "topk_merge")
    return out_s, out_i
