"""CPU ORACLE (test infrastructure, NOT product code) — two-tower forward, losses and gradients.

A numpy restatement of the reference's training hot path, each function citing the lines it follows:
  * tower forward  : src/models/two_tower.py:98-134 (UserTower), :238-281 (ItemTower)
  * layer stack    : src/models/two_tower.py:56-72  (Linear -> act -> BatchNorm1d -> Dropout, then Linear)
  * similarity     : src/models/two_tower.py:380-404
  * explicit loss  : src/models/two_tower.py:406-451
  * in-batch loss  : src/models/two_tower.py:453-479
  * mixed loss     : src/training/trainers/two_tower.py:134  (0.7 explicit + 0.3 in-batch)
Gradients are the analytic derivatives of those formulas (what torch autograd computes for the
reference).  PINNED: ``tests/golden/make_golden.py`` runs the imported reference (torch CPU) on seeded
inputs and stores inputs/outputs/grads under ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks this
module against every stored vector.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may
import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
NORM_EPS = 1e-12

try:  # scipy is in the image; fall back to math.erf vectorised
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


def act_fwd(name: str, z: np.ndarray) -> np.ndarray:
    """two_tower.py:77-86 — unknown names fall back to ReLU."""
    if name == "gelu":
        return 0.5 * z * (1.0 + _erf(z / math.sqrt(2.0)))
    if name == "leaky_relu":
        return np.where(z > 0, z, 0.1 * z)
    if name == "tanh":
        return np.tanh(z)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-z))
    return np.maximum(z, 0)


def act_bwd(name: str, z: np.ndarray, da: np.ndarray) -> np.ndarray:
    if name == "gelu":
        cdf = 0.5 * (1.0 + _erf(z / math.sqrt(2.0)))
        pdf = np.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)
        return da * (cdf + z * pdf)
    if name == "leaky_relu":
        return da * np.where(z > 0, 1.0, 0.1)
    if name == "tanh":
        t = np.tanh(z)
        return da * (1 - t * t)
    if name == "sigmoid":
        s = 1.0 / (1.0 + np.exp(-z))
        return da * s * (1 - s)
    return da * (z > 0)


class TowerOracle:
    """One tower.  ``params`` uses the reference state_dict keys (mlp.{0,4,8,..}.weight [out,in], .bias,
    mlp.{2,6,..}.weight/.bias/.running_mean/.running_var, embeddings.<name>.weight)."""

    def __init__(self, params: Dict[str, np.ndarray], num_hidden: int, activation: str = "relu",
                 dropout_rate: float = 0.0, dtype=np.float64):
        self.p = {k: np.array(v, dtype=dtype) for k, v in params.items() if "num_batches" not in k}
        self.L = num_hidden
        self.act = activation
        self.drop = dropout_rate
        self.dtype = dtype
        self.cache = None

    # ---- forward (two_tower.py:112-132) -------------------------------------------------------
    def forward(self, numerical: np.ndarray, categorical: Optional[Dict[str, np.ndarray]] = None,
                training: bool = True, dropout_masks: Optional[List[np.ndarray]] = None,
                update_running: bool = False, content: Optional[np.ndarray] = None) -> np.ndarray:
        x = np.asarray(numerical, dtype=self.dtype)
        fields = []
        embs = []
        if categorical:
            for name, idx in categorical.items():  # caller's dict order (two_tower.py:116)
                key = f"embeddings.{name}.weight"
                if key in self.p:
                    embs.append(self.p[key][np.asarray(idx)])
                    fields.append((name, np.asarray(idx), self.p[key].shape[1]))
        self.content_cache = None
        if content is not None and "content_projection.0.weight" in self.p:
            # ItemTower content branch (two_tower.py:184-191 Linear -> ReLU -> Dropout -> Linear, appended after the
            # categorical embeddings :264-266); dropout 0 / eval
            c = np.asarray(content, dtype=self.dtype)
            zc = c @ self.p["content_projection.0.weight"].T + self.p["content_projection.0.bias"]
            h = np.maximum(zc, 0)
            embs.append(h @ self.p["content_projection.3.weight"].T + self.p["content_projection.3.bias"])
            self.content_cache = (c, zc, h)
        if embs:
            x = np.concatenate([x] + embs, axis=-1)
        layers = []
        for l in range(self.L):
            W, b = self.p[f"mlp.{4*l}.weight"], self.p[f"mlp.{4*l}.bias"]
            g, be = self.p[f"mlp.{4*l+2}.weight"], self.p[f"mlp.{4*l+2}.bias"]
            z = x @ W.T + b
            a = act_fwd(self.act, z)
            if training:
                mu = a.mean(0)
                var = a.var(0)  # biased, used for normalisation
                if update_running:
                    n = a.shape[0]
                    rm, rv = f"mlp.{4*l+2}.running_mean", f"mlp.{4*l+2}.running_var"
                    self.p[rm] = (1 - BN_MOMENTUM) * self.p[rm] + BN_MOMENTUM * mu
                    self.p[rv] = (1 - BN_MOMENTUM) * self.p[rv] + BN_MOMENTUM * var * n / max(n - 1, 1)
            else:
                mu, var = self.p[f"mlp.{4*l+2}.running_mean"], self.p[f"mlp.{4*l+2}.running_var"]
            inv = 1.0 / np.sqrt(var + BN_EPS)
            xhat = (a - mu) * inv
            y = xhat * g + be
            mask = None
            if training and self.drop > 0:
                assert dropout_masks is not None, "oracle needs explicit dropout masks when p>0"
                mask = dropout_masks[l].astype(self.dtype) / (1.0 - self.drop)
                y = y * mask
            layers.append((x, z, a, xhat, inv, mask))
            x = y
        W, b = self.p[f"mlp.{4*self.L}.weight"], self.p[f"mlp.{4*self.L}.bias"]
        o = x @ W.T + b
        nrm = np.maximum(np.sqrt((o * o).sum(-1, keepdims=True)), NORM_EPS)  # F.normalize eps
        e = o / nrm
        self.cache = (layers, x, o, nrm, e, fields, np.asarray(numerical).shape[1], training)
        return e

    # ---- backward -------------------------------------------------------------------------------
    def backward(self, de: np.ndarray) -> Tuple[Dict[str, np.ndarray], np.ndarray]:
        layers, xl, o, nrm, e, fields, num_dim, training = self.cache
        de = np.asarray(de, dtype=self.dtype)
        grads: Dict[str, np.ndarray] = {}
        do = (de - e * (e * de).sum(-1, keepdims=True)) / nrm
        W = self.p[f"mlp.{4*self.L}.weight"]
        grads[f"mlp.{4*self.L}.weight"] = do.T @ xl
        grads[f"mlp.{4*self.L}.bias"] = do.sum(0)
        dx = do @ W
        for l in reversed(range(self.L)):
            x, z, a, xhat, inv, mask = layers[l]
            g = self.p[f"mlp.{4*l+2}.weight"]
            dy = dx * mask if mask is not None else dx
            grads[f"mlp.{4*l+2}.weight"] = (dy * xhat).sum(0)
            grads[f"mlp.{4*l+2}.bias"] = dy.sum(0)
            dxhat = dy * g
            if training:
                n = a.shape[0]
                da = inv / n * (n * dxhat - dxhat.sum(0) - xhat * (dxhat * xhat).sum(0))
            else:
                da = dxhat * inv
            dz = act_bwd(self.act, z, da)
            Wl = self.p[f"mlp.{4*l}.weight"]
            grads[f"mlp.{4*l}.weight"] = dz.T @ x
            grads[f"mlp.{4*l}.bias"] = dz.sum(0)
            dx = dz @ Wl
        # embedding tables: dense [card+1, e] grad, padding_idx 0 row zeroed (two_tower.py:46-50)
        off = num_dim
        for name, idx, width in fields:
            key = f"embeddings.{name}.weight"
            gt = grads.get(key, np.zeros_like(self.p[key]))
            np.add.at(gt, idx, dx[:, off:off + width])
            gt[0] = 0
            grads[key] = gt
            off += width
        if self.content_cache is not None:       # content branch: the last slice of the concatenated input
            c, zc, h = self.content_cache
            W3 = self.p["content_projection.3.weight"]
            dproj = dx[:, off:off + W3.shape[0]]
            grads["content_projection.3.weight"] = dproj.T @ h
            grads["content_projection.3.bias"] = dproj.sum(0)
            dzc = (dproj @ W3) * (zc > 0)
            grads["content_projection.0.weight"] = dzc.T @ c
            grads["content_projection.0.bias"] = dzc.sum(0)
        return grads, dx[:, :num_dim]


# ---- losses -------------------------------------------------------------------------------------
def _logsumexp(x: np.ndarray, axis: int) -> np.ndarray:
    m = x.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(x - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def in_batch_loss(u: np.ndarray, i: np.ndarray, temperature: float, want_grad: bool = False):
    """two_tower.py:467-479: CE(U I^T / T, arange(B)), mean."""
    B = u.shape[0]
    S = (u @ i.T) / temperature
    lse = _logsumexp(S, 1)
    loss = float((lse - np.diag(S)).mean())
    if not want_grad:
        return loss
    P = np.exp(S - lse[:, None])
    P[np.arange(B), np.arange(B)] -= 1.0
    dS = P / B / temperature
    return loss, dS @ i, dS.T @ u


def explicit_loss(u, pos, neg, temperature, user_bias=0.0, item_bias=0.0, want_grad=False):
    """two_tower.py:422-451: logits = [pos_sim + biases, neg_sim (no bias)], CE with label 0, mean.
    neg is [B*R, E] ordered b-major (trainers/two_tower.py:116-117)."""
    B, E = u.shape
    R = neg.shape[0] // B
    n3 = neg.reshape(B, R, E)
    ps = (u * pos).sum(-1) / temperature + user_bias + item_bias
    ns = (u[:, None, :] * n3).sum(-1) / temperature
    logits = np.concatenate([ps[:, None], ns], axis=1)
    lse = _logsumexp(logits, 1)
    loss = float((lse - logits[:, 0]).mean())
    if not want_grad:
        return loss
    P = np.exp(logits - lse[:, None])
    P[:, 0] -= 1.0
    dl = P / B
    dps, dns = dl[:, 0], dl[:, 1:]
    du = (dps[:, None] * pos + (dns[:, :, None] * n3).sum(1)) / temperature
    dpos = dps[:, None] * u / temperature
    dneg = (dns[:, :, None] * u[:, None, :] / temperature).reshape(B * R, E)
    dbias = float(dps.sum())
    return loss, du, dpos, dneg, dbias


def mixed_loss(u, pos, neg, temperature, user_bias=0.0, item_bias=0.0):
    """trainers/two_tower.py:124-134."""
    return 0.7 * explicit_loss(u, pos, neg, temperature, user_bias, item_bias) + \
        0.3 * in_batch_loss(u, pos, temperature)


def embedding_touched_rows(idx: np.ndarray) -> np.ndarray:
    """Rows of a table that receive a non-zero-able gradient: unique(idx) minus padding row 0
    (nn.Embedding(padding_idx=0) dense backward, two_tower.py:46-50)."""
    u = np.unique(np.asarray(idx).reshape(-1))
    return u[u != 0].astype(np.int64)
