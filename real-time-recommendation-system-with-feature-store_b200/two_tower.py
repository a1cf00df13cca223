"""Drop-in B200 mirror of the reference's `src/models/two_tower.py`: UserTower (:12-134), ItemTower (:137-281),
TwoTowerModel (:284-546), create_two_tower_model (:549-595).

Same constructor signatures, attribute names and state-dict keys (`mlp.{0,2,4,...}.*`, `embeddings.<name>.weight`,
`content_projection.{0,3}.*`): the torch sub-modules are kept as PARAMETER CONTAINERS (so `state_dict`, `.to`, the
reference's init and its checkpoints all keep working), but `forward` never calls them — every operation runs in the
hand-written sm_100a kernels of libb200rec.so through `b200rec.ops`.  CUDA tensors only; no CPU path.

`precision`: "fp32" (default; split-bf16 x6 products on tcgen05, reference-grade numerics) or "bf16" (2e-2 budget).
Dropout uses our own Philox stream (bit-parity with torch's generator is impossible in a fused kernel; eval mode and
dropout_rate=0 are exactly comparable).
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from . import kernels as K
from . import ops

logger = logging.getLogger("b200rec")

_TERMS = {"fp32": 6, "bf16": 1}


def _get_activation(activation: str) -> nn.Module:
    activations = {
        "relu": nn.ReLU(),
        "gelu": nn.GELU(),
        "leaky_relu": nn.LeakyReLU(0.1),
        "tanh": nn.Tanh(),
        "sigmoid": nn.Sigmoid(),
    }
    return activations.get(activation, nn.ReLU())  # unknown names fall back to ReLU (two_tower.py:86)


def _act_id(activation: str) -> int:
    return K.ACT_IDS[activation] if activation in ("relu", "gelu", "leaky_relu", "tanh", "sigmoid") else 0


class _TowerBase(nn.Module):
    """Shared machinery of the two towers: embeddings dict + [Linear, act, BatchNorm1d, Dropout]xL + Linear."""

    def _build(self, input_dim: int, embedding_dim: int, hidden_layers: List[int], dropout_rate: float,
               activation: str, categorical_features: Optional[Dict[str, int]],
               embedding_dims: Optional[Dict[str, int]] = None) -> None:
        self.input_dim = input_dim
        self.embedding_dim = embedding_dim
        self.categorical_features = categorical_features or {}
        self.dropout_rate = float(dropout_rate)
        self.activation = activation
        self.precision = "fp32"
        self._seed_counter = 0
        self.embeddings = nn.ModuleDict()
        total_embedding_dim = 0
        for feat_name, cardinality in self.categorical_features.items():
            embed_dim = min(50, (cardinality + 1) // 2)  # the reference's heuristic (two_tower.py:45,175)
            if embedding_dims and feat_name in embedding_dims:
                embed_dim = int(embedding_dims[feat_name])  # B200 extension: 16-byte aligned rows (64 / 128 wide)
            self.embeddings[feat_name] = nn.Embedding(cardinality + 1, embed_dim, padding_idx=0)
            self.embeddings[feat_name].weight._b200_table = True   # FlatAdam: row-flagged dense update (trainer.py)
            total_embedding_dim += embed_dim
        self._total_embedding_dim = total_embedding_dim
        return None

    def _build_mlp(self, total_input_dim: int, hidden_layers: List[int], dropout_rate: float, activation: str,
                   embedding_dim: int) -> None:
        layers: List[nn.Module] = []
        prev_dim = total_input_dim
        for hidden_dim in hidden_layers:
            layers.extend([nn.Linear(prev_dim, hidden_dim), _get_activation(activation), nn.BatchNorm1d(hidden_dim),
                           nn.Dropout(dropout_rate)])
            prev_dim = hidden_dim
        layers.append(nn.Linear(prev_dim, embedding_dim))
        self.mlp = nn.Sequential(*layers)
        self._num_hidden = len(hidden_layers)

    def _init_weights(self) -> None:
        for module in self.modules():
            if isinstance(module, nn.Linear):
                nn.init.xavier_uniform_(module.weight)
                if module.bias is not None:
                    nn.init.zeros_(module.bias)
            elif isinstance(module, nn.Embedding):
                nn.init.normal_(module.weight, mean=0, std=0.01)

    # ------------------------------------------------------------------ kernels
    def _next_seed(self) -> int:
        self._seed_counter += 1
        base = int(torch.initial_seed())
        dp = getattr(self, "dp", None)  # data parallel: every replica draws its own dropout masks
        rank_mix = (dp.rank + 1) * 0xA24BAED4963EE407 if dp is not None else 0
        return (base * 0x9E3779B97F4A7C15 + self._seed_counter * 0xD1B54A32D192ED03 + (id(self) >> 4) + rank_mix) & 0xFFFFFFFFFFFFFFFF

    def _concat_inputs(self, numerical: torch.Tensor, categorical: Optional[Dict[str, torch.Tensor]],
                       extra: Optional[torch.Tensor]) -> torch.Tensor:
        ops.require_cuda(numerical)
        idx, tables = [], []
        if categorical:
            for name, t in categorical.items():      # caller's dict order, as the reference (two_tower.py:113-117)
                if name in self.embeddings:
                    idx.append(t)
                    tables.append(self.embeddings[name].weight)
        feats = numerical
        if idx:
            feats = ops.GatherConcatFn.apply(numerical, len(idx), *idx, *tables)
        if extra is not None:
            feats = torch.cat([feats, extra], dim=-1)  # content branch only (low-traffic path)
        return feats

    def _fused_ok(self) -> bool:
        """The fused per-layer kernels (csrc/mlp_fused.cuh) stage one block's per-feature constants in shared memory
        (<= 512 features) and normalise rows that live in one tile (embedding_dim <= 128); wider towers run the
        unfused kernel chain below (B200REC_MLP=unfused forces it, for A/B checks)."""
        import os
        if os.environ.get("B200REC_MLP") == "unfused":
            return False
        widths = [self.mlp[4 * l].out_features for l in range(self._num_hidden)]
        return all(w <= 512 for w in widths) and self.mlp[4 * self._num_hidden].out_features <= 128

    def _run_mlp(self, x: torch.Tensor) -> torch.Tensor:
        terms = _TERMS[self.precision]
        act = _act_id(self.activation)
        if self._fused_ok():
            if self.training and x.shape[0] < 2 and self._num_hidden > 0:
                raise ValueError(f"Expected more than 1 value per channel when training, got input size "
                                 f"{[x.shape[0], self.mlp[0].out_features]}")
            p = self.dropout_rate if self.training else 0.0
            seeds = [self._next_seed() if p > 0 else 0 for _ in range(self._num_hidden)]
            bns = [self.mlp[4 * l + 2] for l in range(self._num_hidden)]
            params = []
            for l in range(self._num_hidden):
                lin, bn = self.mlp[4 * l], self.mlp[4 * l + 2]
                params += [lin.weight, lin.bias, bn.weight, bn.bias]
            last = self.mlp[4 * self._num_hidden]
            params += [last.weight, last.bias]
            np_fwd = 3 if terms == 6 else 1
            np_bwd = {6: 3, 3: 2, 1: 1}[ops._bwd_terms(terms)]
            spec = ops.MLPSpec(act, self.training, p, seeds, bns, np_fwd, np_bwd, getattr(self, "dp", None))
            return ops.TowerMLPFn.apply(x, spec, *params)
        for l in range(self._num_hidden):
            lin, bn = self.mlp[4 * l], self.mlp[4 * l + 2]
            z = ops.LinearFn.apply(x, lin.weight, lin.bias, terms)
            if self.training and z.shape[0] < 2:
                raise ValueError(f"Expected more than 1 value per channel when training, got input size {list(z.shape)}")
            p = self.dropout_rate if self.training else 0.0
            x = ops.ActBNDropFn.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, act, self.training,
                                      bn.eps, bn.momentum if bn.momentum is not None else 0.1, p,
                                      self._next_seed() if p > 0 else 0, getattr(self, "dp", None))
            if self.training:
                bn.num_batches_tracked += 1
        last = self.mlp[4 * self._num_hidden]
        o = ops.LinearFn.apply(x, last.weight, last.bias, terms)
        return ops.NormalizeFn.apply(o)


class UserTower(_TowerBase):
    """User tower (reference two_tower.py:12-134)."""

    def __init__(self, input_dim: int, embedding_dim: int = 128, hidden_layers: List[int] = [512, 256, 128],
                 dropout_rate: float = 0.2, activation: str = "relu",
                 categorical_features: Optional[Dict[str, int]] = None,
                 embedding_dims: Optional[Dict[str, int]] = None):
        super().__init__()
        self._build(input_dim, embedding_dim, hidden_layers, dropout_rate, activation, categorical_features,
                    embedding_dims)
        self._build_mlp(input_dim + self._total_embedding_dim, hidden_layers, dropout_rate, activation, embedding_dim)
        self._init_weights()

    def forward(self, numerical_features: torch.Tensor,
                categorical_features: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        feats = self._concat_inputs(numerical_features, categorical_features, None)
        return self._run_mlp(feats)


class ItemTower(_TowerBase):
    """Item tower (reference two_tower.py:137-281) with the optional content-embedding projection."""

    def __init__(self, input_dim: int, embedding_dim: int = 128, hidden_layers: List[int] = [512, 256, 128],
                 dropout_rate: float = 0.2, activation: str = "relu",
                 categorical_features: Optional[Dict[str, int]] = None, use_content_embedding: bool = True,
                 content_embedding_dim: int = 768, embedding_dims: Optional[Dict[str, int]] = None):
        super().__init__()
        self._build(input_dim, embedding_dim, hidden_layers, dropout_rate, activation, categorical_features,
                    embedding_dims)
        self.use_content_embedding = use_content_embedding
        total = self._total_embedding_dim
        if use_content_embedding:
            self.content_projection = nn.Sequential(nn.Linear(content_embedding_dim, 256), nn.ReLU(),
                                                    nn.Dropout(dropout_rate), nn.Linear(256, 128))
            total += 128
        self._build_mlp(input_dim + total, hidden_layers, dropout_rate, activation, embedding_dim)
        self._init_weights()

    def forward(self, numerical_features: torch.Tensor,
                categorical_features: Optional[Dict[str, torch.Tensor]] = None,
                content_embeddings: Optional[torch.Tensor] = None) -> torch.Tensor:
        extra = None
        if content_embeddings is not None and self.use_content_embedding:
            terms = _TERMS[self.precision]
            l0, l3 = self.content_projection[0], self.content_projection[3]
            z = ops.LinearFn.apply(content_embeddings, l0.weight, l0.bias, terms)
            p = self.dropout_rate if self.training else 0.0
            h = ops.ActDropFn.apply(z, K.ACT_IDS["relu"], p, self._next_seed() if p > 0 else 0)
            extra = ops.LinearFn.apply(h, l3.weight, l3.bias, terms)
        feats = self._concat_inputs(numerical_features, categorical_features, extra)
        return self._run_mlp(feats)


class TwoTowerModel(nn.Module):
    """Two-Tower model (reference two_tower.py:284-546)."""

    def __init__(self, user_tower: UserTower, item_tower: ItemTower, temperature: float = 0.05,
                 use_bias: bool = True):
        super().__init__()
        self.user_tower = user_tower
        self.item_tower = item_tower
        self.temperature = temperature
        if use_bias:
            self.user_bias = nn.Parameter(torch.zeros(1))
            self.item_bias = nn.Parameter(torch.zeros(1))
        else:
            self.register_parameter("user_bias", None)
            self.register_parameter("item_bias", None)

    # precision applies to both towers and the loss GEMMs
    @property
    def precision(self) -> str:
        return self.user_tower.precision

    @precision.setter
    def precision(self, value: str) -> None:
        if value not in _TERMS:
            raise ValueError(f"precision must be one of {list(_TERMS)}")
        self.user_tower.precision = value
        self.item_tower.precision = value

    def forward(self, user_features: Dict[str, torch.Tensor], item_features: Dict[str, torch.Tensor],
                compute_loss: bool = False,
                negative_items: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        user_embedding = self.user_tower(user_features.get("numerical", torch.empty(0)),
                                         user_features.get("categorical", {}))
        item_embedding = self.item_tower(item_features.get("numerical", torch.empty(0)),
                                         item_features.get("categorical", {}),
                                         item_features.get("content_embeddings", None))
        outputs = {"user_embedding": user_embedding, "item_embedding": item_embedding}
        outputs["similarity"] = self.compute_similarity(user_embedding, item_embedding)
        if compute_loss:
            if negative_items is not None:
                neg = self.item_tower(negative_items.get("numerical", torch.empty(0)),
                                      negative_items.get("categorical", {}),
                                      negative_items.get("content_embeddings", None))
                outputs["loss"] = self.contrastive_loss(user_embedding, item_embedding, neg)
            else:
                outputs["loss"] = self.in_batch_negative_loss(user_embedding, item_embedding)
        return outputs

    def compute_similarity(self, user_embedding: torch.Tensor, item_embedding: torch.Tensor) -> torch.Tensor:
        return ops.RowDotFn.apply(user_embedding, item_embedding, self.user_bias, self.item_bias,
                                  1.0 / self.temperature)

    def contrastive_loss(self, user_embedding: torch.Tensor, pos_item_embedding: torch.Tensor,
                         neg_item_embedding: torch.Tensor) -> torch.Tensor:
        batch_size = user_embedding.shape[0]
        if neg_item_embedding.shape[0] <= batch_size:
            # the reference concatenates a [B,1] with a [B] tensor here and raises (two_tower.py:439-443)
            raise RuntimeError("Tensors must have same number of dimensions: got 2 and 1 "
                               "(contrastive_loss needs more than one negative per sample)")
        return ops.ExplicitCEFn.apply(user_embedding, pos_item_embedding, neg_item_embedding, self.user_bias,
                                      self.item_bias, 1.0 / self.temperature)

    def in_batch_negative_loss(self, user_embedding: torch.Tensor, item_embedding: torch.Tensor) -> torch.Tensor:
        terms = _TERMS[self.precision]
        dp = getattr(self, "dp", None)
        if dp is not None:
            # data parallel: the negatives are the WHOLE global batch, as in a single process (two_tower.py:453-479).
            # Returns this replica's share sum_local(...) / B_global; the replicas' shares add up to the loss.
            b = user_embedding.shape[0]
            items_all = dp.all_gather_rows(item_embedding)
            return ops.InBatchCEFn.apply(user_embedding, items_all, 1.0 / self.temperature, terms, dp.rank * b,
                                         dp.world * b)
        return ops.InBatchCEFn.apply(user_embedding, item_embedding, 1.0 / self.temperature, terms, 0,
                                     user_embedding.shape[0])

    def get_user_embeddings(self, user_features: Dict[str, torch.Tensor]) -> torch.Tensor:
        return self.user_tower(user_features.get("numerical", torch.empty(0)), user_features.get("categorical", {}))

    def get_item_embeddings(self, item_features: Dict[str, torch.Tensor]) -> torch.Tensor:
        return self.item_tower(item_features.get("numerical", torch.empty(0)), item_features.get("categorical", {}),
                               item_features.get("content_embeddings", None))

    def save_model(self, path: str) -> None:
        torch.save({"user_tower_state": self.user_tower.state_dict(),
                    "item_tower_state": self.item_tower.state_dict(),
                    "temperature": self.temperature, "user_bias": self.user_bias, "item_bias": self.item_bias}, path)
        logger.info("Saved model checkpoint to %s", path)

    def load_model(self, path: str) -> None:
        checkpoint = torch.load(path, map_location="cpu", weights_only=False)
        self.user_tower.load_state_dict(checkpoint["user_tower_state"])
        self.item_tower.load_state_dict(checkpoint["item_tower_state"])
        self.temperature = checkpoint["temperature"]
        if checkpoint.get("user_bias") is not None:
            dev = next(self.parameters()).device
            self.user_bias = nn.Parameter(checkpoint["user_bias"].detach().to(dev))
            self.item_bias = nn.Parameter(checkpoint["item_bias"].detach().to(dev))
        logger.info("Loaded model checkpoint from %s", path)


def create_two_tower_model(config: Dict[str, Any]) -> TwoTowerModel:
    """Factory with the reference's config keys and defaults (two_tower.py:549-595)."""
    user_config = config.get("user_tower", {})
    item_config = config.get("item_tower", {})
    user_tower = UserTower(input_dim=user_config.get("input_dim", 50), embedding_dim=config.get("embedding_dim", 128),
                           hidden_layers=user_config.get("hidden_layers", [512, 256, 128]),
                           dropout_rate=user_config.get("dropout_rate", 0.2),
                           activation=user_config.get("activation", "relu"),
                           categorical_features=user_config.get("categorical_features", {}))
    item_tower = ItemTower(input_dim=item_config.get("input_dim", 50), embedding_dim=config.get("embedding_dim", 128),
                           hidden_layers=item_config.get("hidden_layers", [512, 256, 128]),
                           dropout_rate=item_config.get("dropout_rate", 0.2),
                           activation=item_config.get("activation", "relu"),
                           categorical_features=item_config.get("categorical_features", {}),
                           use_content_embedding=item_config.get("use_content_embedding", True))
    model = TwoTowerModel(user_tower=user_tower, item_tower=item_tower, temperature=config.get("temperature", 0.05),
                          use_bias=config.get("use_bias", True))
    logger.info("Created Two-Tower model")
    return model
