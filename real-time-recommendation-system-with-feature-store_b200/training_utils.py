"""Mirror of the reference's `src/training/utils.py` (factory :14-71, get_device :74, count_parameters :88,
set_seed :100) on top of the B200 modules."""
from __future__ import annotations

import random
from typing import Any, Dict, Optional

import numpy as np
import torch

from .two_tower import ItemTower, TwoTowerModel, UserTower


def create_two_tower_model_for_training(user_feature_dim: int, item_feature_dim: int,
                                        config: Optional[Dict[str, Any]] = None) -> TwoTowerModel:
    config = config or {}
    embedding_dim = config.get("embedding_dim", 64)
    hidden_layers = config.get("hidden_layers", [128, 64])
    dropout_rate = config.get("dropout_rate", 0.2)
    activation = config.get("activation", "relu")
    temperature = config.get("temperature", 0.1)
    use_bias = config.get("use_bias", True)
    user_tower = UserTower(input_dim=user_feature_dim, embedding_dim=embedding_dim, hidden_layers=hidden_layers,
                           dropout_rate=dropout_rate, activation=activation,
                           categorical_features=config.get("user_categorical_features"),
                           embedding_dims=config.get("embedding_dims"))
    item_tower = ItemTower(input_dim=item_feature_dim, embedding_dim=embedding_dim, hidden_layers=hidden_layers,
                           dropout_rate=dropout_rate, activation=activation,
                           categorical_features=config.get("item_categorical_features"), use_content_embedding=False,
                           embedding_dims=config.get("embedding_dims"))
    model = TwoTowerModel(user_tower=user_tower, item_tower=item_tower, temperature=temperature, use_bias=use_bias)
    if config.get("sparse_tables"):
        # B200 extension (hard part 5): id-embedding tables train ROW-SPARSE — Adam moments and
        # updates only on the rows a step touched, no weight decay on untouched rows — instead of the reference's dense
        # Adam over every row (28 B x 6.4 G parameters per step for the 50 M-row table of BASELINE config 4)
        for tower in (user_tower, item_tower):
            for emb in tower.embeddings.values():
                emb.weight._b200_sparse = True
    return model


def get_device(prefer_gpu: bool = True) -> str:
    if not torch.cuda.is_available():
        raise RuntimeError("b200rec has no CPU path: a CUDA (sm_100a) device is required")
    return "cuda"


def count_parameters(model: torch.nn.Module) -> int:
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def set_seed(seed: int = 42) -> None:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
