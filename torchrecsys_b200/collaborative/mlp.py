"""MLP tower with the reference's module interface (collaborative/mlp.py:7-115):
concat[user, item, metadata_f...] -> (Linear -> BatchNorm1d -> ReLU) x n -> Linear(-> 1), (B, 1).

Parameter / buffer names match the reference (``user``, ``item``, ``metadata_embeddings.{f}``,
``fcs.{l}``, ``bns.{l}``, ``output_layer``).  The dense layers are real ``nn.Linear`` /
``nn.BatchNorm1d`` modules so the weights bind to torch optimizers."""
from typing import List, Optional

import torch

from ..embeddings.init_embeddings import ScaledEmbedding


class MLP(torch.nn.Module):
    def __init__(self, n_users, n_items, n_metadata, n_factors, use_metadata=True,
                 use_batch_norm: bool = True, hidden_layers: Optional[List[int]] = None,
                 use_cuda=False):
        super().__init__()
        self.n_users, self.n_items, self.n_metadata = n_users, n_items, n_metadata
        self.n_factors, self.use_metadata, self.use_cuda = n_factors, use_metadata, use_cuda
        self.use_batch_norm = use_batch_norm
        self.hidden_layers = list(hidden_layers) if hidden_layers is not None else [1024, 128]
        self.input_shape = n_factors * 2
        if use_metadata:
            self.n_distinct_metadata = len(n_metadata)
            self.input_shape += n_factors * self.n_distinct_metadata
        self.user = ScaledEmbedding(n_users, n_factors, sparse=True)
        self.item = ScaledEmbedding(n_items, n_factors, sparse=True)
        if use_metadata:
            self.metadata_embeddings = torch.nn.ModuleList(
                ScaledEmbedding(size, n_factors, sparse=True) for size in n_metadata.values())
        self.fcs = torch.nn.ModuleList()
        if use_batch_norm:
            self.bns = torch.nn.ModuleList()
        width = self.input_shape
        for h in self.hidden_layers:
            self.fcs.append(torch.nn.Linear(width, h))
            if use_batch_norm:
                self.bns.append(torch.nn.BatchNorm1d(h))
            width = h
        self.output_layer = torch.nn.Linear(width, 1)

    @property
    def n_meta_features(self) -> int:
        return len(self.n_metadata) if (self.use_metadata and self.n_metadata) else 0

    def forward(self, batch, user_key, item_key, metadata_key=None):
        raise NotImplementedError("the MLP tower's tcgen05 forward is not wired yet")
