"""Second-order factorization machine with the reference's module interface
(collaborative/fm.py:5-101): sigmoid(sum_k w_k + 1/2 sum_d [(sum_k e_kd)^2 - sum_k e_kd^2]) over
the fields user, item, metadata_f; returned as (B,).  Parameter names / order as the reference
(``user``, ``item``, ``linear_user``, ``linear_item``, ``metadata.{f}``, ``linear_metadata.{f}``)."""
import torch

from ..embeddings.init_embeddings import ScaledEmbedding
from .. import _lib
from ._base import SparseScorer


class FM(SparseScorer):
    NET = _lib.NET_FM
    USER = ("user", "linear_user")
    ITEM = ("item", "linear_item")
    META = ("metadata", "linear_metadata")

    def __init__(self, n_users, n_items, n_metadata, n_factors, use_metadata=True, use_cuda=False):
        super().__init__()
        self.n_users, self.n_items, self.n_metadata = n_users, n_items, n_metadata
        self.n_factors, self.use_metadata, self.use_cuda = n_factors, use_metadata, use_cuda
        self.n_input = n_users + n_items
        self.user = ScaledEmbedding(n_users, n_factors, sparse=True)
        self.item = ScaledEmbedding(n_items, n_factors, sparse=True)
        self.linear_user = ScaledEmbedding(n_users, 1, sparse=True)
        self.linear_item = ScaledEmbedding(n_items, 1, sparse=True)
        if use_metadata:
            self.n_distinct_metadata = len(n_metadata)
            self.metadata = torch.nn.ModuleList(
                ScaledEmbedding(size, n_factors, sparse=True) for size in n_metadata.values())
            self.linear_metadata = torch.nn.ModuleList(
                ScaledEmbedding(size, 1, sparse=True) for size in n_metadata.values())

    def forward(self, batch, user_key, item_key, metadata_key=None):
        return self._score(batch, user_key, item_key, metadata_key)
