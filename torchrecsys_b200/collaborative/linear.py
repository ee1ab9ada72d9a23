"""LightFM-style linear scorer with the reference's module interface (collaborative/linear.py:8-80).

score = <user, item + sum_f metadata_f> + user_bias + item_bias, returned as (B, 1).  Parameter
names and registration order match the reference (``metadata.{f}``, ``user``, ``item``,
``user_bias``, ``item_bias``) so ``state_dict()`` and optimizer param indices are interchangeable."""
import torch

from ..embeddings.init_embeddings import ScaledEmbedding, ZeroEmbedding
from .. import _lib
from ._base import SparseScorer


class Linear(SparseScorer):
    NET = _lib.NET_LINEAR
    USER = ("user", "user_bias")
    ITEM = ("item", "item_bias")
    META = ("metadata", None)

    def __init__(self, n_users, n_items, n_metadata, n_factors, use_metadata=True, use_cuda=False):
        super().__init__()
        self.n_users, self.n_items, self.n_metadata = n_users, n_items, n_metadata
        self.n_factors, self.use_metadata, self.use_cuda = n_factors, use_metadata, use_cuda
        if use_metadata:
            self.metadata = torch.nn.ModuleList(
                ScaledEmbedding(size, n_factors, sparse=True) for size in n_metadata.values())
        self.user = ScaledEmbedding(n_users, n_factors, sparse=True)
        self.item = ScaledEmbedding(n_items, n_factors, sparse=True)
        self.user_bias = ZeroEmbedding(n_users, 1, sparse=True)
        self.item_bias = ZeroEmbedding(n_items, 1, sparse=True)

    def forward(self, batch, user_key, item_key, metadata_key=None):
        return self._score(batch, user_key, item_key, metadata_key).view(-1, 1)
