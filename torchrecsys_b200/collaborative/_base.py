"""Shared host logic of the sparse scorers (Linear, FM): batch canonicalisation and the
``trs_model`` view of a module's parameters (+ optimizer state) handed to libtrs_b200."""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from .. import _lib


def canonical_meta(meta: Optional[torch.Tensor], n_meta: int) -> Optional[torch.Tensor]:
    """One id per metadata feature, ``[B, F]`` int64 contiguous.

    The reference loader pads each feature's bag: ``(B, L)`` for one feature, ``(B, F, L)`` for
    several (dataset/dataset.py:281-287), and the scorers only ever read bag element 0
    (linear.py:71-75, fm.py:74-79).  Both layouts and the canonical one are accepted."""
    if n_meta == 0 or meta is None:
        return None
    if meta.dim() == 3:
        meta = meta[:, :, 0]
    elif meta.dim() == 2 and meta.shape[1] != n_meta:
        meta = meta[:, :1]
    elif meta.dim() == 1:
        meta = meta[:, None]
    if meta.shape[1] != n_meta:
        raise ValueError(f"metadata batch has {meta.shape[1]} features, model has {n_meta}")
    return meta.long().contiguous()


class SparseScorer(nn.Module):
    """Base of Linear / FM: owns the tables, builds the C-ABI model view, runs the forward kernel."""

    NET = -1
    # (embedding attr, width-1 attr) of the user and item id spaces; metadata ModuleLists
    USER = ("user", None)
    ITEM = ("item", None)
    META = (None, None)

    def _tables(self):
        def pair(names, f=None):
            emb, lin = names
            get = (lambda a: getattr(self, a)) if f is None else (lambda a: getattr(self, a)[f])
            return get(emb).weight, (get(lin).weight if lin else None)

        user, item = pair(self.USER), pair(self.ITEM)
        metas = [pair(self.META, f) for f in range(self.n_meta_features)] if self.use_metadata else []
        return user, item, metas

    @property
    def n_meta_features(self) -> int:
        return len(self.n_metadata) if (self.use_metadata and self.n_metadata) else 0

    def abi_model(self, state: Optional[Dict[torch.Tensor, dict]] = None, keys=(None, None)) -> _lib.Model:
        """``state`` is ``optimizer.state``; ``keys`` names its (s0, s1) tensors."""

        def table(emb, lin):
            def st(p, k):
                return None if (p is None or state is None or k is None) else state[p][k]
            return _lib.make_table(emb, st(emb, keys[0]), st(emb, keys[1]),
                                   lin, st(lin, keys[0]), st(lin, keys[1]))

        user, item, metas = self._tables()
        return _lib.make_model(self.NET, self.n_factors, table(*user), table(*item),
                               [table(*m) for m in metas])

    def _score(self, batch, user_key, item_key, metadata_key):
        user, item = batch[user_key], batch[item_key]
        if not user.is_cuda:
            raise RuntimeError("torchrecsys_b200 has no CPU fallback: move the model and batch to a "
                               "CUDA device (use_cuda=True)")
        meta = canonical_meta(batch.get(metadata_key) if metadata_key else None, self.n_meta_features)
        if self.n_meta_features and meta is None:
            raise KeyError(f"model uses metadata but batch has no '{metadata_key}'")
        return _lib.scores(self.abi_model(), user.long().contiguous(), item.long().contiguous(), meta)
