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


def check_ids(ids: torch.Tensor, n_rows: int, what: str) -> None:
    """The reference raises IndexError from aten::embedding for an id outside the table; an unchecked id here would be
    an out-of-bounds device read.  One validation kernel + one read-back (this is the user-facing forward, not the
    fused training path, whose splits are validated once in TorchRecSys.__init__)."""
    if ids.numel() == 0:
        return
    bad = torch.zeros(1, dtype=torch.int32, device=ids.device)
    _lib.count_bad_ids(ids, n_rows, bad)
    if int(bad.item()):
        raise IndexError(f"{int(bad.item())} {what} ids fall outside [0, {n_rows})")


class _ScoreFn(torch.autograd.Function):
    """``trs_scores`` with a backward: the gradient of every looked-up row comes back as the sparse COO tensor an
    ``nn.Embedding(sparse=True)`` backward would produce (uncoalesced, one entry per lookup), so ``loss.backward()``
    and any torch optimizer work on the result (reference: model.py:188-200)."""

    @staticmethod
    def forward(ctx, scorer, user, item, meta, *weights):
        ctx.scorer, ctx.n_weights = scorer, len(weights)
        ctx.save_for_backward(user, item, meta if meta is not None else user.new_empty(0))
        ctx.has_meta = meta is not None
        return _lib.scores(scorer.abi_model(), user, item, meta)

    @staticmethod
    def backward(ctx, grad_out):
        scorer = ctx.scorer
        user, item, meta = ctx.saved_tensors
        meta = meta if ctx.has_meta else None
        (u_emb, u_lin), (i_emb, i_lin), metas = scorer._tables()
        g_user, g_item, g_meta, g_lu, g_li, g_lm = _lib.scores_backward(
            scorer.abi_model(), user, item, meta, grad_out.contiguous().float(), u_lin is not None, i_lin is not None,
            [lin is not None for _, lin in metas])

        def coo(ids, values, weight):
            if values.dim() == 1:
                values = values[:, None]
            return torch.sparse_coo_tensor(ids[None], values, weight.shape, check_invariants=False)

        grads = {id(u_emb): coo(user, g_user, u_emb), id(i_emb): coo(item, g_item, i_emb)}
        if u_lin is not None:
            grads[id(u_lin)] = coo(user, g_lu, u_lin)
        if i_lin is not None:
            grads[id(i_lin)] = coo(item, g_li, i_lin)
        for f, (m_emb, m_lin) in enumerate(metas):
            grads[id(m_emb)] = coo(meta[:, f].contiguous(), g_meta[f], m_emb)
            if m_lin is not None:
                grads[id(m_lin)] = coo(meta[:, f].contiguous(), g_lm[f], m_lin)
        out = tuple(grads.get(id(w)) if need else None
                    for w, need in zip(scorer._weights(), ctx.needs_input_grad[4:]))
        return (None, None, None, None) + out


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

    def _weights(self):
        """Every table of the scorer, in a fixed order (the inputs autograd tracks)."""
        (u_emb, u_lin), (i_emb, i_lin), metas = self._tables()
        out = [u_emb, u_lin, i_emb, i_lin]
        for m_emb, m_lin in metas:
            out += [m_emb, m_lin]
        return [w for w in out if w is not None]

    def _score(self, batch, user_key, item_key, metadata_key):
        user, item = batch[user_key], batch[item_key]
        if not user.is_cuda:
            raise RuntimeError("torchrecsys_b200 has no CPU fallback: move the model and batch to a "
                               "CUDA device (use_cuda=True)")
        meta = canonical_meta(batch.get(metadata_key) if metadata_key else None, self.n_meta_features)
        if self.n_meta_features and meta is None:
            raise KeyError(f"model uses metadata but batch has no '{metadata_key}'")
        user, item = user.long().contiguous(), item.long().contiguous()
        if not getattr(self, "_trusted_ids", False):  # AutogradEpochRunner: the splits were validated once, up front
            check_ids(user, self.n_users, "user")
            check_ids(item, self.n_items, "item")
            if meta is not None:
                for f, size in enumerate(self.n_metadata.values()):
                    check_ids(meta[:, f].contiguous(), size, "metadata")
        weights = self._weights()
        if torch.is_grad_enabled() and any(w.requires_grad for w in weights):
            return _ScoreFn.apply(self, user, item, meta, *weights)
        return _lib.scores(self.abi_model(), user, item, meta)
