"""torch-CPU port of the reference scorers -- the CPU baseline arm.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/cf_oracle.py header).  The reference's
``use_cuda=False`` path is ~300 lines of Python that dispatch to stock ATen CPU kernels
(``nn.Embedding(sparse=True)``, elementwise ops, ``nn.Linear``, ``BatchNorm1d``, autograd,
``torch.optim.{SparseAdam,Adagrad,SGD}``).  ``/root/reference`` does not exist on the GPU
box, so this file restates those modules op for op -- same ATen op stream, same
``state_dict`` keys -- and is what ``bench.py`` times as ``cpu_baseline`` (kind "port")
and under ``--impl reference``.  ``tests/test_oracle_golden.py`` pins it against golden
vectors recorded from the live reference (oracle/make_golden.py).

Reference lines followed: collaborative/linear.py:24-80, collaborative/fm.py:21-101,
collaborative/mlp.py:9-115, embeddings/init_embeddings.py:43-50,90-97, helper/loss.py:5-9,
model.py:171-200 (two passes sharing user ids; zero_grad/backward/step/item).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
from torch import nn


def _embedding(n: int, d: int, zero: bool = False) -> nn.Embedding:
    e = nn.Embedding(n, d, sparse=True)
    with torch.no_grad():
        if zero:
            e.weight.zero_()
        else:
            e.weight.normal_(0, 1.0 / d)
    return e


class LinearPort(nn.Module):
    def __init__(self, n_users, n_items, meta_sizes: Sequence[int], d):
        super().__init__()
        if meta_sizes:
            self.metadata = nn.ModuleList([_embedding(s, d) for s in meta_sizes])
        self.user = _embedding(n_users, d)
        self.item = _embedding(n_items, d)
        self.user_bias = _embedding(n_users, 1, zero=True)
        self.item_bias = _embedding(n_items, 1, zero=True)
        self.n_meta = len(meta_sizes)

    def forward(self, user, item, meta=None):
        v = self.item(item)
        for f in range(self.n_meta):
            v = v + self.metadata[f](meta[:, f])
        dot = (self.user(user) * v).sum(1).view(-1, 1)
        return dot + self.user_bias(user) + self.item_bias(item)


class FMPort(nn.Module):
    def __init__(self, n_users, n_items, meta_sizes: Sequence[int], d):
        super().__init__()
        self.user = _embedding(n_users, d)
        self.item = _embedding(n_items, d)
        self.linear_user = _embedding(n_users, 1)
        self.linear_item = _embedding(n_items, 1)
        if meta_sizes:
            self.metadata = nn.ModuleList([_embedding(s, d) for s in meta_sizes])
            self.linear_metadata = nn.ModuleList([_embedding(s, 1) for s in meta_sizes])
        self.n_meta = len(meta_sizes)

    def forward(self, user, item, meta=None):
        B = user.shape[0]
        fields = [self.user(user), self.item(item)]
        lin = [self.linear_user(user), self.linear_item(item)]
        for f in range(self.n_meta):
            fields.append(self.metadata[f](meta[:, f]))
            lin.append(self.linear_metadata[f](meta[:, f]))
        e = torch.stack(fields, dim=1)
        pair = (e.sum(dim=1).pow(2) - e.pow(2).sum(dim=1)).sum(1) * 0.5
        first = torch.cat(lin, dim=1).sum(1).reshape(B)
        return torch.sigmoid(first + pair)


class MLPPort(nn.Module):
    def __init__(self, n_users, n_items, meta_sizes: Sequence[int], d,
                 hidden: Optional[List[int]] = None, batch_norm: bool = True):
        super().__init__()
        hidden = list(hidden) if hidden is not None else [1024, 128]
        self.user = _embedding(n_users, d)
        self.item = _embedding(n_items, d)
        if meta_sizes:
            self.metadata_embeddings = nn.ModuleList([_embedding(s, d) for s in meta_sizes])
        self.fcs = nn.ModuleList()
        if batch_norm:
            self.bns = nn.ModuleList()
        width = d * (2 + len(meta_sizes))
        for h in hidden:
            self.fcs.append(nn.Linear(width, h))
            if batch_norm:
                self.bns.append(nn.BatchNorm1d(h))
            width = h
        self.output_layer = nn.Linear(width, 1)
        self.n_meta, self.batch_norm = len(meta_sizes), batch_norm

    def forward(self, user, item, meta=None):
        cols = [self.user(user), self.item(item)]
        for f in range(self.n_meta):
            cols.append(self.metadata_embeddings[f](meta[:, f]))
        h = torch.cat(cols, dim=1)
        for l, fc in enumerate(self.fcs):
            h = fc(h)
            if self.batch_norm:
                h = self.bns[l](h)
            h = torch.relu(h)
        return self.output_layer(h)


def hinge(pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
    return torch.clamp(neg - pos + 1.0, min=0.0).mean()


def make_net(net_type: str, n_users, n_items, meta_sizes, d, **kw) -> nn.Module:
    cls = {"linear": LinearPort, "fm": FMPort, "mlp": MLPPort}[net_type]
    return cls(n_users, n_items, list(meta_sizes), d, **kw)


def make_optimizer(kind: str, net: nn.Module, lr: Optional[float] = None):
    if kind == "sparse_adam":
        return torch.optim.SparseAdam(list(net.parameters()), lr=lr or 1e-3)
    if kind == "adagrad":
        return torch.optim.Adagrad(net.parameters(), lr=lr or 1e-2)
    if kind == "sgd":
        return torch.optim.SGD(net.parameters(), lr=lr or 1e-3)
    raise ValueError(kind)


def train_step(net: nn.Module, opt, batch: Dict[str, torch.Tensor]) -> float:
    """Loop body of the reference ``fit`` (model.py:274-284) on an explicit batch."""
    pos = net(batch["user"], batch["pos"], batch.get("pos_meta"))
    neg = net(batch["user"], batch["neg"], batch.get("neg_meta"))
    loss = hinge(pos, neg)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss.item()


def predict_topk(net: nn.Module, user_id: int, n_items: int, k: int, chunk: int = 4096) -> torch.Tensor:
    """Reference ``predict`` without its per-chunk pandas frame (model.py:383-450)."""
    outs = []
    with torch.no_grad():
        for lo in range(0, n_items, chunk):
            items = torch.arange(lo, min(lo + chunk, n_items))
            users = torch.full_like(items, user_id)
            outs.append(net(users, items))
    scores = torch.cat(outs, 0).squeeze().float()
    return torch.sort(scores, descending=True, stable=True)[1][:k]
