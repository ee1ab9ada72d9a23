"""Baseline multi-GPU training path: row-sharded Linear training with ``torch.distributed`` all_to_all exchanges
driven from the host (ids -> owners, rows back, gradient rows -> owners).  This was the round-1 path; the product
path is ``sharded.ShardedLinearTrainer`` (peer-mapped shards, one persistent kernel, no collective per step).
It stays as (a) the NCCL baseline the peer-mapped kernel is measured against and (b) the device-agnostic statement
of the routing semantics that the gloo world-size-2 CPU test runs.

    owner(row) = row % G, local row = row // G.  Every rank draws its own B samples.  Per step:
      1. route   the rank's 3B lookups are ordered by (owner, id space); counts go round in one small
                 all_to_all, the local-row ids in a second
      2. gather  each OWNER gathers the requested rows + biases of ITS shard and the rows travel back
      3. compute trs_linear_rows_step on the rank's samples (hinge over the GLOBAL batch)
      4. return  gradient rows go to the owners in a fourth all_to_all
      5. update  each owner runs trs_sparse_row_update on what it received (stable sort by row, duplicates summed
                 in (rank, lookup) order, then the optimizer)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------------
# routing
# ------------------------------------------------------------------------------------------------------
@dataclass
class Route:
    order: torch.Tensor        # [n] permutation: lookups sorted by (owner, space)
    send_splits: List[int]     # lookups this rank sends to each owner (host ints)
    recv_splits: List[int]     # lookups each rank sends to THIS owner
    recv_rows: torch.Tensor    # [n_recv] local row ids requested from this owner (rank-major, then space, then lookup order)
    by_space: torch.Tensor     # [n_recv] permutation grouping the received requests by id space (stable)
    space_sizes: List[int]     # requests per id space on this owner (host ints)


def _a2a(out: torch.Tensor, inp: torch.Tensor, out_splits: List[int], in_splits: List[int], group) -> None:
    dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)


def make_route(ids: torch.Tensor, space: torch.Tensor, n_spaces: int, world: int, group=None) -> Route:
    """ids: global row ids of this rank's lookups, space: their id space (0 user, 1 item, ...)."""
    owner = ids % world
    key = owner * n_spaces + space
    order = torch.sort(key, stable=True)[1]
    send_counts = torch.bincount(key, minlength=world * n_spaces).view(world, n_spaces)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    # the one host synchronisation of a step: all_to_all needs its split sizes on the host
    counts = torch.stack([send_counts, recv_counts]).cpu()
    in_splits = counts[0].sum(1).tolist()
    out_splits = counts[1].sum(1).tolist()
    recv_rows = torch.empty(sum(out_splits), dtype=ids.dtype, device=ids.device)
    _a2a(recv_rows, (ids[order] // world).contiguous(), out_splits, in_splits, group)
    # the space of each received request follows from the counts: per source rank, space 0 block then space 1 ...
    recv_space = torch.repeat_interleave(torch.arange(n_spaces).repeat(world), counts[1].reshape(-1))
    by_space = torch.sort(recv_space, stable=True)[1].to(ids.device, non_blocking=True)
    return Route(order, in_splits, out_splits, recv_rows, by_space, counts[1].sum(0).tolist())


def exchange_back(route: Route, payload: torch.Tensor, group=None) -> torch.Tensor:
    """Owner -> requester: payload[n_recv, W] (one row per received request) -> [n, W] in LOOKUP order."""
    got = torch.empty((sum(route.send_splits), payload.shape[1]), dtype=payload.dtype, device=payload.device)
    _a2a(got, payload.contiguous(), route.send_splits, route.recv_splits, group)
    out = torch.empty_like(got)
    out[route.order] = got
    return out


def exchange_forward(route: Route, payload: torch.Tensor, group=None) -> torch.Tensor:
    """Requester -> owner: payload[n, W] in lookup order -> [n_recv, W] aligned with route.recv_rows."""
    got = torch.empty((sum(route.recv_splits), payload.shape[1]), dtype=payload.dtype, device=payload.device)
    _a2a(got, payload[route.order].contiguous(), route.recv_splits, route.send_splits, group)
    return got


# ------------------------------------------------------------------------------------------------------
# row-sharded Linear training
# ------------------------------------------------------------------------------------------------------
class RoutedLinearTrainer:
    """Linear scorer (collaborative/linear.py) with user / item tables row-sharded over the ranks of ``group``.

    ``tables`` = {"user": (emb, bias), "item": (emb, bias)}: this rank's shards, fp32 [ceil(n/G), dim] / [.., 1]
    (rows r*G + rank of the global table).  ``gather / compute / update`` default to the CUDA kernels."""

    def __init__(self, n_users: int, n_items: int, dim: int, optimizer: str = "sparse_adam", lr: float = 1e-3,
                 device=None, group=None, seed: int = 1234, betas=(0.9, 0.999), eps: Optional[float] = None,
                 hooks: Optional[dict] = None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dim, self.n_users, self.n_items = dim, n_users, n_items
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.kind, self.lr, self.betas = optimizer, lr, betas
        self.eps = eps if eps is not None else (1e-8 if optimizer == "sparse_adam" else 1e-10)
        self.step = 0
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.tables, self.state = {}, {}
        for name, n in (("user", n_users), ("item", n_items)):
            rows = (n - self.rank + self.world - 1) // self.world
            # init = the rows this rank owns of the table a single process would draw: N(0, 1/dim), zero biases
            emb = torch.empty((max(rows, 1), dim), dtype=torch.float32, device=self.device)
            emb.normal_(0.0, 1.0 / dim) if self.device.type == "cuda" else emb.copy_(
                torch.randn((max(rows, 1), dim), generator=g) / dim)
            bias = torch.zeros((max(rows, 1), 1), dtype=torch.float32, device=self.device)
            self.tables[name] = (emb, bias)
            n_state = {"sgd": 0, "adagrad": 1, "sparse_adam": 2}[optimizer]
            self.state[name] = [(torch.zeros_like(emb), torch.zeros_like(bias)) for _ in range(n_state)]
        hooks = hooks or {}
        self._gather = hooks.get("gather", self._gather_cuda)
        self._compute = hooks.get("compute", self._compute_cuda)
        self._update = hooks.get("update", self._update_cuda)

    # ---- CUDA hooks ------------------------------------------------------------------------------
    def _gather_cuda(self, name: str, rows: torch.Tensor) -> torch.Tensor:
        from . import _lib
        emb, bias = self.tables[name]
        out = torch.empty((rows.shape[0], self.dim + 1), dtype=torch.float32, device=rows.device)
        if rows.numel():
            out[:, :self.dim] = _lib.embed_gather_sum(emb, rows)
            out[:, self.dim:] = _lib.embed_gather_sum(bias, rows)
        return out

    def _compute_cuda(self, u, vp, vn, inv_batch):
        from . import _lib
        D = self.dim
        g_u, g_vp, g_vn, g_bp, g_bn, hsum = _lib.linear_rows_step(
            u[:, :D].contiguous(), vp[:, :D].contiguous(), vn[:, :D].contiguous(), u[:, D].contiguous(),
            vp[:, D].contiguous(), vn[:, D].contiguous(), inv_batch)
        zero = torch.zeros_like(g_bp)
        pack = lambda g, b: torch.cat([g, b[:, None]], 1)
        return pack(g_u, zero), pack(g_vp, g_bp), pack(g_vn, g_bn), hsum

    def _scale(self) -> float:
        import math
        t = self.step + 1
        if self.kind == "sparse_adam":
            return self.lr * math.sqrt(1 - self.betas[1] ** t) / (1 - self.betas[0] ** t)
        return self.lr

    def _update_cuda(self, name: str, rows: torch.Tensor, grads: torch.Tensor) -> None:
        from . import _lib
        emb, bias = self.tables[name]
        st = self.state[name]
        s0 = st[0] if len(st) > 0 else (None, None)
        s1 = st[1] if len(st) > 1 else (None, None)
        table = _lib.make_table(emb, s0[0], s1[0], bias, s0[1], s1[1])
        kind = {"sgd": _lib.OPT_SGD, "adagrad": _lib.OPT_ADAGRAD, "sparse_adam": _lib.OPT_SPARSE_ADAM}[self.kind]
        scale = torch.tensor([self._scale()], dtype=torch.float64).float().to(rows.device)
        optim = _lib.Optim(kind, 0, self.betas[0], self.betas[1], self.eps, scale.data_ptr())
        g_lin = None if name == "user" else grads[:, self.dim].contiguous()  # d user_bias == 0
        _lib.sparse_row_update(table, self.dim, rows, grads[:, :self.dim].contiguous(), g_lin, optim, 0)

    # ---- one step --------------------------------------------------------------------------------
    def train_step(self, user: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
        """user / pos / neg: this rank's B samples (global ids).  Returns the rank's hinge sum (device scalar);
        the global batch-mean loss is ``all_reduce(sum) / (G*B)``."""
        B = user.shape[0]
        ids = torch.cat([user, pos, neg])
        space = torch.cat([torch.zeros_like(user), torch.ones_like(pos), torch.ones_like(neg)])
        route = make_route(ids, space, 2, self.world, self.group)
        nu = route.space_sizes[0]
        req = route.recv_rows[route.by_space]                       # user requests first, then item requests
        payload = torch.empty((req.shape[0], self.dim + 1), dtype=torch.float32, device=ids.device)
        payload[route.by_space] = torch.cat([self._gather("user", req[:nu]), self._gather("item", req[nu:])])
        rows = exchange_back(route, payload, self.group)            # [3B, dim+1] in lookup order
        g_u, g_vp, g_vn, hsum = self._compute(rows[:B], rows[B:2 * B], rows[2 * B:], 1.0 / (B * self.world))
        grads = exchange_forward(route, torch.cat([g_u, g_vp, g_vn]), self.group)[route.by_space]
        self._update("user", req[:nu], grads[:nu])
        self._update("item", req[nu:], grads[nu:])
        self.step += 1
        return hsum

    # ---- helpers for tests / checkpoints -----------------------------------------------------------
    def gather_full(self, name: str) -> Tuple[torch.Tensor, torch.Tensor]:
        """All-gather a sharded table back into the single-process [n, dim] layout (state_dict interchange)."""
        emb, bias = self.tables[name]
        n = self.n_users if name == "user" else self.n_items
        rows = (n + self.world - 1) // self.world
        pad = lambda t: torch.cat([t, t.new_zeros((rows - t.shape[0],) + t.shape[1:])]) if t.shape[0] < rows else t[:rows]
        embs = [torch.empty((rows, self.dim), dtype=emb.dtype, device=emb.device) for _ in range(self.world)]
        bs = [torch.empty((rows, 1), dtype=bias.dtype, device=bias.device) for _ in range(self.world)]
        dist.all_gather(embs, pad(emb).contiguous(), group=self.group)
        dist.all_gather(bs, pad(bias).contiguous(), group=self.group)
        full_e = torch.stack(embs, 1).reshape(rows * self.world, self.dim)[:n]
        full_b = torch.stack(bs, 1).reshape(rows * self.world, 1)[:n]
        return full_e, full_b


