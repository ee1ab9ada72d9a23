"""Host driver of the fused training path: binds a torch optimizer to the row-wise CUDA update and
runs whole epochs through libtrs_b200 without a per-step host sync.

Replaces, per step, the reference's ``forward x2 -> hinge_loss -> zero_grad -> backward ->
optimizer.step() -> loss.item()`` (model.py:274-284, 188-200)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import _lib


@dataclass
class OptBinding:
    kind: int                 # _lib.OPT_*
    keys: tuple               # names of (s0, s1) in optimizer.state[p]
    lr: float
    beta1: float = 0.0
    beta2: float = 0.0
    eps: float = 0.0
    lr_decay: float = 0.0
    step0: int = 0            # steps already taken (per-parameter `step`, identical for all tables)


def _same_hparams(groups: List[dict], names) -> dict:
    first = {n: groups[0].get(n) for n in names}
    for g in groups[1:]:
        if any(g.get(n) != first[n] for n in names):
            raise NotImplementedError(
                "the fused path needs identical hyper-parameters in every param group; got differing "
                + ", ".join(names))
    return first


def _get_step(state: dict) -> int:
    s = state.get("step", 0)
    return int(s.item()) if torch.is_tensor(s) else int(s)


_ADAM_WARNED = False


def _warn_adam_once() -> None:
    """torch.optim.Adam cannot run on these models at all (sparse gradients); it is accepted and bound to
    the row-wise kernel, whose semantics are SparseAdam's, not dense Adam's -- say so once."""
    global _ADAM_WARNED
    if not _ADAM_WARNED:
        _ADAM_WARNED = True
        import warnings
        warnings.warn("torch.optim.Adam is applied ROW-WISE with torch.optim.SparseAdam semantics: only rows looked up "
                      "in a step are updated (no moment decay of untouched rows), denominator sqrt(v) + eps with the "
                      "bias corrections folded into the step size.  Dense Adam itself raises on sparse gradients.",
                      stacklevel=3)


def bind_optimizer(optimizer: torch.optim.Optimizer, params: List[torch.nn.Parameter]) -> OptBinding:
    """Recognise the optimizer and make sure ``optimizer.state[p]`` holds torch-compatible state
    tensors for every sparse table (created lazily exactly as torch's own ``step()`` would)."""
    groups = optimizer.param_groups
    name = type(optimizer).__name__
    if name in ("SparseAdam", "Adam"):
        if name == "Adam":
            _warn_adam_once()
        hp = _same_hparams(groups, ("lr", "betas", "eps", "maximize", "weight_decay", "amsgrad"))
        if hp.get("weight_decay") or hp.get("amsgrad") or hp.get("maximize"):
            raise NotImplementedError(f"{name}: weight_decay / amsgrad / maximize are not supported on sparse tables")
        for p in params:
            st = optimizer.state[p]
            if "exp_avg" not in st:
                st["step"] = 0 if name == "SparseAdam" else torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        b = OptBinding(_lib.OPT_SPARSE_ADAM, ("exp_avg", "exp_avg_sq"), float(hp["lr"]),
                       float(hp["betas"][0]), float(hp["betas"][1]), float(hp["eps"]))
    elif name == "Adagrad":
        hp = _same_hparams(groups, ("lr", "lr_decay", "eps", "weight_decay", "maximize"))
        if hp.get("weight_decay") or hp.get("maximize"):
            raise NotImplementedError("Adagrad: weight_decay is not compatible with sparse gradients (as in torch)")
        for p in params:
            st = optimizer.state[p]
            if "sum" not in st:  # torch creates it in __init__, keep for safety
                st["step"] = torch.tensor(0.0)
                st["sum"] = torch.full_like(p, groups[0].get("initial_accumulator_value", 0.0))
        b = OptBinding(_lib.OPT_ADAGRAD, ("sum", None), float(hp["lr"]), eps=float(hp["eps"]),
                       lr_decay=float(hp["lr_decay"]))
    elif name == "SGD":
        hp = _same_hparams(groups, ("lr", "momentum", "weight_decay", "nesterov", "maximize", "dampening"))
        if hp.get("momentum") or hp.get("weight_decay") or hp.get("nesterov") or hp.get("maximize"):
            raise NotImplementedError("SGD: momentum / weight_decay / nesterov are not row-sparse; "
                                      "only plain SGD takes the fused path")
        b = OptBinding(_lib.OPT_SGD, (None, None), float(hp["lr"]))
    else:
        raise NotImplementedError(
            f"optimizer {name} has no fused row-wise update; use SparseAdam, Adagrad, SGD (or Adam, "
            "which is applied row-wise to the sparse tables)")
    for p in params:
        if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
            raise RuntimeError("parameters must be contiguous fp32 CUDA tensors")
        for k in b.keys:
            if k is not None and optimizer.state[p][k].device != p.device:
                optimizer.state[p][k] = optimizer.state[p][k].to(p.device)
    steps = {_get_step(optimizer.state[p]) for p in params} if b.kind != _lib.OPT_SGD else {0}
    b.step0 = max(steps)
    return b


def step_scales(b: OptBinding, n_steps: int) -> List[float]:
    """The per-step scalar of trs_optim.step_scale, in python doubles as torch computes it."""
    out = []
    for i in range(n_steps):
        t = b.step0 + i + 1
        if b.kind == _lib.OPT_SPARSE_ADAM:
            out.append(b.lr * math.sqrt(1 - b.beta2 ** t) / (1 - b.beta1 ** t))
        elif b.kind == _lib.OPT_ADAGRAD:
            out.append(b.lr / (1 + (t - 1) * b.lr_decay))
        else:
            out.append(b.lr)
    return out


def advance_steps(optimizer, params, b: OptBinding, n_steps: int) -> None:
    if b.kind == _lib.OPT_SGD:
        return
    for p in params:
        st = optimizer.state[p]
        if torch.is_tensor(st["step"]):
            st["step"] += n_steps
        else:
            st["step"] = st["step"] + n_steps
    b.step0 += n_steps


def _bump_version(net) -> None:
    """The kernels write parameters through raw pointers: torch's tensor version counters do not see it.  Caches keyed
    on "the model has not changed" (the prepared item operand of predict) read this counter as well."""
    net._trs_version = getattr(net, "_trs_version", 0) + 1


PLAN_FUSED_MAX = 16384  # plan.cu FS_MAX: lookups per (step, id space) the single-CTA plan kernel holds


def plan_launches(model, batch_size: int, n_steps: int) -> int:
    """CUDA kernels one trs_plan_build launches (bench.py's gpu_launches)."""
    n_meta = model.n_meta
    if 2 * batch_size <= PLAN_FUSED_MAX:
        return 1  # one CTA per (step, id space): sort + work items + flags
    dirty = 2 if (n_steps > 1 and model.net != _lib.NET_MLP) else 0
    passes = lambda rows: max(1, -(-max(1, (rows - 1).bit_length()) // 8))
    # (histogram, scatter) per radix pass and id space + one work-item scan per space
    return 2 * (passes(model.user.n_rows) + passes(model.item.n_rows)
                + sum(passes(model.meta[f].n_rows) for f in range(n_meta))) + 2 + n_meta + dirty


MAX_STEPS_PER_CALL = 32768  # the plan kernels index steps with gridDim.y (<= 65535)


class _Chunked:
    """An epoch with more steps than one plan launch can index is cut into consecutive runs of whole steps."""

    def run(self, samples: Dict[str, torch.Tensor], batch_size: int) -> torch.Tensor:
        n = samples["user"].shape[0]
        per_call = MAX_STEPS_PER_CALL * batch_size
        if n <= per_call:
            return self._run(samples, batch_size)
        losses = [self._run({k: v[lo:lo + per_call] for k, v in samples.items()}, batch_size)
                  for lo in range(0, n, per_call)]
        return torch.cat(losses)


class EpochRunner(_Chunked):
    """Runs the steps of one epoch's (already ordered) samples through the fused kernel."""

    def __init__(self, net, optimizer):
        self.net = net
        self.optimizer = optimizer
        self.params = [p for p in net.parameters()]
        self.binding = bind_optimizer(optimizer, self.params)
        self.launches = 0  # CUDA kernels launched by this runner (bench.py reports it)

    def reserve(self, n_samples: int, batch_size: int) -> None:
        """Pre-sizes the device memory pool for epochs of n_samples: the plan, its scratch and the training
        workspace are allocated once and returned to torch's caching allocator, so that no epoch -- not even the
        first -- pays a cudaMalloc between its launches."""
        dev = self.params[0].device
        n_samples = min(n_samples, MAX_STEPS_PER_CALL * batch_size)  # longer epochs run in calls of this size
        model = self.net.abi_model(self.optimizer.state, self.binding.keys)
        shape = _lib.Epoch(None, None, None, None, None, n_samples, batch_size, 0)  # the size queries read no ids
        held = [torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=dev)
                for nbytes in _lib.train_buffer_bytes(model, shape)]  # all three alive at once, like in run()
        del held  # back to torch's pool

    def _run(self, samples: Dict[str, torch.Tensor], batch_size: int) -> torch.Tensor:
        """samples: device tensors user/pos/neg[/pos_meta/neg_meta] of one epoch.  Returns the
        per-step batch-mean hinge losses (device, [n_steps]); nothing here syncs with the host."""
        b = self.binding
        dev = samples["user"].device
        model = self.net.abi_model(self.optimizer.state, b.keys)
        epoch = _lib.make_epoch(samples["user"], samples["pos"], samples["neg"],
                                samples.get("pos_meta"), samples.get("neg_meta"), batch_size)
        n = samples["user"].shape[0]
        n_steps = -(-n // batch_size)
        # plan and workspace are allocated once per epoch SHAPE: torch's caching allocator hands the same
        # blocks back every epoch (reserve() takes the cudaMalloc out of the first one as well).  The plan is
        # launched first: the device builds it while the host prepares the rest.
        plan = _lib.plan_build(model, epoch, dev)
        scales = torch.tensor(step_scales(b, n_steps), dtype=torch.float64).to(torch.float32)
        scales = scales.to(dev, non_blocking=True)
        optim = _lib.Optim(b.kind, 0, b.beta1, b.beta2, b.eps, scales.data_ptr())
        ws = _lib.train_workspace(model, epoch, dev)
        loss = torch.empty(n_steps, dtype=torch.float32, device=dev)
        _lib.train_steps(model, epoch, optim, plan, ws, 0, n_steps, loss)
        self.launches += 1 + plan_launches(model, batch_size, n_steps)
        advance_steps(self.optimizer, self.params, b, n_steps)
        _bump_version(self.net)
        return loss


class MlpEpochRunner(_Chunked):
    """EpochRunner for net_type='mlp': every step runs the tower forward/backward on the tcgen05 GEMMs,
    the dense SGD / Adagrad update and the row-wise embedding update inside libtrs_b200
    (trs_mlp_train_steps); the host only launches."""

    def __init__(self, net, optimizer):
        self.net = net
        self.optimizer = optimizer
        self.params = [p for p in net.parameters()]
        self.binding = bind_optimizer(optimizer, self.params)
        if self.binding.kind == _lib.OPT_SPARSE_ADAM:
            raise NotImplementedError(
                "net_type='mlp' has dense parameters: use Adagrad or SGD (torch's SparseAdam rejects dense "
                "gradients and Adam rejects the sparse ones, SURVEY.md D2)")
        # gradient buffers of the dense tower parameters: workspace of the fused step, NOT installed as p.grad
        # (the user's module is left as it is; `dense_grads()` hands them out on request)
        self.grads = {p: torch.zeros_like(p) for p in net.dense_parameters()}
        self.launches = 0

    def dense_grads(self) -> Dict[torch.nn.Parameter, torch.Tensor]:
        """The dense gradients of the LAST step run (read-only view of the step's workspace)."""
        return self.grads

    @staticmethod
    def check_batchnorm_rows(n: int, batch_size: int, width: int) -> None:
        """torch's BatchNorm1d refuses a training batch of one row (torch: nn/functional.py _verify_batch_size); so does
        the reference, at the step that meets it.  Here the epoch is one launch: refuse it up front."""
        if n > 0 and (batch_size == 1 or n % batch_size == 1):
            raise ValueError("Expected more than 1 value per channel when training, got input size "
                             f"torch.Size([1, {width}])")

    def _run(self, samples: Dict[str, torch.Tensor], batch_size: int) -> torch.Tensor:
        b = self.binding
        dev = samples["user"].device
        key = b.keys[0]
        if self.net.use_batch_norm:
            self.check_batchnorm_rows(samples["user"].shape[0], batch_size, self.net.hidden_layers[0])
        model = self.net.abi_model(self.optimizer.state, b.keys)
        mlp = self.net.abi_mlp(self.grads, self.optimizer.state if key else None, key)
        epoch = _lib.make_epoch(samples["user"], samples["pos"], samples["neg"],
                                samples.get("pos_meta"), samples.get("neg_meta"), batch_size)
        n = samples["user"].shape[0]
        n_steps = -(-n // batch_size)
        scales = torch.tensor(step_scales(b, n_steps), dtype=torch.float64).to(torch.float32)
        scales = scales.to(dev, non_blocking=True)
        optim = _lib.Optim(b.kind, 0, b.beta1, b.beta2, b.eps, scales.data_ptr())
        plan = _lib.plan_build(model, epoch, dev)
        ws = _lib.mlp_train_workspace(model, mlp, epoch, dev)
        loss = torch.empty(n_steps, dtype=torch.float32, device=dev)
        _lib.mlp_train_steps(model, mlp, epoch, optim, plan, ws, 0, n_steps, loss)
        nl = len(self.net.hidden_layers)
        bn = int(self.net.use_batch_norm)
        # per step: gather, weight cast, per layer (gemm, bn finalize, bn+relu), hinge, loss; backward per
        # layer (reduce, finalize, apply, db reduce, wgrad gemm, dW reduce, dgrad gemm); dense update, staging,
        # row update
        self.launches += n_steps * (2 + nl * (2 + bn) + 2 + nl * (5 + 2 * bn) + 3)
        advance_steps(self.optimizer, self.params, b, n_steps)
        if self.net.use_batch_norm:
            for m in self.net.bns:
                m.num_batches_tracked += 2 * n_steps  # two forward passes per step (model.py:173-183)
        _bump_version(self.net)
        return loss


class AutogradEpochRunner(_Chunked):
    """ANY torch optimizer (SGD with momentum, a user-defined one, ...): the reference's own loop body
    (model.py:274-284) -- ``forward x2 -> hinge_loss -> zero_grad -> backward -> optimizer.step()`` -- with the scorer
    kernels under autograd (collaborative/_base.py: _ScoreFn; the MLP tower's torch ops) and the sparse gradients
    coalesced before the step, as SURVEY.md §8(b) prescribes for optimizers the fused row-wise update does not know.
    One Python iteration per step; the loss stays on the device (no per-step ``.item()``)."""

    def __init__(self, net, optimizer):
        self.net = net
        self.optimizer = optimizer
        self.params = [p for p in net.parameters()]
        self.launches = 0

    def _run(self, samples: Dict[str, torch.Tensor], batch_size: int) -> torch.Tensor:
        from .helper.loss import hinge_loss
        n = samples["user"].shape[0]
        names = {"user": "user_id", "pos": "pos_item_id", "neg": "neg_item_id", "pos_meta": "pos_metadata_id",
                 "neg_meta": "neg_metadata_id"}
        losses = []
        self.net._trusted_ids = True
        try:
            for lo in range(0, n, batch_size):
                batch = {names[k]: v[lo:lo + batch_size] for k, v in samples.items() if k in names}
                has_meta = "pos_metadata_id" in batch
                pos = self.net.forward(batch, "user_id", "pos_item_id", "pos_metadata_id" if has_meta else None)
                neg = self.net.forward(batch, "user_id", "neg_item_id", "neg_metadata_id" if has_meta else None)
                loss = hinge_loss(pos, neg)
                self.optimizer.zero_grad()
                loss.backward()
                for p in self.params:
                    if p.grad is not None and p.grad.is_sparse:
                        p.grad = p.grad.coalesce()
                self.optimizer.step()
                losses.append(loss.detach())
        finally:
            self.net._trusted_ids = False
        _bump_version(self.net)
        return torch.stack(losses) if losses else torch.empty(0, device=samples["user"].device)


def make_runner(net, optimizer, is_mlp: bool):
    """The fused runner when the optimizer has a row-wise kernel, otherwise the autograd loop (with a warning)."""
    try:
        return (MlpEpochRunner if is_mlp else EpochRunner)(net, optimizer)
    except NotImplementedError as e:
        import warnings
        warnings.warn(f"{e} -- falling back to the per-step autograd loop (forward kernels + torch autograd + "
                      f"{type(optimizer).__name__}.step() on coalesced sparse gradients): correct, but far slower than "
                      "the fused path.", stacklevel=3)
        return AutogradEpochRunner(net, optimizer)
