"""``TorchRecSys`` -- the reference's orchestrator API (model.py:18-452) on the B200 hot path.

Same constructor, ``fit`` / ``evaluate`` / ``predict`` / ``forward`` / ``backward`` signatures,
prints and ``parameters()`` / ``state_dict()`` surface.  What changes underneath:

* ``fit`` keeps the training split on the device, shuffles there, draws dynamic negatives with the
  Philox kernel and runs each epoch as ONE persistent fused kernel (engine.py); the only host
  sync is the end-of-epoch loss read-back for the reference's progress line.
* ``evaluate`` runs the fused scorer+hinge+pairwise-AUC kernel; ``predict`` scores one user
  against all items on the device.
* there is no CPU path: ``use_cuda=False`` still builds the host objects (data processing,
  parameters) but every compute entry point raises.

Supersets of the reference: ``hidden_layers`` / ``use_batch_norm`` reach the MLP,
``seed`` fixes the Philox key, ``Adam`` is accepted and applied row-wise (D2)."""
from __future__ import annotations

from typing import List, Optional

import pandas as pd
import torch
import torch.profiler

from . import _lib
from .collaborative.fm import FM
from .collaborative.linear import Linear
from .collaborative.mlp import MLP
from .dataset.dataset import FastDataLoader, ProcessData
from .engine import make_runner
from .evaluate.metrics import Metrics
from .helper.cuda import gpu
from .helper.loss import hinge_loss

_NO_CPU = ("torchrecsys_b200 has no CPU fallback: construct TorchRecSys(..., use_cuda=True) on a "
           "machine with a CUDA device (the reference's use_cuda=False path is the CPU oracle)")


class TorchRecSys(torch.nn.Module):
    def __init__(self, dataset: pd.DataFrame, user_id_col: str, item_id_col: str,
                 n_factors: int = 80, net_type: str = "linear",
                 metadata_id_col: Optional[List[str]] = None, split_ratio: float = 0.8,
                 dynamic_neg_sampling: bool = False, use_amp: bool = False, use_cuda: bool = False,
                 debug: bool = False, path: str = "./",
                 hidden_layers: Optional[List[int]] = None, use_batch_norm: bool = True,
                 seed: int = 1234):
        super().__init__()
        self.path, self.debug = path, debug
        self.dynamic_neg_sampling = dynamic_neg_sampling
        self.use_amp, self.use_cuda = use_amp, use_cuda
        self.grad_scaler = None  # bf16 tensor-core path needs no loss scaling
        self.seed = int(seed)
        self._epochs_seen = 0

        self.data_processor = ProcessData(dataset=dataset, user_id_col=user_id_col,
                                          item_id_col=item_id_col, metadata_id_col=metadata_id_col,
                                          split_ratio=split_ratio,
                                          dynamic_neg_sampling=dynamic_neg_sampling)
        self.data_processor.prepare_data()
        self.config = self.data_processor.config
        self.n_users = self.config.get("num_users")
        self.n_items = self.config.get("num_items")
        self.metadata_size = self.config.get("num_metadata")
        self.metadata_name = metadata_id_col if hasattr(self.data_processor, "metadata_id_col") else None
        self.n_factors, self.net_type = n_factors, net_type
        self.use_metadata = bool(self.metadata_name)
        self.hidden_layers, self.use_batch_norm = hidden_layers, use_batch_norm
        self._dev_cache = {}
        self._init_net(net_type)
        if use_cuda:
            self._validate_ids()

    # ------------------------------------------------------------------------------------
    def _init_net(self, net_type: str = "linear") -> None:
        assert net_type in ("linear", "mlp", "neucf", "fm", "lstm"), \
            'Net type must be one of "linear", "mlp", "neu", "ease" or "lstm"'
        common = dict(n_users=self.n_users, n_items=self.n_items, n_metadata=self.metadata_size,
                      n_factors=self.n_factors, use_metadata=self.use_metadata, use_cuda=self.use_cuda)
        if net_type == "linear":
            print("Linear Collaborative Filtering")
            self.net = Linear(**common)
        elif net_type == "mlp":
            print("Multi Layer Perceptron")
            self.net = MLP(hidden_layers=self.hidden_layers, use_batch_norm=self.use_batch_norm, **common)
        elif net_type == "fm":
            print("Factorization Machine")
            self.net = FM(**common)
        else:
            raise NotImplementedError(f"{net_type} is not implemented (nor is it in the reference)")
        self.net = gpu(self.net, self.use_cuda)

    def _require_cuda(self) -> torch.device:
        if not self.use_cuda or not torch.cuda.is_available():
            raise RuntimeError(_NO_CPU)
        return next(self.net.parameters()).device

    def _device_split(self, which: str) -> dict:
        """The train / test split as device tensors with the kernels' key names (cached)."""
        if which not in self._dev_cache:
            dev = self._require_cuda()
            src = getattr(self.data_processor, which)
            names = {"user_id": "user", "pos_item_id": "pos", "neg_item_id": "neg",
                     "pos_metadata_id": "pos_meta", "neg_metadata_id": "neg_meta"}
            self._dev_cache[which] = {names[k]: v.to(dev).contiguous() for k, v in src.items() if k in names}
            if self.data_processor.item_meta is not None and "item_meta" not in self._dev_cache:
                self._dev_cache["item_meta"] = torch.from_numpy(self.data_processor.item_meta).to(dev)
        return self._dev_cache[which]

    def _validate_ids(self) -> None:
        """ids must be dense 0-based; the reference dies with IndexError inside
        aten::embedding, here one kernel checks every split once, up front."""
        dev = self._require_cuda()
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        for which in ("train_data", "test_data"):
            d = self._device_split(which)
            if not d or d["user"].numel() == 0:
                continue
            _lib.count_bad_ids(d["user"], self.n_users, bad)
            _lib.count_bad_ids(d["pos"], self.n_items, bad)
            if "neg" in d:
                _lib.count_bad_ids(d["neg"], self.n_items, bad)
            if "pos_meta" in d:
                for f, size in enumerate(self.metadata_size.values()):
                    _lib.count_bad_ids(d["pos_meta"][:, f].contiguous(), size, bad)
        if int(bad.item()):
            raise IndexError(f"{int(bad.item())} user/item/metadata ids fall outside [0, n): ids must "
                             "be dense 0-based integers, as the reference requires")

    def _with_negatives(self, d: dict, first_index: int) -> dict:
        """Dynamic negatives (dataset.py:435-447) drawn on the device: Philox counter = position
        in the stream of samples this model has consumed."""
        if not self.dynamic_neg_sampling:
            return d
        item_meta = self._dev_cache.get("item_meta") if self.use_metadata else None
        neg, neg_meta = _lib.philox_negatives(self.seed, first_index, d["pos"], self.n_items, item_meta)
        out = dict(d, neg=neg)
        if neg_meta is not None:
            out["neg_meta"] = neg_meta
        return out

    # ------------------------------------------------------------------------------------
    def forward(self, net, batch):
        positive = gpu(net.forward(batch, user_key="user_id", item_key="pos_item_id",
                                   metadata_key="pos_metadata_id"), self.use_cuda)
        negative = gpu(net.forward(batch, user_key="user_id", item_key="neg_item_id",
                                   metadata_key="neg_metadata_id"), self.use_cuda)
        return positive, negative

    def backward(self, loss_value, optimizer):
        optimizer.zero_grad()
        loss_value.backward()
        optimizer.step()
        return loss_value.item()

    # ------------------------------------------------------------------------------------
    def fit(self, optimizer, epochs=10, batch_size=512, profile_epochs: int = 0):
        dev = self._require_cuda()
        base = self._device_split("train_data")
        n = base["user"].shape[0] if base else 0
        # the fused row-wise update for SGD / Adagrad / SparseAdam (/ Adam); the per-step autograd loop for anything
        # else (engine.make_runner)
        runner = make_runner(self.net, optimizer, self.net_type == "mlp")
        self._last_runner = runner
        if n and hasattr(runner, "reserve"):
            runner.reserve(n, batch_size)  # plan / workspace blocks of an epoch of this shape, once
        keys = list(base.keys()) if base else []
        for epoch in range(epochs):
            self.net = self.net.train()

            def one_epoch():
                if n == 0:
                    return 0.0
                # the loader's shuffle (dataset.py:369-373) and batch slicing (:420-427) on the device: a Philox-keyed
                # permutation and ONE gather launch for all id columns -- no host randperm, no H2D copy
                perm = self._epoch_permutation(n, dev)
                samples = dict(zip(keys, _lib.gather_rows([base[k] for k in keys], perm)))
                samples = self._with_negatives(samples, self._epochs_seen * n)
                losses = runner.run(samples, batch_size)
                self._epochs_seen += 1
                return float(losses.mean().item())  # unweighted mean over batches (model.py:287)

            if profile_epochs > 0 and epoch == 0:
                print(f"\n--- Starting Profiling for Epoch {epoch + 1} ---")
                acts = [torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]
                with torch.profiler.profile(activities=acts, record_shapes=True, profile_memory=True,
                                            with_stack=True) as prof:
                    avg_loss = one_epoch()
                print("--- Profiler Results (First Epoch) ---")
                print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=20))
            else:
                avg_loss = one_epoch()
            print(f"|--- Epoch {epoch + 1}/{epochs} --- Training Loss: {avg_loss:.4f}")

    def _epoch_permutation(self, n: int, dev, epoch_index: Optional[int] = None) -> torch.Tensor:
        """The shuffle of the ``epoch_index``-th epoch this model trains (default: the next one): a pure function of
        (seed, epoch index, n), so a run can be replayed."""
        e = self._epochs_seen if epoch_index is None else epoch_index
        seed = (self.seed * 0x9E3779B97F4A7C15 + e + 1) & 0xFFFFFFFFFFFFFFFF
        if n <= (1 << 23):
            return _lib.epoch_shuffle(seed, n, dev)
        g = torch.Generator(device=dev).manual_seed(seed & 0x7FFFFFFFFFFFFFFF)   # beyond one sort call: torch's device sort
        return torch.randperm(n, device=dev, generator=g)

    # ------------------------------------------------------------------------------------
    def evaluate(self, batch_size=512, eval_metrics=["loss", "auc"]):
        self._require_cuda()
        self.net = self.net.eval()
        test = self._device_split("test_data")
        self.last_eval = {}
        if not test or test["user"].numel() == 0:
            print("|--- No test data to evaluate.")
            return
        test = self._with_negatives(test, (1 << 40) + self._epochs_seen * test["user"].shape[0])
        if self.net_type == "mlp":
            # eval-mode BatchNorm is row-independent: score the whole split once, then the per-batch
            # statistics of model.py:315-324
            net = self.net
            pos = _lib.mlp_forward(net.abi_model(), net.abi_mlp(), test["user"], test["pos"], test.get("pos_meta"))
            neg = _lib.mlp_forward(net.abi_model(), net.abi_mlp(), test["user"], test["neg"], test.get("neg_meta"))
            hinge = torch.clamp(neg - pos + 1.0, min=0.0).split(batch_size)
            wins = (pos > neg).float().split(batch_size)
            loss = torch.stack([h.mean() for h in hinge])
            auc = torch.stack([w.mean() for w in wins])
        else:
            epoch = _lib.make_epoch(test["user"], test["pos"], test["neg"], test.get("pos_meta"),
                                    test.get("neg_meta"), batch_size)
            loss, auc, pos, neg = _lib.eval_pairwise(self.net.abi_model(), epoch, want_scores="roc_auc" in eval_metrics)
        results = {"loss": loss, "auc": auc}
        for metric in eval_metrics:
            if metric in results:
                value = float(results[metric].mean().item())  # unweighted mean over batches
            elif metric == "roc_auc":
                # sort-based ROC-AUC over the WHOLE split (not per batch): every positive score against every
                # negative score, ties counted half (csrc/extra.cu: trs_sorted_auc)
                value = self._roc_auc(pos, neg)
            else:
                value = 0
            self.last_eval[metric] = value
            print(f"|--- Testing {metric}: {value:.4f}")

    @staticmethod
    def _roc_auc(pos: torch.Tensor, neg: torch.Tensor) -> float:
        limit = 1 << 22  # trs_sorted_auc sorts at most 2^23 scores per call: beyond that, an evenly strided sample
        if pos.numel() > limit:
            stride = -(-pos.numel() // limit)
            pos, neg = pos.reshape(-1)[::stride].contiguous(), neg.reshape(-1)[::stride].contiguous()
        return float(_lib.sorted_auc(pos.reshape(-1).float(), neg.reshape(-1).float()).item())

    # ------------------------------------------------------------------------------------
    def predict(self, user_id: int, top_k: int = 10, prediction_batch_size: int = 4096):
        """Top-k item ids for one user, best first, as a CPU int64 tensor (model.py:341-452).
        Ties are broken towards the lower item id (the reference's sort is unstable, D10).
        ``prediction_batch_size`` is accepted for compatibility; all items are scored in one pass."""
        dev = self._require_cuda()
        self.net = self.net.eval()
        if not 0 <= int(user_id) < self.n_users:  # the reference: IndexError out of aten::embedding
            raise IndexError(f"user_id {user_id} is outside [0, {self.n_users})")
        if self._topk_kernel_ok(top_k):
            return self.predict_batch(torch.tensor([int(user_id)]), top_k)[0]
        return self._predict_exact(int(user_id), top_k, dev)

    # the tensor-core top-k kernel packs [row, bias] into <= 240 bf16 columns (csrc/topk.cu: check_topk) and keeps
    # at most 128 results per user; everything else takes the exact fp32 path
    TOPK_MAX_FACTORS = 239

    def _topk_kernel_ok(self, top_k: int) -> bool:
        return self.net_type != "mlp" and 1 <= top_k <= 128 and self.n_factors <= self.TOPK_MAX_FACTORS

    def _predict_exact(self, user_id: int, top_k: int, dev) -> torch.Tensor:
        """Score every item in fp32 and sort (stable, descending): the MLP tower, top_k > 128, and users
        the tensor-core path flags as overflowing."""
        items = torch.arange(self.n_items, device=dev)
        users = torch.full_like(items, user_id)
        meta = self._dev_cache.get("item_meta") if self.use_metadata else None
        if self.net_type == "mlp":
            scores = _lib.mlp_forward(self.net.abi_model(), self.net.abi_mlp(), users, items, meta)
        else:
            scores = _lib.scores(self.net.abi_model(), users, items, meta)
        order = torch.sort(scores, descending=True, stable=True)[1]
        return order[:top_k].cpu()

    def predict_batch(self, user_ids, top_k: int = 10, exclude_seen: bool = False) -> torch.Tensor:
        """Top-k item ids for MANY users at once -> CPU int64 ``[n_users, top_k]`` (new entry point; the
        reference API is one user per call, model.py:341).  Linear / FM: user tiles x all items on the
        tensor cores with the top-k fused into the epilogue, exact fp32 re-scoring of the candidates
        (csrc/topk.cu); same indices and tie-break as ``predict``.  ``exclude_seen=True`` leaves out the items a
        user interacted with in the training split (rows with fewer than top_k unseen items are padded with -1)."""
        dev = self._require_cuda()
        self.net = self.net.eval()
        users = torch.as_tensor(user_ids, dtype=torch.int64).reshape(-1)
        if users.numel() and (int(users.min()) < 0 or int(users.max()) >= self.n_users):
            raise IndexError(f"user ids must lie in [0, {self.n_users})")
        users = users.to(dev).contiguous()
        if exclude_seen:
            return self._predict_unseen(users, top_k, dev)
        if not self._topk_kernel_ok(top_k):
            return torch.stack([self._predict_exact(int(u), top_k, dev) for u in users.tolist()]) if users.numel() \
                else torch.empty((0, min(top_k, self.n_items)), dtype=torch.int64)
        if self.use_metadata and "item_meta" not in self._dev_cache:
            self._device_split("train_data")
        meta = self._dev_cache.get("item_meta") if self.use_metadata else None
        if not hasattr(self, "_topk_cache"):
            self._topk_cache = _lib.TopkCache()
        params = list(self.net.parameters())  # unchanged tables -> the prepared bf16 item operand is reused
        version = (getattr(self.net, "_trs_version", 0),) + tuple((p._version, p.data_ptr()) for p in params)
        idx, _, over = _lib.predict_topk(self.net.abi_model(), users, top_k, meta, cache=self._topk_cache,
                                         cache_key=version)
        idx = idx.cpu()
        for q in torch.nonzero(over.cpu()).flatten().tolist():
            idx[q] = self._predict_exact(int(users[q]), top_k, dev)
        if top_k > self.n_items:
            idx = idx[:, :self.n_items]
        return idx

    def _seen_pairs(self, dev):
        """Sorted unique (user * n_items + item) codes of the training split, on the device (cached)."""
        if "seen" not in self._dev_cache:
            tr = self._device_split("train_data")
            self._dev_cache["seen"] = torch.unique(tr["user"] * self.n_items + tr["pos"]) if tr else \
                torch.empty(0, dtype=torch.int64, device=dev)
        return self._dev_cache["seen"]

    def _predict_unseen(self, users: torch.Tensor, top_k: int, dev) -> torch.Tensor:
        """Ask the ranking kernel for top_k + (most items any requested user has seen) candidates, drop the seen ones,
        keep the first top_k -- the order among unseen items is the order of the unfiltered ranking."""
        seen = self._seen_pairs(dev)
        lo = torch.searchsorted(seen, users * self.n_items)
        hi = torch.searchsorted(seen, (users + 1) * self.n_items)
        most = int((hi - lo).max().item()) if users.numel() else 0
        want = min(top_k + most, self.n_items)
        if self._topk_kernel_ok(want):
            cand = self.predict_batch(users.cpu(), want).to(dev)
        else:
            cand = torch.stack([self._predict_exact(int(u), want, dev) for u in users.tolist()]).to(dev) if users.numel() \
                else torch.empty((0, want), dtype=torch.int64, device=dev)
        is_seen = torch.isin(users[:, None] * self.n_items + cand, seen) | (cand < 0)
        order = torch.sort(is_seen.to(torch.int8), dim=1, stable=True)[1]          # unseen first, ranking order kept
        out = torch.gather(torch.where(is_seen, torch.full_like(cand, -1), cand), 1, order)[:, :top_k]
        return out.cpu()
