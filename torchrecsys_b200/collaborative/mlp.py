"""MLP tower with the reference's module interface (collaborative/mlp.py:7-115):
concat[user, item, metadata_f...] -> (Linear -> BatchNorm1d -> ReLU) x n -> Linear(-> 1), (B, 1).

Parameter / buffer names match the reference (``user``, ``item``, ``metadata_embeddings.{f}``,
``fcs.{l}``, ``bns.{l}``, ``output_layer``).  The dense layers are real ``nn.Linear`` /
``nn.BatchNorm1d`` modules so the weights bind to torch optimizers; the arithmetic runs in
libtrs_b200: tcgen05 GEMMs (bf16 operands, fp32 accumulation) with BatchNorm / ReLU kernels around
them (csrc/gemm.cu, csrc/mlp.cu)."""
from typing import Dict, List, Optional

import torch

from .. import _lib
from ..embeddings.init_embeddings import ScaledEmbedding
from ._base import canonical_meta, check_ids


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

    # ---- views handed to the C ABI -----------------------------------------------------------
    def sparse_parameters(self) -> List[torch.nn.Parameter]:
        out = [self.user.weight, self.item.weight]
        if self.n_meta_features:
            out += [e.weight for e in self.metadata_embeddings]
        return out

    def dense_parameters(self) -> List[torch.nn.Parameter]:
        sparse = {id(p) for p in self.sparse_parameters()}
        return [p for p in self.parameters() if id(p) not in sparse]

    def abi_model(self, state: Optional[Dict[torch.Tensor, dict]] = None, keys=(None, None)) -> _lib.Model:
        def table(p):
            s0 = None if (state is None or keys[0] is None) else state[p][keys[0]]
            s1 = None if (state is None or keys[1] is None) else state[p][keys[1]]
            return _lib.make_table(p, s0, s1)

        metas = [table(e.weight) for e in self.metadata_embeddings] if self.n_meta_features else []
        return _lib.make_model(_lib.NET_MLP, self.n_factors, table(self.user.weight), table(self.item.weight), metas)

    def abi_mlp(self, grads: Optional[Dict[torch.Tensor, torch.Tensor]] = None,
                state: Optional[Dict[torch.Tensor, dict]] = None, key: Optional[str] = None) -> _lib.Mlp:
        """``grads[p]``: fp32 buffer receiving p's gradient; ``state[p][key]``: dense optimizer state."""
        if len(self.hidden_layers) > _lib.MAX_LAYERS:
            raise RuntimeError(f"at most {_lib.MAX_LAYERS} hidden layers are supported")
        m = _lib.Mlp()
        m.n_layers, m.use_bn = len(self.hidden_layers), int(self.use_batch_norm)
        ptr = lambda t: _lib._ptr(t, torch.float32)
        g = (lambda p: ptr(grads[p])) if grads is not None else (lambda p: None)
        s0 = (lambda p: ptr(state[p][key])) if (state is not None and key) else (lambda p: None)
        for l, fc in enumerate(self.fcs):
            m.hidden[l] = fc.out_features
            m.W[l], m.b[l] = ptr(fc.weight), ptr(fc.bias)
            m.dW[l], m.db[l] = g(fc.weight), g(fc.bias)
            m.s0W[l], m.s0b[l] = s0(fc.weight), s0(fc.bias)
            if self.use_batch_norm:
                bn = self.bns[l]
                m.gamma[l], m.beta[l] = ptr(bn.weight), ptr(bn.bias)
                m.running_mean[l], m.running_var[l] = ptr(bn.running_mean), ptr(bn.running_var)
                m.dgamma[l], m.dbeta[l] = g(bn.weight), g(bn.bias)
                m.s0gamma[l], m.s0beta[l] = s0(bn.weight), s0(bn.bias)
        o = self.output_layer
        m.w_out, m.b_out = ptr(o.weight), ptr(o.bias)
        m.dw_out, m.db_out = g(o.weight), g(o.bias)
        m.s0w_out, m.s0b_out = s0(o.weight), s0(o.bias)
        return m

    def forward(self, batch, user_key, item_key, metadata_key=None):
        """(B, 1) scores.  In ``train()`` mode BatchNorm uses this batch's statistics and updates the
        running ones, as one reference forward pass does (mlp.py:107-111)."""
        user, item = batch[user_key], batch[item_key]
        if not user.is_cuda:
            raise RuntimeError("torchrecsys_b200 has no CPU fallback: move the model and batch to a "
                               "CUDA device (use_cuda=True)")
        meta = canonical_meta(batch.get(metadata_key) if metadata_key else None, self.n_meta_features)
        if self.n_meta_features and meta is None:
            raise KeyError(f"model uses metadata but batch has no '{metadata_key}'")
        user, item = user.long().contiguous(), item.long().contiguous()
        if not getattr(self, "_trusted_ids", False):
            check_ids(user, self.n_users, "user")
            check_ids(item, self.n_items, "item")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # A graph is being recorded (TorchRecSys.forward -> hinge_loss -> backward, model.py:171-200): the tower
            # is 4 library GEMMs + BatchNorm, and autograd needs their saved activations -- run the reference's own
            # op sequence (mlp.py:93-113) on the device.  fit() / evaluate() / predict() never come here: they use
            # the fused tcgen05 path below and in trs_mlp_train_steps.
            cols = [self.user(user), self.item(item)]
            for f in range(self.n_meta_features):
                cols.append(self.metadata_embeddings[f](meta[:, f]))
            h = torch.cat(cols, dim=1)
            for l, fc in enumerate(self.fcs):
                h = fc(h)
                if self.use_batch_norm:
                    h = self.bns[l](h)
                h = torch.relu(h)
            return self.output_layer(h)
        stats = self.training and self.use_batch_norm
        out = _lib.mlp_forward(self.abi_model(), self.abi_mlp(), user, item, meta, batch_stats=stats)
        if stats:
            for bn in self.bns:
                bn.num_batches_tracked += 1
        return out.view(-1, 1)
