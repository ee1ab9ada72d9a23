"""Embedding tables that own the parameters the CUDA kernels read and update in place.

Mirrors reference embeddings/init_embeddings.py:5-50 (ScaledEmbedding: N(0, 1/D), i.e. std =
1/embedding_dim) and :53-97 (ZeroEmbedding).  They stay real ``nn.Embedding`` modules (fp32,
row-major, ``sparse=True`` at the call sites) so ``parameters()`` / ``state_dict()`` and torch
optimizers bind exactly as with the reference."""
import torch
from torch import nn


class _InitEmbedding(nn.Embedding):
    def _fill(self, w: torch.Tensor) -> None:
        raise NotImplementedError

    def reset_parameters(self) -> None:
        with torch.no_grad():
            self._fill(self.weight)
            if self.padding_idx is not None:
                self.weight[self.padding_idx].zero_()


class ScaledEmbedding(_InitEmbedding):
    def _fill(self, w):
        w.normal_(0.0, 1.0 / self.embedding_dim)


class ZeroEmbedding(_InitEmbedding):
    def _fill(self, w):
        w.zero_()
