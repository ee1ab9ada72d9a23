"""Pairwise hinge of the reference (helper/loss.py:5-9).  Inside ``fit`` / ``evaluate`` the loss is
fused into the CUDA kernels; this torch expression exists for user code that calls it directly."""
import torch


def hinge_loss(positive, negative):
    return torch.clamp(negative - positive + 1.0, min=0.0).mean()
