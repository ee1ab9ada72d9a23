"""Evaluation metrics with the reference's interface (evaluate/metrics.py:6-31)."""
import numpy as np
import torch


class Metrics:
    def hit_rate(self, y_hat, y_pred):
        """Share of rows of ``y_pred`` (recommended ids, [n, k]) containing at least one id of the
        matching row of ``y_hat`` (held-out ids, [n, m], padded with values that never occur)."""
        truth = np.asarray(y_hat.cpu() if torch.is_tensor(y_hat) else y_hat)
        recs = np.asarray(y_pred.cpu() if torch.is_tensor(y_pred) else y_pred)
        hit = (recs[:, :, None] == truth[:, None, :]).any(axis=(1, 2))
        return hit.sum() / recs.shape[0]

    def auc_score(self, positive, negative):
        """Pairwise accuracy ``#(pos > neg) / len(pos)`` (strict), what the reference calls AUC."""
        return (positive > negative).sum() / len(positive)
