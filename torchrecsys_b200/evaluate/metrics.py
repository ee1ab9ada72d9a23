"""Evaluation metrics with the reference's interface (evaluate/metrics.py:6-31)."""
import torch


class Metrics:
    def hit_rate(self, y_hat, y_pred):
        """Share of rows of ``y_pred`` (recommended ids, [n, k]) containing at least one id of the
        matching row of ``y_hat`` (held-out ids, [n, m], padded with values that never occur) --
        evaluate/metrics.py:6-20, with tensor ops on whatever device the inputs live on (no numpy round trip:
        ``predict_batch`` output can stay on the GPU)."""
        truth = torch.as_tensor(y_hat)
        recs = torch.as_tensor(y_pred).to(truth.device)
        if recs.shape[0] == 0:
            return float("nan")
        hit = (recs[:, :, None] == truth[:, None, :]).flatten(1).any(dim=1)
        return float(hit.sum().item()) / recs.shape[0]

    def roc_auc(self, positive, negative):
        """Sort-based ROC-AUC of positive vs negative scores on the device (csrc/extra.cu: trs_sorted_auc): the
        probability that a random positive outranks a random negative, ties counted half -- what
        ``sklearn.metrics.roc_auc_score`` returns (the reference's legacy helper/evaluate.py:8-18 calls it)."""
        from .. import _lib
        return float(_lib.sorted_auc(positive.float(), negative.float()).item())

    def auc_score(self, positive, negative):
        """Pairwise accuracy ``#(pos > neg) / len(pos)`` (strict), what the reference calls AUC."""
        return (positive > negative).sum() / len(positive)

    def precision_recall_at_k(self, user_ids, recommended, truth_users, truth_items):
        """Mean precision@k and recall@k over the scored users that have at least one held-out item
        (the reference's legacy ``precision_recall_k``, helper/evaluate.py:53-76: per user
        ``|truth ∩ top-k| / k`` and ``|truth ∩ top-k| / |truth|``), computed with tensor ops on the device the
        inputs live on -- ``recommended`` is what ``TorchRecSys.predict_batch`` returns.

        user_ids [Q]: user of each row of ``recommended`` [Q, k] (item ids, -1 = padding);
        truth_users / truth_items [N]: the interaction pairs that count as hits."""
        user_ids = torch.as_tensor(user_ids).long()
        recommended = torch.as_tensor(recommended).long().to(user_ids.device)
        truth_users = torch.as_tensor(truth_users).long().to(user_ids.device)
        truth_items = torch.as_tensor(truth_items).long().to(user_ids.device)
        k = recommended.shape[1]
        m = int(max(int(recommended.max()) if recommended.numel() else 0,
                    int(truth_items.max()) if truth_items.numel() else 0)) + 1
        pairs = torch.unique(truth_users * m + truth_items)
        n_users = int(max(int(user_ids.max()) if user_ids.numel() else 0,
                          int(truth_users.max()) if truth_users.numel() else 0)) + 1
        truth_count = torch.bincount(torch.div(pairs, m, rounding_mode="floor"), minlength=n_users)[user_ids]
        hit = torch.isin(user_ids[:, None] * m + recommended, pairs) & (recommended >= 0)
        n_match = hit.sum(dim=1).double()
        has_truth = truth_count > 0
        if not bool(has_truth.any()):
            return float("nan"), float("nan")
        precision = (n_match[has_truth] / k).mean()
        recall = (n_match[has_truth] / truth_count[has_truth].double()).mean()
        return float(precision), float(recall)
