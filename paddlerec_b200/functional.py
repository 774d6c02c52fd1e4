"""Loss / metric helpers with Paddle's semantics (host-side plumbing, plain torch ops)."""
from __future__ import annotations

import numpy as np
import torch


def log_loss(pred: torch.Tensor, label: torch.Tensor, epsilon: float = 1e-4) -> torch.Tensor:
    """paddle.nn.functional.log_loss: -y*log(p+eps) - (1-y)*log(1-p+eps), eps=1e-4 by default
    (called at models/rank/deepfm/dygraph_model.py:53-58)."""
    return -label * torch.log(pred + epsilon) - (1.0 - label) * torch.log(1.0 - pred + epsilon)


def log_loss_mean(pred: torch.Tensor, label: torch.Tensor, epsilon: float = 1e-4) -> torch.Tensor:
    """paddle.mean(log_loss(pred, label)) — what every CTR DygraphModel.create_loss computes
    (deepfm/dygraph_model.py:53-58).  On CUDA one fused reduction kernel each way
    (b200rec_log_loss_mean_fwd/_bwd); elsewhere the composition of the two torch ops."""
    if pred.is_cuda:
        from . import ops
        return ops.log_loss_mean(pred, label, epsilon)
    return log_loss(pred, label.to(torch.float32), epsilon).mean()


class Auc:
    """paddle.metric.Auc("ROC", num_thresholds=4095): bucketed positive/negative histograms and a
    trapezoid sweep.  Unlike the reference (`.numpy()` every step, dygraph_model.py:83-84) the
    histograms stay on the device; only `accumulate()` synchronises."""

    def __init__(self, curve: str = "ROC", num_thresholds: int = 4095):
        self.num_thresholds = num_thresholds
        self._pos = None
        self._neg = None

    def reset(self):
        self._pos = None
        self._neg = None

    def update(self, preds, labels):
        """preds: [B,2] (P(neg), P(pos)) or [B]/[B,1] P(pos); labels: [B,1] or [B]."""
        if isinstance(preds, np.ndarray):
            preds = torch.from_numpy(preds)
        if isinstance(labels, np.ndarray):
            labels = torch.from_numpy(labels)
        p = preds[:, 1] if (preds.dim() == 2 and preds.shape[1] == 2) else preds.reshape(-1)
        y = labels.reshape(-1).to(p.device)
        nb = self.num_thresholds + 1
        if self._pos is None:
            self._pos = torch.zeros(nb, dtype=torch.int64, device=p.device)
            self._neg = torch.zeros(nb, dtype=torch.int64, device=p.device)
        if p.is_cuda and p.dtype == torch.float32:
            from . import ops       # one kernel, integer atomics, no host sync
            ops.raw_auc_update(p.detach(), y, self._pos, self._neg, self.num_thresholds)
            return
        idx = (p.detach().float() * self.num_thresholds).to(torch.int64).clamp_(0, nb - 1)
        # index_add_ into preallocated histograms: no data-dependent shape, hence no host sync
        # (boolean indexing / bincount would synchronise every step like the reference's .numpy())
        is_pos = (y != 0).to(torch.int64)
        self._pos.index_add_(0, idx, is_pos)
        self._neg.index_add_(0, idx, 1 - is_pos)

    def stats(self):
        """(stat_pos, stat_neg) int64 tensors — what utils_single.py:160-206 all-reduces."""
        return self._pos, self._neg

    def accumulate(self) -> float:
        if self._pos is None:
            return 0.0
        pos = self._pos.cpu().numpy().astype(np.float64)
        neg = self._neg.cpu().numpy().astype(np.float64)
        return auc_from_stats(pos, neg)


def auc_from_stats(pos: np.ndarray, neg: np.ndarray) -> float:
    tot_pos = tot_neg = 0.0
    auc = 0.0
    for idx in range(len(pos) - 1, -1, -1):
        tot_pos_prev, tot_neg_prev = tot_pos, tot_neg
        tot_pos += pos[idx]
        tot_neg += neg[idx]
        auc += abs(tot_neg - tot_neg_prev) * (tot_pos + tot_pos_prev) / 2.0
    return auc / tot_pos / tot_neg if tot_pos > 0.0 and tot_neg > 0.0 else 0.0


def softmax_cross_entropy(logits: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """paddle.nn.functional.cross_entropy(input, label) with hard labels [B,1] (per-sample, no
    reduction; models/rank/dlrm/dygraph_model.py:58-62 takes the mean)."""
    return torch.nn.functional.cross_entropy(logits, label.reshape(-1).to(torch.int64),
                                             reduction="none").unsqueeze(1)


class Accuracy:
    """paddle.metric.Accuracy() top-1: `compute(pred, label)` -> per-sample correctness,
    `update(correct)` accumulates, `accumulate()` -> running accuracy.  Counters stay on the device
    until accumulate()."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._correct = None
        self._total = 0

    def compute(self, pred: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        return (pred.argmax(dim=1, keepdim=True) == label.reshape(-1, 1).to(pred.device)).to(torch.float32)

    def update(self, correct: torch.Tensor):
        c = correct.detach().sum()
        self._correct = c if self._correct is None else self._correct + c
        self._total += correct.shape[0]
        return self

    def accumulate(self) -> float:
        return float(self._correct) / self._total if self._total else 0.0
