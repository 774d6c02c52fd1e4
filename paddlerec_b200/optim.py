"""Optimizers with paddle.optimizer's surface (`step()`, `clear_grad()`), where the embedding
tables are updated row-wise from their SelectedRows gradients by the C-ABI kernels
(b200rec_sparse_adam / _sgd / _adagrad) and the small dense parameters by torch's fused optimizers.

Reference call sites: Adam  models/rank/deepfm/dygraph_model.py:61-65 (dygraph: non-lazy) and
models/rank/deepfm/static_model.py:101-103 (lazy_mode=True); SGD + PiecewiseDecay
models/rank/din/dygraph_model.py:64-73; ClipGradByGlobalNorm models/rank/dcn_v2/dygraph_model.py:81-88.
Difference kept on purpose: only lazy (row-wise) table updates are implemented —
non-lazy Adam would stream all V rows every step (77 GB at V=1e8, D=16).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from . import ops


class PiecewiseDecay:
    """paddle.optimizer.lr.PiecewiseDecay(boundaries, values)."""

    def __init__(self, boundaries, values):
        assert len(values) == len(boundaries) + 1
        self.boundaries, self.values = list(boundaries), list(values)
        self.last_epoch = 0

    def __call__(self) -> float:
        for b, v in zip(self.boundaries, self.values):
            if self.last_epoch < b:
                return v
        return self.values[-1]

    def step(self):
        self.last_epoch += 1


class ClipGradByGlobalNorm:
    def __init__(self, clip_norm: float):
        self.clip_norm = float(clip_norm)


def _is_sparse(p: torch.Tensor) -> bool:
    return getattr(p, "is_sparse_table", False)


def _valid_rows(sr: ops.SelectedRows) -> torch.Tensor:
    n = sr.value.shape[0]
    mask = torch.arange(n, device=sr.value.device) < sr.num[0]
    return torch.where(mask.unsqueeze(1), sr.value, torch.zeros((), device=sr.value.device))


class _Base:
    def __init__(self, learning_rate, parameters: Iterable[torch.Tensor], grad_clip=None):
        params = list(parameters)
        self._sparse: List[torch.Tensor] = [p for p in params if _is_sparse(p)]
        self._dense: List[torch.Tensor] = [p for p in params if not _is_sparse(p) and p.requires_grad]
        self._lr = learning_rate
        self._clip: Optional[ClipGradByGlobalNorm] = grad_clip
        self.step_count = 0
        # set by sharded.DistributedOptimizer: sums the squared norm of the LOCAL table-shard
        # gradients over the ranks (the dense gradients are already all-reduced, i.e. identical)
        self.sparse_sq_reduce = None

    #: Paddle's optimizer.step() never advances an LRScheduler and the reference loop
    #: (tools/trainer.py:151-153) never calls scheduler.step(), so the reference DIN run stays at
    #: values[0] forever.  Faithful default: do not step; set True for per-step decay.
    lr_auto_step = False

    def _maybe_step_lr(self) -> None:
        if self.lr_auto_step and hasattr(self._lr, "step"):
            self._lr.step()

    def get_lr(self) -> float:
        return float(self._lr()) if callable(self._lr) else float(self._lr)

    @staticmethod
    def _wait_tower() -> None:
        """Weight gradients of the tower may still be in flight on its side stream."""
        from . import tower
        tower.wait_pending()

    def clear_grad(self) -> None:
        self._wait_tower()
        for p in self._dense:
            p.grad = None
        for p in self._sparse:
            p.grad_rows = None

    zero_grad = clear_grad

    def global_grad_norm(self) -> torch.Tensor:
        """sqrt(sum of squares over EVERY gradient) — ClipGradByGlobalNorm's norm
        (models/rank/dcn_v2/dygraph_model.py:81-88).  With row-sharded tables each rank holds only
        its shard's SelectedRows, so that part is summed over the ranks (one scalar all-reduce)
        before the sqrt; every replica then applies the same scale."""
        self._wait_tower()
        dev = (self._dense + self._sparse)[0].device
        dense_sq = torch.zeros((), device=dev)
        for p in self._dense:
            if p.grad is not None:
                dense_sq = dense_sq + p.grad.float().square().sum()
        sparse_sq = torch.zeros((), device=dev)
        for p in self._sparse:
            sr = getattr(p, "grad_rows", None)
            if sr is not None:
                sparse_sq = sparse_sq + _valid_rows(sr).square().sum()
        if self.sparse_sq_reduce is not None:
            sparse_sq = self.sparse_sq_reduce(sparse_sq)
        return (dense_sq + sparse_sq).sqrt()

    def _apply_clip(self) -> None:
        if self._clip is None:
            return
        norm = self.global_grad_norm()
        scale = self._clip.clip_norm / torch.clamp(norm, min=self._clip.clip_norm)
        for p in self._dense:
            if p.grad is not None:
                p.grad.mul_(scale)
        for p in self._sparse:
            sr = getattr(p, "grad_rows", None)
            if sr is not None:
                sr.value.mul_(scale)


    def _table_parts(self, p, sr):
        """(W, SelectedRows) restricted to the TRAINABLE columns of a table.  A GPUBox / PS table
        (`cvm_stat_cols = 2`, rows = [show, click, embedding...]) receives the batch's show/click
        counts in its first two gradient columns (cvm_grad, wide_deep/net.py:87-88); the accessor
        ACCUMULATES them (show += pushed show, click += pushed click) instead of descending on them
        (CtrCommonAccessor of models/rank/slot_dnn/config_online.yaml:57-79)."""
        k = int(getattr(p, "cvm_stat_cols", 0))
        if not k:
            return p.data, sr, 0
        stat = ops.SelectedRows(sr.rows, sr.value[:, :k], sr.num, sr.height, ncols=k)
        ops.raw_sparse_sgd(p.data[:, :k], stat, -1.0)
        main = ops.SelectedRows(sr.rows, sr.value[:, k:], sr.num, sr.height, ncols=sr.cols - k)
        return p.data[:, k:], main, k

    def _apply_regularizers(self) -> None:
        """Per-parameter L2Decay set through ParamAttr (nn.Linear(weight_l2_decay=c), reference:
        models/rank/dcn_v2/net.py:166-168): grad += c * param.  Paddle's Optimizer.apply_gradients
        clips first and appends the regularisation term afterwards, so this runs after _apply_clip."""
        for p in self._dense:
            c = getattr(p, "l2_decay", 0.0)
            if c and p.grad is not None:
                p.grad.add_(p.detach(), alpha=c)

    def _prepare_grads(self) -> None:
        self._wait_tower()
        self._apply_clip()
        self._apply_regularizers()


class SGD(_Base):
    def __init__(self, learning_rate, parameters, grad_clip=None):
        super().__init__(learning_rate, parameters, grad_clip)
        self._torch = torch.optim.SGD(self._dense, lr=self.get_lr()) if self._dense else None

    @torch.no_grad()
    def step(self) -> None:
        self._prepare_grads()
        lr = self.get_lr()
        if self._torch is not None:
            for g in self._torch.param_groups:
                g["lr"] = lr
            self._torch.step()
        for p in self._sparse:
            sr = getattr(p, "grad_rows", None)
            if sr is not None:
                W, sr, _ = self._table_parts(p, sr)
                ops.raw_sparse_sgd(W, sr, lr)
        self.step_count += 1
        self._maybe_step_lr()


class Adam(_Base):
    """Adam with lazy (row-wise) updates of the sparse tables."""

    def __init__(self, learning_rate=0.001, parameters=(), beta1=0.9, beta2=0.999, epsilon=1e-8,
                 lazy_mode=True, grad_clip=None):
        super().__init__(learning_rate, parameters, grad_clip)
        if not lazy_mode and self._sparse:
            raise NotImplementedError(
                "non-lazy Adam over a sparse table is not supported (see module docstring)")
        self.beta1, self.beta2, self.eps = beta1, beta2, epsilon
        self._torch = (torch.optim.Adam(self._dense, lr=self.get_lr(), betas=(beta1, beta2),
                                        eps=epsilon, fused=self._dense[0].is_cuda)
                       if self._dense else None)
        self._m = {}
        self._v = {}

    def moments(self, p):
        """(m, v) with the row stride of `p`: views INTO the table slots when the table keeps its
        optimizer state in-slot (nn.FusedTable), separate zero-initialised arrays otherwise."""
        inslot = getattr(p, "inslot_moments", None)
        if inslot is not None:
            G = inslot[1] - inslot[0]
            return p.data[:, inslot[0]:inslot[0] + G], p.data[:, inslot[1]:inslot[1] + G]
        k = id(p)
        if k not in self._m:
            self._m[k] = torch.zeros_like(p.data)
            self._v[k] = torch.zeros_like(p.data)
        return self._m[k], self._v[k]

    @torch.no_grad()
    def step(self) -> None:
        self._prepare_grads()
        self.step_sparse(prepared=True)
        self.step_dense(prepared=True)

    @torch.no_grad()
    def step_sparse(self, prepared: bool = False) -> None:
        """The row-wise table updates (first half of step(); sharded.DistributedOptimizer runs it
        beside the dense all-reduce).  Without clipping / regularisers on the tables the two halves
        are independent."""
        if not prepared:
            assert self._clip is None, "split stepping is not defined with gradient clipping"
        lr = self.get_lr()
        self.step_count += 1
        b1p = self.beta1 ** self.step_count
        b2p = self.beta2 ** self.step_count
        for p in self._sparse:
            sr = getattr(p, "grad_rows", None)
            if sr is not None:
                m, v = self.moments(p)
                W, sr, k = self._table_parts(p, sr)
                ops.raw_sparse_adam(W, m[:, k:], v[:, k:], sr, lr, self.beta1, self.beta2, self.eps,
                                    b1p, b2p)

    @torch.no_grad()
    def step_dense(self, prepared: bool = False) -> None:
        """Second half of step(): the replicated dense parameters (torch's fused Adam)."""
        self._wait_tower()
        if not prepared:
            self._apply_regularizers()
        lr = self.get_lr()
        if self._torch is not None:
            for g in self._torch.param_groups:
                g["lr"] = lr
            self._torch.step()
        self._maybe_step_lr()


class SparseAdaGrad(_Base):
    """SparseAdaGradSGDRule of the PS/GPUBox path (models/rank/slot_dnn/config_online.yaml:57-79):
    one g2sum accumulator per ROW; dense parameters fall back to Adam like the PS config does."""

    def __init__(self, learning_rate=0.05, parameters=(), initial_g2sum=3.0,
                 weight_bounds=(-10.0, 10.0), dense_learning_rate=0.001):
        super().__init__(learning_rate, parameters, None)
        self.g0 = initial_g2sum
        self.lo, self.hi = weight_bounds
        self._torch = (torch.optim.Adam(self._dense, lr=dense_learning_rate,
                                        fused=self._dense[0].is_cuda) if self._dense else None)
        self._g2 = {}

    def g2sum(self, p):
        k = id(p)
        if k not in self._g2:
            self._g2[k] = torch.zeros(p.shape[0], device=p.device)
        return self._g2[k]

    @torch.no_grad()
    def step(self) -> None:
        self._wait_tower()
        self._apply_regularizers()
        if self._torch is not None:
            self._torch.step()
        for p in self._sparse:
            sr = getattr(p, "grad_rows", None)
            if sr is not None:
                W, sr, _ = self._table_parts(p, sr)
                ops.raw_sparse_adagrad(W, self.g2sum(p), sr, self.get_lr(), self.g0, self.lo,
                                       self.hi)
        self.step_count += 1
