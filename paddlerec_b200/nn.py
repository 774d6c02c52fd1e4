"""Host-side building blocks with Paddle's layer semantics, backed by the b200rec C ABI.

What each class mirrors (reference = /root/reference):
  Embedding   paddle.nn.Embedding(V, D, padding_idx, sparse=True)    e.g. models/rank/deepfm/net.py:66-86
  Linear      paddle.nn.Linear: y = x @ W + b with W:[in,out]        e.g. models/rank/deepfm/net.py:156-163
  matmul precision: the tower GEMMs are the only tensor-core work on this path (cuBLAS through
  torch); `set_matmul_precision` picks fp32 (exact, SIMT), tf32, or bf16x3 (3 bf16 tensor-core
  GEMMs on a hi/lo split, ~2^-16 relative error — meets the 1e-4 logit bar).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as tnn

from . import ops

_PRECISION = "fp32"


def set_matmul_precision(mode: str) -> None:
    """'fp32' | 'tf32' | 'bf16x3' for every Linear / CrossNet GEMM of this package."""
    global _PRECISION
    if mode not in ("fp32", "tf32", "bf16x3"):
        raise ValueError("matmul precision must be fp32, tf32 or bf16x3")
    _PRECISION = mode


def get_matmul_precision() -> str:
    return _PRECISION


def _split_bf16(a: torch.Tensor):
    hi = a.to(torch.bfloat16)
    lo = (a - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


def mm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a @ b for fp32 operands in the configured precision (no autograd: used inside Functions)."""
    if _PRECISION == "fp32":
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            return torch.mm(a, b)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
    if _PRECISION == "tf32":
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            return torch.mm(a, b)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
    a_hi, a_lo = _split_bf16(a)
    b_hi, b_lo = _split_bf16(b)
    out = torch.mm(a_hi, b_hi, out_dtype=torch.float32)
    out += torch.mm(a_hi, b_lo, out_dtype=torch.float32)
    out += torch.mm(a_lo, b_hi, out_dtype=torch.float32)
    return out


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        y = mm(x, W)
        if b is not None:
            y += b
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = dy.contiguous()
        dx = mm(dy, W.t()) if ctx.needs_input_grad[0] else None
        dW = mm(x.t(), dy) if ctx.needs_input_grad[1] else None
        db = dy.sum(0) if ctx.needs_input_grad[2] else None
        return dx, dW, db


class Linear(tnn.Module):
    """paddle.nn.Linear: weight is [in_features, out_features] (transposed w.r.t. torch)."""

    def __init__(self, in_features: int, out_features: int, weight_std: Optional[float] = None,
                 xavier: bool = False, bias: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = tnn.Parameter(torch.empty(in_features, out_features))
        self.bias = tnn.Parameter(torch.zeros(out_features)) if bias else None
        if xavier:  # paddle XavierUniform: U(-sqrt(6/(fan_in+fan_out)), +)
            lim = math.sqrt(6.0 / (in_features + out_features))
            tnn.init.uniform_(self.weight, -lim, lim)
        elif weight_std is not None:
            tnn.init.normal_(self.weight, 0.0, weight_std)
        else:  # paddle.nn.Linear default: XavierUniform as well
            lim = math.sqrt(6.0 / (in_features + out_features))
            tnn.init.uniform_(self.weight, -lim, lim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lead = x.shape[:-1]
        y = _LinearFn.apply(x.reshape(-1, self.in_features), self.weight, self.bias)
        return y.reshape(*lead, self.out_features)


_HOOKS = {}


def _autograd_hook(device) -> torch.Tensor:
    """A requires_grad scalar per device: keeps autograd walking into a gather whose table has no
    dense grad (its backward returns None for it, so nothing is ever accumulated)."""
    h = _HOOKS.get(device)
    if h is None:
        h = torch.zeros(1, device=device, requires_grad=True)
        _HOOKS[device] = h
    return h


class Embedding(tnn.Module):
    """paddle.nn.Embedding(num_embeddings, embedding_dim, padding_idx, sparse=True).

    The table is a Parameter that never gets a dense `.grad`: backward leaves a merged
    `ops.SelectedRows` in `self.grad_rows`, consumed by paddlerec_b200.optim."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None,
                 init_std: Optional[float] = None, init: str = "truncated_normal",
                 device=None):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        if padding_idx is not None and padding_idx < 0:
            padding_idx = num_embeddings + padding_idx  # paddle wraps negative padding_idx
        self.padding_idx = padding_idx
        w = torch.empty(num_embeddings, embedding_dim, device=device)
        self.weight = tnn.Parameter(w, requires_grad=False)
        # the SelectedRows gradient lives ON the parameter tensor (like Paddle's param.grad), so an
        # optimizer built from `model.parameters()` finds it without knowing about modules
        self.weight.is_sparse_table = True
        self.weight.grad_rows = None
        self.reset_parameters(init, init_std)

    @property
    def grad_rows(self) -> Optional[ops.SelectedRows]:
        return getattr(self.weight, "grad_rows", None)

    @property
    def pad(self) -> int:
        return -1 if self.padding_idx is None else int(self.padding_idx)

    @torch.no_grad()
    def reset_parameters(self, init: str, std: Optional[float]) -> None:
        w = self.weight
        if init == "truncated_normal":  # paddle TruncatedNormal: resample outside +-2 sigma
            tnn.init.trunc_normal_(w, 0.0, std or 1.0, -2.0 * (std or 1.0), 2.0 * (std or 1.0))
        elif init == "uniform":  # paddle Uniform(): U(-1, 1)
            w.uniform_(-1.0, 1.0)
        elif init == "xavier_uniform":
            lim = math.sqrt(6.0 / (self.num_embeddings + self.embedding_dim))
            w.uniform_(-lim, lim)
        elif init == "zeros":
            w.zero_()
        elif init == "empty":
            pass
        else:
            raise ValueError(init)
        if self.padding_idx is not None and init != "empty":
            w[self.padding_idx].zero_()  # paddle dygraph zeroes the padding row at construction

    def accept(self, sr: ops.SelectedRows) -> None:
        if self.grad_rows is not None:
            raise RuntimeError("Embedding used twice in one step: merge of SelectedRows grads "
                               "is not implemented (clear_grad between steps)")
        self.weight.grad_rows = sr

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        return ops.gather(self.weight, ids, self.pad, self, _autograd_hook(ids.device))

    def clear_grad(self) -> None:
        self.weight.grad_rows = None

    def _apply(self, fn, *a, **kw):  # keep the tags if .to()/.cuda() replaces the tensor
        out = super()._apply(fn, *a, **kw)
        self.weight.is_sparse_table = True
        if not hasattr(self.weight, "grad_rows"):
            self.weight.grad_rows = None
        return out


class EmbeddingPair:
    """Grad sink for the fused DeepFM kernel, which reads TWO tables ([V,1] and [V,D]) with the
    same ids: one grouping pass serves both SelectedRows."""

    def __init__(self, emb: Embedding, emb_one: Embedding):
        self.emb, self.emb_one = emb, emb_one

    def groups_for(self, ids, V, pad):
        return ops.raw_group_ids(ids, V, pad)

    def accept(self, sr_w: ops.SelectedRows, sr_w1: ops.SelectedRows) -> None:
        self.emb.accept(sr_w)
        self.emb_one.accept(sr_w1)


def sparse_tables(model: tnn.Module) -> List[Embedding]:
    return [m for m in model.modules() if isinstance(m, Embedding)]
