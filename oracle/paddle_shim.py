"""A torch-backed stand-in for the handful of `paddle` APIs that the reference's rank `net.py` and
reader files use — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Purpose: PaddlePaddle is not installed (and not installable) here, but the reference's model
definitions are plain Python.  `install()` registers fake `paddle`, `paddle.nn`,
`paddle.nn.functional`, ... modules in sys.modules so that e.g.
/root/reference/models/rank/deepfm/net.py can be imported UNMODIFIED and executed on CPU in
float64.  tests/golden/make_golden.py uses this to mint the golden vectors that pin
oracle/nets.py to the reference's own composition of operators.  What remains restated (and
therefore unpinned against real Paddle kernels) is only the per-op semantics below, each of which
follows Paddle's public API documentation:
    Embedding(padding_idx)   zero output row / no gradient for id == padding_idx
    Linear                   y = x @ W + b,  W: [in, out]
    reshape                  a 0 in `shape` copies the input dimension
    softmax                  default axis -1
    Dropout                  upscale_in_train; identity in eval()
Never imported by the product (paddlerec_b200/) or by anything that runs on the GPU box.
"""
from __future__ import annotations

import math
import sys
import types

import torch
import torch.nn as tnn


# ---- initializers ------------------------------------------------------------------------------
class _Init:
    def __call__(self, t: torch.Tensor):
        raise NotImplementedError


class TruncatedNormal(_Init):
    def __init__(self, mean=0.0, std=1.0, a=-2.0, b=2.0):
        self.mean, self.std = mean, std

    def __call__(self, t):
        with torch.no_grad():
            tnn.init.trunc_normal_(t, self.mean, self.std, self.mean - 2 * self.std,
                                   self.mean + 2 * self.std)


class Normal(_Init):
    def __init__(self, mean=0.0, std=1.0):
        self.mean, self.std = mean, std

    def __call__(self, t):
        with torch.no_grad():
            t.normal_(self.mean, self.std)


class Constant(_Init):
    def __init__(self, value=0.0):
        self.value = value

    def __call__(self, t):
        with torch.no_grad():
            t.fill_(self.value)


class Uniform(_Init):
    def __init__(self, low=-1.0, high=1.0):
        self.low, self.high = low, high

    def __call__(self, t):
        with torch.no_grad():
            t.uniform_(self.low, self.high)


def _fans(t):
    if t.dim() < 2:
        return t.numel(), t.numel()
    rec = 1
    for s in t.shape[2:]:
        rec *= s
    return t.shape[0] * rec, t.shape[1] * rec


class XavierUniform(_Init):
    def __call__(self, t):
        fi, fo = _fans(t)
        lim = math.sqrt(6.0 / (fi + fo))
        with torch.no_grad():
            t.uniform_(-lim, lim)


class XavierNormal(_Init):
    def __call__(self, t):
        fi, fo = _fans(t)
        with torch.no_grad():
            t.normal_(0.0, math.sqrt(2.0 / (fi + fo)))


class ParamAttr:
    def __init__(self, name=None, initializer=None, regularizer=None, **kw):
        self.name, self.initializer, self.regularizer = name, initializer, regularizer


class L2Decay:
    def __init__(self, coeff=0.0):
        self.coeff = coeff


def _make_param(shape, attr, default_init):
    t = torch.empty(*shape)  # default dtype (float64 when minting goldens)
    init = attr.initializer if (isinstance(attr, ParamAttr) and attr.initializer) else default_init
    init(t)
    return tnn.Parameter(t)


# ---- layers ------------------------------------------------------------------------------------
class Layer(tnn.Module):
    def add_sublayer(self, name, layer):
        self.add_module(name, layer)  # re-adding a name overwrites it, exactly like Paddle (Q6)
        return layer

    def create_parameter(self, shape, attr=None, dtype="float32", default_initializer=None):
        return _make_param(shape, attr, default_initializer or XavierUniform())


class Embedding(Layer):
    def __init__(self, num_embeddings, embedding_dim, padding_idx=None, sparse=False,
                 weight_attr=None, name=None):
        super().__init__()
        if padding_idx is not None and padding_idx < 0:
            padding_idx = num_embeddings + padding_idx
        self._padding_idx = padding_idx
        self.weight = _make_param([num_embeddings, embedding_dim], weight_attr, XavierUniform())
        if padding_idx is not None:
            with torch.no_grad():
                self.weight[padding_idx].zero_()

    def forward(self, x):
        out = self.weight[x]
        if self._padding_idx is not None:
            out = out * (x != self._padding_idx).unsqueeze(-1).to(out.dtype)
        return out


class Linear(Layer):
    def __init__(self, in_features, out_features, weight_attr=None, bias_attr=None, name=None):
        super().__init__()
        self.weight = _make_param([in_features, out_features], weight_attr, XavierUniform())
        self.bias = _make_param([out_features], bias_attr, Constant(0.0))

    def forward(self, x):
        return torch.matmul(x, self.weight) + self.bias


class ReLU(Layer):
    def forward(self, x):
        return torch.relu(x)


class Sigmoid(Layer):
    def forward(self, x):
        return torch.sigmoid(x)


class Dropout(Layer):
    def __init__(self, p=0.5, **kw):
        super().__init__()
        self.p = p

    def forward(self, x):
        return tnn.functional.dropout(x, self.p, self.training)


class BatchNorm1D(Layer):
    """paddle.nn.BatchNorm1D(num_features, momentum=0.9, epsilon=1e-05): state `weight`, `bias`,
    `_mean`, `_variance`; train mode normalises with the batch mean and the biased batch variance
    and moves the running statistics by (1 - momentum)."""

    def __init__(self, num_features, momentum=0.9, epsilon=1e-05, **kw):
        super().__init__()
        self.weight = tnn.Parameter(torch.ones(num_features))
        self.bias = tnn.Parameter(torch.zeros(num_features))
        self.register_buffer("_mean", torch.zeros(num_features))
        self.register_buffer("_variance", torch.ones(num_features))
        self._momentum, self._epsilon = momentum, epsilon

    def forward(self, x):
        if self.training:
            mu = x.mean(0)
            var = ((x - mu) ** 2).mean(0)
            with torch.no_grad():
                self._mean.mul_(self._momentum).add_(mu.detach() * (1 - self._momentum))
                self._variance.mul_(self._momentum).add_(var.detach() * (1 - self._momentum))
        else:
            mu, var = self._mean, self._variance
        return (x - mu) / torch.sqrt(var + self._epsilon) * self.weight + self.bias


class LayerList(tnn.ModuleList):
    pass


class ParameterList(tnn.ParameterList):
    pass


# ---- functional --------------------------------------------------------------------------------
def _reshape(x, shape):
    shape = [x.shape[i] if s == 0 else s for i, s in enumerate(shape)]
    return x.reshape(*shape)


def _sum(x, axis=None, dtype=None, keepdim=False):
    return x.sum() if axis is None else x.sum(axis, keepdim=keepdim)


def _softmax(x, axis=-1):
    return torch.softmax(x, dim=axis)


def install():
    """Register the fake `paddle` package.  Idempotent."""
    if "paddle" in sys.modules and getattr(sys.modules["paddle"], "__b200rec_shim__", False):
        return sys.modules["paddle"]
    if not hasattr(torch.Tensor, "astype"):
        torch.Tensor.astype = lambda self, dt: self.to(dt)  # noqa: E731  (paddle Tensor API)

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    paddle = mod("paddle")
    paddle.__b200rec_shim__ = True
    nn = mod("paddle.nn")
    F = mod("paddle.nn.functional")
    init = mod("paddle.nn.initializer")
    framework = mod("paddle.framework")
    regularizer = mod("paddle.regularizer")
    io = mod("paddle.io")
    dist = mod("paddle.distributed")
    static = mod("paddle.static")
    paddle.nn, nn.functional, nn.initializer = nn, F, init
    paddle.framework, paddle.regularizer, paddle.io = framework, regularizer, io
    paddle.distributed, paddle.static = dist, static

    for cls in (Layer, Embedding, Linear, ReLU, Sigmoid, Dropout, LayerList, ParameterList,
                BatchNorm1D):
        setattr(nn, cls.__name__, cls)
    nn.Conv1D = type("Conv1D", (Layer,), {})  # imported (never used) by din/net.py:13
    for cls in (TruncatedNormal, Normal, Constant, Uniform, XavierUniform, XavierNormal):
        setattr(init, cls.__name__, cls)
    paddle.ParamAttr = framework.ParamAttr = ParamAttr
    regularizer.L2Decay = L2Decay

    paddle.create_parameter = lambda shape, dtype="float32", name=None, attr=None, \
        is_bias=False, default_initializer=None: _make_param(
            shape, attr, default_initializer or XavierUniform())
    paddle.is_compiled_with_custom_device = lambda *_a, **_k: False
    paddle.concat = lambda x, axis=0: torch.cat(list(x), dim=axis)
    paddle.sum = _sum
    paddle.square = torch.square
    paddle.multiply = lambda x, y: x * y
    paddle.add = lambda x, y: x + y
    paddle.unsqueeze = lambda x, axis: x.unsqueeze(axis)
    paddle.reshape = lambda x, shape: _reshape(x, shape)
    paddle.matmul = torch.matmul
    paddle.tanh = torch.tanh
    paddle.stack = lambda x, axis=0: torch.stack(list(x), dim=axis)
    paddle.transpose = lambda x, perm: x.permute(*perm)
    paddle.scale = lambda x, scale=1.0, bias=0.0: x * scale + bias
    paddle.cast = lambda x, dtype: x.to(dtype)
    paddle.mean = lambda x: x.mean()
    paddle.to_tensor = torch.as_tensor
    # dlrm/net.py:104-111
    paddle.bmm = torch.bmm
    paddle.triu = lambda x, diagonal=0: torch.triu(x, diagonal)
    paddle.tril = lambda x, diagonal=0: torch.tril(x, diagonal)
    paddle.ones_like = torch.ones_like
    paddle.greater_than = lambda x, y: x > y
    paddle.masked_select = lambda x, mask: torch.masked_select(x, mask)
    F.sigmoid = torch.sigmoid
    F.softmax = _softmax
    F.relu = torch.relu

    class IterableDataset(torch.utils.data.IterableDataset):
        pass

    io.IterableDataset = IterableDataset
    dist.get_rank = lambda: 0
    dist.get_world_size = lambda: 1
    return paddle


def import_reference_net(model: str, reference_root: str = "/root/reference"):
    """Import /root/reference/models/rank/<model>/net.py (unmodified) on top of the shim."""
    import importlib.util
    import os

    install()
    path = os.path.join(reference_root, "models", "rank", model, "net.py")
    spec = importlib.util.spec_from_file_location("ref_%s_net" % model, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m
