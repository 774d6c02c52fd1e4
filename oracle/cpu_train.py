"""CPU training step of DeepFM used as the reported CPU baseline (TEST/BENCH INFRASTRUCTURE — see
oracle/__init__.py).  It is the oracle's forward (oracle/nets.py, i.e. the reference's
models/rank/deepfm/net.py op for op) driven the way tools/trainer.py:142-154 drives it:
clear_grad -> forward -> log_loss -> backward -> optimizer step, in fp32 on all host threads.

Two concessions that FAVOUR the CPU, both needed to make V=1e8 runnable at all on a host:
  * the two tables produce sparse gradients (torch's analogue of Paddle's SelectedRows) and are
    updated with a lazy Adam (touched rows only, multi-threaded index ops) — the reference's
    dygraph path would run a non-lazy Adam over all V rows every step
    (deepfm/dygraph_model.py:61-65);
  * inputs are already-parsed tensors (no Python text reader in the timed region).
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F

from . import nets


class CpuDeepFM:
    def __init__(self, V, D, Dn=13, Fs=26, fc=(400, 400, 400), lr=1e-3, seed=12345):
        g = torch.Generator().manual_seed(seed)
        std = 0.1 / D ** 0.5
        self.V, self.D, self.Dn, self.Fs, self.fc = V, D, Dn, Fs, list(fc)
        self.W = torch.empty(V, D).uniform_(-2 * std, 2 * std, generator=g).requires_grad_(True)
        self.W1 = torch.empty(V, 1).uniform_(-2 * std, 2 * std, generator=g).requires_grad_(True)
        with torch.no_grad():
            self.W[0].zero_()
            self.W1[0].zero_()
        self.p = {
            "fm.dense_w_one": (torch.randn(Dn, generator=g) * std).requires_grad_(True),
            "fm.dense_w": (torch.randn(1, Dn, D, generator=g) * std).requires_grad_(True),
        }
        sizes = [(Fs + Dn) * D] + self.fc + [1]
        for i in range(len(sizes) - 1):
            self.p["dnn.linear_%d.weight" % i] = (torch.randn(sizes[i], sizes[i + 1], generator=g)
                                                  / sizes[i] ** 0.5).requires_grad_(True)
            self.p["dnn.linear_%d.bias" % i] = torch.zeros(sizes[i + 1], requires_grad=True)
        self.opt_dense = torch.optim.Adam(list(self.p.values()), lr=lr)
        # lazy Adam state of the two tables (touched rows only; see _lazy_adam below).
        # torch.optim.SparseAdam is NOT used: its sparse-dense adds run single-threaded and made the
        # optimizer 56 % of the CPU step, which has nothing to do with the reference's algorithm.
        self.lr, self.t = lr, 0
        # .zero_() (not torch.zeros) so every page is touched now, not inside the timed steps
        self.mom = {id(w): (torch.empty_like(w).zero_(), torch.empty_like(w).zero_())
                    for w in (self.W, self.W1)}

    def forward(self, ids, dense):
        """deepfm_forward of oracle/nets.py with the two lookups producing sparse grads."""
        keep = (ids != 0).unsqueeze(-1).to(torch.float32)
        e1 = F.embedding(ids, self.W1, sparse=True) * keep
        e = F.embedding(ids, self.W, sparse=True) * keep
        y1 = e1.sum(1) + (dense * self.p["fm.dense_w_one"]).unsqueeze(2).sum(1)
        feat = torch.cat([e, dense.unsqueeze(2) * self.p["fm.dense_w"]], 1)
        y2 = 0.5 * (feat.sum(1).square() - feat.square().sum(1)).sum(1, keepdim=True)
        y_dnn = nets.mlp_relu(self.p, "dnn.", feat.reshape(feat.shape[0], -1), len(self.fc) + 1)
        return torch.sigmoid(y1 + y2 + y_dnn)

    @torch.no_grad()
    def _lazy_adam(self, w, b1=0.9, b2=0.999, eps=1e-8):
        """Adam(lazy_mode=True): merge the SelectedRows-like sparse gradient, update only the rows
        present (oracle/optim.py:adam_lazy with multi-threaded torch index ops)."""
        g = w.grad.coalesce()
        rows, val = g.indices()[0], g.values()
        m, v = self.mom[id(w)]
        mr = m.index_select(0, rows).mul_(b1).add_(val, alpha=1 - b1)
        vr = v.index_select(0, rows).mul_(b2).addcmul_(val, val, value=1 - b2)
        c2 = (1 - b2 ** self.t) ** 0.5
        step = mr / (vr.sqrt() + eps * c2) * (self.lr * c2 / (1 - b1 ** self.t))
        m.index_copy_(0, rows, mr)
        v.index_copy_(0, rows, vr)
        w.index_add_(0, rows, step, alpha=-1.0)
        w.grad = None

    def step(self, ids, dense, label):
        self.opt_dense.zero_grad(set_to_none=True)
        pred = self.forward(ids, dense)
        loss = nets.log_loss(pred, label).mean()
        loss.backward()
        self.opt_dense.step()
        self.t += 1
        self._lazy_adam(self.W)
        self._lazy_adam(self.W1)
        return float(loss.detach())


def time_steps(model: CpuDeepFM, batches, steps: int, warmup: int):
    """Returns (seconds for `steps` steps, last loss)."""
    loss = 0.0
    for i in range(warmup):
        loss = model.step(*batches[i % len(batches)])
    t0 = time.perf_counter()
    for i in range(steps):
        loss = model.step(*batches[(warmup + i) % len(batches)])
    return time.perf_counter() - t0, loss
