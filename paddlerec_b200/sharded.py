"""Row-cyclic sharding of the embedding tables over the GPUs of one box, with an NCCL all-to-all on
the id -> owner and row -> requester exchanges.

Behavioural spec in the reference: the PSGPU / HeterPS pull-push of the GPUBox trainer
(tools/static_gpubox_trainer.py:152-159 builds the per-GPU HBM tables, :244-259 runs
pull_sparse -> fwd/bwd -> push_sparse per batch; tools/run_gpubox.sh:7-45 launches it).  Here:

    owner(id) = id mod world,  local_row(id) = id div world        (padding id 0 lives on rank 0)
    forward   bucketize ids by owner  ->  all-to-all(counts)  ->  all-to-all(ids)
              owner: b200rec_gather on its shard  ->  all-to-all(rows) back
              requester: the SAME fused kernel as the single-GPU path (b200rec_embed_fm_fwd) over
              the received rows, indexed by the bucket permutation
    backward  requester: b200rec_embed_fm_bwd emits per-slot gradients in bucket order
              all-to-all(grads) to the owners  ->  owner: group_ids + segment_reduce on local rows
              -> SelectedRows on the local shard -> row-wise optimizer
    dense     parameters are replicated; their gradients are all-reduced in one flat bucket.

The collective plumbing is torch.distributed (NCCL on GPUs; gloo in the CPU unit tests, which
inject a torch stand-in for the kernel set — the product default `ops` has no CPU path).
"""
from __future__ import annotations

import contextlib
import math
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as tnn

from . import nn as bnn
from . import ops as _cuda_ops
from . import optim


def _timed(name: str):
    """CUDA-event pair around a collective when bench.py collects per-kernel times (CUDA only)."""
    return _cuda_ops._Timed(name) if torch.cuda.is_available() else contextlib.nullcontext()


def shard_rows(V: int, rank: int, world: int) -> int:
    """Number of global ids with id % world == rank."""
    return (V - rank + world - 1) // world if V > rank else 0


@dataclass
class ExchangePlan:
    n: int
    perm: torch.Tensor        # int64 [n]: slot of each position in bucket order
    inv_perm: torch.Tensor    # int32 [n]
    send_splits: List[int]    # ids sent to each rank
    recv_splits: List[int]    # ids received from each rank
    recv_ids: torch.Tensor    # int64 [m]: local rows requested from this rank (owner view)
    owner_groups: object = None   # grouping of recv_ids by local row, if it was precomputed


class _PendingPlan:
    """Bucketing + count exchange of a FUTURE batch, running on a side stream."""

    def __init__(self, ids, send_ids, perm, inv_perm, host_counts, event):
        self.ids, self.send_ids, self.perm, self.inv_perm = ids, send_ids, perm, inv_perm
        self.host_counts, self.event = host_counts, event


class ShardExchange:
    """The all-to-all choreography; shared by every table that is looked up with the same ids.

    The split sizes of the id exchange must be known on the host.  `plan()` alone therefore costs a
    device->host sync per step, which also drains the launch queue (the CPU can no longer run
    ahead of the GPU).  `prefetch(ids_next)` removes it: bucketing, the count all-to-all and the
    D2H copy of the NEXT batch run on a side stream while the current step computes; by the time
    `plan(ids_next)` is called the counts are already on the host."""

    def __init__(self, V: int, rank: int, world: int, group=None, kernels=_cuda_ops,
                 owner_pad: Optional[int] = None):
        self.V, self.rank, self.world, self.group, self.k = V, rank, world, group, kernels
        self._pending = None
        self._side = None
        # local padding row of the tables behind this exchange (None = unknown): lets the prefetch
        # also run the OWNER-side grouping of the received ids (a CUB sort) off the critical path
        self.owner_pad = owner_pad

    def _bucketize_and_count(self, ids):
        send_ids, perm, inv_perm, counts = self.k.raw_shard_bucketize(ids, self.world, self.V)
        recv_counts = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts, group=self.group)
        return send_ids, perm, inv_perm, torch.stack([counts, recv_counts])

    def prefetch(self, ids: torch.Tensor) -> None:
        """Start planning the exchange of a future batch (CUDA only; a no-op on CPU tensors)."""
        if not ids.is_cuda:
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=ids.device)
        flat = ids.reshape(-1)
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            send_ids, perm, inv_perm, both = self._bucketize_and_count(flat)
            host = torch.empty(both.shape, dtype=both.dtype, pin_memory=True)
            host.copy_(both, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._side)
        for t in (flat, send_ids, perm, inv_perm, both):
            t.record_stream(self._side)
        self._pending = _PendingPlan(ids, send_ids, perm, inv_perm, host, ev)

    def finish_prefetch(self) -> None:
        """Second half of the prefetch (call it later in the step, e.g. after backward was
        launched): once the counts are on the host, run the id all-to-all of the NEXT batch on the
        side stream too, so the next forward starts directly with the owner-side gather."""
        pend = self._pending
        if pend is None or getattr(pend, "recv_ids", None) is not None:
            return
        # Every rank must issue its collectives in the same order, so this must not depend on a
        # rank-local "is it ready yet" test: always wait for the counts (the GPU is busy with the
        # backward pass that was already enqueued, so the host wait costs no GPU time).
        pend.event.synchronize()
        send_splits, recv_splits = pend.host_counts[0].tolist(), pend.host_counts[1].tolist()
        with torch.cuda.stream(self._side):
            recv_ids = torch.empty(sum(recv_splits), dtype=torch.int64, device=pend.send_ids.device)
            dist.all_to_all_single(recv_ids, pend.send_ids, recv_splits, send_splits,
                                   group=self.group)
            groups = None
            if self.owner_pad is not None:
                # what owner_reduce() needs in the backward of the NEXT step: done here, on the side
                # stream, while the current backward runs (own scratch buffer: ws_tag)
                groups = self.k.raw_group_ids(recv_ids, max(shard_rows(self.V, self.rank, self.world), 1),
                                              self.owner_pad, ws_tag="group_side")
            ev = torch.cuda.Event()
            ev.record(self._side)
        recv_ids.record_stream(self._side)
        pend.recv_ids, pend.splits, pend.event2 = recv_ids, (send_splits, recv_splits), ev
        pend.groups = groups

    def plan(self, ids: torch.Tensor) -> ExchangePlan:
        pend, self._pending = self._pending, None
        flat = ids.reshape(-1)
        n = flat.numel()
        if pend is not None and pend.ids is ids:
            send_ids, perm, inv_perm = pend.send_ids, pend.perm, pend.inv_perm
            if getattr(pend, "recv_ids", None) is not None:       # fully prefetched
                torch.cuda.current_stream().wait_event(pend.event2)
                send_splits, recv_splits = pend.splits
                cur = torch.cuda.current_stream()
                for t in (pend.recv_ids, perm, inv_perm):
                    t.record_stream(cur)
                groups = getattr(pend, "groups", None)
                if groups is not None:
                    for t in (groups.unique_ids, groups.seg_offsets, groups.sorted_pos, groups.num):
                        t.record_stream(cur)
                return ExchangePlan(n, perm, inv_perm, send_splits, recv_splits, pend.recv_ids,
                                    groups)
            pend.event.synchronize()                              # normally long complete
            torch.cuda.current_stream().wait_event(pend.event)
            send_splits, recv_splits = pend.host_counts[0].tolist(), pend.host_counts[1].tolist()
        else:
            send_ids, perm, inv_perm, both = self._bucketize_and_count(flat)
            both = both.cpu()                                     # host sync (no prefetch)
            send_splits, recv_splits = both[0].tolist(), both[1].tolist()
        recv_ids = torch.empty(sum(recv_splits), dtype=torch.int64, device=ids.device)
        dist.all_to_all_single(recv_ids, send_ids, recv_splits, send_splits, group=self.group)
        return ExchangePlan(n, perm, inv_perm, send_splits, recv_splits, recv_ids)

    def pull(self, plan: ExchangePlan, shard: torch.Tensor, local_pad: int,
             D: Optional[int] = None) -> torch.Tensor:
        """Owner-side gather of the first D columns (default: all) + rows back to the requesters:
        returns [n, D] in bucket order."""
        rows_out = self.k.raw_gather(shard, plan.recv_ids, local_pad, D)
        rows_in = torch.empty(plan.n, rows_out.shape[-1], dtype=shard.dtype, device=shard.device)
        with _timed("nccl_a2a_rows"):
            dist.all_to_all_single(rows_in, rows_out, plan.send_splits, plan.recv_splits,
                                   group=self.group)
        return rows_in

    def push(self, plan: ExchangePlan, grads_bucket_order: torch.Tensor) -> torch.Tensor:
        """Per-slot gradients [n, D] (bucket order) -> owners: returns [m, D] aligned with
        plan.recv_ids."""
        D = grads_bucket_order.shape[1]
        out = torch.empty(plan.recv_ids.numel(), D, dtype=grads_bucket_order.dtype,
                          device=grads_bucket_order.device)
        with _timed("nccl_a2a_grads"):
            dist.all_to_all_single(out, grads_bucket_order.contiguous(), plan.recv_splits,
                                   plan.send_splits, group=self.group)
        return out

    def owner_reduce(self, plan: ExchangePlan, grads: torch.Tensor, V_loc: int, local_pad: int):
        """Merge the received gradients by local row -> SelectedRows on the local shard."""
        groups = plan.owner_groups
        if groups is None or local_pad != self.owner_pad:
            groups = self.k.raw_group_ids(plan.recv_ids, max(V_loc, 1), local_pad)
        rows = self.k.raw_segment_reduce(grads.contiguous(), groups.seg_offsets, groups.sorted_pos,
                                         groups.num, groups.n)
        return self.k.SelectedRows(groups.unique_ids, rows, groups.num, V_loc)


class _ShardedLookup(torch.autograd.Function):
    """paddle.nn.Embedding forward/backward over a row-cyclically sharded table: the generic
    (non-FM) lookup used by DCN-V2 / Wide&Deep / DIN when their tables are sharded."""

    @staticmethod
    def forward(ctx, ids, emb, _hook):
        ex, k = emb.exchange, emb.exchange.k
        plan = ex.plan(ids)
        rows = ex.pull(plan, emb.weight, emb.pad)                  # [n, D] in bucket order
        out = k.raw_gather(rows, plan.perm, -1)                    # back to position order
        ctx.plan, ctx.emb = plan, emb
        return out.reshape(*ids.shape, rows.shape[1])

    @staticmethod
    def backward(ctx, dout):
        plan, emb = ctx.plan, ctx.emb
        ex, k = emb.exchange, emb.exchange.k
        D = dout.shape[-1]
        g = k.raw_gather(dout.reshape(-1, D).contiguous(), plan.inv_perm.to(torch.int64), -1)
        g = ex.push(plan, g)
        emb.accept(ex.owner_reduce(plan, g, emb.num_embeddings, emb.pad))
        return None, None, None


class ShardedEmbedding(bnn.Embedding):
    """The local shard ([ceil((V-rank)/world), D]) of a row-cyclically sharded table.  `forward`
    has paddle.nn.Embedding's signature; the all-to-all exchange happens inside."""

    def __init__(self, num_embeddings, embedding_dim, padding_idx, rank, world, init_std=None,
                 init="truncated_normal", device=None, group=None, kernels=_cuda_ops):
        local_pad = None
        if padding_idx is not None and padding_idx % world == rank:
            local_pad = padding_idx // world
        super().__init__(shard_rows(num_embeddings, rank, world), embedding_dim, local_pad,
                         init_std=init_std, init=init, device=device)
        self.global_rows, self.rank, self.world = num_embeddings, rank, world
        self.exchange = ShardExchange(num_embeddings, rank, world, group, kernels,
                                      owner_pad=(-1 if local_pad is None else local_pad))

    def forward(self, ids):
        return _ShardedLookup.apply(ids, self, bnn._autograd_hook(ids.device))


@torch.no_grad()
def shard_embeddings(model: tnn.Module, rank: int, world: int, group=None, min_rows: int = 0,
                     kernels=_cuda_ops) -> tnn.Module:
    """Replace every `bnn.Embedding` of `model` whose table has >= min_rows rows by the local
    shard of a row-cyclic ShardedEmbedding (rows rank, rank+world, ... of the original, so a
    replicated initialisation stays consistent), then broadcast the dense parameters.  This is how
    DCN-V2 (BASELINE config 3), Wide&Deep and large-vocabulary DIN run on N GPUs."""
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if type(child) is bnn.Embedding and child.num_embeddings >= min_rows:
                sh = ShardedEmbedding(child.num_embeddings, child.embedding_dim, child.padding_idx,
                                      rank, world, init="empty", device=child.weight.device,
                                      group=group, kernels=kernels)
                sh.weight.copy_(child.weight[rank::world])
                setattr(parent, name, sh)
    sync_dense_parameters(model, group)
    return model


class _ShardedEmbedFM(torch.autograd.Function):
    """DeepFM's FM block over sharded tables: exchange + the single-GPU fused kernels."""

    @staticmethod
    def forward(ctx, ids, dense, dense_w, dense_w1, fm):
        k, ex = fm.k, fm.exchange
        B, F = ids.shape
        plan = ex.plan(ids)
        D = fm.sparse_feature_dim
        slot_ids = plan.perm.reshape(B, F)
        if fm.fused:   # ONE exchange of [emb | w1 | pad] rows instead of two
            tab = fm._fused
            rows = ex.pull(plan, tab.weight, tab.pad, tab.grad_cols)
            feat, y1, y2, S = k.raw_embed_fm_fwd(rows, None, slot_ids, dense,
                                                 dense_w.reshape(-1, D), dense_w1.reshape(-1), -1,
                                                 D=D)
        else:
            rows = ex.pull(plan, fm.embedding.weight, fm.embedding.pad)
            rows1 = ex.pull(plan, fm.embedding_one.weight, fm.embedding_one.pad)
            feat, y1, y2, S = k.raw_embed_fm_fwd(rows, rows1, slot_ids, dense,
                                                 dense_w.reshape(-1, D), dense_w1.reshape(-1), -1)
        ctx.save_for_backward(dense, feat, S)
        ctx.plan, ctx.fm, ctx.F = plan, fm, F
        ctx.dense_w_shape = dense_w.shape
        return feat, y1.unsqueeze(1), y2.unsqueeze(1)

    @staticmethod
    def backward(ctx, dfeat, dy1, dy2):
        dense, feat, S = ctx.saved_tensors
        fm, plan, F = ctx.fm, ctx.plan, ctx.F
        k, ex = fm.k, fm.exchange
        B = feat.shape[0]
        dev = feat.device
        gy1 = dy1.reshape(-1).contiguous() if dy1 is not None else torch.zeros(B, device=dev)
        gy2 = dy2.reshape(-1).contiguous() if dy2 is not None else torch.zeros(B, device=dev)
        if dfeat is not None:
            dfeat = dfeat.contiguous()
        n = plan.n
        iota, num = fm.trivial_groups(n, dev)
        # seg_offsets = iota, sorted_pos = inv_perm: row k of the output is the gradient of slot k
        if fm.fused:
            tab = fm._fused
            dW, _, ddense_w, ddense_w1 = k.raw_embed_fm_bwd(feat, S, dfeat, gy1, gy2, dense, iota,
                                                            plan.inv_perm, num, F,
                                                            fused_cols=tab.grad_cols)
            g = ex.push(plan, dW[:n])
            tab.accept_fused(ex.owner_reduce(plan, g, tab.num_embeddings, tab.pad))
        else:
            dW, dW1, ddense_w, ddense_w1 = k.raw_embed_fm_bwd(feat, S, dfeat, gy1, gy2, dense, iota,
                                                              plan.inv_perm, num, F)
            g = ex.push(plan, dW[:n])
            g1 = ex.push(plan, dW1[:n].unsqueeze(1))
            emb, emb1 = fm.embedding, fm.embedding_one
            emb.accept(ex.owner_reduce(plan, g, emb.num_embeddings, emb.pad))
            emb1.accept(ex.owner_reduce(plan, g1, emb1.num_embeddings, emb1.pad))
        return None, None, ddense_w.reshape(ctx.dense_w_shape), ddense_w1, None


class ShardedFusedTable(bnn.FusedTable):
    """Local shard of a row-cyclically sharded FusedTable."""

    def __init__(self, num_embeddings, embedding_dim, padding_idx, rank, world, init_std=None,
                 device=None):
        local_pad = None
        if padding_idx is not None and padding_idx % world == rank:
            local_pad = padding_idx // world
        super().__init__(shard_rows(num_embeddings, rank, world), embedding_dim, local_pad,
                         init_std=init_std, device=device)
        self.global_rows, self.rank, self.world = num_embeddings, rank, world


class ShardedFM(bnn.FusedTableOwner):
    """FM of models/rank/deepfm/net.py:52-139 with both tables sharded (same state_dict names;
    `embedding*.weight` hold the LOCAL shard).  With `fused_table` (default when D+1 <= 32) the
    shard is a FusedTable and every exchange moves [emb | w1 | pad] rows in ONE all-to-all."""

    def __init__(self, sparse_feature_number, sparse_feature_dim, dense_feature_dim,
                 sparse_num_field, rank, world, group=None, device="cuda", kernels=_cuda_ops,
                 fused_table=None):
        super().__init__()
        self.k = kernels
        self.sparse_feature_number = sparse_feature_number
        self.sparse_feature_dim = sparse_feature_dim
        std = 0.1 / math.sqrt(float(sparse_feature_dim))
        self.fused = (sparse_feature_dim + 1 <= 32) if fused_table is None else bool(fused_table)
        if self.fused:
            self._fused = ShardedFusedTable(sparse_feature_number, sparse_feature_dim, 0, rank,
                                            world, init_std=std, device=device)
        else:
            self.embedding_one = ShardedEmbedding(sparse_feature_number, 1, 0, rank, world,
                                                  init_std=std, device=device)
            self.embedding = ShardedEmbedding(sparse_feature_number, sparse_feature_dim, 0, rank,
                                              world, init_std=std, device=device)
        self.dense_w_one = tnn.Parameter(torch.empty(dense_feature_dim, device=device))
        self.dense_w = tnn.Parameter(
            torch.empty(1, dense_feature_dim, sparse_feature_dim, device=device))
        tnn.init.trunc_normal_(self.dense_w_one, 0.0, std, -2 * std, 2 * std)
        tnn.init.trunc_normal_(self.dense_w, 0.0, std, -2 * std, 2 * std)
        self.exchange = ShardExchange(sparse_feature_number, rank, world, group, kernels,
                                      owner_pad=(0 if rank == 0 else -1))   # padding id 0 lives on rank 0
        self._trivial = {}

    def table_grad_dense(self):
        if self.fused:
            return self._fused.grad_dense()
        return (self.embedding.grad_rows.to_dense(), self.embedding_one.grad_rows.to_dense())

    def trivial_groups(self, n: int, device):
        """seg_offsets = 0..n and num = {n, n}: "every slot is its own segment" (cached per n so
        the backward issues no allocation / host->device copy for them)."""
        key = (n, str(device))
        hit = self._trivial.get(key)
        if hit is None:
            hit = (torch.arange(n + 1, dtype=torch.int32, device=device),
                   torch.full((2,), n, dtype=torch.int32, device=device))
            self._trivial[key] = hit
        return hit

    def forward(self, sparse_inputs, dense_inputs):
        ids = (torch.cat(list(sparse_inputs), dim=1) if isinstance(sparse_inputs, (list, tuple))
               else sparse_inputs)
        feat, y1, y2 = _ShardedEmbedFM.apply(ids, dense_inputs, self.dense_w, self.dense_w_one, self)
        return y1, y2, feat


class ShardedDeepFMLayer(tnn.Module):
    """DeepFMLayer (net.py:21-49) with sharded tables and a replicated tower."""

    def __init__(self, sparse_feature_number, sparse_feature_dim, dense_feature_dim,
                 sparse_num_field, layer_sizes, rank, world, group=None, device="cuda",
                 kernels=_cuda_ops, fused_table=None):
        super().__init__()
        from .rank.deepfm import net
        self.fm = ShardedFM(sparse_feature_number, sparse_feature_dim, dense_feature_dim,
                            sparse_num_field, rank, world, group, device, kernels, fused_table)
        self.dnn = net.DNN(sparse_feature_number, sparse_feature_dim, dense_feature_dim,
                           dense_feature_dim + sparse_num_field, layer_sizes, device=device)
        self.bias = tnn.Parameter(torch.zeros(1, device=device))
        self.world, self.group = world, group
        sync_dense_parameters(self, group)

    def forward(self, sparse_inputs, dense_inputs):
        y1, y2, feat = self.fm(sparse_inputs, dense_inputs)
        return torch.sigmoid(y1 + y2 + self.dnn(feat))

    def prefetch(self, next_sparse_inputs) -> None:
        """Hint: the ids of the NEXT batch (the very tensor that will be passed to forward)."""
        self.fm.exchange.prefetch(next_sparse_inputs)

    def finish_prefetch(self) -> None:
        self.fm.exchange.finish_prefetch()


def dense_parameters(model: tnn.Module) -> List[torch.Tensor]:
    return [p for p in model.parameters() if not getattr(p, "is_sparse_table", False)]


@torch.no_grad()
def sync_dense_parameters(model: tnn.Module, group=None) -> None:
    """Replicas must start identical: broadcast rank 0's dense parameters."""
    for p in dense_parameters(model):
        dist.broadcast(p.data, src=dist.get_global_rank(group, 0) if group is not None else 0,
                       group=group)


class DistributedOptimizer:
    """Wraps a paddlerec_b200.optim optimizer: all-reduces (SUM) the dense gradients in one flat
    bucket before the step.  Callers scale the local loss by 1/world (`scale_loss`) so that the
    summed dense gradients and the owner-summed table gradients both equal the gradient of the
    global-batch mean loss — what fleet.distributed_model does at tools/trainer.py:113-118."""

    def __init__(self, inner, model: tnn.Module, world: int, group=None):
        self.inner, self.world, self.group = inner, world, group
        self._dense = [p for p in dense_parameters(model) if p.requires_grad]
        self._side = None
        if world > 1 and hasattr(inner, "sparse_sq_reduce"):
            # ClipGradByGlobalNorm (DCN-V2): the table gradients live on their owners, so the
            # squared norm of that part is one scalar all-reduce; the dense part is identical on
            # every rank after the bucket all-reduce below.
            inner.sparse_sq_reduce = self._sum_over_ranks

    def _sum_over_ranks(self, t: torch.Tensor) -> torch.Tensor:
        t = t.clone()
        dist.all_reduce(t, group=self.group)
        return t

    def scale_loss(self, loss: torch.Tensor) -> torch.Tensor:
        return loss / self.world

    def clear_grad(self):
        self.inner.clear_grad()

    def _allreduce_dense(self):
        grads = [p.grad for p in self._dense if p.grad is not None]
        if grads and self.world > 1:
            with _timed("nccl_allreduce_dense"):
                flat = torch.cat([g.reshape(-1) for g in grads])
                dist.all_reduce(flat, group=self.group)
                off = 0
                for g in grads:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()

    def step(self):
        inner = self.inner
        split = (hasattr(inner, "step_sparse") and getattr(inner, "_clip", None) is None
                 and self.world > 1 and self._dense and self._dense[0].is_cuda)
        if not split:       # clipping needs the reduced gradients first (global norm)
            self._allreduce_dense()
            inner.step()
            return
        # The dense all-reduce runs on a side stream while the row-wise table update (the large
        # part of the optimizer) runs on the main stream; they touch disjoint tensors.
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            self._allreduce_dense()
        inner.step_sparse()
        main.wait_stream(self._side)
        inner.step_dense()


def create_sharded_deepfm(config, device, rank, world, group=None, kernels=_cuda_ops):
    return ShardedDeepFMLayer(
        config.get("hyper_parameters.sparse_feature_number"),
        config.get("hyper_parameters.sparse_feature_dim"),
        config.get("hyper_parameters.dense_input_dim"),
        config.get("hyper_parameters.sparse_inputs_slots") - 1,
        config.get("hyper_parameters.fc_sizes"), rank, world, group, device, kernels)


def create_optimizer(model, config, world=None, group=None):
    lr = config.get("hyper_parameters.optimizer.learning_rate", 0.001)
    inner = optim.Adam(learning_rate=lr, parameters=model.parameters(), lazy_mode=True)
    world = world if world is not None else dist.get_world_size(group)
    return DistributedOptimizer(inner, model, world, group)
