"""Wide&Deep network — mirrors the reference's models/rank/wide_deep/net.py (WideDeepLayer :21-101).
The 26 per-slot lookups into the ONE shared table (no padding_idx, Uniform init, Q9) are a single
b200rec_gather over [B,26]; the deep MLP is the tensor-core tower.

`sync_mode="gpubox"` is the reference's PSGPU branch (:80-88): the table rows are
[show, click, embedding(D)] (`sparse_embedding(size=[V, D+2])`), every looked-up row goes through
`continuous_value_model(emb, show_click, use_cvm=False)` which drops the two statistic columns in
forward and, in backward, puts the sample's (show, click) into those two gradient columns so the
table accumulates them.  Here: uint64 feasigns are folded to rows on the device
(b200rec_hash_keys, per-slot salted), the lookup is the same b200rec_gather (or the row-sharded
exchange of paddlerec_b200.sharded after shard_embeddings(): the PSGPU pull/push of
tools/static_gpubox_trainer.py:152-159,244-259), CVM is b200rec_cvm_fwd/_bwd and the optimizers
accumulate the statistic columns instead of descending on them (optim._table_parts)."""
from __future__ import annotations

import math

import torch
import torch.nn as tnn

from ... import nn as bnn
from ... import ops
from ... import tower


class WideDeepLayer(tnn.Module):
    def __init__(self, sparse_feature_number, sparse_feature_dim, dense_feature_dim, num_field,
                 layer_sizes, sync_mode="", device="cuda"):
        super().__init__()
        self.sparse_feature_dim, self.dense_feature_dim = sparse_feature_dim, dense_feature_dim
        self.num_field, self.layer_sizes, self.sync_mode = num_field, layer_sizes, sync_mode
        self.wide_part = bnn.Linear(dense_feature_dim, 1,
                                    weight_std=1.0 / math.sqrt(dense_feature_dim)).to(device)
        self.sparse_feature_number = sparse_feature_number
        gpubox = sync_mode == "gpubox"
        self.embedding = bnn.Embedding(sparse_feature_number,
                                       sparse_feature_dim + (2 if gpubox else 0), padding_idx=None,
                                       init="uniform", device=device)
        if gpubox:   # rows = [show, click, embedding]: statistics start at zero and are accumulated
            with torch.no_grad():
                self.embedding.weight[:, :2].zero_()
            self.embedding.weight.cvm_stat_cols = 2
        sizes = [sparse_feature_dim * num_field + dense_feature_dim] + list(layer_sizes) + [1]
        self._mlp_layers = []
        for i in range(len(layer_sizes) + 1):
            linear = bnn.Linear(sizes[i], sizes[i + 1],
                                weight_std=1.0 / math.sqrt(sizes[i])).to(device)
            self.add_module("linear_%d" % i, linear)
            self._mlp_layers.append(linear)
            if i < len(layer_sizes):
                act = tnn.ReLU()
                self.add_module("act_%d" % i, act)
                self._mlp_layers.append(act)

    def forward(self, sparse_inputs, dense_inputs, show_click=None, feasigns=False):
        """gpubox: `show_click` [B,2] is required; with `feasigns` the sparse inputs are raw uint64
        keys (int64 bit patterns) that are hashed to rows on the device, slot-salted."""
        wide_output = self.wide_part(dense_inputs)                                # net.py:75
        ids = (torch.cat(list(sparse_inputs), dim=1) if isinstance(sparse_inputs, (list, tuple))
               else sparse_inputs)
        B, F = ids.shape
        if self.sync_mode == "gpubox":                                            # :80-88
            if show_click is None:
                raise ValueError("gpubox branch needs show_click [B,2]")
            if feasigns:
                slot = torch.arange(F, dtype=torch.int32, device=ids.device).repeat(B)
                ids = ops.raw_hash_keys(ids.reshape(-1), self.sparse_feature_number, slot).reshape(B, F)
            rows = self.embedding(ids).reshape(B * F, self.sparse_feature_dim + 2)
            per_row = show_click.to(torch.float32).repeat_interleave(F, dim=0)
            emb = ops.continuous_value_model(rows, per_row, False).reshape(B, -1)
        else:
            emb = self.embedding(ids).reshape(B, -1)                              # :90-93
        deep = torch.cat([emb, dense_inputs], dim=1)                              # :95
        if bnn.get_matmul_precision() == "bf16x3" and deep.is_cuda:
            linears = [m for m in self._mlp_layers if isinstance(m, bnn.Linear)]
            deep = tower.mlp(deep, [m.weight for m in linears], [m.bias for m in linears])
        else:
            for layer in self._mlp_layers:
                deep = layer(deep)
        if deep.is_cuda:
            return ops.sum_sigmoid(wide_output, deep)                             # :99-101, fused
        return torch.sigmoid(wide_output + deep)                                  # :99-101
