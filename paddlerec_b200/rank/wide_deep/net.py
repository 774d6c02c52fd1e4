"""Wide&Deep network — mirrors the reference's models/rank/wide_deep/net.py (WideDeepLayer :21-101).
The 26 per-slot lookups into the ONE shared table (no padding_idx, Uniform init, Q9) are a single
b200rec_gather over [B,26]; the deep MLP is the tensor-core tower."""
from __future__ import annotations

import math

import torch
import torch.nn as tnn

from ... import nn as bnn
from ... import tower


class WideDeepLayer(tnn.Module):
    def __init__(self, sparse_feature_number, sparse_feature_dim, dense_feature_dim, num_field,
                 layer_sizes, sync_mode="", device="cuda"):
        super().__init__()
        self.sparse_feature_dim, self.dense_feature_dim = sparse_feature_dim, dense_feature_dim
        self.num_field, self.layer_sizes, self.sync_mode = num_field, layer_sizes, sync_mode
        self.wide_part = bnn.Linear(dense_feature_dim, 1,
                                    weight_std=1.0 / math.sqrt(dense_feature_dim)).to(device)
        self.embedding = bnn.Embedding(sparse_feature_number, sparse_feature_dim, padding_idx=None,
                                       init="uniform", device=device)
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

    def forward(self, sparse_inputs, dense_inputs, show_click=None):
        if self.sync_mode == "gpubox":
            raise NotImplementedError("gpubox branch: use paddlerec_b200.sharded (CVM: see DESIGN.md)")
        wide_output = self.wide_part(dense_inputs)                                # net.py:75
        ids = (torch.cat(list(sparse_inputs), dim=1) if isinstance(sparse_inputs, (list, tuple))
               else sparse_inputs)
        emb = self.embedding(ids).reshape(ids.shape[0], -1)                       # :90-93
        deep = torch.cat([emb, dense_inputs], dim=1)                              # :95
        if bnn.get_matmul_precision() == "bf16x3" and deep.is_cuda:
            linears = [m for m in self._mlp_layers if isinstance(m, bnn.Linear)]
            deep = tower.mlp(deep, [m.weight for m in linears], [m.bias for m in linears])
        else:
            for layer in self._mlp_layers:
                deep = layer(deep)
        return torch.sigmoid(wide_output + deep)                                  # :99-101
