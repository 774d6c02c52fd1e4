"""DLRM network — same classes / constructor arguments / state_dict names as the reference's
models/rank/dlrm/net.py (DLRMLayer :23-118, MLPLayer :121-171).

CUDA path: the 26 per-slot lookups into the one table (no padding_idx, TruncatedNormal) are a single
b200rec_gather over [B,26]; `bmm(T, T^T)` + triu/tril/masked_select + `concat([x, Zflat])`
(net.py:103-115) is ONE kernel, b200rec_dot_interact (K6), that never materialises the [B,27,27]
product; the two MLPs are library GEMMs with torch's batch-norm.
Quirks kept: MLPLayer's guard `i != len(units_list) - 1` is always true, so EVERY layer — the last
one producing the 2 class scores included — is Linear -> ReLU -> BatchNorm1D (net.py:134-151);
with self_interaction=True the extra diagonal positions hold 0, not <e_i, e_i> (net.py:105-113).
"""
from __future__ import annotations

import math

import torch
import torch.nn as tnn

from ... import nn as bnn
from ... import ops


class MLPLayer(tnn.Module):
    def __init__(self, input_shape, units_list=None, activation=None, device="cuda"):
        super().__init__()
        if units_list is None:
            units_list = [128, 128, 64]
        units_list = [input_shape] + list(units_list)
        self.units_list = units_list
        self.activation = activation
        self.mlp = []
        for i, unit in enumerate(units_list[:-1]):
            dense = bnn.Linear(unit, units_list[i + 1], weight_std=1.0 / math.sqrt(unit),
                               truncated=True).to(device)
            self.add_module("dense_%d" % i, dense)
            relu = tnn.ReLU()
            self.add_module("relu_%d" % i, relu)
            norm = bnn.BatchNorm1D(units_list[i + 1]).to(device)
            self.add_module("norm_%d" % i, norm)
            self.mlp += [dense, relu, norm]

    def forward(self, inputs):
        outputs = inputs
        for layer in self.mlp:
            outputs = layer(outputs)
        return outputs


class DLRMLayer(tnn.Module):
    def __init__(self, dense_feature_dim, bot_layer_sizes, sparse_feature_number,
                 sparse_feature_dim, top_layer_sizes, num_field, sync_mode=None,
                 self_interaction=False, device="cuda"):
        super().__init__()
        self.dense_feature_dim = dense_feature_dim
        self.bot_layer_sizes = bot_layer_sizes
        self.sparse_feature_number = sparse_feature_number
        self.sparse_feature_dim = sparse_feature_dim
        self.top_layer_sizes = top_layer_sizes
        self.num_field = num_field
        self.self_interaction = self_interaction
        if bot_layer_sizes[-1] != sparse_feature_dim:
            raise ValueError("the bottom MLP must end at sparse_feature_dim (%d != %d): its output is "
                             "stacked with the embedding rows" % (bot_layer_sizes[-1], sparse_feature_dim))
        self.bot_mlp = MLPLayer(dense_feature_dim, bot_layer_sizes, activation="relu", device=device)
        self.concat_size = ops.dot_interact_width(num_field + 1, 0, self_interaction)
        self.top_mlp = MLPLayer(self.concat_size + sparse_feature_dim, top_layer_sizes, device=device)
        self.embedding = bnn.Embedding(sparse_feature_number, sparse_feature_dim, padding_idx=None,
                                       init="truncated_normal", init_std=1.0, device=device)

    def forward(self, sparse_inputs, dense_inputs):
        x = self.bot_mlp(dense_inputs)                                            # net.py:86
        ids = (torch.cat(list(sparse_inputs), dim=1) if isinstance(sparse_inputs, (list, tuple))
               else sparse_inputs)
        emb = self.embedding(ids)                                                 # [B, 26, d]  :89-94
        T = torch.cat([emb, x.unsqueeze(1)], dim=1)                               # :97-100
        R = ops.dot_interact(T, self.self_interaction)                            # :103-115 (K6)
        return self.top_mlp(R)                                                    # :117
