"""DIN network — mirrors the reference's models/rank/din/net.py (DINLayer.__init__ :21-137,
forward :139-184): seven independent embedding tables (Q8), an attention unit MLP
(4E -> 80 -> 40 -> 1, sigmoid) over [h, t, h-t, h*t], additive -1e9 mask BEFORE the E^-0.5 scale
(Q7), softmax over the history, weighted-sum pooling, output MLP, + item bias.

CUDA path: the seven lookups are b200rec_gather (sorted segment-reduce backward); the target
`*_seq` lookups exploit the reader's tiling (dinReader.py:85-90 repeats the target id L times) and
gather ONE row per sample; the attention unit + masked softmax + pooling forward is the fused
b200rec_din_attn_fwd kernel when available (ops.din_attention), with the [B,L,4E] concat never
materialised.

Q6 (sub-layer name collision): the reference registers the attention linears and the output-MLP
linears under the same names `linear_0..2`, so the attention linears vanish from its
parameters()/state_dict() and are never trained.  Here they live under `attention.linear_{i}`;
`faithful_frozen_attention=True` (default) keeps them out of the optimizer exactly like the
reference; set it False to train them (their gradients are parity-tested either way).
"""
from __future__ import annotations

import torch
import torch.nn as tnn

from ... import nn as bnn
from ... import ops


class _MLP3(tnn.Module):
    """Linear-Sigmoid-Linear-Sigmoid-Linear with the reference's `linear_%d` names."""

    def __init__(self, sizes):
        super().__init__()
        for i in range(3):
            self.add_module("linear_%d" % i, bnn.Linear(sizes[i], sizes[i + 1], xavier=True))

    def forward(self, x):
        x = torch.sigmoid(self.linear_0(x))
        x = torch.sigmoid(self.linear_1(x))
        return self.linear_2(x)


class DINLayer(tnn.Module):
    def __init__(self, item_emb_size, cat_emb_size, act, is_sparse, use_DataLoader, item_count,
                 cat_count, device="cuda", faithful_frozen_attention=True,
                 tiled_target_seq=True, fused_attention=True):
        super().__init__()
        self.fused_attention = fused_attention
        self.item_emb_size, self.cat_emb_size = item_emb_size, cat_emb_size
        self.item_count, self.cat_count = item_count, cat_count
        self.tiled_target_seq = tiled_target_seq

        def table(n, d, init="xavier_uniform"):
            return bnn.Embedding(n, d, padding_idx=None, init=init, device=device)

        self.hist_item_emb_attr = table(item_count, item_emb_size)
        self.hist_cat_emb_attr = table(cat_count, cat_emb_size)
        self.target_item_emb_attr = table(item_count, item_emb_size)
        self.target_cat_emb_attr = table(cat_count, cat_emb_size)
        self.target_item_seq_emb_attr = table(item_count, item_emb_size)
        self.target_cat_seq_emb_attr = table(cat_count, cat_emb_size)
        self.item_b_attr = table(item_count, 1, init="zeros")

        E = item_emb_size + cat_emb_size
        self.firInDim = self.firOutDim = E
        self.attention = _MLP3([4 * E, 80, 40, 1]).to(device)          # net.py:84-104
        # absent from a reference-produced checkpoint (Q6): checkpoint.set_state_dict tolerates it
        self.optional_state_prefixes = ("attention.",)
        if faithful_frozen_attention:
            for p in self.attention.parameters():
                p.requires_grad_(False)
        self.linearCon = bnn.Linear(E, E, xavier=True).to(device)      # :111-119
        con = _MLP3([2 * E, 80, 40, 1]).to(device)                     # :121-137
        self.linear_0, self.linear_1, self.linear_2 = con.linear_0, con.linear_1, con.linear_2

    def attention_pool(self, hist, tseq, mask):
        """net.py:155-173 on already gathered rows.  hist [B,L,E]; tseq [B,E] (tiled) or [B,L,E];
        mask [B,L,1] (0 / -1e9, any dtype).  Returns [B,E]."""
        att = self.attention
        if (ops.HAVE_DIN_ATTN and self.fused_attention and hist.is_cuda and tseq.dim() == 2
                and hist.shape[2] % 4 == 0 and hist.shape[2] <= 128
                and att.linear_0.out_features == 80 and att.linear_1.out_features == 40):
            return ops.din_attention(hist.contiguous(), tseq.contiguous(), mask,
                                     att.linear_0.weight, att.linear_0.bias, att.linear_1.weight,
                                     att.linear_1.bias, att.linear_2.weight, att.linear_2.bias)
        if tseq.dim() == 2:
            tseq = tseq.unsqueeze(1).expand_as(hist)
        concat = torch.cat([hist, tseq, hist - tseq, hist * tseq], dim=2)
        a = att(concat)
        a = a + mask.to(a.dtype)
        a = a.transpose(1, 2) * (self.firInDim ** -0.5)
        w = torch.softmax(a, dim=-1)
        return torch.matmul(w, hist).reshape(-1, self.firInDim)

    def forward(self, hist_item_seq, hist_cat_seq, target_item, target_cat, label, mask,
                target_item_seq, target_cat_seq):
        hist = torch.cat([self.hist_item_emb_attr(hist_item_seq),
                          self.hist_cat_emb_attr(hist_cat_seq)], dim=2)           # :141-149
        target = torch.cat([self.target_item_emb_attr(target_item),
                            self.target_cat_emb_attr(target_cat)], dim=1)         # :143-153
        if self.tiled_target_seq:   # one row per sample instead of L identical rows
            tseq = torch.cat([self.target_item_seq_emb_attr(target_item_seq[:, 0].contiguous()),
                              self.target_cat_seq_emb_attr(target_cat_seq[:, 0].contiguous())], 1)
        else:
            tseq = torch.cat([self.target_item_seq_emb_attr(target_item_seq),
                              self.target_cat_seq_emb_attr(target_cat_seq)], dim=2)
        item_b = self.item_b_attr(target_item)                                    # :147
        output = self.attention_pool(hist, tseq, mask)
        concat = self.linearCon(output)                                           # :175-176
        x = torch.cat([concat, target], dim=1)                                    # :178
        x = torch.sigmoid(self.linear_0(x))
        x = torch.sigmoid(self.linear_1(x))
        x = self.linear_2(x)
        return x + item_b                                                         # :183
