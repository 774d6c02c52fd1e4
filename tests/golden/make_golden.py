"""Mint the golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN net.py files
(/root/reference/models/rank/*/net.py, unmodified) in float64 on top of oracle/paddle_shim.py.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
Each .npz holds: params (state_dict names, `param/<name>`), inputs (`in/<name>`), the forward output
(`out/pred`), the scalar loss (`out/loss`) and d loss / d param for every parameter (`grad/<name>`).
Seeds are fixed; edge cases are injected on purpose: padding id 0, a sample made only of padding
ids, duplicate ids inside a sample and across the batch, DIN histories of length 1 and full length.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import nets, paddle_shim  # noqa: E402


def criteo_batch(g, B, V, F=26, Dn=13):
    ids = torch.randint(1, V, (B, F), generator=g)
    ids[0, 3] = 0                     # one padding id
    ids[1, :] = 0                     # an all-padding sample
    ids[2, 5] = ids[2, 4]             # duplicate inside a sample
    ids[3:, 0] = ids[3, 0]            # duplicate across the batch
    dense = torch.rand(B, Dn, generator=g, dtype=torch.float64)
    dense[torch.rand(B, Dn, generator=g) < 0.3] = 0.0
    label = (torch.rand(B, 1, generator=g) < 0.29).to(torch.int64)
    return ids, dense, label


def save(name, params, inputs, pred, loss, grads):
    out = {}
    for k, v in params.items():
        out["param/" + k] = v.detach().numpy()
    for k, v in inputs.items():
        out["in/" + k] = v.detach().numpy()
    out["out/pred"] = pred.detach().numpy()
    out["out/loss"] = loss.detach().numpy()
    for k, v in grads.items():
        out["grad/" + k] = v.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (path, len(out), os.path.getsize(path) / 1024))


def grads_of(loss, named):
    gs = torch.autograd.grad(loss, list(named.values()), allow_unused=True)
    return {k: (torch.zeros_like(p) if g is None else g).detach()
            for (k, p), g in zip(named.items(), gs)}


def check_oracle(tag, pred_ref, pred_oracle):
    err = (pred_ref - pred_oracle).abs().max().item()
    assert err < 1e-12, (tag, err)
    print("  oracle vs reference-on-shim (%s): max abs diff %.3e" % (tag, err))


def make_deepfm(D, tag):
    ref = paddle_shim.import_reference_net("deepfm")
    torch.manual_seed(12345)
    V, F, Dn, fc = 211, 26, 13, [32, 16]
    layer = ref.DeepFMLayer(V, D, Dn, F, fc)
    g = torch.Generator().manual_seed(777)
    ids, dense, label = criteo_batch(g, 12, V)
    sparse_inputs = [ids[:, i:i + 1] for i in range(F)]
    named = dict(layer.named_parameters())
    pred = layer(sparse_inputs, dense)
    loss = nets.log_loss(pred, label.to(pred.dtype)).mean()
    check_oracle("deepfm D=%d" % D, pred, nets.deepfm_forward(named, sparse_inputs, dense, len(fc)))
    save("deepfm_%s" % tag, named, {"ids": ids, "dense": dense, "label": label}, pred, loss,
         grads_of(loss, named))


def make_dcn_v2(mix, stacked, tag):
    ref = paddle_shim.import_reference_net("dcn_v2")
    torch.manual_seed(2024)
    V, D, F, Dn, fc = 157, 4, 26, 13, [24, 16]
    layer = ref.DCN_V2Layer(V, D, Dn, F, fc, 2, stacked, mix, 6, 4)
    layer.eval()  # Dropout(0.5) after every sublayer in train mode (Q5): parity is defined in eval
    g = torch.Generator().manual_seed(4242)
    ids, dense, label = criteo_batch(g, 10, V)
    sparse_inputs = [ids[:, i:i + 1] for i in range(F)]
    named = dict(layer.named_parameters())
    pred = layer(sparse_inputs, dense)
    loss = nets.log_loss(pred, label.to(pred.dtype)).mean()
    check_oracle("dcn_v2 " + tag, pred,
                 nets.dcn_v2_forward(named, sparse_inputs, dense, n_fc=len(fc), cross_num=2,
                                     is_stacked=stacked, use_low_rank_mixture=mix, num_experts=4))
    save("dcn_v2_%s" % tag, named, {"ids": ids, "dense": dense, "label": label}, pred, loss,
         grads_of(loss, named))


def make_din():
    ref = paddle_shim.import_reference_net("din")
    torch.manual_seed(31337)
    item_count, cat_count, E2 = 97, 23, 8   # item/cat emb size 8 -> E = 16
    layer = ref.DINLayer(E2, E2, "sigmoid", False, False, item_count, cat_count)
    g = torch.Generator().manual_seed(99)
    B, L = 6, 7
    lens = [1, 7, 3, 5, 7, 2]
    hist_item = torch.randint(1, item_count, (B, L), generator=g)
    hist_cat = torch.randint(1, cat_count, (B, L), generator=g)
    mask = torch.zeros(B, L, 1, dtype=torch.int64)
    for b, n in enumerate(lens):   # dinReader.py:70-101: pad with id 0, mask -1e9 (as int64)
        hist_item[b, n:] = 0
        hist_cat[b, n:] = 0
        mask[b, n:, 0] = int(-1e9)
    target_item = torch.randint(1, item_count, (B,), generator=g)
    target_cat = torch.randint(1, cat_count, (B,), generator=g)
    target_item[1] = hist_item[1, 0]   # target equal to a history item
    label = (torch.rand(B, 1, generator=g) < 0.5).to(torch.float64)
    target_item_seq = target_item.unsqueeze(1).repeat(1, L)
    target_cat_seq = target_cat.unsqueeze(1).repeat(1, L)
    named = dict(layer.named_parameters())
    # the attention-unit linears are hidden from named_parameters() by the name collision (Q6)
    att = [m for m in layer.attention_layer if hasattr(m, "weight")]
    for i, m in enumerate(att):
        named["att.linear_%d.weight" % i] = m.weight
        named["att.linear_%d.bias" % i] = m.bias
    # give item_b non-zero values so that path is exercised (reference initialises it to 0)
    with torch.no_grad():
        named["item_b_attr.weight"].normal_(0, 0.1)
    args = (hist_item, hist_cat, target_item, target_cat, label, mask, target_item_seq,
            target_cat_seq)
    logit = layer(*args)
    loss = nets.bce_with_logits(logit, label)
    check_oracle("din", logit, nets.din_forward(named, *args))
    save("din", named,
         {"hist_item": hist_item, "hist_cat": hist_cat, "target_item": target_item,
          "target_cat": target_cat, "label": label, "mask": mask,
          "lens": torch.tensor(lens)}, logit, loss, grads_of(loss, named))


def make_wide_deep():
    ref = paddle_shim.import_reference_net("wide_deep")
    torch.manual_seed(555)
    V, D, F, Dn, fc = 131, 8, 26, 13, [32, 16]
    layer = ref.WideDeepLayer(V, D, Dn, F, fc)
    g = torch.Generator().manual_seed(321)
    ids, dense, label = criteo_batch(g, 9, V)   # id 0 is a REAL row here (no padding_idx, Q9)
    sparse_inputs = [ids[:, i:i + 1] for i in range(F)]
    named = dict(layer.named_parameters())
    pred = layer(sparse_inputs, dense)
    loss = nets.log_loss(pred, label.to(pred.dtype)).mean()
    check_oracle("wide_deep", pred, nets.wide_deep_forward(named, sparse_inputs, dense, len(fc)))
    save("wide_deep", named, {"ids": ids, "dense": dense, "label": label}, pred, loss,
         grads_of(loss, named))


def make_dlrm(self_interaction, tag):
    """dlrm/net.py in TRAIN mode (BatchNorm uses batch statistics); also stores the running
    statistics after the step under `buf/<name>`-style params (`_mean`, `_variance`)."""
    ref = paddle_shim.import_reference_net("dlrm")
    torch.manual_seed(777)
    V, D, F, Dn, bot, top = 257, 8, 26, 13, [32, 16, 8], [32, 16, 2]
    layer = ref.DLRMLayer(Dn, bot, V, D, top, F, self_interaction=self_interaction)
    layer.train()
    g = torch.Generator().manual_seed(99)
    ids, dense, label = criteo_batch(g, 11, V)  # no padding_idx: id 0 is a real row
    sparse_inputs = [ids[:, i:i + 1] for i in range(F)]
    named = dict(layer.named_parameters())
    pred = layer(sparse_inputs, dense)           # raw [B, 2] scores
    loss = nets.softmax_cross_entropy(pred, label)
    check_oracle("dlrm_" + tag, pred, nets.dlrm_forward(named, sparse_inputs, dense, n_bot=len(bot),
                                                        n_top=len(top),
                                                        self_interaction=self_interaction))
    params = dict(named)
    params.update({k: v for k, v in layer.named_buffers()})   # running stats AFTER this step
    save("dlrm_" + tag, params, {"ids": ids, "dense": dense, "label": label}, pred, loss,
         grads_of(loss, named))


if __name__ == "__main__":
    torch.set_default_dtype(torch.float64)
    make_deepfm(9, "d9")      # the reference's own D (config.yaml:51): scalar row path
    make_deepfm(16, "d16")    # perf D: 128-bit row path
    make_dcn_v2(False, True, "v2_stacked")
    make_dcn_v2(True, False, "mix_parallel")
    make_din()
    make_wide_deep()
    make_dlrm(False, "pairs")
    make_dlrm(True, "self")
