"""Shared helpers for the tests: golden loading and oracle drivers (CPU, float64/float32)."""
import os

import numpy as np
import torch

from oracle import nets

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {"param": {}, "in": {}, "out": {}, "grad": {}}
    for k in z.files:
        group, _, key = k.partition("/")
        out[group][key] = z[k]
    return out


def to_params(d, dtype=torch.float64, requires_grad=True):
    return {k: torch.tensor(v, dtype=dtype).requires_grad_(requires_grad) for k, v in d.items()}


def slots(ids):
    ids = torch.as_tensor(ids)
    return [ids[:, i:i + 1] for i in range(ids.shape[1])]


def n_fc(params, prefix):
    return len([k for k in params if k.startswith(prefix + "linear_") and k.endswith(".weight")])


def oracle_run(model, g, dtype=torch.float64):
    """Run the oracle on a golden's params+inputs; returns (pred, loss, grads dict)."""
    p = to_params(g["param"], dtype)
    i = g["in"]
    if model == "deepfm":
        dense = torch.tensor(i["dense"], dtype=dtype)
        pred = nets.deepfm_forward(p, slots(i["ids"]), dense, n_fc(p, "dnn.") - 1)
        loss = nets.log_loss(pred, torch.tensor(i["label"], dtype=dtype)).mean()
    elif model.startswith("dcn_v2"):
        dense = torch.tensor(i["dense"], dtype=dtype)
        mix = any("U_list" in k for k in p)
        stacked = p["fc.weight"].shape[0] == p["DNN_.linear_%d.weight" % (n_fc(p, "DNN_.") - 1)].shape[1]
        pred = nets.dcn_v2_forward(p, slots(i["ids"]), dense, n_fc=n_fc(p, "DNN_."), cross_num=2,
                                   is_stacked=stacked, use_low_rank_mixture=mix, num_experts=4)
        loss = nets.log_loss(pred, torch.tensor(i["label"], dtype=dtype)).mean()
    elif model == "din":
        L = i["hist_item"].shape[1]
        ti, tc = torch.tensor(i["target_item"]), torch.tensor(i["target_cat"])
        label = torch.tensor(i["label"], dtype=dtype)
        pred = nets.din_forward(p, torch.tensor(i["hist_item"]), torch.tensor(i["hist_cat"]), ti, tc,
                                label, torch.tensor(i["mask"]), ti.unsqueeze(1).repeat(1, L),
                                tc.unsqueeze(1).repeat(1, L))
        loss = nets.bce_with_logits(pred, label)
    elif model == "wide_deep":
        dense = torch.tensor(i["dense"], dtype=dtype)
        pred = nets.wide_deep_forward(p, slots(i["ids"]), dense, n_fc(p, "") - 1)
        loss = nets.log_loss(pred, torch.tensor(i["label"], dtype=dtype)).mean()
    elif model == "dlrm":
        dense = torch.tensor(i["dense"], dtype=dtype)
        n_bot = len([k for k in p if k.startswith("bot_mlp.dense_") and k.endswith(".weight")])
        n_top = len([k for k in p if k.startswith("top_mlp.dense_") and k.endswith(".weight")])
        N = i["ids"].shape[1] + 1
        d = p["embedding.weight"].shape[1]
        self_int = p["top_mlp.dense_0.weight"].shape[0] == d + N * (N + 1) // 2
        # the golden's `_mean` / `_variance` are the running statistics AFTER the step: not inputs
        stats = {k: p.pop(k) for k in list(p) if k.endswith("._mean") or k.endswith("._variance")}
        pred = nets.dlrm_forward(p, slots(i["ids"]), dense, n_bot=n_bot, n_top=n_top,
                                 self_interaction=self_int)
        loss = nets.softmax_cross_entropy(pred, torch.tensor(i["label"]))
    else:
        raise ValueError(model)
    gs = torch.autograd.grad(loss, list(p.values()), allow_unused=True)
    grads = {k: (torch.zeros_like(v) if gg is None else gg) for (k, v), gg in zip(p.items(), gs)}
    return pred.detach(), loss.detach(), grads


def _np(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float64)


def rel_err(a, b):
    a = _np(a)
    b = _np(b)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def elementwise_excess(a, b, rtol=1e-4, floor_frac=1e-4):
    """Largest amount by which |a - b| exceeds rtol * |b| + floor, floor = floor_frac * rms(b): an
    ELEMENT-WISE bound (rel_err above is a tensor max-norm bound, lenient on small entries).  The
    floor is tied to the tensor's own scale so entries that are zero up to rounding do not fail
    and a wrong row cannot hide.  <= 0 means every entry is within tolerance."""
    a = _np(a)
    b = _np(b)
    floor = floor_frac * float(np.sqrt(np.mean(np.square(b)))) + 1e-30
    return float((np.abs(a - b) - rtol * np.abs(b) - floor).max())
