#!/usr/bin/env python
"""BASELINE config 4: DIN AmazonElec-shape (item 63001, cat 801, E=64+64, history length 100).
Times (CUDA events) the fused attention-pooling forward kernel pair (K4) alone and one full
training step of DINLayer (gathers, K4 forward and fused backward, SGD); prints JSON lines."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paddlerec_b200 import ops  # noqa: E402
from paddlerec_b200.rank.din import net  # noqa: E402
from paddlerec_b200.rank.din.dygraph_model import DygraphModel  # noqa: E402


def timeit(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--L", type=int, default=100)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = "cuda"
    B, L, E = a.B, a.L, 128
    g = torch.Generator().manual_seed(12345)
    dm = DygraphModel()
    cfg = {"hyper_parameters.item_count": 63001, "hyper_parameters.cat_count": 801,
           "hyper_parameters.optimizer.learning_rate_base_lr": 0.85}
    model = dm.create_model(cfg)
    for p in model.attention.parameters():
        p.requires_grad_(True)
    opt = dm.create_optimizer(model, cfg)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    hist_item = torch.randint(1, 63001, (B, L), generator=g)
    hist_cat = torch.randint(1, 801, (B, L), generator=g)
    mask = torch.zeros(B, L, 1, dtype=torch.int64)
    ar = torch.arange(L).unsqueeze(0)
    pad = ar >= lens.unsqueeze(1)
    hist_item[pad] = 0
    hist_cat[pad] = 0
    mask[pad] = int(-1e9)
    ti = torch.randint(1, 63001, (B,), generator=g)
    tc = torch.randint(1, 801, (B,), generator=g)
    label = (torch.rand(B, generator=g) < 0.5).float()
    batch = [hist_item, hist_cat, ti, tc, label, mask, ti.unsqueeze(1).repeat(1, L),
             tc.unsqueeze(1).repeat(1, L)]
    batch = [t.to(dev) for t in batch]

    hist = torch.randn(B, L, E, device=dev) * 0.1
    tseq = torch.randn(B, E, device=dev) * 0.1
    att = model.attention
    params = [att.linear_0.weight, att.linear_0.bias, att.linear_1.weight, att.linear_1.bias,
              att.linear_2.weight, att.linear_2.bias]
    ms = timeit(lambda: ops.raw_din_attn_fwd(hist, tseq, batch[5], *params), a.iters)
    flop = 2.0 * B * L * (2 * E * 80 + 80 * 40 + 40)
    print(json.dumps({"kernel": "din_attn_fwd (K4: scores + softmax/pool)", "B": B, "L": L, "E": E,
                      "ms": ms, "TFLOPs_fp32": flop / ms / 1e9,
                      "positions_per_s": B * L / ms * 1e3,
                      "gather_GBps_equiv": B * (2 * L * 8 + L * 4 * E) / ms / 1e6}), flush=True)
    with torch.no_grad():
        comp = timeit(lambda: ops._din_attention_composite(hist[:512], tseq[:512], batch[5][:512],
                                                           *params), a.iters) * (B / 512)
    print(json.dumps({"kernel": "torch composite of din/net.py:155-173 (extrapolated from 512 "
                                "samples)", "ms": comp}), flush=True)

    def step():
        opt.clear_grad()
        loss, _, _ = dm.train_forward(model, None, batch, cfg)
        loss.backward()
        opt.step()
    ms = timeit(step, a.iters)
    print(json.dumps({"what": "DIN train step (gathers, K4 fwd + fused bwd, SGD)", "B": B, "L": L,
                      "ms": ms, "samples_per_s": B / ms * 1e3}), flush=True)


if __name__ == "__main__":
    main()
