#!/usr/bin/env python
"""BASELINE config 3 (single-GPU leg): DCN-V2 Criteo-shape training step — 26-slot gather,
CrossNetV2 x2 (GEMM + fused K3 epilogues), MLP tower, Adam with global-norm clipping.
D=40 is the reference's embedding size (models/rank/dcn_v2/config.yaml); the cross input is 39*D."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paddlerec_b200 import nn as bnn  # noqa: E402
from paddlerec_b200.rank.dcn_v2.dygraph_model import DygraphModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--V", type=int, default=100_000_000)
    ap.add_argument("--D", type=int, default=40)
    ap.add_argument("--mix", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--precision", default="bf16x3")
    a = ap.parse_args()
    bnn.set_matmul_precision(a.precision)
    cfg = {"hyper_parameters.sparse_feature_number": a.V, "hyper_parameters.sparse_feature_dim": a.D,
           "hyper_parameters.fc_sizes": [768, 768], "hyper_parameters.dense_input_dim": 13,
           "hyper_parameters.sparse_inputs_slots": 27, "hyper_parameters.cross_num": 2,
           "hyper_parameters.is_Stacked": True, "hyper_parameters.use_low_rank_mixture": bool(a.mix),
           "hyper_parameters.low_rank": 256, "hyper_parameters.num_experts": 4,
           "hyper_parameters.optimizer.learning_rate": 1e-4}
    torch.manual_seed(12345)
    dm = DygraphModel()
    model = dm.create_model(cfg)
    model.eval()   # Dropout(0.5) off: timing of the deterministic path (same kernels)
    opt = dm.create_optimizer(model, cfg)
    g = torch.Generator().manual_seed(1)
    batches = []
    for _ in range(4):
        ids = torch.randint(1, a.V, (a.B, 26), generator=g)
        ids[torch.rand(a.B, 26, generator=g) < 0.02] = 0
        dense = torch.log1p(torch.rand(a.B, 13, generator=g) * 10)
        label = (torch.rand(a.B, 1, generator=g) < 0.29).long()
        batches.append(tuple(t.cuda() for t in (label, ids, dense)))

    def step(i):
        opt.clear_grad()
        loss, _, _ = dm.train_forward(model, None, batches[i % 4], cfg)
        loss.backward()
        opt.step()
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(json.dumps({"what": "DCN-V2 train step (stacked, CrossNet%s x2, fc [768,768])" %
                              ("Mix" if a.mix else "V2"), "B": a.B, "V": a.V, "D": a.D,
                      "precision": a.precision, "ms": ms, "samples_per_s": a.B / ms * 1e3}))


if __name__ == "__main__":
    main()
