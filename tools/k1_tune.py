#!/usr/bin/env python
"""Times the fused gather+FM forward (K1) for each (unroll, row-cache policy) build-time variant,
two-table vs fused-slot layout, V=1e8, D=16, B=65536.  One subprocess per env setting."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    from paddlerec_b200 import ops
    dev, B, F, Dn, D, V = "cuda", 65536, 26, 13, 16, 100_000_000
    g = torch.Generator().manual_seed(1)
    W = torch.empty(V, D, device=dev).uniform_(-0.05, 0.05)
    W1 = torch.empty(V, 1, device=dev).uniform_(-0.05, 0.05)
    Wf = torch.zeros(V, 32, device=dev)
    Wf[:, :D] = W
    Wf[:, D] = W1[:, 0]
    dw = torch.randn(Dn, D, device=dev) * 0.05
    dw1 = torch.randn(Dn, device=dev) * 0.05
    ids = [torch.randint(1, V, (B, F), generator=g).to(dev) for _ in range(4)]
    den = [torch.rand(B, Dn, generator=g).to(dev) for _ in range(4)]

    def t(fn):
        for i in range(3):
            fn(i % 4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            fn(i % 4)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20
    # correctness of this variant on a slice (bit-exact gathers, fp64 sums)
    feat, y1, y2, S = ops.raw_embed_fm_fwd(Wf, None, ids[0], den[0], dw, dw1, 0, D=D)
    n = 4096
    rows = Wf[ids[0][:n]]                                   # [n,F,32]
    live = (ids[0][:n] != 0).unsqueeze(2)
    fe = torch.cat([rows[..., :D] * live, den[0][:n].unsqueeze(2) * dw.unsqueeze(0)], 1)
    ok = bool(torch.equal(feat[:n], fe))
    Sd = fe.double().sum(1)
    y2r = 0.5 * (Sd.square() - fe.double().square().sum(1)).sum(1)
    y1r = (rows[..., D].double() * live[..., 0]).sum(1) + (den[0][:n].double() * dw1.double()).sum(1)
    err = max(float((y2[:n].double() - y2r).abs().max()), float((y1[:n].double() - y1r).abs().max()),
              float((S[:n].double() - Sd).abs().max()))
    alg = B * (F * 8 + Dn * 4 + F * 4 * D + F * 4 + (F + Dn) * 4 * D + 8)
    a = t(lambda i: ops.raw_embed_fm_fwd(W, W1, ids[i], den[i], dw, dw1, 0))
    b = t(lambda i: ops.raw_embed_fm_fwd(Wf, None, ids[i], den[i], dw, dw1, 0, D=D))
    print(json.dumps({"unroll": os.environ.get("B200REC_K1_UNROLL"),
                      "cache": os.environ.get("B200REC_K1_CACHE"),
                      "ctas_per_sm": os.environ.get("B200REC_K1_CTAS"),
                      "debug_mask": os.environ.get("B200REC_K1_DEBUG"),
                      "tma_variant": os.environ.get("B200REC_K1_TMA"),
                      "minb": os.environ.get("B200REC_K1_MINB"),
                      "warp_sched": os.environ.get("B200REC_K1_WARP"),
                      "feat_exact": ok, "max_abs_err": err,
                      "two_tables_ms": a, "two_tables_GBps": alg / a / 1e6,
                      "fused_slots_ms": b, "fused_slots_GBps": alg / b / 1e6}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        if len(sys.argv) > 1 and sys.argv[1] == "minb":
            for u, mb, ctas in (("13", "4", "8"), ("13", "5", "5"), ("13", "5", "10"), ("13", "6", "6"),
                                ("13", "6", "12"), ("8", "6", "6"), ("8", "8", "8"), ("8", "8", "16")):
                env = dict(os.environ, B200REC_K1_UNROLL=u, B200REC_K1_MINB=mb, B200REC_K1_CTAS=ctas)
                subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
            sys.exit(0)
        if len(sys.argv) > 1 and sys.argv[1] == "warp":
            # CTA-tiled kernel vs the warp-scheduled one (dynamic groups, no CTA barriers)
            for w, u, ctas in (("0", "13", "8"), ("1", "13", "4"), ("1", "13", "8"), ("1", "13", "16"),
                               ("1", "8", "4"), ("1", "8", "8"), ("1", "8", "16")):
                env = dict(os.environ, B200REC_K1_WARP=w, B200REC_K1_UNROLL=u, B200REC_K1_CTAS=ctas)
                subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
            sys.exit(0)
        if len(sys.argv) > 1 and sys.argv[1] == "tma":
            for tma in ("0", "1"):
                env = dict(os.environ, B200REC_K1_TMA=tma)
                subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
            sys.exit(0)
        if len(sys.argv) > 1 and sys.argv[1] == "decompose":
            # where does K1's time go?  0 = full kernel, 1 = no feat stores, 2 = no table loads
            for dbg in ("0", "1", "2", "3"):
                env = dict(os.environ, B200REC_K1_DEBUG=dbg, B200REC_K1_TMA="0")
                subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
            sys.exit(0)
        for u, ctas in (("8", "4"), ("8", "5"), ("8", "8"), ("13", "3"), ("13", "4"), ("13", "8"),
                        ("26", "2")):
            for c in ("0", "1"):
                env = dict(os.environ, B200REC_K1_UNROLL=u, B200REC_K1_CACHE=c,
                           B200REC_K1_CTAS=ctas)
                subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
