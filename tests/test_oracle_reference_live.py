"""Live cross-check (authoring container only: needs /root/reference): the reference's own
models/rank/*/net.py, executed unmodified on oracle/paddle_shim.py with FRESH seeds and shapes,
against oracle/nets.py — forward and every parameter gradient, float64.  Skipped on the GPU box."""
import os

import pytest
import torch

REF = "/root/reference/models/rank"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


@pytest.fixture(autouse=True)
def _f64():
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(prev)


def _grads(loss, named):
    gs = torch.autograd.grad(loss, list(named.values()), allow_unused=True, retain_graph=True)
    return {k: (torch.zeros_like(p) if g is None else g) for (k, p), g in zip(named.items(), gs)}


def _same(ref_out, ora_out, named):
    assert (ref_out - ora_out).abs().max() < 1e-12
    g1 = _grads(ref_out.square().sum(), named)
    g2 = _grads(ora_out.square().sum(), named)
    for k in g1:
        assert (g1[k] - g2[k]).abs().max() < 1e-10, k


@pytest.mark.parametrize("seed,D,B", [(1, 5, 3), (2, 12, 9)])
def test_deepfm_live(seed, D, B):
    from oracle import nets, paddle_shim
    ref = paddle_shim.import_reference_net("deepfm")
    torch.manual_seed(seed)
    V, fc = 61, [7, 5, 3]
    layer = ref.DeepFMLayer(V, D, 13, 26, fc)
    ids = [torch.randint(0, V, (B, 1)) for _ in range(26)]
    dense = torch.rand(B, 13)
    named = dict(layer.named_parameters())
    _same(layer(ids, dense), nets.deepfm_forward(named, ids, dense, len(fc)), named)


@pytest.mark.parametrize("mix,stacked", [(False, False), (True, True)])
def test_dcn_v2_live(mix, stacked):
    from oracle import nets, paddle_shim
    ref = paddle_shim.import_reference_net("dcn_v2")
    torch.manual_seed(11)
    V, D, B, fc = 43, 3, 5, [9, 6]
    layer = ref.DCN_V2Layer(V, D, 13, 26, fc, 3, stacked, mix, 4, 2)
    layer.eval()
    ids = [torch.randint(0, V, (B, 1)) for _ in range(26)]
    dense = torch.rand(B, 13)
    named = dict(layer.named_parameters())
    out = nets.dcn_v2_forward(named, ids, dense, n_fc=len(fc), cross_num=3, is_stacked=stacked,
                              use_low_rank_mixture=mix, num_experts=2)
    _same(layer(ids, dense), out, named)


def test_din_and_wide_deep_live():
    from oracle import nets, paddle_shim
    torch.manual_seed(5)
    ref = paddle_shim.import_reference_net("wide_deep")
    V, D, B, fc = 37, 6, 4, [8, 4]
    layer = ref.WideDeepLayer(V, D, 13, 26, fc)
    ids = [torch.randint(0, V, (B, 1)) for _ in range(26)]
    dense = torch.rand(B, 13)
    named = dict(layer.named_parameters())
    _same(layer(ids, dense), nets.wide_deep_forward(named, ids, dense, len(fc)), named)

    refd = paddle_shim.import_reference_net("din")
    layer = refd.DINLayer(4, 4, "sigmoid", False, False, 29, 7)
    B, L = 3, 5
    hi, hc = torch.randint(0, 29, (B, L)), torch.randint(0, 7, (B, L))
    ti, tc = torch.randint(0, 29, (B,)), torch.randint(0, 7, (B,))
    mask = torch.zeros(B, L, 1, dtype=torch.int64)
    mask[1, 3:] = int(-1e9)
    named = dict(layer.named_parameters())
    for i, m in enumerate([m for m in layer.attention_layer if hasattr(m, "weight")]):
        named["att.linear_%d.weight" % i], named["att.linear_%d.bias" % i] = m.weight, m.bias
    args = (hi, hc, ti, tc, None, mask, ti.unsqueeze(1).repeat(1, L), tc.unsqueeze(1).repeat(1, L))
    _same(layer(*args), nets.din_forward(named, *args), named)
    if os.path.exists("tmp.txt"):
        os.remove("tmp.txt")
