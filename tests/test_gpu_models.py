"""Model-level parity on the GPU: the net.py-shaped layers of paddlerec_b200 (CUDA kernels through
the C ABI) against the committed golden vectors — forward logits/probabilities, the loss, and the
gradient of every parameter (north-star bar: 1e-4 relative, fp32)."""
import numpy as np
import pytest
import torch

from tests.util import elementwise_excess, load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def load_state(layer, g, extra=None):
    sd = layer.state_dict()
    with torch.no_grad():
        for k, v in g["param"].items():
            if k in sd:
                sd[k].copy_(torch.tensor(v, dtype=torch.float32))
            elif extra and k in extra:
                extra[k].copy_(torch.tensor(v, dtype=torch.float32))
            else:
                raise KeyError(k)
    return layer


def grads_of(layer, extra=None):
    out = {}
    for k, p in layer.named_parameters():
        if hasattr(p, "fused_D"):    # fused [emb | w1 | pad] slots -> the reference's two keys
            D, names = p.fused_D, p.fused_names
            parent = k.rsplit(".", 2)[0] + "." if k.count(".") >= 2 else ""
            dense = (p.grad_rows.to_dense().cpu().numpy() if p.grad_rows is not None
                     else np.zeros((p.shape[0], D + 1)))
            out[parent + names[0]] = dense[:, :D]
            out[parent + names[1]] = dense[:, D:D + 1]
        elif getattr(p, "is_sparse_table", False):
            sr = p.grad_rows
            out[k] = sr.to_dense().cpu().numpy() if sr is not None else np.zeros(tuple(p.shape))
        else:
            out[k] = (p.grad if p.grad is not None else torch.zeros_like(p)).cpu().numpy()
    for k, p in (extra or {}).items():
        out[k] = (p.grad if p.grad is not None else torch.zeros_like(p)).cpu().numpy()
    return out


def check(g, pred, loss, grads, tol=TOL):
    assert rel_err(pred.detach().cpu().numpy(), g["out"]["pred"]) < tol
    # the north-star bar read element-wise: every logit/probability within tol of its own value
    assert elementwise_excess(pred.detach().cpu().numpy(), g["out"]["pred"], rtol=tol) <= 0
    assert abs(float(loss) - float(g["out"]["loss"])) < tol * max(1.0, abs(float(g["out"]["loss"])))
    for k, ref in g["grad"].items():
        assert k in grads, k
        if np.abs(ref).max() < 1e-10:   # mathematically zero (e.g. softmax shift invariance)
            assert np.abs(grads[k]).max() < 1e-7, k
        else:
            assert rel_err(grads[k], ref) < tol, (k, rel_err(grads[k], ref))


@pytest.mark.parametrize("name", ["deepfm_d9", "deepfm_d16"])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("as_list", [True, False])
@pytest.mark.parametrize("fused", [True, False])
def test_deepfm_golden(name, precision, as_list, fused):
    from paddlerec_b200 import functional as BF
    from paddlerec_b200 import nn as bnn
    from paddlerec_b200.rank.deepfm import net
    g = load_golden(name)
    V, D = g["param"]["fm.embedding.weight"].shape
    fc = [g["param"]["dnn.linear_%d.weight" % i].shape[1] for i in range(2)]
    bnn.set_matmul_precision(precision)
    try:
        layer = load_state(net.DeepFMLayer(V, D, 13, 26, fc, fused_table=fused), g)
        ids = torch.tensor(g["in"]["ids"], device="cuda")
        dense = torch.tensor(g["in"]["dense"], dtype=torch.float32, device="cuda")
        label = torch.tensor(g["in"]["label"], dtype=torch.float32, device="cuda")
        sparse = [ids[:, i:i + 1] for i in range(26)] if as_list else ids
        pred = layer(sparse, dense)
        loss = BF.log_loss(pred, label).mean()
        loss.backward()
        check(g, pred, loss, grads_of(layer))
    finally:
        bnn.set_matmul_precision("fp32")


def _criteo_inputs(g):
    ids = torch.tensor(g["in"]["ids"], device="cuda")
    dense = torch.tensor(g["in"]["dense"], dtype=torch.float32, device="cuda")
    label = torch.tensor(g["in"]["label"], dtype=torch.float32, device="cuda")
    return ids, dense, label


@pytest.mark.parametrize("name,mix,stacked", [("dcn_v2_v2_stacked", False, True),
                                              ("dcn_v2_mix_parallel", True, False)])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_dcn_v2_golden(name, mix, stacked, precision):
    from paddlerec_b200 import functional as BF
    from paddlerec_b200 import nn as bnn
    from paddlerec_b200.rank.dcn_v2 import net
    g = load_golden(name)
    V, D = g["param"]["embedding.weight"].shape
    fc = [g["param"]["DNN_.linear_%d.weight" % i].shape[1] for i in range(2)]
    bnn.set_matmul_precision(precision)
    try:
        layer = net.DCN_V2Layer(V, D, 13, 26, fc, 2, stacked, mix, 6, 4)
        load_state(layer, g)
        layer.eval()      # parity is defined without dropout
        ids, dense, label = _criteo_inputs(g)
        pred = layer([ids[:, i:i + 1] for i in range(26)], dense)
        loss = BF.log_loss(pred, label).mean()
        loss.backward()
        check(g, pred, loss, grads_of(layer))
    finally:
        bnn.set_matmul_precision("fp32")


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_wide_deep_golden(precision):
    from paddlerec_b200 import functional as BF
    from paddlerec_b200 import nn as bnn
    from paddlerec_b200.rank.wide_deep import net
    g = load_golden("wide_deep")
    V, D = g["param"]["embedding.weight"].shape
    fc = [g["param"]["linear_%d.weight" % i].shape[1] for i in range(2)]
    bnn.set_matmul_precision(precision)
    try:
        layer = load_state(net.WideDeepLayer(V, D, 13, 26, fc), g)
        ids, dense, label = _criteo_inputs(g)
        pred = layer([ids[:, i:i + 1] for i in range(26)], dense)
        loss = BF.log_loss(pred, label).mean()
        loss.backward()
        check(g, pred, loss, grads_of(layer))
        # id 0 is a real row here (no padding_idx): it must receive gradient
        assert np.abs(grads_of(layer)["embedding.weight"][0]).max() > 0
    finally:
        bnn.set_matmul_precision("fp32")


@pytest.mark.parametrize("name,self_interaction", [("dlrm_pairs", False), ("dlrm_self", True)])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_dlrm_golden(name, self_interaction, precision):
    """Train-mode forward (BatchNorm on batch statistics), softmax cross-entropy, every gradient and
    the running statistics after the step, against the reference's own dlrm/net.py (golden).

    Tolerance: 1e-4 in fp32 (the package default).  In the opt-in 'bf16x3' GEMM mode the six
    BatchNorm layers divide by a standard deviation estimated from this golden's 11 samples, which
    amplifies the ~2^-16 relative error of the dropped lo*lo products: a CPU emulation of the split
    GEMMs gives 5e-5 on the scores and up to 4e-4 on the gradients, so that mode is held to 2e-3."""
    from paddlerec_b200 import functional as BF
    from paddlerec_b200 import nn as bnn
    from paddlerec_b200.rank.dlrm import net
    g = load_golden(name)
    V, D = g["param"]["embedding.weight"].shape
    bot = [g["param"]["bot_mlp.dense_%d.weight" % i].shape[1] for i in range(3)]
    top = [g["param"]["top_mlp.dense_%d.weight" % i].shape[1] for i in range(3)]
    stats = {k: v for k, v in g["param"].items() if k.endswith("._mean") or k.endswith("._variance")}
    g2 = dict(g)
    g2["param"] = {k: v for k, v in g["param"].items() if k not in stats}
    bnn.set_matmul_precision(precision)
    try:
        layer = load_state(net.DLRMLayer(13, bot, V, D, top, 26, self_interaction=self_interaction), g2)
        layer.train()
        ids, dense, label = _criteo_inputs(g)
        pred = layer([ids[:, i:i + 1] for i in range(26)], dense)
        assert pred.shape == (ids.shape[0], 2)
        loss = BF.softmax_cross_entropy(pred, label.to(torch.int64)).mean()
        loss.backward()
        tol = TOL if precision == "fp32" else 2e-3
        check(g2, pred, loss, grads_of(layer), tol=tol)
        sd = layer.state_dict()
        for k, v in stats.items():          # Paddle's rule: biased batch variance, momentum 0.9
            assert rel_err(sd[k], v) < tol, k
        assert np.abs(grads_of(layer)["embedding.weight"][0]).max() > 0   # id 0 is a real row
        # eval mode runs on the running statistics and is deterministic
        layer.eval()
        with torch.no_grad():
            e1 = layer(ids, dense)
            e2 = layer(ids, dense)
        assert torch.equal(e1, e2) and not torch.allclose(e1, pred.detach())
    finally:
        bnn.set_matmul_precision("fp32")


@pytest.mark.parametrize("tiled", [True, False])
def test_din_golden(tiled):
    import torch.nn.functional as F

    from paddlerec_b200.rank.din import net
    g = load_golden("din")
    item_count, e2 = g["param"]["hist_item_emb_attr.weight"].shape
    cat_count = g["param"]["hist_cat_emb_attr.weight"].shape[0]
    layer = net.DINLayer(e2, e2, "sigmoid", False, False, item_count, cat_count,
                         faithful_frozen_attention=False, tiled_target_seq=tiled)
    remap = {k: k.replace("att.", "attention.") for k in g["param"]}
    g2 = dict(g)
    g2["param"] = {remap[k]: v for k, v in g["param"].items()}
    g2["grad"] = {remap[k]: v for k, v in g["grad"].items()}
    load_state(layer, g2)
    i = g["in"]
    dev = "cuda"
    L = i["hist_item"].shape[1]
    ti = torch.tensor(i["target_item"], device=dev)
    tc = torch.tensor(i["target_cat"], device=dev)
    label = torch.tensor(i["label"], dtype=torch.float32, device=dev)
    logit = layer(torch.tensor(i["hist_item"], device=dev), torch.tensor(i["hist_cat"], device=dev),
                  ti, tc, label, torch.tensor(i["mask"], device=dev),
                  ti.unsqueeze(1).repeat(1, L), tc.unsqueeze(1).repeat(1, L))
    loss = F.binary_cross_entropy_with_logits(logit, label)
    loss.backward()
    check(g2, logit, loss, grads_of(layer))


def test_din_faithful_frozen_attention_default():
    """by default the attention-unit linears are NOT trainable (the reference loses
    them from parameters() through a sub-layer name collision)."""
    from paddlerec_b200.rank.din import net
    layer = net.DINLayer(8, 8, "sigmoid", False, False, 50, 11)
    assert all(not p.requires_grad for p in layer.attention.parameters())
    assert layer.linear_0.weight.requires_grad


def test_wide_deep_gpubox_branch_vs_oracle():
    """The reference's PSGPU branch (wide_deep/net.py:80-88): [show, click, emb] rows, CVM with
    use_cvm=False, show/click pushed through the gradient and ACCUMULATED by the optimizer; uint64
    feasigns hashed to rows on the device (bit-exact vs oracle/readers.hash_keys)."""
    from oracle import nets, readers
    from paddlerec_b200 import functional as BF
    from paddlerec_b200 import optim
    from paddlerec_b200.rank.wide_deep import net
    from tests.util import slots
    g = torch.Generator().manual_seed(31)
    V, D, B, F = 211, 8, 37, 26
    layer = net.WideDeepLayer(V, D, 13, F, [32, 16], sync_mode="gpubox", device="cuda")
    assert layer.embedding.weight.shape == (V, D + 2)
    with torch.no_grad():    # give the statistic columns some history
        layer.embedding.weight[:, :2] = torch.randint(0, 5, (V, 2), generator=g).float().cuda()
    keys = torch.randint(0, 2 ** 62, (B, F), generator=g)
    keys[3, 5] = keys[7, 5]                                   # the same feasign twice in a slot
    dense = torch.rand(B, 13, generator=g)
    label = (torch.rand(B, 1, generator=g) < 0.4).float()
    show_click = torch.stack([torch.ones(B), label[:, 0]], 1)  # static_model.py:88-94
    slot = np.tile(np.arange(F, dtype=np.int32), B)
    rows_ref = readers.hash_keys(keys.numpy().reshape(-1), V, slot).astype(np.int64).reshape(B, F)
    p = {k: v.detach().cpu().double().requires_grad_(True) for k, v in layer.state_dict().items()}
    pred_ref = nets.wide_deep_forward_gpubox(p, slots(torch.from_numpy(rows_ref)), dense.double(), 2)
    nets.log_loss(pred_ref, label.double()).mean().backward()

    W0 = layer.embedding.weight.detach().clone()
    opt = optim.SGD(0.1, layer.parameters())
    pred = layer(keys.cuda(), dense.cuda(), show_click.cuda(), feasigns=True)
    BF.log_loss(pred, label.cuda()).mean().backward()
    assert rel_err(pred, pred_ref.detach()) < TOL
    dW = layer.embedding.grad_rows.to_dense().cpu().double()
    assert rel_err(dW[:, 2:], p["embedding.weight"].grad[:, 2:]) < TOL
    # cvm_grad: the two leading gradient columns of a row are the sums of (show, click) over its
    # positions in the batch
    want_stats = torch.zeros(V, 2, dtype=torch.float64)
    want_stats.index_add_(0, torch.from_numpy(rows_ref.reshape(-1)),
                          show_click.double().repeat_interleave(F, dim=0))
    assert torch.equal(dW[:, :2], want_stats)
    for k, v in layer.named_parameters():
        if k != "embedding.weight" and v.grad is not None:
            assert rel_err(v.grad, p[k].grad) < TOL, k
    opt.step()
    W1 = layer.embedding.weight.detach().cpu().double()
    assert torch.equal(W1[:, :2], W0[:, :2].cpu().double() + want_stats)      # accumulated, exact
    assert rel_err(W1[:, 2:], W0[:, 2:].cpu().double() - 0.1 * p["embedding.weight"].grad[:, 2:]) < 1e-6


def test_fused_seqpool_cvm_gpu_vs_oracle():
    """ops.fused_seqpool_cvm itself (tools/utils/static_ps/model_util.py:411-415): multi-hot bags,
    sum-pool + CVM in both modes, forward and the table gradient incl. the show/click columns."""
    from oracle import nets
    from paddlerec_b200 import nn as bnn
    g = torch.Generator().manual_seed(8)
    V, D, B, F = 101, 6, 19, 4
    emb = bnn.Embedding(V, D + 2, padding_idx=0, init="uniform", device="cuda")
    with torch.no_grad():
        emb.weight[:, :2].abs_()
    lens = torch.randint(0, 4, (B * F,), generator=g)
    lens[5] = 0
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    keys = torch.randint(0, V, (int(offsets[-1]),), generator=g)
    show_click = torch.stack([torch.ones(B), (torch.rand(B, generator=g) < 0.3).float()], 1)
    W = emb.weight.detach().cpu().double()
    for use_cvm in (False, True):
        emb.clear_grad()
        Wd = W.clone().requires_grad_(True)
        rows = nets.embedding(Wd, keys, 0)
        pooled = torch.zeros(B * F, D + 2, dtype=torch.float64)
        bag = torch.repeat_interleave(torch.arange(B * F), lens)
        pooled = pooled.index_add(0, bag, rows)
        want = nets.cvm(pooled, use_cvm).reshape(B, F, -1)
        gy = torch.randn(want.shape, generator=g, dtype=torch.float64)
        # cvm_grad is not the derivative: only the embedding columns follow autograd
        (want[..., (2 if use_cvm else 0):] * gy[..., (2 if use_cvm else 0):]).sum().backward()
        got = emb.forward_seqpool_cvm(keys.cuda(), offsets.cuda(), F, show_click.cuda(), use_cvm)
        assert rel_err(got, want.detach()) < 1e-5
        (got * gy.float().cuda()).sum().backward()
        dW = emb.grad_rows.to_dense().cpu().double()
        assert rel_err(dW[:, 2:], Wd.grad[:, 2:]) < 1e-5
        live = keys != 0
        want_stats = torch.zeros(V, 2, dtype=torch.float64)
        want_stats.index_add_(0, keys[live], show_click.double().repeat_interleave(F, dim=0)[bag][live])
        assert rel_err(dW[:, :2], want_stats) < 1e-6 and not dW[0].any()
