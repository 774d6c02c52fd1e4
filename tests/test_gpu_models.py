"""Model-level parity on the GPU: the net.py-shaped layers of paddlerec_b200 (CUDA kernels through
the C ABI) against the committed golden vectors — forward logits/probabilities, the loss, and the
gradient of every parameter (north-star bar: 1e-4 relative, fp32)."""
import numpy as np
import pytest
import torch

from tests.util import load_golden, rel_err

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
        if getattr(p, "is_sparse_table", False):
            sr = p.grad_rows
            out[k] = sr.to_dense().cpu().numpy() if sr is not None else np.zeros(tuple(p.shape))
        else:
            out[k] = (p.grad if p.grad is not None else torch.zeros_like(p)).cpu().numpy()
    for k, p in (extra or {}).items():
        out[k] = (p.grad if p.grad is not None else torch.zeros_like(p)).cpu().numpy()
    return out


def check(g, pred, loss, grads, tol=TOL):
    assert rel_err(pred.detach().cpu().numpy(), g["out"]["pred"]) < tol
    assert abs(float(loss) - float(g["out"]["loss"])) < tol * max(1.0, abs(float(g["out"]["loss"])))
    for k, ref in g["grad"].items():
        assert k in grads, k
        if np.abs(ref).max() == 0:
            assert np.abs(grads[k]).max() == 0, k
        else:
            assert rel_err(grads[k], ref) < tol, (k, rel_err(grads[k], ref))


@pytest.mark.parametrize("name", ["deepfm_d9", "deepfm_d16"])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("as_list", [True, False])
def test_deepfm_golden(name, precision, as_list):
    from paddlerec_b200 import functional as BF
    from paddlerec_b200 import nn as bnn
    from paddlerec_b200.rank.deepfm import net
    g = load_golden(name)
    V, D = g["param"]["fm.embedding.weight"].shape
    fc = [g["param"]["dnn.linear_%d.weight" % i].shape[1] for i in range(2)]
    bnn.set_matmul_precision(precision)
    try:
        layer = load_state(net.DeepFMLayer(V, D, 13, 26, fc), g)
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
