"""Checkpoint layout and format (paddlerec_b200/checkpoint.py) — CPU only."""
import os
import pickle

import numpy as np
import pytest
import torch

from paddlerec_b200 import checkpoint, optim
from paddlerec_b200 import nn as bnn


def _dlrm():
    from paddlerec_b200.rank.dlrm import net
    torch.manual_seed(3)
    return net.DLRMLayer(13, [8, 4], 50, 4, [8, 2], 26, device="cpu")


def test_pdparams_is_a_plain_pickle_of_numpy_arrays_with_the_name_table(tmp_path):
    """What paddle.save(layer.state_dict()) writes: loadable with nothing but pickle + numpy."""
    m = _dlrm()
    path = str(tmp_path / "rec.pdparams")
    checkpoint.save_pdparams(m.state_dict(), path)
    with open(path, "rb") as fh:
        raw = pickle.load(fh)
    names = set(m.state_dict())
    assert set(raw) == names | {checkpoint.NAME_TABLE_KEY}
    assert raw[checkpoint.NAME_TABLE_KEY] == {k: k for k in names}
    assert all(isinstance(raw[k], np.ndarray) for k in names)
    # the reference's names and layouts: Linear weight [in, out]; BatchNorm running statistics
    assert raw["bot_mlp.dense_0.weight"].shape == (13, 8) and "top_mlp.norm_1._variance" in raw
    assert raw["embedding.weight"].dtype == np.float32

    m2 = _dlrm()
    with torch.no_grad():
        for p in m2.parameters():
            p.add_(1.0)
    checkpoint.set_state_dict(m2, checkpoint.load_pdparams(path))
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


def test_pdparams_reads_paddle_side_variants_and_legacy_files(tmp_path):
    m = _dlrm()
    want = checkpoint.to_numpy_state(m.state_dict())
    # (name, ndarray) pairs — what paddle's tensor reducer leaves when a Tensor is pickled directly
    p1 = str(tmp_path / "pairs.pdparams")
    with open(p1, "wb") as fh:
        pickle.dump({k: ("param_%d" % i, v) for i, (k, v) in enumerate(want.items())}, fh, protocol=2)
    got = checkpoint.load_pdparams(p1)
    assert set(got) == set(want) and all(np.array_equal(got[k], want[k]) for k in want)
    # torch-zip files written by earlier revisions of this repo
    p2 = str(tmp_path / "legacy.pdparams")
    torch.save(m.state_dict(), p2)
    got = checkpoint.load_pdparams(p2)
    assert all(np.array_equal(got[k], want[k]) for k in want)
    with open(str(tmp_path / "junk.pdparams"), "wb") as fh:
        pickle.dump([1, 2, 3], fh)
    with pytest.raises(ValueError, match="state_dict"):
        checkpoint.load_pdparams(str(tmp_path / "junk.pdparams"))
    # a missing / extra key is reported like Layer.set_state_dict(strict) would
    bad = dict(want)
    bad.pop("embedding.weight")
    with pytest.raises(RuntimeError, match="embedding.weight"):
        checkpoint.set_state_dict(_dlrm(), bad)


def test_fused_table_exports_the_reference_tables(tmp_path):
    """nn.FusedTable keeps [emb | w1 | Adam moments] in one slot; the checkpoint must still hold
    the reference's `embedding.weight [V,D]` and `embedding_one.weight [V,1]`."""
    from paddlerec_b200.rank.deepfm import net
    torch.manual_seed(0)
    layer = net.DeepFMLayer(101, 9, 13, 26, [16, 8], device="cpu")
    path = str(tmp_path / "rec.pdparams")
    checkpoint.save_pdparams(layer.state_dict(), path)
    got = checkpoint.load_pdparams(path)
    assert got["fm.embedding.weight"].shape == (101, 9)
    assert got["fm.embedding_one.weight"].shape == (101, 1)
    assert not any("slots" in k or "moment" in k for k in got)
    layer2 = net.DeepFMLayer(101, 9, 13, 26, [16, 8], device="cpu")
    checkpoint.set_state_dict(layer2, got)
    for (k, a), (_, b) in zip(layer.state_dict().items(), layer2.state_dict().items()):
        assert torch.equal(a, b), k


def test_optimizer_state_round_trip_resumes_the_same_trajectory(tmp_path):
    """Dense part on CPU (the sparse moments need the CUDA row kernels): two steps, save, one more
    step == load into a fresh optimizer, one more step."""
    def make():
        torch.manual_seed(5)
        lin = bnn.Linear(6, 3, weight_std=0.3, weight_l2_decay=0.01)
        sched = optim.PiecewiseDecay([2], [0.1, 0.01])
        opt = optim.Adam(sched, lin.parameters())
        opt.lr_auto_step = True     # per-step decay is opt-in (the reference never steps it)
        return lin, opt

    def step(lin, opt, seed):
        g = torch.Generator().manual_seed(seed)
        opt.clear_grad()
        x = torch.randn(5, 6, generator=g)
        (lin.weight * x.t().mean(1, keepdim=True)).sum().add(lin.bias.sum()).backward()
        opt.step()

    lin, opt = make()
    step(lin, opt, 1)
    step(lin, opt, 2)
    checkpoint.save_pdparams(lin.state_dict(), str(tmp_path / "rec.pdparams"))
    checkpoint.save_pdopt(opt, lin, str(tmp_path / "rec.pdopt"))
    step(lin, opt, 3)

    lin2, opt2 = make()
    checkpoint.set_state_dict(lin2, checkpoint.load_pdparams(str(tmp_path / "rec.pdparams")))
    checkpoint.load_pdopt(opt2, lin2, str(tmp_path / "rec.pdopt"))
    assert opt2.step_count == 2 and opt2.get_lr() == 0.01     # scheduler epoch restored
    step(lin2, opt2, 3)
    assert torch.equal(lin.weight, lin2.weight) and torch.equal(lin.bias, lin2.bias)
    with pytest.raises(ValueError, match="optimizer is SGD"):
        checkpoint.load_pdopt(optim.SGD(0.1, lin2.parameters()), lin2, str(tmp_path / "rec.pdopt"))


def test_runner_save_and_load_use_the_reference_layout(tmp_path):
    from paddlerec_b200 import runner
    m = _dlrm()
    opt = optim.SGD(0.1, [p for p in m.parameters() if not getattr(p, "is_sparse_table", False)])
    runner.save_model(m, opt, str(tmp_path), 7)
    assert sorted(os.listdir(tmp_path / "7")) == ["rec.pdopt", "rec.pdparams"]
    m2 = _dlrm()
    with torch.no_grad():
        m2.embedding.weight.zero_()
    runner.load_model(str(tmp_path / "7"), m2, optimizer=opt)
    assert torch.equal(m.embedding.weight, m2.embedding.weight)


def test_reference_din_checkpoint_without_attention_keys_loads():
    """a Paddle-produced DIN rec.pdparams has no attention.linear_* keys (the reference's
    name collision hides them from state_dict()); strict loading must accept that and keep the
    seeded attention weights, while a genuinely missing key still fails."""
    from paddlerec_b200.rank.din import net
    torch.manual_seed(3)
    m = net.DINLayer(8, 8, "sigmoid", True, True, 50, 7, device="cpu")
    full = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    ref_like = {k: v for k, v in full.items() if not k.startswith("attention.")}
    assert len(ref_like) < len(full)
    torch.manual_seed(4)
    m2 = net.DINLayer(8, 8, "sigmoid", True, True, 50, 7, device="cpu")
    att_before = m2.attention.linear_0.weight.detach().clone()
    checkpoint.set_state_dict(m2, ref_like)
    assert torch.equal(m2.attention.linear_0.weight, att_before)
    assert torch.equal(m2.linearCon.weight, m.linearCon.weight)
    broken = dict(ref_like)
    broken.pop("linearCon.weight")
    with pytest.raises(RuntimeError):
        checkpoint.set_state_dict(m2, broken)
