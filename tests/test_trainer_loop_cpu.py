"""HOST logic of the trainer / infer loops on a box without a GPU: the tools/trainer.py- and
tools/infer.py-shaped loops, checkpoints and readers run end to end with the torch stand-ins of
tests/cpu_kernels.py patched over the CUDA kernels (test doubles; the kernels themselves are
covered by the -m gpu parity tests).  What this pins: loop structure, metric plumbing, checkpoint
layout and resume, reader selection — for every model directory whose kernels have a stand-in."""
import os

import numpy as np
import pytest
import torch

from tests import cpu_kernels

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "paddlerec_b200")


@pytest.fixture
def cpu_engine(monkeypatch):
    from paddlerec_b200 import ops, runner
    from paddlerec_b200.rank.deepfm import dygraph_model as criteo_base

    cpu_kernels.install(monkeypatch, ops)
    monkeypatch.setattr(runner, "_require_cuda", lambda: None)
    monkeypatch.setattr(criteo_base.DygraphModel, "device", "cpu")
    return runner


def _config(runner, model, tmp_path, **over):
    d = os.path.join(PKG, "rank", model)
    cfg = runner.load_yaml(os.path.join(d, "config.yaml"))
    cfg["config_abs_dir"] = d
    cfg["runner.model_save_path"] = str(tmp_path / "out")
    cfg["runner.infer_load_path"] = str(tmp_path / "out")
    cfg.update(over)
    return cfg


@pytest.mark.parametrize("model,extra", [
    ("deepfm", {}),
    ("wide_deep", {"runner.epochs": 2, "runner.infer_start_epoch": 0, "runner.infer_end_epoch": 2}),
    ("dlrm", {"runner.epochs": 2, "runner.infer_end_epoch": 2,
              "hyper_parameters.optimizer.learning_rate": 0.01}),
])
def test_train_then_infer_every_saved_epoch(cpu_engine, tmp_path, model, extra):
    runner = cpu_engine
    cfg = _config(runner, model, tmp_path, **{"runner.train_batch_size": 16, "runner.infer_batch_size": 16,
                                              "hyper_parameters.sparse_feature_number": 20011,
                                              "runner.reader_type": "PackedReader", **extra})
    losses, metric_values, net = runner.train(cfg)
    epochs = cfg["runner.epochs"]
    assert len(losses) == 5 * epochs and np.all(np.isfinite(losses))
    assert "auc" in metric_values
    for e in range(epochs):
        assert sorted(os.listdir(tmp_path / "out" / str(e))) == ["rec.pdopt", "rec.pdparams"]
    cfg["runner.infer_start_epoch"], cfg["runner.infer_end_epoch"] = 0, epochs
    results = runner.infer(cfg)
    assert sorted(results) == list(range(epochs))
    assert all(0.0 <= r["auc"] <= 1.0 for r in results.values())
    if epochs > 1:      # memorising 80 samples: the later checkpoint ranks the train set better
        assert results[epochs - 1]["auc"] > results[0]["auc"]
    # the last checkpoint is the live model
    from paddlerec_b200 import checkpoint
    saved = checkpoint.load_pdparams(str(tmp_path / "out" / str(epochs - 1) / "rec.pdparams"))
    for k, v in net.state_dict().items():
        assert np.array_equal(saved[k], v.numpy()), k


def test_resume_from_checkpoint_continues_the_same_trajectory(cpu_engine, tmp_path):
    """runner.model_init_path (trainer.py:106-107) restores parameters AND rec.pdopt, so
    2 epochs == 1 epoch + resume for 1 more (same batches, lazy Adam moments in the slots)."""
    runner = cpu_engine
    base = {"runner.train_batch_size": 16, "hyper_parameters.sparse_feature_number": 20011,
            "runner.print_interval": 100}
    full = _config(runner, "deepfm", tmp_path / "a", **{**base, "runner.epochs": 2})
    losses_full, _, net_full = runner.train(full)
    first = _config(runner, "deepfm", tmp_path / "b", **{**base, "runner.epochs": 1})
    runner.train(first)
    second = _config(runner, "deepfm", tmp_path / "c", **{
        **base, "runner.epochs": 2, "last_epoch": 0,
        "runner.model_init_path": str(tmp_path / "b" / "out" / "0")})
    losses_second, _, net_resumed = runner.train(second)
    assert np.allclose(losses_second, losses_full[5:], rtol=1e-6, atol=0)
    for (k, a), (_, b) in zip(net_full.state_dict().items(), net_resumed.state_dict().items()):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-8), k


def test_infer_without_checkpoint_or_data_fails_loudly(cpu_engine, tmp_path):
    runner = cpu_engine
    cfg = _config(runner, "deepfm", tmp_path, **{"hyper_parameters.sparse_feature_number": 1001})
    with pytest.raises(FileNotFoundError):
        runner.infer(cfg)
    runner.train({**cfg, "runner.epochs": 1, "runner.train_batch_size": 40})
    with pytest.raises(RuntimeError, match="test_dataloader is null"):
        runner.infer({**cfg, "runner.infer_batch_size": 500, "runner.infer_end_epoch": 1})


def test_the_product_refuses_to_run_without_cuda(tmp_path):
    from paddlerec_b200 import runner
    if torch.cuda.is_available():
        pytest.skip("CUDA box")
    d = os.path.join(PKG, "rank", "deepfm")
    cfg = runner.load_yaml(os.path.join(d, "config.yaml"))
    cfg["config_abs_dir"] = d
    with pytest.raises(RuntimeError, match="CUDA device only"):
        runner.train(cfg)
    with pytest.raises(RuntimeError, match="CUDA device only"):
        runner.infer(cfg)


@pytest.mark.parametrize("use_cvm", [False, True])
def test_fused_seqpool_cvm_from_parsed_lod_text(cpu_engine, use_cvm):
    """Multi-hot path end to end on the host side: variable-length `slot:value` text -> native LoD
    parser -> ONE pooled lookup + ONE CVM for all slots (ops.fused_seqpool_cvm), forward and the
    table gradient (show/click written into the two leading columns) against the oracle."""
    from oracle import nets
    from paddlerec_b200 import dataio, nn as bnn

    rng = np.random.default_rng(11)
    slots = ["s%d" % i for i in range(4)]
    V, D, B = 50, 6, 9
    lines = []
    for _ in range(B):
        toks = ["y:%d" % rng.integers(0, 2)]
        for s in slots:
            toks += ["%s:%d" % (s, rng.integers(0, V)) for _ in range(rng.integers(0, 5))]
        lines.append(" ".join(toks))
    sch = dataio.SlotSchema(sparse_slots=tuple(slots), label_slot="y", dense_slot=None)
    label, keys, offsets, _ = dataio.parse_slot_text_lod("\n".join(lines), sch)
    emb = bnn.Embedding(V, D + 2, padding_idx=0, init_std=0.5, device="cpu")
    with torch.no_grad():
        emb.weight[:, :2] = torch.rand(V, 2) * 5          # accumulated show / click statistics
        emb.weight[0] = 0
    show_click = torch.rand(B, 2) * 3
    out = emb.forward_seqpool_cvm(torch.from_numpy(keys), torch.from_numpy(offsets), len(slots),
                                  show_click, use_cvm)
    assert out.shape == (B, len(slots), D + 2 if use_cvm else D)
    # oracle: per bag sum of rows (padding key 0 -> zeros), then the CVM column transform
    W = emb.weight.detach().double().requires_grad_(True)
    ref = []
    for b in range(B * len(slots)):
        ks = torch.from_numpy(keys[offsets[b]:offsets[b + 1]])
        ref.append((W[ks] * (ks != 0).unsqueeze(1)).sum(0))
    ref = nets.cvm(torch.stack(ref), use_cvm).reshape(out.shape)
    assert torch.allclose(out.double(), ref, rtol=1e-6, atol=1e-6)
    gout = torch.randn(out.shape)
    (out * gout).sum().backward()
    (ref * gout.double()).sum().backward()
    dW = emb.grad_rows.to_dense().double()
    touched = np.unique(keys[keys != 0])
    # embedding columns: the oracle's autograd; show/click columns: the batch's show_click, summed
    # over the bags a key occurs in (Paddle's cvm_grad semantics), not d(log)/d(show)
    assert torch.allclose(dW[touched][:, 2:], W.grad[touched][:, 2:], rtol=1e-5, atol=1e-6)
    want_sc = torch.zeros(V, 2, dtype=torch.float64)
    for b in range(B * len(slots)):
        for k in keys[offsets[b]:offsets[b + 1]]:
            if k != 0:
                want_sc[k] += show_click[b // len(slots)].double()
    assert torch.allclose(dW[:, :2], want_sc, rtol=1e-6, atol=1e-6)
    assert not dW[0].any()
    with pytest.raises(ValueError, match="multiple of n_slots"):
        emb.forward_seqpool_cvm(torch.from_numpy(keys), torch.from_numpy(offsets), 5, show_click, use_cvm)
