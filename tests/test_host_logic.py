"""Host-side logic that needs no GPU: yaml flattening / overrides (utils_single.py:57-86,
trainer.py:55-65), the Criteo and DIN readers, AUC, LR schedule, the oracle's optimizer rules."""
import os

import numpy as np
import pytest
import torch

from paddlerec_b200 import functional as BF
from paddlerec_b200 import optim, runner

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "paddlerec_b200")


def test_yaml_flatten_and_overrides():
    cfg = runner.load_yaml(os.path.join(PKG, "rank", "deepfm", "config.yaml"))
    assert cfg["runner.train_batch_size"] == 2 and cfg["runner.epochs"] == 3
    assert cfg["hyper_parameters.sparse_feature_number"] == 1000001
    assert cfg["hyper_parameters.sparse_feature_dim"] == 9
    assert cfg["hyper_parameters.optimizer.learning_rate"] == 0.001
    assert cfg["hyper_parameters.fc_sizes"] == [512, 256, 128, 32]
    runner.apply_overrides(cfg, ["runner.train_batch_size=8", "runner.use_gpu=False",
                                 "hyper_parameters.optimizer.learning_rate=0.01",
                                 "hyper_parameters.fc_sizes=[8,4]", "runner.model_save_path=xx"])
    assert cfg["runner.train_batch_size"] == 8 and cfg["runner.use_gpu"] is False
    assert cfg["hyper_parameters.optimizer.learning_rate"] == 0.01
    assert cfg["hyper_parameters.fc_sizes"] == [8, 4] and cfg["runner.model_save_path"] == "xx"


def test_named_list_flattening():
    flat = runner.flatten_yaml({"runner": [{"name": "a", "x": 1}], "hyper_parameters": {"y": {"z": 2}}})
    assert flat == {"runner.a.name": "a", "runner.a.x": 1, "hyper_parameters.y.z": 2}


def test_criteo_reader_contract():
    from paddlerec_b200.rank.deepfm import criteo_reader
    d = os.path.join(PKG, "rank", "deepfm", "data", "sample_data", "train")
    ds = criteo_reader.RecDataset([os.path.join(d, f) for f in os.listdir(d)], config={})
    samples = list(ds)
    assert len(samples) == 80
    s = samples[0]
    assert len(s) == 28 and all(a.dtype == np.int64 and a.shape == (1,) for a in s[:27])
    assert s[27].dtype == np.float32 and s[27].shape == (13,)
    assert any((np.concatenate(x[1:27]) == 0).any() for x in samples)   # padding id present
    # missing slot -> padding id 0
    out = ds.parse_line("click:1 dense_feature:0.5 3:77")
    assert out[0][0] == 1 and out[3][0] == 77 and out[1][0] == 0 and out[27].shape == (1,)
    loader_cfg = {"runner.train_data_dir": "data/sample_data/train", "runner.train_batch_size": 2,
                  "runner.train_reader_path": "criteo_reader",
                  "config_abs_dir": os.path.join(PKG, "rank", "deepfm")}
    batch = next(iter(runner.create_data_loader(loader_cfg)))
    assert len(batch) == 28 and batch[1].shape == (2, 1) and batch[27].shape == (2, 13)


def test_packed_reader_type_yields_the_same_batches_as_the_dataloader():
    cfg = {"runner.train_data_dir": "data/sample_data/train", "runner.train_batch_size": 16,
           "runner.train_reader_path": "criteo_reader",
           "config_abs_dir": os.path.join(PKG, "rank", "deepfm")}
    plain = list(runner.create_data_loader(cfg))
    packed = list(runner.create_data_loader({**cfg, "runner.reader_type": "PackedReader"}))
    assert len(plain) == len(packed) == 5
    for a, b in zip(plain, packed):
        label, ids, dense = b
        assert torch.equal(label, a[0]) and torch.equal(ids, torch.cat(a[1:27], 1))
        assert torch.equal(dense, a[27])
    with pytest.raises(ValueError, match="reader_type"):
        runner.create_data_loader({**cfg, "runner.reader_type": "QueueDataset"})


def test_din_reader_contract(tmp_path):
    from paddlerec_b200.rank.din import reader
    p = tmp_path / "d.txt"
    p.write_text("1 2 3;4 5 6;7;8;1\n9;10;11;12;0\n3 4;5 6;7;8;1\n1 1 1 1;2 2 2 2;5;6;0\n")
    ds = reader.RecDataset([str(p)], {"runner.train_batch_size": 2})
    out = list(ds)
    assert len(out) == 4
    lens = [int((s[5] == 0).sum()) for s in out]
    assert lens == sorted(lens)                                  # sorted by history length
    s = out[1]
    L = s[0].shape[0]
    assert s[5].shape == (L, 1) and s[5].dtype == np.int64
    assert (s[6] == s[2]).all() and s[6].shape == (L,)           # target id tiled L times
    first = out[0]
    assert first[0].tolist() == [9, 0] and first[5].reshape(-1).tolist() == [0, int(-1e9)]


def test_auc_matches_sklearn():
    from sklearn.metrics import roc_auc_score
    g = torch.Generator().manual_seed(0)
    y = (torch.rand(5000, generator=g) < 0.3).long()
    p = (torch.rand(5000, generator=g) * 0.6 + 0.3 * y).clamp(0, 1)
    m = BF.Auc()
    m.update(torch.stack([1 - p, p], 1)[:2500], y[:2500].reshape(-1, 1))
    m.update(p[2500:].reshape(-1, 1), y[2500:])
    assert abs(m.accumulate() - roc_auc_score(y.numpy(), p.numpy())) < 2e-3   # 4095 buckets
    assert BF.Auc().accumulate() == 0.0


def test_piecewise_decay():
    lr = optim.PiecewiseDecay([3], [0.85, 0.2])
    vals = []
    for _ in range(5):
        vals.append(lr())
        lr.step()
    assert vals == [0.85, 0.85, 0.85, 0.2, 0.2]


def test_oracle_optimizer_rules_against_torch():
    from oracle import optim as oo
    g = torch.Generator().manual_seed(1)
    W = torch.randn(20, 4, generator=g, dtype=torch.float64)
    grad = torch.randn(20, 4, generator=g, dtype=torch.float64)
    p = W.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=1e-2)
    Wr, m, v = W.numpy(), np.zeros((20, 4)), np.zeros((20, 4))
    for t in (1, 2, 3):
        p.grad = grad.clone()
        opt.step()
        Wr, m, v = oo.adam_lazy(Wr, m, v, np.arange(20), grad.numpy(), 1e-2, 0.9, 0.999, 1e-8, t)
    np.testing.assert_allclose(Wr, p.detach().numpy(), rtol=1e-10, atol=1e-12)
    ids = np.array([3, 5, 3, 0, 5, 5])
    u, merged = oo.merge_rows(ids, np.ones((6, 2)), padding_idx=0)
    assert u.tolist() == [3, 5] and merged[:, 0].tolist() == [2.0, 3.0]


def test_l2_decay_regularizer_is_added_after_clipping():
    """ParamAttr(regularizer=L2Decay(c)) (dcn_v2/net.py:166-168): grad <- clip(grad) + c*w."""
    from paddlerec_b200 import nn as bnn, optim

    torch.manual_seed(0)
    lin = bnn.Linear(4, 3, weight_std=0.5, weight_l2_decay=0.25)
    assert lin.weight.l2_decay == 0.25 and not hasattr(lin.bias, "l2_decay")
    w0, b0 = lin.weight.detach().clone(), lin.bias.detach().clone()
    gw, gb = torch.randn(4, 3) * 5, torch.randn(3) * 5
    lin.weight.grad, lin.bias.grad = gw.clone(), gb.clone()
    opt = optim.SGD(0.1, lin.parameters(), grad_clip=optim.ClipGradByGlobalNorm(1.0))
    opt.step()
    from oracle import optim as ooptim

    ow, ob = ooptim.clip_then_l2([gw.numpy(), gb.numpy()], [w0.numpy(), b0.numpy()], [0.25, 0.0], 1.0)
    assert np.allclose(lin.weight.detach().numpy(), w0.numpy() - 0.1 * ow, atol=1e-6)
    assert np.allclose(lin.bias.detach().numpy(), b0.numpy() - 0.1 * ob, atol=1e-6)
    # the DCN-V2 tower carries the reference's coefficient
    from paddlerec_b200.rank.dcn_v2 import net as dcn_net

    dnn = dcn_net.DNNLayer(4, 13, 26, [8, 8])
    assert [getattr(p, "l2_decay", 0.0) for n, p in dnn.named_parameters() if n.endswith("weight")] == [1e-7, 1e-7]


@pytest.mark.parametrize("model", ["deepfm", "dcn_v2", "wide_deep", "dlrm", "din"])
def test_every_model_directory_is_a_complete_plugin(model):
    """doc/model_develop.md:3-45: a model directory holds net.py, dygraph_model.py (class
    DygraphModel with the seven methods), a reader and a yaml whose data paths resolve."""
    d = os.path.join(PKG, "rank", model)
    cfg = runner.load_yaml(os.path.join(d, "config.yaml"))
    cfg["config_abs_dir"] = d
    dm = runner.load_dy_model_class(d)
    for meth in ("create_model", "create_feeds", "create_loss", "create_optimizer", "create_metrics",
                 "train_forward", "infer_forward"):
        assert callable(getattr(dm, meth)), meth
    for mode in ("train", "test"):
        batch = next(iter(runner.create_data_loader(cfg, mode)))
        bs = cfg["runner.train_batch_size" if mode == "train" else "runner.infer_batch_size"]
        assert len(batch[0]) == bs
    if model == "din":
        assert len(batch) == 8 and batch[5].dtype == torch.int64 and batch[5].shape[2] == 1
        assert batch[0].shape == batch[6].shape                    # target ids tiled to [B, L]
    else:
        assert len(batch) == 28 and batch[27].shape == (bs, 13)


def test_din_packed_reader_type():
    d = os.path.join(PKG, "rank", "din")
    cfg = runner.load_yaml(os.path.join(d, "config.yaml"))
    cfg["config_abs_dir"] = d
    plain = list(runner.create_data_loader(cfg))
    native = list(runner.create_data_loader({**cfg, "runner.reader_type": "PackedReader",
                                             "runner.packed_format": "din"}))
    assert len(plain) == len(native) == 3
    assert all(torch.equal(x, y) for a, b in zip(plain, native) for x, y in zip(a, b))


def test_dcn_v2_reader_log_transform_and_native_schema():
    """dcn_v2/reader.py:55-64: dense = log(v+1), `slot:` with an empty value is skipped — the Python
    mirror and the native parser (dataio.CRITEO_DCN_V2) agree bit for bit."""
    from paddlerec_b200 import dataio
    from paddlerec_b200.rank.dcn_v2 import reader

    lines = ["click:1 dense_feature:0.5 dense_feature:0 " + " ".join("dense_feature:%r" % (i * 0.37) for i in range(11))
             + " 1:7 2: 3:9", "click:0 4:5 26:11"]
    ds = reader.RecDataset([], config={})
    rows = [ds.parse_line(l) for l in lines]
    assert rows[0][2][0] == 0 and rows[0][3][0] == 9                  # `2:` skipped -> padding id
    assert rows[0][27][0] == np.float32(np.log(1.5)) and rows[0][27][1] == 0
    assert rows[1][27].tolist() == [0.0] * 13
    label, ids, dense = dataio.parse_slot_text("\n".join(lines), dataio.CRITEO_DCN_V2)
    assert np.array_equal(ids, np.stack([np.concatenate(r[1:27]) for r in rows]))
    assert np.array_equal(dense, np.stack([r[27] for r in rows]))
    assert label[:, 0].tolist() == [1, 0]
    with pytest.raises(dataio.B200RecIOError, match="bad integer"):
        dataio.parse_slot_text(lines[0], dataio.CRITEO)               # plain schema: `2:` is an error
    cfg = {"runner.train_data_dir": "../deepfm/data/sample_data/train", "runner.train_batch_size": 16,
           "runner.train_reader_path": "reader", "config_abs_dir": os.path.join(PKG, "rank", "dcn_v2")}
    plain = list(runner.create_data_loader(cfg))
    packed = list(runner.create_data_loader({**cfg, "runner.reader_type": "PackedReader",
                                             "runner.packed_schema": "criteo_dcn_v2"}))
    for a, b in zip(plain, packed):
        assert torch.equal(b[1], torch.cat(a[1:27], 1)) and torch.equal(b[2], a[27])


def test_optimizer_step_does_not_advance_the_lr_scheduler():
    """Paddle's optimizer.step() never steps an LRScheduler and tools/trainer.py:151-153 never calls
    scheduler.step(): the reference DIN run stays at values[0].  Per-step decay is opt-in."""
    lin = torch.nn.Linear(3, 2)
    sched = optim.PiecewiseDecay([1], [0.5, 0.05])
    opt = optim.SGD(sched, lin.parameters())
    for _ in range(3):
        opt.clear_grad()
        lin(torch.ones(1, 3)).sum().backward()
        opt.step()
    assert opt.get_lr() == 0.5 and sched.last_epoch == 0
    opt.lr_auto_step = True
    opt.clear_grad()
    lin(torch.ones(1, 3)).sum().backward()
    opt.step()
    assert opt.get_lr() == 0.05


def test_peer_memory_exchange_default_is_limited_to_validated_world_sizes(monkeypatch):
    """sharded.p2p_enabled: on for <= 4 ranks unless B200REC_P2P says otherwise; bench.py names the
    exchange it will actually use in config.parallelism."""
    import importlib

    from paddlerec_b200 import sharded
    bench = importlib.import_module("bench")
    monkeypatch.delenv("B200REC_P2P", raising=False)
    assert [sharded.p2p_enabled(w) for w in (2, 4, 8)] == [True, True, False]
    assert "peer memory" in bench.exchange_name(4) and bench.exchange_name(8) == "NCCL all-to-all"
    monkeypatch.setenv("B200REC_P2P", "0")
    assert not sharded.p2p_enabled(2) and bench.exchange_name(2) == "NCCL all-to-all"
    monkeypatch.setenv("B200REC_P2P", "1")
    assert sharded.p2p_enabled(8) and "peer memory" in bench.exchange_name(8)
