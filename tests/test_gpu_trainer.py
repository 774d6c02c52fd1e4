"""BASELINE config 1: the reference's own plumbing case (DeepFM, config.yaml, bs=2, sample data)
driven through the tools/trainer.py-shaped loop on the GPU, checked step by step against the CPU
oracle running the same loop (same initial parameters, same batches, lazy Adam)."""
import os

import numpy as np
import pytest
import torch

from oracle import nets
from oracle import optim as oo
from tests.util import rel_err

pytestmark = pytest.mark.gpu
PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "paddlerec_b200")


def _config(tmp_path):
    from paddlerec_b200 import runner
    path = os.path.join(PKG, "rank", "deepfm", "config.yaml")
    cfg = runner.load_yaml(path)
    cfg["config_abs_dir"] = os.path.dirname(path)
    cfg["runner.model_save_path"] = str(tmp_path / "out")
    return cfg


def test_config1_loop_matches_oracle(tmp_path):
    from paddlerec_b200 import runner
    cfg = _config(tmp_path)
    cfg["hyper_parameters.sparse_feature_number"] = 1000001
    dm = runner.load_dy_model_class(cfg["config_abs_dir"])
    torch.manual_seed(12345)
    model = dm.create_model(cfg)
    opt = dm.create_optimizer(model, cfg)
    loader = runner.create_data_loader(cfg, "train")
    # oracle twin (float64, CPU) with lazy Adam on the tables and dense Adam elsewhere
    p = {k: v.detach().cpu().double() for k, v in model.state_dict().items()}
    mom = {k: (np.zeros(v.shape), np.zeros(v.shape)) for k, v in p.items()}
    metrics, _ = dm.create_metrics()
    n_steps = 6
    for step, batch in enumerate(loader):
        if step == n_steps:
            break
        opt.clear_grad()
        loss, metrics, _ = dm.train_forward(model, metrics, batch, cfg)
        loss.backward()
        opt.step()
        # oracle step
        q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        ids = torch.cat([b.reshape(-1, 1) for b in batch[1:27]], 1)
        pred = nets.deepfm_forward(q, [ids[:, i:i + 1] for i in range(26)], batch[27].double(), 4)
        oloss = nets.log_loss(pred, batch[0].reshape(-1, 1).double()).mean()
        oloss.backward()
        assert abs(float(loss) - float(oloss)) < 1e-4 * max(1.0, abs(float(oloss))), step
        t = step + 1
        for k in p:
            g = q[k].grad
            if g is None:
                continue
            if k.startswith("fm.embedding"):
                rows = torch.unique(ids[ids != 0]).numpy()
                W, m, v = oo.adam_lazy(p[k].numpy(), mom[k][0], mom[k][1], rows, g.numpy()[rows],
                                       1e-3, 0.9, 0.999, 1e-8, t)
            else:
                W, m, v = oo.adam_lazy(p[k].numpy(), mom[k][0], mom[k][1], slice(None), g.numpy(),
                                       1e-3, 0.9, 0.999, 1e-8, t)
            p[k], mom[k] = torch.from_numpy(W), (m, v)
    sd = model.state_dict()
    for k in ("dnn.linear_0.weight", "fm.dense_w", "fm.dense_w_one"):
        assert rel_err(sd[k], p[k]) < 1e-4, k
    touched = torch.unique(ids[ids != 0])
    assert rel_err(sd["fm.embedding.weight"][touched.cuda()], p["fm.embedding.weight"][touched]) < 1e-4
    assert 0.0 <= metrics[0].accumulate() <= 1.0


def test_trainer_cli_runs_and_checkpoints(tmp_path):
    from paddlerec_b200 import runner
    cfg = _config(tmp_path)
    cfg["runner.epochs"] = 1
    losses, metric_values, model = runner.train(cfg, max_steps=12)
    assert len(losses) == 12 and all(np.isfinite(losses))
    assert "auc" in metric_values
    ckpt = os.path.join(cfg["runner.model_save_path"], "0")
    assert os.path.exists(os.path.join(ckpt, "rec.pdparams"))
    dm = runner.load_dy_model_class(cfg["config_abs_dir"])
    model2 = dm.create_model(cfg)
    runner.load_model(ckpt, model2)
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k


def test_packed_reader_trains_identically_to_the_dataloader(tmp_path):
    """runner.reader_type=PackedReader (native parser, packed batches in pinned memory) must give
    the same training trajectory as the per-sample Python reader: same batches, same seed."""
    from paddlerec_b200 import runner
    runs = []
    for reader_type in ("DataLoader", "PackedReader"):
        cfg = _config(tmp_path)
        cfg["runner.epochs"] = 1
        cfg["runner.train_batch_size"] = 16
        cfg["runner.reader_type"] = reader_type
        cfg["hyper_parameters.sparse_feature_number"] = 1000001
        losses, metric_values, _ = runner.train(cfg, max_steps=5, save=False)
        runs.append((losses, metric_values["auc"]))
    assert len(runs[0][0]) == 5
    # same batches in the same order => the same trajectory (different batches differ at 1e-1)
    assert np.allclose(runs[0][0], runs[1][0], rtol=1e-5, atol=0)
    assert abs(runs[0][1] - runs[1][1]) < 1e-6


def test_dlrm_model_dir_trains(tmp_path):
    """The DLRM plugin directory (net.py + dygraph_model.py + yaml) through the trainer loop:
    finite, decreasing loss on the bundled sample; AUC and accuracy metrics; checkpoint names are
    the reference's (BatchNorm `_mean` / `_variance` included)."""
    from paddlerec_b200 import runner
    path = os.path.join(PKG, "rank", "dlrm", "config.yaml")
    cfg = runner.load_yaml(path)
    cfg["config_abs_dir"] = os.path.dirname(path)
    cfg["runner.model_save_path"] = str(tmp_path / "out")
    cfg["runner.epochs"] = 2
    cfg["runner.train_batch_size"] = 16
    cfg["runner.reader_type"] = "PackedReader"
    cfg["hyper_parameters.optimizer.learning_rate"] = 0.01
    losses, metric_values, model = runner.train(cfg)
    assert len(losses) == 10 and all(np.isfinite(losses))
    assert np.mean(losses[5:]) < np.mean(losses[:5])
    assert set(metric_values) == {"auc", "accuracy"} and 0.0 <= metric_values["accuracy"] <= 1.0
    sd = model.state_dict()
    for k in ("embedding.weight", "bot_mlp.dense_0.weight", "bot_mlp.norm_3._variance",
              "top_mlp.dense_2.bias", "top_mlp.norm_2._mean"):
        assert k in sd, k
    assert os.path.exists(os.path.join(cfg["runner.model_save_path"], "1", "rec.pdparams"))
