"""DygraphModel adapter for DeepFM — the reference's plugin surface
(models/rank/deepfm/dygraph_model.py:23-98; contract in doc/model_develop.md:3-45): the same seven
methods with the same argument meaning, over CUDA tensors and the b200rec kernels.

Differences that are deliberate:
  * create_feeds makes ONE [B,26] int64 + ONE [B,13] f32 host->device copy instead of 28;
  * the AUC histograms stay on the device (no `.numpy()` sync per step);
  * Adam updates the two tables lazily (touched rows only).
"""
from __future__ import annotations

import numpy as np
import torch

from ... import functional as BF
from ... import optim
from . import net


def _to_host_tensor(b):
    if isinstance(b, torch.Tensor):
        return b
    return torch.from_numpy(np.asarray(b))


class DygraphModel:
    device = "cuda"

    def create_model(self, config):
        """The yaml's hyper_parameters -> DeepFMLayer on `self.device` (reference :25-38).  The label
        occupies one of `sparse_inputs_slots`, hence the -1."""
        sparse_feature_number = config.get("hyper_parameters.sparse_feature_number")
        sparse_feature_dim = config.get("hyper_parameters.sparse_feature_dim")
        fc_sizes = config.get("hyper_parameters.fc_sizes")
        dense_feature_dim = config.get("hyper_parameters.dense_input_dim")
        sparse_input_slot = config.get("hyper_parameters.sparse_inputs_slots")
        return net.DeepFMLayer(sparse_feature_number, sparse_feature_dim, dense_feature_dim,
                               sparse_input_slot - 1, fc_sizes, device=self.device)

    def create_feeds(self, batch_data, config):
        """batch_data: [label, C1..C26, dense] as the reader/DataLoader yields them (host arrays or
        tensors, each [B] or [B,1]) or the packed form (label[B,1] i64, ids[B,26] i64, dense[B,13])."""
        dense_feature_dim = config.get("hyper_parameters.dense_input_dim")
        if len(batch_data) == 3:
            label, ids, dense = (_to_host_tensor(b) for b in batch_data)
        else:
            cols = [_to_host_tensor(b).reshape(-1, 1).to(torch.int64) for b in batch_data[:-1]]
            label, ids = cols[0], torch.cat(cols[1:], dim=1)
            dense = _to_host_tensor(batch_data[-1])
        dense = dense.to(torch.float32).reshape(-1, dense_feature_dim)
        dev = self.device
        label = label.reshape(-1, 1).to(dev, non_blocking=True)
        ids = ids.to(dev, non_blocking=True)
        dense = dense.to(dev, non_blocking=True)
        return label, ids, dense

    def create_loss(self, pred, label):
        """Mean log-loss with Paddle's epsilon 1e-4 on the sigmoid output (reference :53-58)."""
        return BF.log_loss_mean(pred, label)

    def create_optimizer(self, dy_model, config):
        """Adam over every parameter; the two tables are updated row-wise from their SelectedRows
        gradients (lazy_mode — the reference's dygraph Adam :61-65 is non-lazy, its static twin
        static_model.py:101-103 lazy; a non-lazy pass over V=1e8 rows would move 77 GB per step)."""
        lr = config.get("hyper_parameters.optimizer.learning_rate", 0.001)
        return optim.Adam(learning_rate=lr, parameters=dy_model.parameters(), lazy_mode=True)

    def create_metrics(self):
        """One ROC-AUC metric named "auc" (reference :68-72); its histograms live on the device."""
        return [BF.Auc("ROC")], ["auc"]

    def train_forward(self, dy_model, metrics_list, batch_data, config):
        """-> (loss, metrics_list, print_dict) as doc/model_develop.md:35-39 requires; the metric
        update reads the detached prediction without leaving the device (reference :75-87)."""
        label, sparse_tensor, dense_tensor = self.create_feeds(batch_data, config)
        pred = dy_model.forward(sparse_tensor, dense_tensor)
        loss = self.create_loss(pred, label)
        if metrics_list:
            metrics_list[0].update(preds=pred.detach(), labels=label)
        return loss, metrics_list, {"loss": loss}

    def infer_forward(self, dy_model, metrics_list, batch_data, config):
        """-> (metrics_list, None): forward under no_grad + metric update (reference :89-98)."""
        label, sparse_tensor, dense_tensor = self.create_feeds(batch_data, config)
        with torch.no_grad():
            pred = dy_model.forward(sparse_tensor, dense_tensor)
        if metrics_list:
            metrics_list[0].update(preds=pred, labels=label)
        return metrics_list, None
