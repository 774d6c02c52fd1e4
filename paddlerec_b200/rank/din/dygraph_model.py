"""DygraphModel adapter for DIN (reference: models/rank/din/dygraph_model.py:21-115): BCE-with-logits
loss, SGD with PiecewiseDecay([410000], [base_lr, 0.2]), AUC on sigmoid(logit)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from ... import functional as BF
from ... import optim
from . import net


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.asarray(x))


class DygraphModel:
    device = "cuda"

    def create_model(self, config):
        g = config.get
        return net.DINLayer(g("hyper_parameters.item_emb_size", 64),
                            g("hyper_parameters.cat_emb_size", 64),
                            g("hyper_parameters.act", "sigmoid"),
                            g("hyper_parameters.is_sparse", False),
                            g("hyper_parameters.use_DataLoader", False),
                            g("hyper_parameters.item_count", 63001),
                            g("hyper_parameters.cat_count", 801), device=self.device)

    def create_feeds(self, batch, config):
        dev = self.device
        b = [_t(x).to(dev, non_blocking=True) for x in batch]
        label = b[4].reshape(-1, 1).to(torch.float32)
        return b[0], b[1], b[2].reshape(-1), b[3].reshape(-1), label, b[5], b[6], b[7]

    def create_loss(self, raw_pred, label):
        return F.binary_cross_entropy_with_logits(raw_pred, label, reduction="mean")

    def create_optimizer(self, dy_model, config):
        base_lr = config.get("hyper_parameters.optimizer.learning_rate_base_lr")
        lr = optim.PiecewiseDecay(boundaries=[410000], values=[base_lr, 0.2])
        return optim.SGD(learning_rate=lr, parameters=dy_model.parameters())

    def create_metrics(self):
        return [BF.Auc("ROC")], ["auc"]

    def train_forward(self, dy_model, metrics_list, batch_data, config):
        feeds = self.create_feeds(batch_data, config)
        raw_pred = dy_model.forward(*feeds)
        loss = self.create_loss(raw_pred, feeds[4])
        if metrics_list:
            metrics_list[0].update(preds=torch.sigmoid(raw_pred.detach()), labels=feeds[4])
        return loss, metrics_list, {"loss": loss}

    def infer_forward(self, dy_model, metrics_list, batch_data, config):
        feeds = self.create_feeds(batch_data, config)
        with torch.no_grad():
            raw_pred = dy_model.forward(*feeds)
        if metrics_list:
            metrics_list[0].update(preds=torch.sigmoid(raw_pred), labels=feeds[4])
        return metrics_list, None
