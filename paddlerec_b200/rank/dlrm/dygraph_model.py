"""DygraphModel adapter for DLRM (reference: models/rank/dlrm/dygraph_model.py:24-110): two-class
scores, softmax cross-entropy, Adam, AUC + top-1 accuracy.  Feeds are the Criteo ones."""
from __future__ import annotations

import torch

from ... import functional as BF
from ... import optim
from ..deepfm.dygraph_model import DygraphModel as _CriteoBase
from . import net


class DygraphModel(_CriteoBase):
    def create_model(self, config):
        g = config.get
        return net.DLRMLayer(
            dense_feature_dim=g("hyper_parameters.dense_input_dim"),
            bot_layer_sizes=g("hyper_parameters.bot_layer_sizes"),
            sparse_feature_number=g("hyper_parameters.sparse_feature_number"),
            sparse_feature_dim=g("hyper_parameters.sparse_feature_dim"),
            top_layer_sizes=g("hyper_parameters.top_layer_sizes"),
            num_field=g("hyper_parameters.num_field"), self_interaction=False, device=self.device)

    def create_loss(self, raw_predict_2d, label):
        return BF.softmax_cross_entropy(raw_predict_2d, label).mean()            # :58-62

    def create_optimizer(self, dy_model, config):
        lr = config.get("hyper_parameters.optimizer.learning_rate", 0.001)
        return optim.Adam(learning_rate=lr, parameters=dy_model.parameters(), lazy_mode=True)

    def create_metrics(self):
        return [BF.Auc("ROC"), BF.Accuracy()], ["auc", "accuracy"]

    def _update_metrics(self, metrics_list, raw_pred_2d, label):
        predict_2d = torch.softmax(raw_pred_2d.detach(), dim=1)                  # :88
        if metrics_list:
            metrics_list[0].update(preds=predict_2d, labels=label)
            metrics_list[1].update(metrics_list[1].compute(pred=predict_2d, label=label))

    def train_forward(self, dy_model, metrics_list, batch_data, config):
        label, sparse_tensor, dense_tensor = self.create_feeds(batch_data, config)
        raw_pred_2d = dy_model.forward(sparse_tensor, dense_tensor)
        loss = self.create_loss(raw_pred_2d, label)
        self._update_metrics(metrics_list, raw_pred_2d, label)
        return loss, metrics_list, {"loss": loss}

    def infer_forward(self, dy_model, metrics_list, batch_data, config):
        label, sparse_tensor, dense_tensor = self.create_feeds(batch_data, config)
        with torch.no_grad():
            raw_pred_2d = dy_model.forward(sparse_tensor, dense_tensor)
        self._update_metrics(metrics_list, raw_pred_2d, label)
        return metrics_list, None
