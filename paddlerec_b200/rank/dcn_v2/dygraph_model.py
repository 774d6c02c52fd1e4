"""DygraphModel adapter for DCN-V2 (reference: models/rank/dcn_v2/dygraph_model.py:22-137)."""
from __future__ import annotations

import torch

from ... import functional as BF
from ... import optim
from ..deepfm.dygraph_model import DygraphModel as _CriteoBase
from . import net


class DygraphModel(_CriteoBase):
    def create_model(self, config):
        g = config.get
        return net.DCN_V2Layer(
            g("hyper_parameters.sparse_feature_number"), g("hyper_parameters.sparse_feature_dim"),
            g("hyper_parameters.dense_input_dim"), g("hyper_parameters.sparse_inputs_slots") - 1,
            g("hyper_parameters.fc_sizes"), g("hyper_parameters.cross_num"),
            g("hyper_parameters.is_Stacked", None), g("hyper_parameters.use_low_rank_mixture", None),
            g("hyper_parameters.low_rank", 32), g("hyper_parameters.num_experts", 4),
            device=self.device)

    def create_optimizer(self, dy_model, config):
        lr = config.get("hyper_parameters.optimizer.learning_rate", 0.001)
        clip = optim.ClipGradByGlobalNorm(
            config.get("hyper_parameters.optimizer.clip_by_norm", 10.0))   # dygraph_model.py:83-87
        return optim.Adam(learning_rate=lr, parameters=dy_model.parameters(), grad_clip=clip)

    def train_forward(self, dy_model, metrics_list, batch_data, config):
        label, sparse_tensor, dense_tensor = self.create_feeds(batch_data, config)
        pred = dy_model.forward(sparse_tensor, dense_tensor)
        log_loss = self.create_loss(pred, label)
        if metrics_list:
            metrics_list[0].update(preds=pred.detach(), labels=label)
        return log_loss, metrics_list, {"log_loss": log_loss}
