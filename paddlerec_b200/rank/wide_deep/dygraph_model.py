"""DygraphModel adapter for Wide&Deep (reference: models/rank/wide_deep/dygraph_model.py:23-100)."""
from __future__ import annotations

from ..deepfm.dygraph_model import DygraphModel as _CriteoBase
from . import net


class DygraphModel(_CriteoBase):
    def create_model(self, config):
        g = config.get
        return net.WideDeepLayer(
            g("hyper_parameters.sparse_feature_number"), g("hyper_parameters.sparse_feature_dim"),
            g("hyper_parameters.dense_input_dim"), g("hyper_parameters.sparse_inputs_slots") - 1,
            g("hyper_parameters.fc_sizes"), device=self.device)
