"""DCN-V2 reader — the reference's models/rank/dcn_v2/reader.py:21-89: the Criteo `slot:value`
format of DeepFM with two differences: dense values are log(v + 1) (:63-64) and a sparse token with
an empty value is skipped (:55-57).  Native equivalent: dataio.CRITEO_DCN_V2."""
from __future__ import annotations

import numpy as np

from ..deepfm.criteo_reader import DENSE_DIM, DENSE_SLOT, SPARSE_SLOTS
from ..deepfm.criteo_reader import RecDataset as _CriteoDataset


class RecDataset(_CriteoDataset):
    def parse_line(self, line: str):
        sparse = [[] for _ in SPARSE_SLOTS]
        dense = []
        for tok in line.strip().split(" "):
            slot, _, val = tok.partition(":")
            if slot == DENSE_SLOT:
                dense.append(np.log(float(val) + 1))
            elif slot in self._index:
                if val == "":
                    continue
                sparse[self._index[slot]].append(int(val))
        out = [np.asarray(v if v else [0], dtype=np.int64) for v in sparse]
        out.append(np.asarray(dense if dense else [0.0] * DENSE_DIM, dtype=np.float32))
        return out
