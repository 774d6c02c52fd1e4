"""Reader for the Criteo `slot:feasign` text format — same contract as the reference's
models/rank/deepfm/criteo_reader.py:21-103: every sample is [label(int64[1]), 26 x int64[1],
dense(float32[13])]; a slot missing on a line is filled with the padding id 0.

Format (doc/custom_reader.md:5-24): space-separated `slot:value` tokens; slots are `click`,
`1`..`26` (hashed feasigns) and `dense_feature` (13 floats, one token each).
"""
from __future__ import annotations

import numpy as np
from torch.utils.data import IterableDataset

SPARSE_SLOTS = ["click"] + [str(i) for i in range(1, 27)]
DENSE_SLOT = "dense_feature"
DENSE_DIM = 13


class RecDataset(IterableDataset):
    def __init__(self, file_list, config=None, rank=0, world_size=1):
        super().__init__()
        files = sorted(file_list)
        use_fleet = bool(config.get("runner.use_fleet", False)) if config else False
        self.inference = bool(config.get("runner.inference", False)) if config else False
        if use_fleet and world_size > 1:  # shard FILES by rank like criteo_reader.py:30-43
            if len(files) < world_size:
                raise ValueError("The number of data files is less than the number of workers")
            files = files[rank::world_size]
        self.file_list = files
        self._index = {s: i for i, s in enumerate(SPARSE_SLOTS)}

    def parse_line(self, line: str):
        sparse = [[] for _ in SPARSE_SLOTS]
        dense = []
        for tok in line.strip().split(" "):
            slot, _, val = tok.partition(":")
            if slot == DENSE_SLOT:
                dense.append(float(val))
            elif slot in self._index:
                sparse[self._index[slot]].append(int(val))
        out = [np.asarray(v if v else [0], dtype=np.int64) for v in sparse]
        out.append(np.asarray(dense if dense else [0.0] * DENSE_DIM, dtype=np.float32))
        return out[1:] if self.inference else out

    def __iter__(self):
        for path in self.file_list:
            with open(path, "r") as fh:
                for line in fh:
                    if line.strip():
                        yield self.parse_line(line)
