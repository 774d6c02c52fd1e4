"""DIN reader — same sample contract as the reference's models/rank/din/dinReader.py:22-144:
lines `hist_items;hist_cats;target_item;target_cat;label`; groups of 20 batches are sorted by
history length, each batch is padded with id 0 to its own max length, the mask is 0 / -1e9 stored
as int64 [L,1], and the target ids are tiled L times.  (The reference also writes the global max
length to ./tmp.txt as a side effect, dinReader.py:39-41; that is not reproduced.)"""
from __future__ import annotations

import numpy as np
from torch.utils.data import IterableDataset


class RecDataset(IterableDataset):
    def __init__(self, file_list, config):
        super().__init__()
        self.file_list = sorted(file_list)
        self.batch_size = config.get("runner.train_batch_size")
        self.group_size = self.batch_size * 20

    def _records(self):
        for path in self.file_list:
            with open(path) as fh:
                for line in fh:
                    parts = line.strip().split(";")
                    if len(parts) < 5:
                        continue
                    yield (parts[0].split(), parts[1].split(), parts[2], parts[3], float(parts[4]))

    def _emit(self, group, drop_remainder):
        group = sorted(group, key=lambda r: len(r[0]))
        end = len(group) - (len(group) % self.batch_size if drop_remainder else 0)
        for i in range(0, end, self.batch_size):
            b = group[i:i + self.batch_size]
            L = max(len(r[0]) for r in b)
            for hist, cats, ti, tc, label in b:
                n = len(hist)
                item = np.asarray([int(x) for x in hist] + [0] * (L - n), dtype=np.int64)
                cat = np.asarray([int(x) for x in cats] + [0] * (L - n), dtype=np.int64)
                mask = np.asarray([0] * n + [int(-1e9)] * (L - n), dtype=np.int64).reshape(L, 1)
                yield [item, cat, np.asarray(int(ti), dtype=np.int64),
                       np.asarray(int(tc), dtype=np.int64), np.asarray(label, dtype=np.float32),
                       mask, np.full(L, int(ti), dtype=np.int64), np.full(L, int(tc), dtype=np.int64)]

    def __iter__(self):
        group = []
        for rec in self._records():
            group.append(rec)
            if len(group) == self.group_size:
                yield from self._emit(group, drop_remainder=False)
                group = []
        if group:
            yield from self._emit(group, drop_remainder=True)
