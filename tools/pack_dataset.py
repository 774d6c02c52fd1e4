#!/usr/bin/env python
"""Parse text datasets ONCE into the packed binary cache that PackedBatchReader streams at memcpy
speed (`runner.reader_type: PackedReader`, `runner.packed_format: packed`).

    python tools/pack_dataset.py --format slot_text --schema criteo --out cache/ data/part-*.txt
    python tools/pack_dataset.py --format criteo_tsv --hash xxh32 --hash-dim 1000001 --out cache/ day_0

One `<input name>.b2r` per input file (dataio.write_packed layout: magic, (n, F, Dn), label, ids,
dense), so the reference's file-level sharding across ranks (criteo_reader.py:30-43) keeps working.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlerec_b200 import dataio  # noqa: E402


def pack_file(path, out_dir, fmt, schema, hash_kind, hash_dim, threads, chunk_bytes=64 << 20):
    rd = dataio.PackedBatchReader([path], batch_size=1 << 20, schema=schema, fmt=fmt, hash_kind=hash_kind,
                                  hash_dim=hash_dim, drop_last=False, threads=threads,
                                  chunk_bytes=chunk_bytes, prefetch=0, as_torch=False)
    parts = [tuple(np.array(a) for a in b) for b in rd]
    if parts:
        label, ids, dense = (np.concatenate([p[k] for p in parts]) for k in range(3))
    else:
        F = 26 if fmt == "criteo_tsv" else schema.n_sparse
        Dn = 13 if fmt == "criteo_tsv" else (schema.dense_dim if schema.dense_slot else 0)
        label, ids, dense = np.zeros((0, 1), np.int64), np.zeros((0, F), np.int64), np.zeros((0, Dn), np.float32)
    out = os.path.join(out_dir, os.path.basename(path) + ".b2r")
    dataio.write_packed(out, label, ids, dense)
    return out, ids.shape[0]


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("files", nargs="+")
    ap.add_argument("--out", required=True)
    ap.add_argument("--format", default="slot_text", choices=["slot_text", "criteo_tsv"])
    ap.add_argument("--schema", default="criteo", choices=["criteo", "criteo_dcn_v2"])
    ap.add_argument("--hash", default="std", choices=["std", "xxh32"])
    ap.add_argument("--hash-dim", type=int, default=1000001)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args(argv)
    os.makedirs(a.out, exist_ok=True)
    schema = {"criteo": dataio.CRITEO, "criteo_dcn_v2": dataio.CRITEO_DCN_V2}[a.schema]
    hash_kind = dataio.HASH_STD if a.hash == "std" else dataio.HASH_XXH32
    total, t0 = 0, time.time()
    for path in a.files:
        out, n = pack_file(path, a.out, a.format, schema, hash_kind, a.hash_dim, a.threads)
        total += n
        print("%s -> %s (%d samples)" % (path, out, n))
    dt = time.time() - t0
    print("%d samples in %.2f s (%.2f M samples/s)" % (total, dt, total / max(dt, 1e-9) / 1e6))
    return total


if __name__ == "__main__":
    main()
