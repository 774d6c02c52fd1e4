"""oracle/readers.py — pure-Python restatement of the reference's text readers.  TEST
INFRASTRUCTURE (see oracle/__init__.py): the checker for libb200rec_io.so, small inputs only.

Unlike the arithmetic in oracle/nets.py this part is PINNED against the real reference: the
raw-Criteo path (`criteo_tsv_lines`, `std_hash_string`) is checked against the reference's own
tools/dataset/parser.cpp compiled unmodified into oracle/_ref/criteo_parser (oracle/Makefile), both
live and through tests/golden/criteo_tsv_parser_cpp.txt; `slot_text_lines` is checked against the
reference's models/rank/deepfm/criteo_reader.py imported from /root/reference when present
(tests/test_dataio.py).  xxHash32 is pinned on the algorithm's published test vectors; the
`xxhash` Python package benchmark_reader.py imports is not installed here.
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple

import numpy as np

M64 = (1 << 64) - 1
M32 = (1 << 32) - 1

CONT_MIN = [0, -3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]           # parser.cpp:37, benchmark_reader.py:23
CONT_DIFF = [20, 603, 100, 50, 64000, 500, 100, 50, 500, 10, 10, 10, 50]  # parser.cpp:39, :25
HASH_DIM = 1000001                                             # parser.cpp:40, benchmark_reader.py:26


# ---- models/rank/deepfm/criteo_reader.py:61-103 ---------------------------------------------------
def slot_text_lines(lines: Sequence[str], sparse_slots: Sequence[str], dense_slot: str, dense_dim: int):
    """One entry per line: ([values of sparse slot 0], ..., [dense values]).  Follows the reader
    statement by statement: strip + split(" ") (:67), split(":") and take [0]/[1] (:70-71), skip
    unknown slots (:72-73), int()/float() (:74-77), pad a slot that did not appear with [0] or
    [0]*dense_dim (:80-89).  Blank lines are skipped (the packed readers' one documented
    difference: the reference would emit an all-padding sample)."""
    slots = list(sparse_slots) + [dense_slot]
    out = []
    for l in lines:
        toks = l.strip().split(" ")
        if toks == [""]:
            continue
        vals: List[list] = [[] for _ in slots]
        for t in toks:
            sf = t.split(":")
            if sf[0] not in slots:
                continue
            i = slots.index(sf[0])
            vals[i].append(float(sf[1]) if sf[0] == dense_slot else int(sf[1]))
        for i, s in enumerate(slots):
            if not vals[i]:
                vals[i] = [0] * dense_dim if s == dense_slot else [0]
        out.append(vals)
    return out


def slot_text_packed(lines, sparse_slots, dense_slot, dense_dim):
    """The fixed-length case as (ids[n, len(sparse_slots)] int64, dense[n, dense_dim] float32)."""
    rows = slot_text_lines(lines, sparse_slots, dense_slot, dense_dim)
    ids = np.asarray([[v[0] for v in r[:-1]] for r in rows], dtype=np.int64).reshape(len(rows), len(sparse_slots))
    dense = np.asarray([r[-1] for r in rows], dtype=np.float64).astype(np.float32).reshape(len(rows), dense_dim)
    return ids, dense


# ---- libstdc++ std::hash<std::string> (what parser.cpp:68 calls) ----------------------------------
def std_hash_string(s: bytes) -> int:
    """64-bit std::_Hash_bytes of libstdc++ (MurmurHash64A-style, seed 0xc70f6907)."""
    mul = ((0xc6a4a793 << 32) + 0x5bd1e995) & M64
    n = len(s)
    h = (0xc70f6907 ^ (n * mul)) & M64
    body = n & ~7
    for i in range(0, body, 8):
        d = struct.unpack_from("<Q", s, i)[0]
        d = (d * mul) & M64
        d ^= d >> 47
        d = (d * mul) & M64
        h ^= d
        h = (h * mul) & M64
    if n & 7:
        d = int.from_bytes(s[body:], "little")
        h ^= d
        h = (h * mul) & M64
    h ^= h >> 47
    h = (h * mul) & M64
    h ^= h >> 47
    return h


# ---- xxHash32 (benchmark_reader.py:52 calls xxhash.xxh32(..).intdigest()) --------------------------
def xxh32(s: bytes, seed: int = 0) -> int:
    P1, P2, P3, P4, P5 = 2654435761, 2246822519, 3266489917, 668265263, 374761393
    rotl = lambda x, r: ((x << r) | (x >> (32 - r))) & M32
    n, p = len(s), 0
    if n >= 16:
        v = [(seed + P1 + P2) & M32, (seed + P2) & M32, seed & M32, (seed - P1) & M32]
        while p <= n - 16:
            for k in range(4):
                w = struct.unpack_from("<I", s, p)[0]
                v[k] = (rotl((v[k] + w * P2) & M32, 13) * P1) & M32
                p += 4
        h = (rotl(v[0], 1) + rotl(v[1], 7) + rotl(v[2], 12) + rotl(v[3], 18)) & M32
    else:
        h = (seed + P5) & M32
    h = (h + n) & M32
    while p + 4 <= n:
        w = struct.unpack_from("<I", s, p)[0]
        h = (rotl((h + w * P3) & M32, 17) * P4) & M32
        p += 4
    while p < n:
        h = (rotl((h + s[p] * P5) & M32, 11) * P1) & M32
        p += 1
    h ^= h >> 15
    h = (h * P2) & M32
    h ^= h >> 13
    h = (h * P3) & M32
    h ^= h >> 16
    return h


# ---- tools/dataset/parser.cpp:43-77 and models/rank/dnn/benchmark_reader.py:39-56 -------------------
def criteo_tsv_lines(lines: Sequence[str], hash_kind: str = "std", hash_dim: int = HASH_DIM
                     ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """(label[n], ids[n,26], dense[n,13] float32, n_skipped).  hash_kind "std": parser.cpp — lines
    without exactly 40 tab-separated columns are skipped (:49-51), dense = (stod(x)-min)/diff or 0
    for an empty column (:55-63), id = std::hash<string>(column) % hash_dim (:68).  "xxh32":
    benchmark_reader.line_process — id = xxh32(str(idx) + column) % hash_dim (:50-52)."""
    label, ids, dense, skipped = [], [], [], 0
    for l in lines:
        l = l.rstrip("\n")
        if l.endswith("\r"):
            l = l[:-1]
        if l == "":
            continue
        cols = l.split("\t")
        if hash_kind == "std" and len(cols) != 40:
            skipped += 1
            continue
        d = []
        for i in range(1, 14):
            d.append(0.0 if cols[i] == "" else (float(cols[i]) - CONT_MIN[i - 1]) / CONT_DIFF[i - 1])
        row = []
        for i in range(14, 40):
            if hash_kind == "std":
                row.append(std_hash_string(cols[i].encode()) % hash_dim)
            else:
                row.append(xxh32((str(i) + cols[i]).encode()) % hash_dim)
        label.append(int(cols[0]))
        ids.append(row)
        dense.append(d)
    n = len(label)
    return (np.asarray(label, np.int64), np.asarray(ids, np.int64).reshape(n, 26),
            np.asarray(dense, np.float64).astype(np.float32).reshape(n, 13), skipped)


def multislot_lines(lines: Sequence[str], slot_is_float: Sequence[bool]):
    """`<count> v1 .. v_count` per slot (tools/dataset/README.MD example; parser.cpp:54-75 output).
    Returns per line a list of per-slot value lists."""
    out = []
    for l in lines:
        toks = l.split()
        if not toks:
            continue
        at, row = 0, []
        for isf in slot_is_float:
            cnt = int(toks[at]); at += 1
            vals = toks[at:at + cnt]; at += cnt
            assert cnt > 0 and len(vals) == cnt
            row.append([float(v) for v in vals] if isf else [int(v) for v in vals])
        assert at == len(toks)
        out.append(row)
    return out


# ---- device-side feasign fold (include/b200rec.h: b200rec_hash_keys) --------------------------------
def hash_keys(keys: np.ndarray, V: int, slot_of_key=None, reserve_zero: bool = True) -> np.ndarray:
    """numpy restatement of the uint64 feasign -> row fold: splitmix64 finaliser of
    key ^ (slot+1)*golden, then `% (V-1) + 1` (row 0 reserved for feasign 0, the readers' padding
    key — criteo_reader.py:48,86-88) or `% V`.  The reference folds on the host with
    xxh32(str(slot)+token) % hash_dim (benchmark_reader.py:50-52); this is the same contract
    (deterministic, slot-salted, uniform over rows) with a 64-bit mixer that needs no string."""
    k = np.asarray(keys).astype(np.uint64)
    z = k.copy()
    if slot_of_key is not None:
        salt = (np.asarray(slot_of_key).astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = z ^ salt
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    if reserve_zero:
        r = np.uint64(1) + z % np.uint64(V - 1)
        r = np.where(k == 0, np.uint64(0), r)
    else:
        r = z % np.uint64(V)
    return r.astype(np.int64)
