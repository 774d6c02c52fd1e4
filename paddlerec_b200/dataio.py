"""Native input pipeline: the reference's text formats -> packed batch arrays.

`libb200rec_io.so` (include/b200rec_io.h, csrc/slot_io.cc — plain C++17, no CUDA) does the parsing
on all host cores with the GIL released; this module binds it with ctypes and feeds batches as the
packed triple `(label[B,1] i64, ids[B,F] i64, dense[B,Dn] f32)` that `DygraphModel.create_feeds`
accepts — three arrays per batch instead of the reference's 28 per SAMPLE
(models/rank/deepfm/criteo_reader.py:92-103, dygraph_model.py:41-50).

  parse_slot_text / parse_slot_text_lod   `slot:value` lines (doc/custom_reader.md:5-24)
  parse_multislot                         `<count> v..` lines (QueueDataset wire format)
  parse_criteo_tsv                        raw Criteo TSV (tools/dataset/parser.cpp,
                                          models/rank/dnn/benchmark_reader.py)
  PackedBatchReader                       files -> fixed-size batches, parsed one chunk ahead on a
                                          background thread into pinned host memory
  write_packed / read_packed              binary cache of the packed arrays (parse once)
"""
from __future__ import annotations

import ctypes
import os
import queue
import subprocess
import threading
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_uint32, c_uint64, c_void_p
from dataclasses import dataclass, field
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import CSRC, INCLUDE_DIR, LIB_DIR

IO_LIB_PATH = os.path.join(LIB_DIR, "libb200rec_io.so")
IO_ABI_VERSION = 1
HASH_STD, HASH_XXH32 = 0, 1
GXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-shared", "-pthread"]


class B200RecIOError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s (rc=%d)" % (msg, code))
        self.code = code


_P = c_void_p
_SIG = {
    "b200rec_io_abi_version": (c_int, []),
    "b200rec_io_last_error": (c_char_p, []),
    "b200rec_io_count_lines": (c_int, [_P, c_size_t, POINTER(c_int64)]),
    "b200rec_io_parse_slot_text": (c_int, [_P, c_size_t, c_char_p, POINTER(c_char_p), c_int, c_char_p,
                                           c_int, _P, _P, _P, c_int64, POINTER(c_int64), c_int]),
    "b200rec_io_parse_slot_text_ex": (c_int, [_P, c_size_t, c_char_p, POINTER(c_char_p), c_int, c_char_p,
                                              c_int, c_int, _P, _P, _P, c_int64, POINTER(c_int64),
                                              c_int]),
    "b200rec_io_parse_slot_text_lod": (c_int, [_P, c_size_t, c_char_p, POINTER(c_char_p), c_int,
                                               c_char_p, c_int, _P, _P, _P, _P, c_int64, c_int64,
                                               POINTER(c_int64), POINTER(c_int64), c_int]),
    "b200rec_io_parse_multislot": (c_int, [_P, c_size_t, POINTER(c_int), c_int, _P, _P, c_int64, _P,
                                           _P, c_int64, c_int64, POINTER(c_int64), POINTER(c_int64),
                                           POINTER(c_int64), c_int]),
    "b200rec_io_parse_criteo_tsv": (c_int, [_P, c_size_t, c_int, c_int64, POINTER(c_double),
                                            POINTER(c_double), _P, _P, _P, c_int64, POINTER(c_int64),
                                            POINTER(c_int64), c_int]),
    "b200rec_io_parse_din": (c_int, [_P, c_size_t, _P, _P, _P, _P, _P, _P, c_int64, c_int64,
                                     POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), c_int]),
    "b200rec_io_hash_std_string": (c_uint64, [c_char_p, c_size_t]),
    "b200rec_io_xxh32": (c_uint32, [c_char_p, c_size_t, c_uint32]),
}
_lib = None


def _sources():
    return [os.path.join(INCLUDE_DIR, "b200rec_io.h"), os.path.join(CSRC, "slot_io.cc")]


def needs_build() -> bool:
    if not os.path.exists(IO_LIB_PATH):
        return True
    t = os.path.getmtime(IO_LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False) -> str:
    """g++ csrc/slot_io.cc -> paddlerec_b200/lib/libb200rec_io.so (in-tree, ~3 s)."""
    if not force and not needs_build():
        return IO_LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = IO_LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [os.environ.get("CXX", "g++"), *GXX_FLAGS, "-I", INCLUDE_DIR, _sources()[1], "-o", tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, IO_LIB_PATH)
    return IO_LIB_PATH


def declared_symbols() -> List[str]:
    import re

    text = open(os.path.join(INCLUDE_DIR, "b200rec_io.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rec_io_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if needs_build():
        try:
            build()
        except (RuntimeError, FileNotFoundError) as e:
            if not os.path.exists(IO_LIB_PATH):
                raise RuntimeError("libb200rec_io.so is missing and could not be built: %s" % e)
    lib = ctypes.CDLL(IO_LIB_PATH)
    for name, (res, args) in _SIG.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.b200rec_io_abi_version() != IO_ABI_VERSION:
        raise RuntimeError("libb200rec_io ABI version mismatch")
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        raise B200RecIOError(rc, load().b200rec_io_last_error().decode("utf-8", "replace"))


def _buf(data) -> Tuple[c_void_p, int, object]:
    """(pointer, length, keep-alive) of bytes / bytearray / memoryview / mmap / uint8 ndarray."""
    if isinstance(data, str):
        data = data.encode("utf-8")
    if isinstance(data, bytes):
        return ctypes.cast(ctypes.c_char_p(data), c_void_p), len(data), data
    arr = np.frombuffer(data, dtype=np.uint8)
    return c_void_p(arr.ctypes.data), arr.size, arr


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else c_void_p(a.ctypes.data)


def _names(names: Sequence[str]):
    arr = (c_char_p * len(names))(*[n.encode() for n in names])
    return arr


@dataclass(frozen=True)
class SlotSchema:
    """Which slots of a `slot:value` line go where.  Default: the Criteo layout every rank model of
    the reference reads (criteo_reader.py:49-53): label `click`, slots `1`..`26`, 13 dense floats."""
    sparse_slots: Tuple[str, ...] = tuple(str(i) for i in range(1, 27))
    label_slot: Optional[str] = "click"
    dense_slot: Optional[str] = "dense_feature"
    dense_dim: int = 13
    dense_log1p: bool = False          # dense = log(v + 1)   (models/rank/dcn_v2/reader.py:63-64)
    skip_empty_sparse: bool = False    # ignore `slot:` with no value (dcn_v2/reader.py:55-57)

    @property
    def flags(self) -> int:
        return (1 if self.dense_log1p else 0) | (2 if self.skip_empty_sparse else 0)

    @property
    def n_sparse(self) -> int:
        return len(self.sparse_slots)


CRITEO = SlotSchema()
CRITEO_DCN_V2 = SlotSchema(dense_log1p=True, skip_empty_sparse=True)


def _line_bound(data) -> int:
    """Cheap upper bound of the number of samples in `data` (newline count + 1)."""
    if isinstance(data, str):
        return data.count("\n") + 1
    if isinstance(data, (bytes, bytearray)):
        return data.count(b"\n") + 1
    return count_lines(data)


def count_lines(data) -> int:
    p, n, _keep = _buf(data)
    out = c_int64(0)
    _check(load().b200rec_io_count_lines(p, n, ctypes.byref(out)))
    return out.value


def _alloc(n: int, schema: SlotSchema, out):
    if out is not None:
        label, ids, dense = out
        assert ids.dtype == np.int64 and ids.flags.c_contiguous and ids.shape[1] == schema.n_sparse
        return label, ids, dense
    label = np.empty((n, 1), np.int64) if schema.label_slot else None
    ids = np.empty((n, schema.n_sparse), np.int64)
    dense = np.empty((n, schema.dense_dim), np.float32) if schema.dense_slot else None
    return label, ids, dense


def parse_slot_text(data, schema: SlotSchema = CRITEO, out=None, threads: int = 0):
    """`slot:value` text -> (label[n,1], ids[n,F], dense[n,Dn]) numpy arrays (views of `out` when
    given: a (label, ids, dense) triple with at least as many rows as the text has lines)."""
    p, nbytes, _keep = _buf(data)
    cap = out[1].shape[0] if out is not None else _line_bound(data)
    label, ids, dense = _alloc(cap, schema, out)
    n = c_int64(0)
    _check(load().b200rec_io_parse_slot_text_ex(
        p, nbytes, schema.label_slot.encode() if schema.label_slot else None,
        _names(schema.sparse_slots), schema.n_sparse,
        schema.dense_slot.encode() if schema.dense_slot else None, schema.dense_dim, schema.flags,
        _np_ptr(label), _np_ptr(ids), _np_ptr(dense), cap, ctypes.byref(n), threads))
    k = n.value
    return (label[:k] if label is not None else None, ids[:k], dense[:k] if dense is not None else None)


def parse_slot_text_lod(data, schema: SlotSchema = CRITEO, threads: int = 0):
    """Variable-length slots -> (label[n,1], keys[K], offsets[n*F+1], dense[n,Dn]); bag n*F+f holds
    keys[offsets[n*F+f]:offsets[n*F+f+1]] — what ops.gather_pool_sum consumes."""
    p, nbytes, _keep = _buf(data)
    cap = count_lines(data)
    keys_cap = max(1, nbytes // 2)  # every key needs >= "s:v" + separator... 2 bytes is a safe floor
    label = np.empty((cap, 1), np.int64) if schema.label_slot else None
    dense = np.empty((cap, schema.dense_dim), np.float32) if schema.dense_slot else None
    keys = np.empty(max(keys_cap, cap * schema.n_sparse), np.int64)
    offsets = np.empty(cap * schema.n_sparse + 1, np.int64)
    n, nk = c_int64(0), c_int64(0)
    _check(load().b200rec_io_parse_slot_text_lod(
        p, nbytes, schema.label_slot.encode() if schema.label_slot else None,
        _names(schema.sparse_slots), schema.n_sparse,
        schema.dense_slot.encode() if schema.dense_slot else None, schema.dense_dim,
        _np_ptr(label), _np_ptr(keys), _np_ptr(offsets), _np_ptr(dense), cap, keys.size,
        ctypes.byref(n), ctypes.byref(nk), threads))
    # `cap` counts every line, the parser drops blank / whitespace-only ones: trim to the n samples
    # it actually produced (the tail rows are uninitialised; offsets would index out of bounds)
    k = n.value
    F = schema.n_sparse
    return (label[:k] if label is not None else None, keys[:nk.value].copy(),
            offsets[:k * F + 1].copy(), dense[:k] if dense is not None else None)


def parse_multislot(data, slot_is_float: Sequence[bool], threads: int = 0):
    """MultiSlot lines -> dict(keys u64[K], key_offsets[n*n_int+1], fvals f32[M],
    float_offsets[n*n_float+1], n)."""
    p, nbytes, _keep = _buf(data)
    cap = count_lines(data)
    n_float = sum(bool(x) for x in slot_is_float)
    n_int = len(slot_is_float) - n_float
    vals_cap = max(1, nbytes // 2)
    keys = np.empty(vals_cap if n_int else 0, np.uint64)
    fvals = np.empty(vals_cap if n_float else 0, np.float32)
    koff = np.empty(cap * n_int + 1, np.int64) if n_int else None
    foff = np.empty(cap * n_float + 1, np.int64) if n_float else None
    flags = (c_int * len(slot_is_float))(*[int(bool(x)) for x in slot_is_float])
    n, nk, nf = c_int64(0), c_int64(0), c_int64(0)
    _check(load().b200rec_io_parse_multislot(
        p, nbytes, flags, len(slot_is_float), _np_ptr(keys) if n_int else None, _np_ptr(koff),
        keys.size, _np_ptr(fvals) if n_float else None, _np_ptr(foff), fvals.size, cap,
        ctypes.byref(n), ctypes.byref(nk), ctypes.byref(nf), threads))
    k = n.value
    return {"n": k, "keys": keys[:nk.value].copy(),
            "key_offsets": koff[:k * n_int + 1].copy() if koff is not None else None,
            "fvals": fvals[:nf.value].copy(),
            "float_offsets": foff[:k * n_float + 1].copy() if foff is not None else None}


def parse_criteo_tsv(data, hash_kind: int = HASH_STD, hash_dim: int = 1000001, cont_min=None,
                     cont_diff=None, out=None, threads: int = 0):
    """Raw Criteo TSV -> (label[n,1], ids[n,26], dense[n,13], n_skipped)."""
    p, nbytes, _keep = _buf(data)
    cap = out[1].shape[0] if out is not None else _line_bound(data)
    label, ids, dense = _alloc(cap, SlotSchema(), out)
    cm = (c_double * 13)(*cont_min) if cont_min is not None else None
    cd = (c_double * 13)(*cont_diff) if cont_diff is not None else None
    n, skipped = c_int64(0), c_int64(0)
    _check(load().b200rec_io_parse_criteo_tsv(p, nbytes, hash_kind, hash_dim, cm, cd, _np_ptr(label),
                                              _np_ptr(ids), _np_ptr(dense), cap, ctypes.byref(n),
                                              ctypes.byref(skipped), threads))
    k = n.value
    return label[:k], ids[:k], dense[:k], skipped.value


def parse_din(data, threads: int = 0):
    """DIN behaviour-log lines -> dict(hist_items, hist_cats int64[K], offsets int64[n+1],
    target_item, target_cat int64[n], label float32[n], n_skipped)."""
    p, nbytes, _keep = _buf(data)
    cap = count_lines(data)
    keys_cap = max(1, nbytes // 2)
    items = np.empty(keys_cap, np.int64)
    cats = np.empty(keys_cap, np.int64)
    offsets = np.empty(cap + 1, np.int64)
    ti, tc = np.empty(cap, np.int64), np.empty(cap, np.int64)
    label = np.empty(cap, np.float32)
    n, nk, skipped = c_int64(0), c_int64(0), c_int64(0)
    _check(load().b200rec_io_parse_din(p, nbytes, _np_ptr(items), _np_ptr(cats), _np_ptr(offsets),
                                       _np_ptr(ti), _np_ptr(tc), _np_ptr(label), cap, keys_cap,
                                       ctypes.byref(n), ctypes.byref(nk), ctypes.byref(skipped), threads))
    k = n.value
    return {"hist_items": items[:nk.value].copy(), "hist_cats": cats[:nk.value].copy(),
            "offsets": offsets[:k + 1].copy(), "target_item": ti[:k].copy(), "target_cat": tc[:k].copy(),
            "label": label[:k].copy(), "n_skipped": skipped.value}


class DinBatchReader:
    """Batches of the reference's DIN reader (models/rank/din/dinReader.py:45-144) built from the
    natively parsed LoD arrays with numpy instead of per-sample Python: records are taken in file
    order in groups of 20*batch_size, each group is stably sorted by history length, cut into
    batches, and every batch is padded with id 0 to ITS OWN max length; the mask is 0 / -1e9 stored
    as int64 [B, L, 1]; the target ids are tiled L times; the tail group drops its incomplete
    batch.  Yields the 8-tuple the DataLoader collate of that reader produces:
    (hist_item[B,L], hist_cat[B,L], target_item[B], target_cat[B], label[B] f32, mask[B,L,1] i64,
     target_item_seq[B,L], target_cat_seq[B,L])."""

    def __init__(self, file_list: Sequence[str], batch_size: int, threads: int = 0,
                 as_torch: bool = True, pin_memory: bool = False):
        self.file_list = sorted(file_list)
        self.batch_size = int(batch_size)
        self.group_size = self.batch_size * 20
        self.threads, self.as_torch, self.pin_memory = threads, as_torch, pin_memory

    def _records(self):
        parts = []
        for path in self.file_list:
            with open(path, "rb") as fh:
                parts.append(parse_din(fh.read(), self.threads))
        lens = np.concatenate([np.diff(p["offsets"]) for p in parts]) if parts else np.zeros(0, np.int64)
        cat = lambda k: np.concatenate([p[k] for p in parts]) if parts else np.zeros(0, np.int64)  # noqa: E731
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        return cat("hist_items"), cat("hist_cats"), offsets, cat("target_item"), cat("target_cat"), cat("label")

    def _batch(self, rec, idx):
        items, cats, offsets, ti, tc, label = rec
        lens = (offsets[idx + 1] - offsets[idx]).astype(np.int64)
        B, L = idx.size, int(lens.max())
        col = np.arange(L, dtype=np.int64)[None, :]
        valid = col < lens[:, None]
        src = (offsets[idx][:, None] + col)[valid]
        hist_item = np.zeros((B, L), np.int64)
        hist_cat = np.zeros((B, L), np.int64)
        hist_item[valid] = items[src]
        hist_cat[valid] = cats[src]
        mask = np.where(valid, 0, int(-1e9)).astype(np.int64).reshape(B, L, 1)
        out = (hist_item, hist_cat, ti[idx], tc[idx], label[idx].astype(np.float32), mask,
               np.repeat(ti[idx][:, None], L, 1), np.repeat(tc[idx][:, None], L, 1))
        if not self.as_torch:
            return out
        import torch

        ts = tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in out)
        if self.pin_memory and torch.cuda.is_available():
            ts = tuple(t.pin_memory() for t in ts)
        return ts

    def __iter__(self):
        rec = self._records()
        n = rec[3].size
        lens = np.diff(rec[2])
        B, G = self.batch_size, self.group_size
        for g0 in range(0, n, G):
            g1 = min(n, g0 + G)
            order = g0 + np.argsort(lens[g0:g1], kind="stable")
            end = (g1 - g0) if g1 - g0 == G else (g1 - g0) - (g1 - g0) % B
            for i in range(0, end, B):
                yield self._batch(rec, order[i:i + B])


def hash_std_string(s: bytes) -> int:
    return int(load().b200rec_io_hash_std_string(s, len(s)))


def xxh32(s: bytes, seed: int = 0) -> int:
    return int(load().b200rec_io_xxh32(s, len(s), seed))


# ---- binary cache --------------------------------------------------------------------------------
_MAGIC = b"B2RPACK1"


def write_packed(path: str, label: np.ndarray, ids: np.ndarray, dense: np.ndarray) -> None:
    """Packed arrays -> one file: magic, (n, F, Dn) int64 header, label, ids, dense (C order)."""
    n, F = ids.shape
    Dn = dense.shape[1]
    with open(path, "wb") as fh:
        fh.write(_MAGIC)
        fh.write(np.asarray([n, F, Dn], np.int64).tobytes())
        fh.write(np.ascontiguousarray(label, np.int64).tobytes())
        fh.write(np.ascontiguousarray(ids, np.int64).tobytes())
        fh.write(np.ascontiguousarray(dense, np.float32).tobytes())


def read_packed(path: str):
    """Memory-maps a write_packed file -> (label[n,1], ids[n,F], dense[n,Dn]) read-only views."""
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    if bytes(mm[:8]) != _MAGIC:
        raise ValueError("%s is not a b200rec packed file" % path)
    n, F, Dn = (int(x) for x in np.frombuffer(mm[8:32], np.int64))
    o = 32
    label = np.frombuffer(mm, np.int64, n, o).reshape(n, 1)
    o += 8 * n
    ids = np.frombuffer(mm, np.int64, n * F, o).reshape(n, F)
    o += 8 * n * F
    dense = np.frombuffer(mm, np.float32, n * Dn, o).reshape(n, Dn)
    return label, ids, dense


# ---- batch reader --------------------------------------------------------------------------------
@dataclass
class PackedBatchReader:
    """Iterates fixed-size packed batches over a list of text files.

    A producer thread reads `chunk_bytes` at a time (cut at the last newline), parses the chunk with
    all cores (the C call releases the GIL) and slices it into batches; `prefetch` parsed batches are
    queued ahead of the consumer.  With `pin_memory` the batch tensors live in pinned host memory,
    so `tensor.to(device, non_blocking=True)` is a true async H2D copy.  File order, line order and
    the drop-last rule are those of the reference's DataLoader(batch_size, drop_last=True) over
    criteo_reader.RecDataset (tools/utils/utils_single.py:89-113); files are sharded by rank the way
    criteo_reader.py:30-43 does when `use_fleet`.
    """
    file_list: Sequence[str]
    batch_size: int
    schema: SlotSchema = CRITEO
    fmt: str = "slot_text"          # "slot_text" | "criteo_tsv" | "packed"
    hash_kind: int = HASH_STD
    hash_dim: int = 1000001
    drop_last: bool = True
    threads: int = 0
    chunk_bytes: int = 16 << 20
    prefetch: int = 4
    pin_memory: bool = False
    as_torch: bool = True
    rank: int = 0
    world_size: int = 1
    shard_files: bool = False
    _files: List[str] = field(default_factory=list, init=False, repr=False)
    _scratch_bufs: Optional[tuple] = field(default=None, init=False, repr=False)

    def __post_init__(self):
        files = sorted(self.file_list)
        if self.shard_files and self.world_size > 1:
            if len(files) < self.world_size:
                raise ValueError("The number of data files is less than the number of workers")
            files = files[self.rank::self.world_size]
        self._files = files
        if self.fmt not in ("slot_text", "criteo_tsv", "packed"):
            raise ValueError("unknown fmt %r" % self.fmt)

    # -- parsing -----------------------------------------------------------------------------------
    def _scratch(self, n: int):
        """Grow-only parse buffers, reused across chunks (fresh arrays would page-fault every time)."""
        cur = self._scratch_bufs
        if cur is None or cur[1].shape[0] < n:
            cap = max(n, int(1.25 * (cur[1].shape[0] if cur is not None else 0)))
            F, Dn = self.schema.n_sparse, (self.schema.dense_dim if self.schema.dense_slot else 0)
            if self.fmt == "criteo_tsv":
                F, Dn = 26, 13
            cur = (np.empty((cap, 1), np.int64), np.empty((cap, F), np.int64),
                   np.empty((cap, Dn), np.float32))
            self._scratch_bufs = cur
        return cur

    def _parse(self, chunk):
        n = count_lines(chunk)
        out = self._scratch(n)
        if self.fmt == "slot_text":
            label, ids, dense = parse_slot_text(chunk, self.schema, out=out, threads=self.threads)
            if label is None:
                label = np.zeros((ids.shape[0], 1), np.int64)
            if dense is None:
                dense = np.zeros((ids.shape[0], 0), np.float32)
            return label, ids, dense
        label, ids, dense, _ = parse_criteo_tsv(chunk, self.hash_kind, self.hash_dim, out=out,
                                                threads=self.threads)
        return label, ids, dense

    def _chunks(self) -> Iterator[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
        """Parsed chunks as views of the scratch buffers — valid until the next chunk is parsed."""
        buf = bytearray(self.chunk_bytes + (1 << 16))
        for path in self._files:
            if self.fmt == "packed":
                yield read_packed(path)
                continue
            have = 0                     # bytes of an unfinished last line carried to the front
            with open(path, "rb", buffering=0) as fh:
                while True:
                    if have + self.chunk_bytes > len(buf):
                        buf.extend(bytes(have + self.chunk_bytes - len(buf)))
                    mv = memoryview(buf)
                    got = fh.readinto(mv[have:have + self.chunk_bytes])
                    if not got:
                        break
                    end = have + got
                    cut = buf.rfind(b"\n", 0, end)
                    if cut < 0:          # a line longer than the chunk: keep reading
                        have = end
                        mv.release()
                        continue
                    yield self._parse(mv[:cut + 1])
                    tail = end - (cut + 1)
                    buf[:tail] = buf[cut + 1:end]
                    have = tail
                    mv.release()
            if have and bytes(buf[:have]).strip():
                yield self._parse(memoryview(buf)[:have])

    def _new_batch(self, B: int, Dn: int, F: int):
        if self.as_torch:
            import torch

            pin = bool(self.pin_memory and torch.cuda.is_available())
            ts = (torch.empty((B, 1), dtype=torch.int64, pin_memory=pin),
                  torch.empty((B, F), dtype=torch.int64, pin_memory=pin),
                  torch.empty((B, Dn), dtype=torch.float32, pin_memory=pin))
            return ts, tuple(t.numpy() for t in ts)
        arrs = (np.empty((B, 1), np.int64), np.empty((B, F), np.int64), np.empty((B, Dn), np.float32))
        return arrs, arrs

    def _batches(self):
        """Every batch owns its memory (pinned when asked): chunk views are copied in as they come,
        which is the one extra pass over the parsed data (268 B per Criteo sample)."""
        B = self.batch_size
        cur = views = None
        have = 0
        for label, ids, dense in self._chunks():
            at, n = 0, ids.shape[0]
            while at < n:
                if cur is None:
                    cur, views = self._new_batch(B, dense.shape[1], ids.shape[1])
                take = min(B - have, n - at)
                views[0][have:have + take] = label[at:at + take]
                views[1][have:have + take] = ids[at:at + take]
                views[2][have:have + take] = dense[at:at + take]
                have += take
                at += take
                if have == B:
                    yield cur
                    cur, have = None, 0
        if have and not self.drop_last:
            yield tuple(t[:have] for t in cur)

    def __iter__(self):
        if self.prefetch <= 0:
            yield from self._batches()
            return
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()
        END = object()

        def produce():
            try:
                for b in self._batches():
                    while not stop.is_set():
                        try:
                            q.put(b, timeout=0.1)
                            break
                        except queue.Full:
                            continue
                    if stop.is_set():
                        return
                q.put(END)
            except BaseException as e:  # surface parse errors on the consumer side
                q.put(e)

        th = threading.Thread(target=produce, name="b200rec-reader", daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is END:
                    break
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            th.join(timeout=5)
