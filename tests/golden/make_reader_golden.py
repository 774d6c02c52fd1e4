"""Mints the reader goldens from the REFERENCE's own code (run in the build container, where
/root/reference exists; the outputs are committed so the tests run anywhere).

  criteo_tsv_sample.tsv        synthetic raw-Criteo lines (seeded; a few malformed on purpose)
  criteo_tsv_parser_cpp.txt    what the reference's tools/dataset/parser.cpp prints for them —
                               produced by oracle/_ref/criteo_parser, i.e. parser.cpp compiled
                               unmodified by oracle/Makefile
  slot_text_sample.txt         synthetic `slot:value` lines with missing / unknown / reordered slots
  slot_text_criteo_reader.npz  what the reference's models/rank/deepfm/criteo_reader.py yields for
                               them (imported from /root/reference on top of oracle/paddle_shim.py)

  din_sample.txt               synthetic DIN behaviour logs (`hist;cats;target;cat;label`)
  din_reader_batches.npz       what the reference's models/rank/din/dinReader.py yields for them at
                               batch_size 3 (one full 60-record group + a 17-record tail), stacked
                               per batch the way the DataLoader collates them

usage: python tests/golden/make_reader_golden.py
"""
import os
import random
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def make_tsv(rng, n=64):
    lines = []
    for i in range(n):
        cols = [str(rng.randint(0, 1))]
        for j in range(13):
            if rng.random() < 0.2:
                cols.append("")
            elif j == 1:
                cols.append(str(rng.randint(-3, 600)))
            else:
                cols.append(str(rng.randint(0, [20, 0, 100, 50, 64000, 500, 100, 50, 500, 10, 10, 10, 50][j])))
        for j in range(26):
            cols.append("" if rng.random() < 0.15 else "%08x" % rng.getrandbits(32))
        if i == 7:
            cols = cols[:-1]           # 39 columns: parser.cpp skips it
        if i == 19:
            cols.append("deadbeef")    # 41 columns: skipped too
        lines.append("\t".join(cols))
    lines.insert(30, "")               # empty line
    return "\n".join(lines) + "\n"


def make_slot_text(rng, n=48):
    lines = []
    for i in range(n):
        toks = ["click:%d" % rng.randint(0, 1)]
        toks += ["dense_feature:%s" % repr(round(rng.random() * rng.choice([0, 1, 1, 6.25]), rng.randint(1, 12)))
                 for _ in range(13)]
        for s in range(1, 27):
            if rng.random() < 0.1:
                continue               # missing slot -> padding id 0
            toks.append("%d:%d" % (s, rng.randint(0, 1000000)))
        if i % 5 == 0:
            toks.insert(rng.randint(0, len(toks)), "27:12345")          # unknown slot, ignored
            toks.insert(rng.randint(0, len(toks)), "user_tag:9")
        if i % 7 == 0:
            head, tail = toks[:14], toks[14:]
            rng.shuffle(tail)          # sparse slots in any order (dense order must be kept)
            toks = tail[:5] + head + tail[5:]
        # (a line WITHOUT the dense slot is not in this golden: criteo_reader.py:84-86 indexes
        #  dense_slots_shape with the slot index 27 and raises IndexError; tests cover our zeros.)
        if i == 13:
            toks = [t for t in toks if not t.startswith("click")]          # label missing -> 0
        lines.append(" ".join(toks))
    lines[20] = "  " + lines[20] + " \t"   # surrounding whitespace is stripped
    lines[21] = lines[21].replace(" 3:", "  3:", 1)  # a double space yields an empty token
    return "\n".join(lines) + "\n"


def make_din(rng, n=77):
    lines = []
    for i in range(n):
        k = min(40, max(1, int(rng.expovariate(1 / 6.0)) + 1))
        hist = " ".join(str(rng.randint(1, 63000)) for _ in range(k))
        cats = " ".join(str(rng.randint(1, 800)) for _ in range(k))
        lines.append("%s;%s;%d;%d;%d" % (hist, cats, rng.randint(1, 63000), rng.randint(1, 800),
                                        rng.randint(0, 1)))
    lines.insert(9, "1 2 3;4 5 6;7")          # fewer than 5 fields: skipped (dinReader.py:52-53)
    return "\n".join(lines) + "\n"


def din_golden(path):
    import importlib.util
    import tempfile

    from oracle import paddle_shim

    paddle_shim.install()
    spec = importlib.util.spec_from_file_location("ref_din_reader", "/root/reference/models/rank/din/dinReader.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())              # the reference writes ./tmp.txt as a side effect
    try:
        ds = mod.RecDataset([path], {"runner.train_batch_size": 3})
        samples = list(ds)
    finally:
        os.chdir(cwd)
    out = {}
    names = ["hist_item", "hist_cat", "target_item", "target_cat", "label", "mask", "target_item_seq",
             "target_cat_seq"]
    for b in range(len(samples) // 3):
        for j, name in enumerate(names):
            out["b%d/%s" % (b, name)] = np.stack([np.asarray(s[j]) for s in samples[3 * b:3 * b + 3]])
    np.savez_compressed(os.path.join(HERE, "din_reader_batches.npz"), **out)
    return len(samples)


def main():
    rng = random.Random(12345)
    tsv = make_tsv(rng)
    with open(os.path.join(HERE, "criteo_tsv_sample.tsv"), "w") as fh:
        fh.write(tsv)
    exe = os.path.join(ROOT, "oracle", "_ref", "criteo_parser")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    out = subprocess.run([exe], input=tsv.encode(), capture_output=True, check=True).stdout
    with open(os.path.join(HERE, "criteo_tsv_parser_cpp.txt"), "wb") as fh:
        fh.write(out)

    text = make_slot_text(rng)
    path = os.path.join(HERE, "slot_text_sample.txt")
    with open(path, "w") as fh:
        fh.write(text)
    from oracle import paddle_shim

    paddle_shim.install()
    sys.path.insert(0, "/root/reference/models/rank/deepfm")
    import criteo_reader  # the reference's reader, unmodified

    ds = criteo_reader.RecDataset([path], config=None)
    ds.inference = False
    samples = list(ds)
    ids = np.stack([np.concatenate(s[:27]) for s in samples]).astype(np.int64)
    dense = np.stack([s[27] for s in samples]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "slot_text_criteo_reader.npz"), ids=ids, dense=dense)
    din_path = os.path.join(HERE, "din_sample.txt")
    with open(din_path, "w") as fh:
        fh.write(make_din(rng))
    print("din: %d samples from the reference reader" % din_golden(din_path))
    print("tsv: %d lines -> %d parsed by parser.cpp; slot text: %d samples" %
          (tsv.count("\n"), out.count(b"\n"), len(samples)))


if __name__ == "__main__":
    main()
