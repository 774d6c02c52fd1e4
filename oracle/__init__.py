"""oracle/ — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this package, and only as the checker or as the timed CPU baseline.  Nothing under
`paddlerec_b200/` imports it; the product path raises if the CUDA library is missing.

PARITY UNPINNED (against real Paddle): the arithmetic of this path lives in PaddlePaddle core, a
pip dependency of the reference that is neither vendored under /root/reference nor installable
here (no network; requirements.txt leaves it unpinned, CI used paddlepaddle==2.0.0rc0,
.travis.yml:32), and the reference ships no numerical tests or golden vectors for it
(SURVEY.md §4, §8c).  What IS pinned: oracle/nets.py agrees to ~1e-12 (float64) with the
reference's own models/rank/{deepfm,dcn_v2,din,wide_deep}/net.py executed UNMODIFIED on top of
oracle/paddle_shim.py (a torch-backed stand-in for the ~40 paddle APIs those files call), and the
committed tests/golden/*.npz were minted from that execution by tests/golden/make_golden.py.
So the composition of operators is the reference's; the per-operator semantics (Embedding
padding, Linear layout, log_loss eps, softmax axis) are restated from Paddle's public docs.

PINNED, in contrast: oracle/readers.py (the input-format row) — the reference's own
tools/dataset/parser.cpp compiles from its single source file (oracle/Makefile -> oracle/_ref/
criteo_parser) and its models/rank/deepfm/criteo_reader.py imports unmodified; the goldens
tests/golden/criteo_tsv_parser_cpp.txt and slot_text_criteo_reader.npz are their outputs.

Files
  nets.py         forward restatements (torch CPU, fp32/fp64), file:line cited per statement
  optim.py        row-wise optimizer rules (numpy)
  paddle_shim.py  the stand-in used to run the reference's net.py here
  readers.py      the reference's text readers (slot:value, raw Criteo TSV, multislot) + the two
                  string hashes they use, pure Python
  interact_ref.c  plain-C restatement of the DLRM dot interaction (fwd/bwd) and the feasign fold
  fm_ref.c        plain-C restatement of the fused FM forward/backward (double accumulation),
                  built by oracle/Makefile into oracle/_build/libfm_ref.so
"""
