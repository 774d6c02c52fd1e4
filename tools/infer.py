#!/usr/bin/env python
"""`python -u tools/infer.py -m <config.yaml> [-o key=value ...]` — same CLI as the reference's
tools/infer.py:38-45,200-202: evaluates every checkpoint `<infer_load_path>/<epoch>/rec.pdparams`
for epochs [infer_start_epoch, infer_end_epoch); the loop lives in paddlerec_b200/runner.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlerec_b200.runner import main  # noqa: E402

if __name__ == "__main__":
    main(mode="infer")
