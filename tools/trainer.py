#!/usr/bin/env python
"""`python -u tools/trainer.py -m <config.yaml> [-o key=value ...]` — same CLI as the reference's
tools/trainer.py:39-46,226-228; the loop lives in paddlerec_b200/runner.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlerec_b200.runner import main  # noqa: E402

if __name__ == "__main__":
    main()
