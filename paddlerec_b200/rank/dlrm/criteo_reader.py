"""DLRM reads the same Criteo slot files as DeepFM (the reference keeps a byte-identical copy of the
reader in every model directory: models/rank/dlrm/criteo_reader.py)."""
from ..deepfm.criteo_reader import RecDataset  # noqa: F401
