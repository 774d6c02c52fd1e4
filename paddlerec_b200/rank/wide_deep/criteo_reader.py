"""Wide&Deep reads the same Criteo slot files as DeepFM (the reference keeps a copy of the reader in
every model directory: models/rank/wide_deep/criteo_reader.py)."""
from ..deepfm.criteo_reader import RecDataset  # noqa: F401
