"""Import-path shim for `from utils.utils import initialize_weights` (models/model_clam.py:4 in the reference)."""
