"""Import-path shim for `from utils.utils import initialize_weights` (models/model_clam.py:4 in the reference); the
reference's other utils modules (file_utils, core_utils, eval_utils, ...) stay importable from its checkout."""
from hipt_abmil_atec23_b200.shim import extend_package_path

extend_package_path(__name__, __path__)
