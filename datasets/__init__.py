"""Makes `datasets` a REGULAR package rooted here, so that the reference's `from datasets.dataset_h5 import ...` /
`from datasets.dataset_generic import ...` (extract_features_fp.py:10, utils/eval_utils.py:19-21) resolve to the reference
checkout's datasets/ directory (a namespace package) instead of the HuggingFace `datasets` distribution installed in
site-packages, which would otherwise shadow it."""
from hipt_abmil_atec23_b200.shim import extend_package_path

extend_package_path(__name__, __path__)
