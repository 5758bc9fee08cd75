"""Import-path shim: the reference's callers do `from HIPT_4K.hipt_4k import HIPT_4K` (extract_features_fp.py:15,
create_heatmaps.py:23).  The implementation lives in hipt_abmil_atec23_b200/; the reference's other HIPT_4K modules
(hipt_heatmap_utils, attention_visualization_utils) stay importable from its checkout."""
from hipt_abmil_atec23_b200.shim import extend_package_path

extend_package_path(__name__, __path__)
