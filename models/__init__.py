"""Import-path shim for `from models.model_clam import CLAM_MB, CLAM_SB` / `from models.model_mil import MIL_fc, MIL_fc_mc`
(utils/core_utils.py:6-7, utils/eval_utils.py:5-6, create_heatmaps.py:15 in the reference); `models.resnet_custom` and the
rest of the reference's package stay importable from its checkout."""
from hipt_abmil_atec23_b200.shim import extend_package_path

extend_package_path(__name__, __path__)
