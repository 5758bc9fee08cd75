"""Import-path shim for `from models.model_clam import CLAM_MB, CLAM_SB` / `from models.model_mil import MIL_fc, MIL_fc_mc`
(utils/core_utils.py:6-7, utils/eval_utils.py:5-6, create_heatmaps.py:15 in the reference)."""
