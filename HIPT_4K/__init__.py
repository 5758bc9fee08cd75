"""Import-path shim: the reference's callers do `from HIPT_4K.hipt_4k import HIPT_4K` (extract_features_fp.py:15,
create_heatmaps.py:23).  The implementation lives in hipt_abmil_atec23_b200/."""
