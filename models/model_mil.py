from hipt_abmil_atec23_b200.model_mil import MIL_fc, MIL_fc_mc  # noqa: F401
