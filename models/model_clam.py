from hipt_abmil_atec23_b200.model_clam import CLAM_MB, CLAM_SB, Attn_Net, Attn_Net_Gated  # noqa: F401
