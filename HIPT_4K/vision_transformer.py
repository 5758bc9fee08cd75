from hipt_abmil_atec23_b200.vision_transformer import (Attention, Block, Mlp, PatchEmbed, VisionTransformer,  # noqa: F401
                                                       vit_base, vit_small, vit_tiny)
