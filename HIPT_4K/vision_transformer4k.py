from hipt_abmil_atec23_b200.vision_transformer4k import VisionTransformer4K, count_parameters, vit4k_xs  # noqa: F401
