from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K  # noqa: F401
from hipt_abmil_atec23_b200.hipt_model_utils import (eval_transforms, get_vit256, get_vit4k, roll_batch2img,  # noqa: F401
                                                     tensorbatch2im)
