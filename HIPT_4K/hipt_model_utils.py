from hipt_abmil_atec23_b200.hipt_model_utils import (HIPT_MEAN, HIPT_STD, eval_transforms, get_vit256,  # noqa: F401
                                                     get_vit4k, roll_batch2img, tensorbatch2im)
