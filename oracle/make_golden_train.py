"""Training-mode goldens of the REFERENCE CLAM_SB (imported unmodified from /root/reference) -> tests/golden/clam_train_reference.pt

    python oracle/make_golden_train.py      # in the build container only

TEST INFRASTRUCTURE.  Pins the instance-clustering branch (models/model_clam.py:116-178) and the training configuration the
reference actually uses (dropout 0.85, docs/README.md:186-193) for the oracle and the CUDA path:
  * eval-mode `instance_eval=True` cases (n_classes 2 / 5, subtyping on / off): instance loss, predictions, targets;
  * train-mode gradients of  bag_weight * CE(logits, label) + (1 - bag_weight) * instance_loss  (train_loop_clam,
    utils/core_utils.py:300-371, bag_weight 0.7) at dropout 0;
  * train-mode outputs and gradients at dropout 0.85 / 0.25 with KNOWN masks: a forward hook on each nn.Dropout of the
    unmodified module replaces its random mask by the mask libhipt_b200's hb_clam_dropout_masks produces for a fixed seed
    (torch's own RNG stream cannot be reproduced inside a kernel: SURVEY.md §7).
"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from make_golden import _import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "clam_train_reference.pt")


def grads_of(mod):
    return {k: (p.grad.clone() if p.grad is not None else None) for k, p in mod.named_parameters()}


def main():
    _, _, rclam, _ = _import_reference()
    sys.path.insert(0, ROOT)
    from hipt_abmil_atec23_b200 import clam_engine
    gold = {"inst_eval": {}, "train": {}}

    # ---------------------------------------------------------------- instance_eval, eval mode
    for name, size_arg, ncls, subtyping, n, label, seed in (("smaller_2c", "hipt_smaller", 2, False, 300, 1, 40),
                                                            ("smaller_2c_sub", "hipt_smaller", 2, True, 300, 0, 41),
                                                            ("medium_5c_sub", "hipt_medium", 5, True, 120, 3, 42),
                                                            ("big_5c", "hipt_big", 5, False, 64, 2, 43)):
        torch.manual_seed(seed)
        mod = rclam.CLAM_SB(size_arg=size_arg, dropout=0.0, n_classes=ncls, subtyping=subtyping, k_sample=8).eval()
        bag = torch.randn(n, 192, generator=torch.Generator().manual_seed(seed + 100))
        with torch.no_grad():
            logits, y_prob, y_hat, a_raw, res = mod(bag, label=torch.tensor([label]), instance_eval=True)
        gold["inst_eval"][name] = {"size_arg": size_arg, "n_classes": ncls, "subtyping": subtyping, "n": n, "label": label,
                                   "model_seed": seed, "bag_seed": seed + 100, "logits": logits.clone(), "a_raw": a_raw.clone(),
                                   "instance_loss": torch.as_tensor(res["instance_loss"]).clone(),
                                   "inst_preds": torch.as_tensor(res["inst_preds"]).clone(),
                                   "inst_labels": torch.as_tensor(res["inst_labels"]).clone()}

    # ---------------------------------------------------------------- training step gradients
    for name, size_arg, ncls, p, inst, subtyping, n, label, seed, mseed in (
            ("smaller_p0", "hipt_smaller", 2, 0.0, False, False, 75, 1, 50, 0),
            ("smaller_p0_inst", "hipt_smaller", 2, 0.0, True, False, 75, 0, 51, 0),
            ("smaller_p85", "hipt_smaller", 2, 0.85, False, False, 75, 1, 52, 1234567),
            ("smaller_p85_inst", "hipt_smaller", 2, 0.85, True, True, 200, 1, 53, 7654321),
            ("big_p25_5c", "hipt_big", 5, 0.25, False, False, 130, 4, 54, 99),
            ("small_p50", "hipt_small", 2, 0.5, True, False, 64, 1, 55, 2 ** 40 + 17)):
        torch.manual_seed(seed)
        mod = rclam.CLAM_SB(size_arg=size_arg, dropout=p, n_classes=ncls, subtyping=subtyping, k_sample=8).train()
        bag = torch.randn(n, 192, generator=torch.Generator().manual_seed(seed + 100))
        L1, D = mod.attention_net[0].out_features, mod.attention_net[-1].attention_c.in_features
        if p > 0:
            m1, ma, mb = clam_engine.dropout_masks(n, L1, D, p, mseed)
            drops = [mod.attention_net[2], mod.attention_net[3].attention_a[2], mod.attention_net[3].attention_b[2]]
            assert all(isinstance(d, nn.Dropout) for d in drops)
            for d, m in zip(drops, (m1, ma, mb)):
                d.register_forward_hook(lambda module, inp, out, m=m: inp[0] * m)      # the module's own code, known mask
        lab = torch.tensor([label])
        logits, y_prob, y_hat, a_raw, res = mod(bag, label=lab, instance_eval=inst)
        loss = F.cross_entropy(logits, lab)
        total = 0.7 * loss + 0.3 * res["instance_loss"] if inst else loss
        total.backward()
        gold["train"][name] = {"size_arg": size_arg, "n_classes": ncls, "dropout": p, "instance_eval": inst, "subtyping": subtyping,
                               "n": n, "label": label, "model_seed": seed, "bag_seed": seed + 100, "mask_seed": mseed,
                               "logits": logits.detach().clone(), "a_raw": a_raw.detach().clone(), "loss": total.detach().clone(),
                               "grads": grads_of(mod)}
    torch.save(gold, OUT)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
