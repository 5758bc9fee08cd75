"""Config 4 of BASELINE.json: CLAM_SB gated-ABMIL on ragged bags of 50-20,000 instances x 192-d, reported as achieved
HBM GB/s (algorithmic bytes: 772 B / instance forward = 192 fp32 features read once + one fp32 score written; SURVEY
§8d).  One launch covers all bags (and all folds).  python tools/bench_clam.py [--size hipt_smaller] [--folds 1]"""
import argparse, json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hipt_abmil_atec23_b200 import clam_engine, _lib
from hipt_abmil_atec23_b200.model_clam import CLAM_SB


def bag_lengths(n_bags=256, lo=50, hi=20000, seed=4):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n_bags, generator=g)
    return torch.exp(math.log(lo) + u * (math.log(hi) - math.log(lo))).long().clamp(lo, hi)


def run(size, folds, reps=20):
    dev = torch.device("cuda:0")
    lens = bag_lengths()
    offs = torch.zeros(lens.numel() + 1, dtype=torch.int32)
    offs[1:] = torch.cumsum(lens, 0)
    total = int(offs[-1])
    feats = torch.randn((total, 192), generator=torch.Generator().manual_seed(5)).to(dev)
    models = []
    for f in range(folds):
        torch.manual_seed(10 + f)
        models.append(CLAM_SB(size_arg=size, dropout=0.0, n_classes=2).eval().to(dev))
    offs_d = offs.to(dev)
    mx = int(lens.max())
    f = lambda: clam_engine.forward_bags(models, feats, offs_d, max_bag_len=mx, want=("logits", "y_prob", "y_hat"))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_alg = total * (192 * 4 + 4 * folds)
    # per-kernel times of a second pass (CUDA events around every launch; the headline `ms` above is taken without them)
    _lib.prof_enable(True)
    for _ in range(reps): f()
    torch.cuda.synchronize()
    kern_us = {k: round(v[0] / v[1] * 1e3, 2) for k, v in _lib.prof_read().items()}
    _lib.prof_enable(False)
    score_us = kern_us.get("clam_scores")
    return {"size": size, "folds": folds, "ms": ms, "kernels_us": kern_us,
            "score_kernel_GBps": bytes_alg / score_us / 1e3 if score_us else None,
            "bags": int(lens.numel()), "instances": total,
            "bags_per_s": lens.numel() / ms * 1e3, "instances_per_s": total / ms * 1e3,
            "algorithmic_GBps": bytes_alg / ms / 1e6, "feature_MB": total * 768 / 1e6}


def run_train(size, n, steps=50):
    """One bag per step as in the reference's train_loop: forward, CE, backward (fused recomputing kernel), Adam (one launch)."""
    import torch.nn.functional as F
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    model = CLAM_SB(size_arg=size, dropout=0.0, n_classes=2).to(dev).train()
    opt = clam_engine.FusedAdam(filter(lambda p: p.requires_grad, model.parameters()), lr=2e-4, weight_decay=1e-5)
    bag = torch.randn((n, 192), generator=torch.Generator().manual_seed(5)).to(dev)
    label = torch.tensor([1], device=dev)
    def step():
        logits, _, _, _, _ = model(bag)
        F.cross_entropy(logits, label).backward()
        opt.step(); opt.zero_grad()
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # the same step without autograd: forward, backward with the cross-entropy fused in, one-launch Adam (clam_engine.TrainStep)
    torch.manual_seed(2)
    model2 = CLAM_SB(size_arg=size, dropout=0.0, n_classes=2).to(dev).train()
    ts = clam_engine.TrainStep(model2, clam_engine.FusedAdam(clam_engine._param_list(model2), lr=2e-4, weight_decay=1e-5), n)
    for _ in range(5): ts.step(bag, label)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps): ts.step(bag, label)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / steps
    return {"size": size, "train_step": True, "instances": n, "ms_per_step": ms, "steps_per_s": 1e3 / ms,
            "lean_ms_per_step": ms2, "lean_steps_per_s": 1e3 / ms2, "algorithmic_GBps": n * 1544 / ms2 / 1e6}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default=None)
    ap.add_argument("--folds", type=int, default=None)
    a = ap.parse_args()
    for size in ([a.size] if a.size else ["hipt_smaller", "hipt_big"]):
        for folds in ([a.folds] if a.folds else [1, 5]):
            print(json.dumps(run(size, folds)))
        for n in (50, 1000, 20000):
            print(json.dumps(run_train(size, n)))
