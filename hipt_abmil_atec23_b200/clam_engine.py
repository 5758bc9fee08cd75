"""Host side of the fused CLAM_SB pooling kernel (hb_clam_sb_forward in include/hipt_b200.h).

`forward_single` serves CLAM_SB.forward (one bag, one model — models/model_clam.py:147-191); `forward_bags` is the
batched form used for slide sets and fold ensembles: many ragged bags and up to 8 weight sets in one launch, the
feature matrix read from HBM once.
"""
import ctypes as C

import torch

from . import _lib


def _gate_module(model):
    return model.attention_net[-1]


def _weights(model, device):
    """The 10 fp32 tensors of one CLAM_SB in the order hb_clam_sb_forward documents."""
    g = _gate_module(model)
    ts = [model.attention_net[0].weight, model.attention_net[0].bias,
          g.attention_a[0].weight, g.attention_a[0].bias, g.attention_b[0].weight, g.attention_b[0].bias,
          g.attention_c.weight, g.attention_c.bias, model.classifiers.weight, model.classifiers.bias]
    out = []
    for t in ts:
        t = t.detach()
        if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=device, dtype=torch.float32).contiguous()
        out.append(t)
    return out


# ------------------------------------------------------------------------------------------------- training-mode dropout
_M32 = 0xFFFFFFFF


def _mul32(h, c):
    """(h * c) mod 2^32 on int64 tensors holding uint32 values (the product itself would overflow int64)."""
    return ((h & 0xFFFF) * c + ((((h >> 16) * c) & 0xFFFF) << 16)) & _M32


def _mix32(h):
    h = h ^ (h >> 16)
    h = _mul32(h, 0x85EBCA6B)
    h = h ^ (h >> 13)
    h = _mul32(h, 0xC2B2AE35)
    return h ^ (h >> 16)


def dropout_keep(seed, instances, unit0, n_units, p):
    """torch restatement of clam_keep() in csrc/hb_clam.cu for chosen instances: [len(instances), n_units] fp32 holding 0 or
    1 / (1 - p) for units unit0 .. unit0 + n_units - 1 (units: [0,L1) ReLU outputs, [L1,L1+D) gate branch a, then branch b).
    `instances`: int64 tensor of instance indices inside the bag (any device)."""
    inst = instances.to(torch.int64).view(-1, 1)
    if p <= 0.0:
        return torch.ones((inst.shape[0], n_units), dtype=torch.float32, device=inst.device)
    if p >= 1.0:
        return torch.zeros((inst.shape[0], n_units), dtype=torch.float32, device=inst.device)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    lo, hi = seed & _M32, seed >> 32
    thresh = min(max(int(float(torch.tensor(p, dtype=torch.float32)) * 4294967296.0 + 0.5), 1), 4294967295)
    unit = torch.arange(unit0, unit0 + n_units, dtype=torch.int64, device=inst.device).view(1, -1)
    a = _mix32((_mul32(inst, 0x9E3779B1) + lo) & _M32)
    h = _mix32(a ^ ((_mul32(unit, 0x7FEB352D) + hi) & _M32))
    scale = float(1.0 / (torch.tensor(1.0, dtype=torch.float32) - torch.tensor(p, dtype=torch.float32)))
    return (h >= thresh).to(torch.float32) * scale


def dropout_masks(n_instances, L1, D, p, seed):
    """(m1 [N,L1], ma [N,D], mb [N,D]) exactly as the kernels apply them — from the C library's own host routine
    (hb_clam_dropout_masks); parity tests multiply them into the reference module's activations."""
    lib = _lib.load()
    m1 = torch.empty((n_instances, L1), dtype=torch.float32)
    ma = torch.empty((n_instances, D), dtype=torch.float32)
    mb = torch.empty((n_instances, D), dtype=torch.float32)
    _lib.check(lib.hb_clam_dropout_masks(n_instances, L1, D, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, m1.data_ptr(), ma.data_ptr(),
                                         mb.data_ptr()))
    return m1, ma, mb


def dropout_p(model):
    """Dropout probability of a CLAM_SB as nn.Dropout will apply it (a YAML `drop_out: true` arrives as p = True = 1.0)."""
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            return float(m.p)
    return 0.0


def draw_seed():
    """A fresh 62-bit dropout seed from torch's CPU generator (torch.manual_seed makes a training run reproducible)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def forward_bags(models, feats, bag_offsets, max_bag_len=None, want=("logits", "y_prob", "y_hat", "m")):
    """models: list of CLAM_SB (same size_arg / n_classes); feats [total, L0] fp32 CUDA; bag_offsets int32 [n_bags+1]
    (CUDA or CPU).  Returns dict with a_raw [n_models, total] and the requested per-bag outputs [n_models, n_bags, ...]."""
    _lib.require_cuda(feats, "features")
    lib = _lib.load()
    dev = feats.device
    if feats.dtype != torch.float32 or not feats.is_contiguous():
        feats = feats.float().contiguous()
    total, L0 = feats.shape
    n_models = len(models)
    m0 = models[0]
    L1 = m0.attention_net[0].out_features
    D = _gate_module(m0).attention_c.in_features
    Cc = m0.classifiers.out_features
    if _gate_module(m0).attention_c.out_features != 1:
        raise RuntimeError("the fused kernel implements the single-branch (CLAM_SB) head")
    if max_bag_len is None:
        off_cpu = bag_offsets.cpu()
        lens = off_cpu[1:] - off_cpu[:-1]
        max_bag_len = int(lens.max().item()) if lens.numel() else 0
    offs = bag_offsets.to(device=dev, dtype=torch.int32).contiguous()
    n_bags = offs.numel() - 1
    with torch.cuda.device(dev):
        _lib.device_check()
        keep = []
        ptrs = []
        for m in models:
            w = _weights(m, dev)
            keep.append(w)
            ptrs += [t.data_ptr() for t in w]
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        a_raw = torch.empty((n_models, total), dtype=torch.float32, device=dev)
        m_out = torch.empty((n_models, n_bags, L1), dtype=torch.float32, device=dev) if "m" in want else None
        logits = torch.empty((n_models, n_bags, Cc), dtype=torch.float32, device=dev) if "logits" in want else None
        y_prob = torch.empty((n_models, n_bags, Cc), dtype=torch.float32, device=dev) if "y_prob" in want else None
        y_hat = torch.empty((n_models, n_bags), dtype=torch.int64, device=dev) if "y_hat" in want else None
        ws_bytes = lib.hb_clam_workspace_bytes(max_bag_len, n_bags, n_models, L1)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        _lib.check(lib.hb_clam_sb_forward(_lib.ptr(feats), _lib.ptr(offs), n_bags, total, max_bag_len, arr, n_models,
                                          L0, L1, D, Cc, _lib.ptr(a_raw), _lib.ptr(m_out), _lib.ptr(logits),
                                          _lib.ptr(y_prob), _lib.ptr(y_hat), _lib.ptr(ws), ws.numel(),
                                          _lib.stream_ptr()))
    return {"a_raw": a_raw, "m": m_out, "logits": logits, "y_prob": y_prob, "y_hat": y_hat}


def forward_single(model, h, attention_only=False):
    """One bag through one model with the reference's return shapes: logits [1,C], Y_prob [1,C], Y_hat [1,1] int64,
    A_raw [1,N], M [1,L1]  (A_raw alone when attention_only)."""
    N = h.shape[0]
    offs = torch.tensor([0, N], dtype=torch.int32)
    want = () if attention_only else ("logits", "y_prob", "y_hat", "m")
    r = forward_bags([model], h, offs, max_bag_len=N, want=want)
    if attention_only:
        return r["a_raw"]
    return r["logits"][0], r["y_prob"][0], r["y_hat"][0].view(1, 1), r["a_raw"], r["m"][0]


# ------------------------------------------------------------------------------------------------------ training step
def supports_fused_backward(model):
    """The fused backward covers the HIPT sizes: 192-d features, L1 <= 128 (hb_clam_sb_backward)."""
    fc = model.attention_net[0]
    g = _gate_module(model)
    D = g.attention_c.in_features
    return (fc.in_features == 192 and fc.out_features <= 128 and fc.out_features % 8 == 0 and D % 4 == 0
            and g.attention_c.out_features == 1)


def _param_list(model):
    g = _gate_module(model)
    return [model.attention_net[0].weight, model.attention_net[0].bias,
            g.attention_a[0].weight, g.attention_a[0].bias, g.attention_b[0].weight, g.attention_b[0].bias,
            g.attention_c.weight, g.attention_c.bias, model.classifiers.weight, model.classifiers.bias]


class ClamSBFunction(torch.autograd.Function):
    """CLAM_SB.forward for one bag as a differentiable op: forward = hb_clam_sb_forward, backward = hb_clam_sb_backward
    (recomputation: only A_raw and M are kept).  Differentiable outputs: logits [1,C], Y_prob [1,C], A_raw [1,N], M [1,L1];
    gradients flow to the 10 weight tensors (not to the bag: features are frozen HIPT_4K embeddings)."""

    @staticmethod
    def forward(ctx, h, drop_p, drop_seed, *params):
        lib = _lib.load()
        dev = h.device
        if h.dtype != torch.float32 or not h.is_contiguous():
            h = h.float().contiguous()
        N, L0 = h.shape
        w = [p.detach().contiguous() for p in params]
        L1, D, Cc = w[0].shape[0], w[2].shape[0], w[8].shape[0]
        offs = torch.tensor([0, N], dtype=torch.int32).to(dev)
        with torch.cuda.device(dev):
            arr = (C.c_void_p * 10)(*[t.data_ptr() for t in w])
            a_raw = torch.empty((1, N), dtype=torch.float32, device=dev)
            m_out = torch.empty((1, L1), dtype=torch.float32, device=dev)
            logits = torch.empty((1, Cc), dtype=torch.float32, device=dev)
            y_prob = torch.empty((1, Cc), dtype=torch.float32, device=dev)
            ws_bytes = lib.hb_clam_workspace_bytes(N, 1, 1, L1)
            ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
            _lib.check(lib.hb_clam_sb_forward_train(_lib.ptr(h), _lib.ptr(offs), 1, N, N, arr, 1, L0, L1, D, Cc, _lib.ptr(a_raw),
                                                    _lib.ptr(m_out), _lib.ptr(logits), _lib.ptr(y_prob), None, _lib.ptr(ws),
                                                    ws.numel(), float(drop_p), int(drop_seed), _lib.stream_ptr()))
        ctx.save_for_backward(h, a_raw, m_out, y_prob, *w)
        ctx.dims = (N, L0, L1, D, Cc)
        ctx.drop = (float(drop_p), int(drop_seed))
        return logits, y_prob, a_raw, m_out

    @staticmethod
    def backward(ctx, g_logits, g_yprob, g_araw, g_m):
        lib = _lib.load()
        h, a_raw, m_out, y_prob, *w = ctx.saved_tensors
        N, L0, L1, D, Cc = ctx.dims
        dev = h.device
        dl = torch.zeros((1, Cc), dtype=torch.float32, device=dev) if g_logits is None else g_logits.float()
        if g_yprob is not None:                                    # softmax backward on the [1, C] logits
            dl = dl + y_prob * (g_yprob - (g_yprob * y_prob).sum(dim=1, keepdim=True))
        dl = dl.contiguous()
        gm = None if g_m is None else g_m.float().contiguous()
        ga = None if g_araw is None else g_araw.float().contiguous()
        with torch.cuda.device(dev):
            grads = [torch.empty_like(t) for t in w]
            warr = (C.c_void_p * 10)(*[t.data_ptr() for t in w])
            garr = (C.c_void_p * 10)(*[t.data_ptr() for t in grads])
            ws = torch.empty(4 + L1, dtype=torch.float32, device=dev)
            _lib.check(lib.hb_clam_sb_backward_train(_lib.ptr(h), N, warr, _lib.ptr(a_raw), _lib.ptr(m_out), _lib.ptr(dl),
                                                     _lib.ptr(gm), _lib.ptr(ga), None, None, None, garr, L0, L1, D, Cc,
                                                     _lib.ptr(ws), ws.numel() * 4, ctx.drop[0], ctx.drop[1], _lib.stream_ptr()))
        return (None, None, None, *grads)


def forward_single_autograd(model, h, drop_p=0.0, drop_seed=0):
    """(logits, Y_prob, Y_hat, A_raw, M) with autograd through the fused kernels (training: main.py -> train_loop).
    drop_p > 0: training-mode dropout with the masks of `drop_seed` in the forward and the backward."""
    logits, y_prob, a_raw, m = ClamSBFunction.apply(h, float(drop_p), int(drop_seed), *_param_list(model))
    y_hat = torch.topk(logits, 1, dim=1)[1]
    return logits, y_prob, y_hat, a_raw, m


def h1_rows(model, h, idx, drop_p=0.0, drop_seed=0):
    """Rows `idx` of the [N, L1] instance features attention_net[0..2] produces (Linear, ReLU, Dropout) — the only part of `h`
    the instance-clustering branch reads (inst_eval / inst_eval_out index_select 2 k_sample rows, model_clam.py:116-145).
    A [len(idx), 192] x [192, L1] product in torch ops, differentiable w.r.t. fc.weight / fc.bias; with active dropout the
    rows carry the same keep mask the fused kernels applied."""
    fc = model.attention_net[0]
    rows = torch.relu(torch.nn.functional.linear(h.index_select(0, idx), fc.weight, fc.bias))
    if drop_p > 0.0:
        rows = rows * dropout_keep(drop_seed, idx, 0, fc.out_features, drop_p)
    return rows


class TrainStep:
    """One optimisation step of train_loop (utils/core_utils.py:409-425: model(data) -> CrossEntropyLoss -> backward ->
    optimizer.step -> zero_grad) for a CLAM_SB on the fused kernels with nothing else in between: forward (work table, scores,
    combine), backward with the loss fused in (hb_clam_sb_backward_ce), Adam (FusedAdam: one launch).  Buffers are allocated
    once; `step(bag, label)` returns the loss as a device scalar (no host synchronisation).  The autograd route
    (CLAM_SB.forward + loss.backward()) gives the same gradients; this one removes its per-step Python / autograd overhead."""

    def __init__(self, model, optimizer, max_instances, seed=0):
        if not supports_fused_backward(model):
            raise RuntimeError("TrainStep covers the HIPT head sizes (192-d features, L1 <= 128)")
        self.model, self.opt = model, optimizer
        self.params = _param_list(model)
        if not all(p.requires_grad for p in self.params):
            raise RuntimeError("TrainStep updates all 10 CLAM_SB tensors: a frozen parameter needs the autograd route")
        self.drop_p = dropout_p(model)                       # nn.Dropout p of the module (0.85 in the reference's final config)
        self.seed, self.n_steps = int(seed), 0
        dev = self.params[0].device
        self.dev = dev
        self.L0, self.L1 = 192, self.params[0].shape[0]
        self.D, self.C = self.params[2].shape[0], self.params[8].shape[0]
        self.maxn = int(max_instances)
        lib = _lib.load()
        self.lib = lib
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        self.a_raw, self.m_out, self.logits, self.y_prob = f32(1, self.maxn), f32(1, self.L1), f32(1, self.C), f32(1, self.C)
        self.loss = torch.empty((), dtype=torch.float32, device=dev)
        self.offs = torch.zeros(2, dtype=torch.int32, device=dev)
        self.ws_f = torch.empty(max(lib.hb_clam_workspace_bytes(self.maxn, 1, 1, self.L1), 16), dtype=torch.uint8, device=dev)
        self.ws_b = f32(4 + self.L1)
        # the gradient buffers are OWNED here: optimizer.zero_grad(set_to_none=True) or a reassigned .grad cannot leave the
        # backward kernel writing into freed memory; step() re-attaches them before the optimizer runs
        self.grads = [torch.zeros_like(p) for p in self.params]
        self.warr = (C.c_void_p * 10)(*[p.data_ptr() for p in self.params])
        self.garr = (C.c_void_p * 10)(*[g.data_ptr() for g in self.grads])

    @torch.no_grad()
    def step(self, bag, label):
        """bag [N, 192] fp32 CUDA, label int64 CUDA tensor with one element."""
        N = bag.shape[0]
        if N > self.maxn or N < 1:
            raise RuntimeError(f"bag of {N} instances outside 1..{self.maxn}")
        if bag.dtype != torch.float32 or not bag.is_contiguous():
            bag = bag.float().contiguous()
        lib, st = self.lib, _lib.stream_ptr()
        drop = self.drop_p if self.model.training else 0.0
        seed = (self.seed + 0x9E3779B97F4A7C15 * (self.n_steps + 1)) & 0x3FFFFFFFFFFFFFFF      # a new mask every step
        self.n_steps += 1
        for p, w in zip(self.params, (self.warr[i] for i in range(10))):
            if p.data_ptr() != w:
                raise RuntimeError("a CLAM_SB parameter was re-allocated after TrainStep was built")
        with torch.cuda.device(self.dev):
            self.offs[1] = N
            _lib.check(lib.hb_clam_sb_forward_train(_lib.ptr(bag), _lib.ptr(self.offs), 1, N, N, self.warr, 1, self.L0, self.L1,
                                                    self.D, self.C, _lib.ptr(self.a_raw), _lib.ptr(self.m_out), _lib.ptr(self.logits),
                                                    _lib.ptr(self.y_prob), None, _lib.ptr(self.ws_f), self.ws_f.numel(), drop, seed, st))
            _lib.check(lib.hb_clam_sb_backward_train(_lib.ptr(bag), N, self.warr, _lib.ptr(self.a_raw), _lib.ptr(self.m_out), None,
                                                     None, None, _lib.ptr(self.logits), _lib.ptr(label), _lib.ptr(self.loss),
                                                     self.garr, self.L0, self.L1, self.D, self.C, _lib.ptr(self.ws_b),
                                                     self.ws_b.numel() * 4, drop, seed, st))
        for p, g in zip(self.params, self.grads):
            if p.grad is not g:
                p.grad = g
        self.last_seed = seed
        self.opt.step()
        return self.loss


class TrialBatchStep:
    """up to 8 independent training trials of one CLAM_SB head size advance one step of train_loop
    each (utils/core_utils.py:409-425: model(bag) -> CrossEntropyLoss -> backward -> Adam step) in SIX launches in total
    (hb_clam_sb_train_step_trials) — the reference runs such trials as separate Ray Tune processes sharing a GPU
    (main.py:40-52), ~60 launches per trial and step.  Trial t has its own model, Adam state (lr, weight decay, betas shared),
    dropout seed and bag; semantics per trial are those of TrainStep + FusedAdam (same kernels' arithmetic).

    models: CLAM_SB modules of one size_arg / n_classes / dropout on one CUDA device, all 10 tensors trainable.
    step(bags, labels): bags = list of [n_t, 192] fp32 CUDA tensors (one per trial), labels int64 CUDA [T]; returns the
    losses [T] as a device tensor (no host synchronisation).  Adam state lives in self.exp_avg / self.exp_avg_sq."""

    def __init__(self, models, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, max_instances=20000, seeds=None):
        self.models = list(models)
        T = len(self.models)
        if not 1 <= T <= 8:
            raise RuntimeError("TrialBatchStep takes 1..8 trials per launch")
        if not all(supports_fused_backward(m) for m in self.models):
            raise RuntimeError("TrialBatchStep covers the HIPT head sizes (192-d features, L1 <= 128)")
        self.params = [_param_list(m) for m in self.models]
        if not all(p.requires_grad for ps in self.params for p in ps):
            raise RuntimeError("TrialBatchStep updates all 10 CLAM_SB tensors of every trial")
        shapes = [tuple(p.shape) for p in self.params[0]]
        if any([tuple(p.shape) for p in ps] != shapes for ps in self.params):
            raise RuntimeError("the trials of one launch share one head size and class count")
        self.drop_p = dropout_p(self.models[0])
        if any(dropout_p(m) != self.drop_p for m in self.models):
            raise RuntimeError("the trials of one launch share one dropout probability")
        self.T = T
        self.dev = self.params[0][0].device
        self.L0, self.L1 = 192, shapes[0][0]
        self.D, self.C = shapes[2][0], shapes[8][0]
        as_list = lambda v: [float(x) for x in v] if isinstance(v, (list, tuple)) else [float(v)] * T
        self.lr, self.wd = as_list(lr), as_list(weight_decay)
        self.betas, self.eps = betas, float(eps)
        self.seeds = [int(s) for s in (seeds if seeds is not None else range(T))]
        self.n_steps = [0] * T
        self.maxn = int(max_instances)
        self.lib = _lib.load()
        dev = self.dev
        self.grads = [[torch.zeros_like(p) for p in ps] for ps in self.params]
        self.exp_avg = [[torch.zeros_like(p) for p in ps] for ps in self.params]
        self.exp_avg_sq = [[torch.zeros_like(p) for p in ps] for ps in self.params]
        flat = lambda groups: (C.c_void_p * (10 * T))(*[t.data_ptr() for g in groups for t in g])
        self.warr, self.garr, self.marr, self.varr = flat(self.params), flat(self.grads), flat(self.exp_avg), flat(self.exp_avg_sq)
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        self.a_raw, self.m_out, self.logits, self.loss = f32(T * self.maxn), f32(T, self.L1), f32(T, self.C), f32(T)
        self.ws = torch.empty(max(self.lib.hb_clam_trials_workspace_bytes(self.maxn, T, self.L1), 16), dtype=torch.uint8, device=dev)
        self.last_seeds = None

    @torch.no_grad()
    def step(self, bags, labels):
        T = self.T
        assert len(bags) == T and labels.numel() == T and labels.dtype == torch.int64 and labels.is_cuda
        lens = [int(b.shape[0]) for b in bags]
        if min(lens) < 1 or max(lens) > self.maxn:
            raise RuntimeError(f"bag sizes {lens} outside 1..{self.maxn}")
        feats = bags[0] if T == 1 else torch.cat([b.float() for b in bags], dim=0)
        if feats.dtype != torch.float32 or not feats.is_contiguous():
            feats = feats.float().contiguous()
        offs_h = [0]
        for n in lens:
            offs_h.append(offs_h[-1] + n)
        offs_host = (C.c_int32 * (T + 1))(*offs_h)
        offs_dev = torch.tensor(offs_h, dtype=torch.int32).to(self.dev, non_blocking=True)
        training = self.models[0].training
        seeds = []
        for t in range(T):
            self.n_steps[t] += 1
            seeds.append((self.seeds[t] + 0x9E3779B97F4A7C15 * self.n_steps[t]) & 0x3FFFFFFFFFFFFFFF)   # TrainStep's schedule
        self.last_seeds = seeds
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.hb_clam_sb_train_step_trials(
                _lib.ptr(feats), _lib.ptr(offs_dev), offs_host, T, self.warr, self.garr, self.marr, self.varr, _lib.ptr(labels),
                (C.c_float * T)(*self.lr), (C.c_float * T)(*self.wd), (C.c_int * T)(*self.n_steps), float(self.betas[0]),
                float(self.betas[1]), self.eps, self.drop_p if training else 0.0, (C.c_uint64 * T)(*seeds), _lib.ptr(self.a_raw),
                _lib.ptr(self.m_out), _lib.ptr(self.logits), _lib.ptr(self.loss), self.L0, self.L1, self.D, self.C,
                _lib.ptr(self.ws), self.ws.numel(), _lib.stream_ptr()))
        return self.loss


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) semantics with every fp32 CUDA tensor of a group updated by
    ONE hb_adam_step launch (the reference's get_optim builds optim.Adam(lr=args.lr, weight_decay=args.reg),
    utils/utils.py:100-107; a CLAM_SB step is 14 tiny tensors -> 14+ launches there)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _launch(self, lib, group, chunk, step):
        n = len(chunk)
        grads = [p.grad.contiguous() for p in chunk]
        parr = (C.c_void_p * n)(*[p.data_ptr() for p in chunk])
        garr = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
        marr = (C.c_void_p * n)(*[self.state[p]["exp_avg"].data_ptr() for p in chunk])
        varr = (C.c_void_p * n)(*[self.state[p]["exp_avg_sq"].data_ptr() for p in chunk])
        narr = (C.c_int * n)(*[p.numel() for p in chunk])
        with torch.cuda.device(chunk[0].device):
            _lib.check(lib.hb_adam_step(parr, garr, marr, varr, narr, n, float(group["lr"]), float(group["betas"][0]),
                                        float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]),
                                        step, _lib.stream_ptr()))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdam updates contiguous fp32 CUDA parameters only (there is no CPU path)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
            # torch.optim.Adam keeps one step count PER PARAMETER (a tensor whose grad was None on some steps lags behind):
            # tensors are bucketed by their own count, one launch per bucket of up to 16
            buckets = {}
            for p in ps:
                buckets.setdefault(self.state[p]["step"] + 1, []).append(p)
            for step, bps in sorted(buckets.items()):
                for i in range(0, len(bps), 16):
                    self._launch(lib, group, bps[i:i + 16], step)
            for p in ps:
                self.state[p]["step"] += 1
        return loss
