"""CLAM attention-MIL heads with the reference's constructors and state_dict keys (models/model_clam.py).

CLAM_SB with gated attention — the model the HIPT-ABMIL pipeline trains and evaluates — runs its inference forward
(`model(h)`, `model(h, attention_only=True)`, `return_features=True`) through the fused ragged-bag CUDA kernel in
csrc/hb_clam.cu, and its training step (autograd enabled, HIPT sizes, any dropout probability: the reference's final
model trains at 0.85) through the same forward plus the fused recomputing backward (clam_engine.ClamSBFunction) with the
dropout masks regenerated inside both kernels.  The instance-clustering branch (`instance_eval=True`) runs on top of the
fused outputs: it needs the scores and 2 k_sample rows of the instance features.  Other feature sizes in training,
gradients w.r.t. the bag, CLAM_MB and the ungated Attn_Net are composed from torch ops on the same device (kept
constructible for checkpoint compatibility, out of scope: SURVEY.md §2.1 #5).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import clam_engine

SIZE_DICT = {"tinier3": [1024, 32, 8], "256": [256, 64, 16], "tinier_resnet18": [512, 64, 16],
             "tinier2_resnet18": [512, 32, 8], "tiny_resnet18": [512, 128, 32], "small_resnet18": [512, 256, 64],
             "tinier": [1024, 64, 16], "tiny128": [1024, 128, 32], "tiny": [1024, 256, 64], "small": [1024, 512, 256],
             "big": [1024, 512, 384], "hipt_big": [192, 128, 64], "hipt_medium": [192, 64, 32],
             "hipt_small": [192, 32, 16], "hipt_smaller": [192, 16, 8], "hipt_smallest": [192, 8, 4]}


def initialize_weights(module):
    """xavier_normal_ weights / zero bias for every Linear (utils/utils.py:217-225)."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_normal_(m.weight)
            m.bias.data.zero_()
        elif isinstance(m, nn.BatchNorm1d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


def _branch(L, D, act, dropout):
    layers = [nn.Linear(L, D), act]
    if dropout > 0:
        layers.append(nn.Dropout(dropout))
    return nn.Sequential(*layers)


class Attn_Net(nn.Module):
    """Ungated attention (two Linear layers); returns (A [N, n_classes], x)."""

    def __init__(self, L=1024, D=256, dropout=0.25, n_classes=1):
        super().__init__()
        layers = [nn.Linear(L, D), nn.Tanh()]
        if dropout > 0:
            layers.append(nn.Dropout(dropout))
        layers.append(nn.Linear(D, n_classes))
        self.module = nn.Sequential(*layers)

    def forward(self, x):
        return self.module(x), x


class Attn_Net_Gated(nn.Module):
    """A = Wc (tanh(Wa x + ba) * sigmoid(Wb x + bb)) + bc; returns (A [N, n_classes], x)."""

    def __init__(self, L=1024, D=256, dropout=0.0, n_classes=1):
        super().__init__()
        self.attention_a = _branch(L, D, nn.Tanh(), dropout)
        self.attention_b = _branch(L, D, nn.Sigmoid(), dropout)
        self.attention_c = nn.Linear(D, n_classes)

    def forward(self, x):
        return self.attention_c(self.attention_a(x).mul(self.attention_b(x))), x


class CLAM_SB(nn.Module):
    def __init__(self, gate=True, size_arg="small", dropout=0.0, k_sample=8, n_classes=2,
                 instance_loss_fn=nn.CrossEntropyLoss(), subtyping=False):
        super().__init__()
        self.size_dict = dict(SIZE_DICT)
        size = self.size_dict[size_arg]
        fc = [nn.Linear(size[0], size[1]), nn.ReLU()]
        if dropout > 0:
            fc.append(nn.Dropout(dropout))
        fc.append((Attn_Net_Gated if gate else Attn_Net)(L=size[1], D=size[2], dropout=dropout, n_classes=1))
        self.attention_net = nn.Sequential(*fc)
        self.classifiers = nn.Linear(size[1], n_classes)
        self.instance_classifiers = nn.ModuleList([nn.Linear(size[1], 2) for _ in range(n_classes)])
        self.k_sample = k_sample
        self.instance_loss_fn = instance_loss_fn
        self.n_classes = n_classes
        self.subtyping = subtyping
        self.gate = gate
        initialize_weights(self)

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.attention_net = self.attention_net.to(device)
        self.classifiers = self.classifiers.to(device)
        self.instance_classifiers = self.instance_classifiers.to(device)

    @staticmethod
    def create_positive_targets(length, device):
        return torch.full((length,), 1, device=device).long()

    @staticmethod
    def create_negative_targets(length, device):
        return torch.full((length,), 0, device=device).long()

    def inst_eval(self, A, h, classifier):
        """In-the-class branch: top-k and bottom-k attended instances as positive / negative (model_clam.py:116-132)."""
        if A.dim() == 1:
            A = A.view(1, -1)
        k = self.k_sample
        top_p = torch.index_select(h, 0, torch.topk(A, k)[1][-1])
        top_n = torch.index_select(h, 0, torch.topk(-A, k, dim=1)[1][-1])
        targets = torch.cat([self.create_positive_targets(k, h.device), self.create_negative_targets(k, h.device)])
        logits = classifier(torch.cat([top_p, top_n], dim=0))
        preds = torch.topk(logits, 1, dim=1)[1].squeeze(1)
        return self.instance_loss_fn(logits, targets), preds, targets

    def inst_eval_out(self, A, h, classifier):
        """Out-of-the-class branch (subtyping): top-k instances are negatives (model_clam.py:135-145)."""
        if A.dim() == 1:
            A = A.view(1, -1)
        k = self.k_sample
        top_p = torch.index_select(h, 0, torch.topk(A, k)[1][-1])
        targets = self.create_negative_targets(k, h.device)
        logits = classifier(top_p)
        preds = torch.topk(logits, 1, dim=1)[1].squeeze(1)
        return self.instance_loss_fn(logits, targets), preds, targets

    # ------------------------------------------------------------------------------------------------ forward
    def _needs_autograd(self, h):
        return torch.is_grad_enabled() and (h.requires_grad or any(p.requires_grad for p in self.parameters()))

    def forward(self, h, label=None, instance_eval=False, return_features=False, attention_only=False):
        if not h.is_cuda:
            raise RuntimeError("CLAM_SB runs only on CUDA (B200): there is no CPU path; move the bag to the GPU")
        drop = clam_engine.dropout_p(self) if self.training else 0.0      # nn.Dropout is the identity in eval()
        grad = self._needs_autograd(h)
        if not self.gate or h.requires_grad or (grad or drop > 0.0) and not clam_engine.supports_fused_backward(self):
            # ungated attention, gradients w.r.t. the bag, or training of a non-HIPT head: plain torch composition
            return self._forward_autograd(h, label, instance_eval, return_features, attention_only)
        seed = clam_engine.draw_seed() if drop > 0.0 else 0
        if grad or drop > 0.0:
            # training step (utils/core_utils.py:409-423 / :300-371): forward and backward both run the fused ragged-bag
            # kernels, dropout masks (model_clam.py:84-85, :50-52) are regenerated from `seed` inside them
            logits, Y_prob, Y_hat, A_raw, M = clam_engine.forward_single_autograd(self, h, drop, seed)
        else:
            res = clam_engine.forward_single(self, h, attention_only=attention_only)
            if attention_only:
                return res
            logits, Y_prob, Y_hat, A_raw, M = res
        if attention_only:
            return A_raw
        results_dict = {}
        if instance_eval:
            results_dict = self._instance_eval_fused(h, A_raw, label, drop, seed)
        if return_features:
            results_dict.update({'features': M})
        return logits, Y_prob, Y_hat, A_raw, results_dict

    def _instance_eval_fused(self, h, A_raw, label, drop, seed):
        """The instance-clustering branch (model_clam.py:156-178) on top of the fused forward: it needs the softmaxed scores
        and 2 * k_sample ROWS of the [N, L1] instance features, which clam_engine.h1_rows recomputes for just those rows."""
        A = F.softmax(A_raw, dim=1)
        k = self.k_sample
        total_inst_loss = 0.0
        all_preds, all_targets = [], []
        inst_labels = F.one_hot(label, num_classes=self.n_classes).squeeze()
        for i, classifier in enumerate(self.instance_classifiers):
            if inst_labels[i].item() == 1:                                    # in-the-class: top-k positive, bottom-k negative
                top_p_ids = torch.topk(A, k)[1][-1]
                top_n_ids = torch.topk(-A, k, dim=1)[1][-1]
                rows = clam_engine.h1_rows(self, h, torch.cat([top_p_ids, top_n_ids]), drop, seed)
                targets = torch.cat([self.create_positive_targets(k, h.device), self.create_negative_targets(k, h.device)])
            elif self.subtyping:                                              # out-of-the-class: top-k are negatives
                rows = clam_engine.h1_rows(self, h, torch.topk(A, k)[1][-1], drop, seed)
                targets = self.create_negative_targets(k, h.device)
            else:
                continue
            logits = classifier(rows)
            preds = torch.topk(logits, 1, dim=1)[1].squeeze(1)
            all_preds.extend(preds.cpu().numpy())
            all_targets.extend(targets.cpu().numpy())
            total_inst_loss += self.instance_loss_fn(logits, targets)
        if self.subtyping:
            total_inst_loss /= len(self.instance_classifiers)
        return {'instance_loss': total_inst_loss, 'inst_labels': np.array(all_targets), 'inst_preds': np.array(all_preds)}

    def _has_active_dropout(self):
        return any(isinstance(m, nn.Dropout) and m.p > 0 for m in self.modules())

    def _forward_autograd(self, h, label, instance_eval, return_features, attention_only):
        """Differentiable composition used for training steps (model_clam.py:147-191)."""
        A, h = self.attention_net(h)
        A = torch.transpose(A, 1, 0)
        if attention_only:
            return A
        A_raw = A
        A = F.softmax(A, dim=1)
        results_dict = {}
        if instance_eval:
            total_inst_loss = 0.0
            all_preds, all_targets = [], []
            inst_labels = F.one_hot(label, num_classes=self.n_classes).squeeze()
            for i, classifier in enumerate(self.instance_classifiers):
                if inst_labels[i].item() == 1:
                    loss, preds, targets = self.inst_eval(A, h, classifier)
                elif self.subtyping:
                    loss, preds, targets = self.inst_eval_out(A, h, classifier)
                else:
                    continue
                all_preds.extend(preds.cpu().numpy())
                all_targets.extend(targets.cpu().numpy())
                total_inst_loss += loss
            if self.subtyping:
                total_inst_loss /= len(self.instance_classifiers)
            results_dict = {'instance_loss': total_inst_loss, 'inst_labels': np.array(all_targets),
                            'inst_preds': np.array(all_preds)}
        M = torch.mm(A, h)
        logits = self.classifiers(M)
        Y_hat = torch.topk(logits, 1, dim=1)[1]
        Y_prob = F.softmax(logits, dim=1)
        if return_features:
            results_dict.update({'features': M})
        return logits, Y_prob, Y_hat, A_raw, results_dict


class CLAM_MB(CLAM_SB):
    """Multi-branch CLAM (one attention branch per class).  Not on the HIPT-ABMIL hot path: torch ops only."""

    def __init__(self, gate=True, size_arg="small", dropout=0.0, k_sample=8, n_classes=2,
                 instance_loss_fn=nn.CrossEntropyLoss(), subtyping=False):
        nn.Module.__init__(self)
        self.size_dict = dict(SIZE_DICT)
        size = self.size_dict[size_arg]
        fc = [nn.Linear(size[0], size[1]), nn.ReLU()]
        if dropout > 0:
            fc.append(nn.Dropout(dropout))
        fc.append((Attn_Net_Gated if gate else Attn_Net)(L=size[1], D=size[2], dropout=dropout, n_classes=n_classes))
        self.attention_net = nn.Sequential(*fc)
        self.classifiers = nn.ModuleList([nn.Linear(size[1], 1) for _ in range(n_classes)])
        self.instance_classifiers = nn.ModuleList([nn.Linear(size[1], 2) for _ in range(n_classes)])
        self.k_sample = k_sample
        self.instance_loss_fn = instance_loss_fn
        self.n_classes = n_classes
        self.subtyping = subtyping
        self.gate = gate
        initialize_weights(self)

    def forward(self, h, label=None, instance_eval=False, return_features=False, attention_only=False):
        A, h = self.attention_net(h)
        A = torch.transpose(A, 1, 0)
        if attention_only:
            return A
        A_raw = A
        A = F.softmax(A, dim=1)
        results_dict = {}
        if instance_eval:
            total_inst_loss = 0.0
            all_preds, all_targets = [], []
            inst_labels = F.one_hot(label, num_classes=self.n_classes).squeeze()
            for i, classifier in enumerate(self.instance_classifiers):
                if inst_labels[i].item() == 1:
                    loss, preds, targets = self.inst_eval(A[i], h, classifier)
                elif self.subtyping:
                    loss, preds, targets = self.inst_eval_out(A[i], h, classifier)
                else:
                    continue
                all_preds.extend(preds.cpu().numpy())
                all_targets.extend(targets.cpu().numpy())
                total_inst_loss += loss
            if self.subtyping:
                total_inst_loss /= len(self.instance_classifiers)
            results_dict = {'instance_loss': total_inst_loss, 'inst_labels': np.array(all_targets),
                            'inst_preds': np.array(all_preds)}
        M = torch.mm(A, h)
        logits = torch.empty(1, self.n_classes, device=h.device, dtype=M.dtype)
        for c in range(self.n_classes):
            logits[0, c] = self.classifiers[c](M[c])
        Y_hat = torch.topk(logits, 1, dim=1)[1]
        Y_prob = F.softmax(logits, dim=1)
        if return_features:
            results_dict.update({'features': M})
        return logits, Y_prob, Y_hat, A_raw, results_dict
