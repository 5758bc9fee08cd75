"""Max-pooling MIL baselines with the reference's interface (models/model_mil.py).

`MIL_fc` is named in the hot-path contract only for API parity: its size table holds a single 1024-d entry
(model_mil.py:11), so it never receives 192-d HIPT bags.  It is a two-layer per-instance MLP followed by a top-1
selection; it is composed from torch ops on the input's device and has no dedicated kernel.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .model_clam import initialize_weights


def _mlp(size, dropout, n_out):
    layers = [nn.Linear(size[0], size[1]), nn.ReLU()]
    if dropout:
        layers.append(nn.Dropout(0.25))
    layers.append(nn.Linear(size[1], n_out))
    return nn.Sequential(*layers)


class MIL_fc(nn.Module):
    def __init__(self, gate=True, size_arg="small", dropout=False, n_classes=2, top_k=1):
        super().__init__()
        assert n_classes == 2
        self.size_dict = {"small": [1024, 512]}
        self.classifier = _mlp(self.size_dict[size_arg], dropout, n_classes)
        initialize_weights(self)
        self.top_k = top_k

    def relocate(self):
        self.classifier.to(torch.device("cuda" if torch.cuda.is_available() else "cpu"))

    def forward(self, h, return_features=False):
        """Returns (top_instance logits [1,2], Y_prob [1,2], Y_hat [1,1], y_probs [N,2], results_dict)
        (model_mil.py:26-43)."""
        net = getattr(self.classifier, "module", self.classifier)     # DataParallel-wrapped in the reference (:28-29)
        feats = net[:-1](h)
        logits = net[-1](feats)
        y_probs = F.softmax(logits, dim=1)
        idx = torch.topk(y_probs[:, 1], self.top_k, dim=0)[1].view(1,)
        top_instance = torch.index_select(logits, 0, idx)
        Y_hat = torch.topk(top_instance, 1, dim=1)[1]
        Y_prob = F.softmax(top_instance, dim=1)
        results = {'features': torch.index_select(feats, 0, idx)} if return_features else {}
        return top_instance, Y_prob, Y_hat, y_probs, results


class MIL_fc_mc(nn.Module):
    """Multi-class variant (model_mil.py:46-93): one shared MLP, top-1 instance over the flattened class scores."""

    def __init__(self, gate=True, size_arg="small", dropout=False, n_classes=2, top_k=1):
        super().__init__()
        assert n_classes > 2
        self.size_dict = {"small": [1024, 512]}
        size = self.size_dict[size_arg]
        fc = [nn.Linear(size[0], size[1]), nn.ReLU()]
        if dropout:
            fc.append(nn.Dropout(0.25))
        self.fc = nn.Sequential(*fc)
        self.classifiers = nn.Linear(size[1], n_classes)
        initialize_weights(self)
        self.top_k = top_k
        self.n_classes = n_classes
        assert self.top_k == 1

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.fc = self.fc.to(device)
        self.classifiers = self.classifiers.to(device)

    def forward(self, h, return_features=False):
        h = self.fc(h)
        logits = self.classifiers(h)
        y_probs = F.softmax(logits, dim=1)
        m = y_probs.view(1, -1).argmax(1)
        top_indices = torch.cat(((m // self.n_classes).view(-1, 1), (m % self.n_classes).view(-1, 1)), dim=1).view(-1, 1)
        top_instance = logits[top_indices[0]]
        Y_hat = top_indices[1]
        Y_prob = y_probs[top_indices[0]]
        results = {'features': torch.index_select(h, 0, top_indices[0])} if return_features else {}
        return top_instance, Y_prob, Y_hat, y_probs, results
