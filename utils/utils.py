"""Shim of the reference's utils/utils.py.  With the reference checkout on sys.path this module IS the reference's own
(collate_features, get_split_loader, print_network, ... : executed into this namespace); standing alone it provides the two
functions the accelerated modules and the training loop need."""
import torch.optim as optim

from hipt_abmil_atec23_b200.shim import chain_load
from hipt_abmil_atec23_b200.model_clam import initialize_weights  # noqa: F401

_reference_file = chain_load(globals(), __import__("utils").__path__, "utils")

if _reference_file is None:
    def get_optim(model, args):
        """Adam / SGD over the trainable parameters (utils/utils.py:100-107 in the reference)."""
        params = filter(lambda p: p.requires_grad, model.parameters())
        if args.opt == "adam":
            return optim.Adam(params, lr=args.lr, weight_decay=args.reg)
        if args.opt == "sgd":
            return optim.SGD(params, lr=args.lr, momentum=0.9, weight_decay=args.reg)
        raise NotImplementedError
