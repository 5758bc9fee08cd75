import torch.optim as optim

from hipt_abmil_atec23_b200.model_clam import initialize_weights  # noqa: F401


def get_optim(model, args):
    """Adam / SGD over the trainable parameters (utils/utils.py:100-107 in the reference)."""
    params = filter(lambda p: p.requires_grad, model.parameters())
    if args.opt == "adam":
        return optim.Adam(params, lr=args.lr, weight_decay=args.reg)
    if args.opt == "sgd":
        return optim.SGD(params, lr=args.lr, momentum=0.9, weight_decay=args.reg)
    raise NotImplementedError
