import torch

from _refshim_core import t


def erf(x):
    return torch.special.erf(t(x))
