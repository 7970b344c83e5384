"""Priors on the hot path: only `Atomref` (an embedding lookup added before the reduction,
reference `torchmdnet/priors/atomref.py`); ZBL / Coulomb / D2 / LearnableAtomref are out of scope
(SURVEY.md section 2 row 18)."""
from typing import Dict, Optional

import torch
from torch import Tensor, nn

__all__ = ["Atomref"]


class BasePrior(nn.Module):
    def __init__(self, dataset=None):
        super().__init__()

    def get_init_args(self):
        return {}

    def pre_reduce(self, x, z, pos, batch, extra_args: Optional[Dict[str, Tensor]]):
        return x

    def post_reduce(self, y, z, pos, batch, box: Optional[Tensor], extra_args: Optional[Dict[str, Tensor]]):
        return y


class Atomref(BasePrior):
    """x' = x + atomref(z)  (reference priors/atomref.py:12-95)."""

    def __init__(self, max_z=None, dataset=None, trainable=False, enable=True):
        super().__init__()
        if max_z is None and dataset is None:
            raise ValueError("Can't instantiate Atomref prior, all arguments are None.")
        if dataset is None:
            atomref = torch.zeros(max_z, 1)
        else:
            atomref = dataset.get_atomref()
            if atomref is None:
                atomref = torch.zeros(100, 1)
        if atomref.ndim == 1:
            atomref = atomref.view(-1, 1)
        self.register_buffer("initial_atomref", atomref)
        self.atomref = nn.Embedding(len(atomref), 1, _freeze=not trainable, _weight=atomref)
        self.enable = enable

    def reset_parameters(self):
        self.atomref.weight.data.copy_(self.initial_atomref)

    def get_init_args(self):
        return dict(max_z=self.initial_atomref.size(0), trainable=self.atomref.weight.requires_grad, enable=self.enable)

    def pre_reduce(self, x, z, pos, batch=None, extra_args=None):
        return x + self.atomref(z) if self.enable else x
