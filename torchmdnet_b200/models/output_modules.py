"""Output heads on the hot path: Scalar (TensorNet) and EquivariantScalar (ET).

API and state-dict keys follow the reference's `torchmdnet/models/output_modules.py:22-163`.  The other
reference heads (DipoleMoment, ElectronicSpatialExtent, VectorOutput, ScalarPlusWeightedCoulomb) are
outside the energy+force hot path and are not provided.
"""
from typing import Optional
from warnings import warn

import torch
from torch import nn

from .utils import MLP, GatedEquivariantBlock, scatter

__all__ = ["Scalar", "EquivariantScalar"]


class OutputModel(nn.Module):
    def __init__(self, allow_prior_model, reduce_op, static_shapes=False):
        super().__init__()
        self.allow_prior_model, self.reduce_op, self.static_shapes = allow_prior_model, reduce_op, static_shapes
        self.dim_size = 0

    def reset_parameters(self):
        pass

    def pre_reduce(self, x, v, z, pos, batch, box: Optional[torch.Tensor] = None):
        raise NotImplementedError

    def reduce(self, x, batch, num_systems: Optional[int] = None):
        """Per-molecule reduction.  `num_systems` avoids the host read of batch.max(); without it the size is
        read once per call outside CUDA-graph capture and frozen inside capture (reference :43-73)."""
        if num_systems is None:
            capturing = x.is_cuda and torch.cuda.is_current_stream_capturing()
            if not capturing:
                self.dim_size = int(batch.max().item()) + 1
            else:
                assert self.dim_size > 0, "Warming up is needed before capturing the model into a CUDA graph"
                warn(f"CUDA graph capture will lock the batch to the current number of samples ({self.dim_size}). "
                     "Changing this will result in a crash")
            num_systems = self.dim_size
        return scatter(x, batch, dim=0, dim_size=num_systems, reduce=self.reduce_op)

    def post_reduce(self, x):
        return x


class Scalar(OutputModel):
    def __init__(self, hidden_channels, activation="silu", allow_prior_model=True, reduce_op="sum", dtype=torch.float,
                 static_shapes=False, **kwargs):
        super().__init__(allow_prior_model=allow_prior_model, reduce_op=reduce_op, static_shapes=static_shapes)
        self.output_network = MLP(hidden_channels, 1, hidden_channels // 2, activation,
                                  num_hidden_layers=kwargs.get("num_layers", 0), dtype=dtype)
        self.reset_parameters()

    def reset_parameters(self):
        self.output_network.reset_parameters()

    def pre_reduce(self, x, v: Optional[torch.Tensor], z, pos, batch, box: Optional[torch.Tensor] = None):
        return self.output_network(x)


class EquivariantScalar(OutputModel):
    def __init__(self, hidden_channels, activation="silu", allow_prior_model=True, reduce_op="sum", dtype=torch.float,
                 static_shapes=False, **kwargs):
        super().__init__(allow_prior_model=allow_prior_model, reduce_op=reduce_op, static_shapes=static_shapes)
        if kwargs.get("num_layers", 0) > 0:
            warn("num_layers is not used in EquivariantScalar")
        self.output_network = nn.ModuleList([
            GatedEquivariantBlock(hidden_channels, hidden_channels // 2, activation=activation, scalar_activation=True, dtype=dtype),
            GatedEquivariantBlock(hidden_channels // 2, 1, activation=activation, dtype=dtype)])
        self.reset_parameters()

    def reset_parameters(self):
        for layer in self.output_network:
            layer.reset_parameters()

    def pre_reduce(self, x, v, z, pos, batch, box: Optional[torch.Tensor] = None):
        for layer in self.output_network:
            x, v = layer(x, v)
        return x + v.sum() * 0
