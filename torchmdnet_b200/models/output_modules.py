"""Output heads on the hot path: Scalar (TensorNet) and EquivariantScalar (ET).

API and state-dict keys follow the reference's `torchmdnet/models/output_modules.py:22-163`.  ScalarPlusWeightedCoulomb
(reference :344-609) is the head of TensorNet2 (host structure + torch ops, GPU parity of the assembled
model pending -- see models/tensornet2.py).  The other reference heads (DipoleMoment, ElectronicSpatialExtent, VectorOutput) are
outside the energy+force hot path and are not provided.
"""
import math
from typing import Optional
from warnings import warn

import torch
from torch import nn

from .utils import MLP, GatedEquivariantBlock, OptimizedDistance, scatter

__all__ = ["Scalar", "EquivariantScalar", "ScalarPlusWeightedCoulomb"]


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
        return scatter(x, batch, dim=0, dim_size=self.resolve_num_systems(batch, num_systems), reduce=self.reduce_op)

    def resolve_num_systems(self, batch, num_systems: Optional[int] = None) -> int:
        if num_systems is None:
            capturing = batch.is_cuda and torch.cuda.is_current_stream_capturing()
            if not capturing:
                self.dim_size = int(batch.max().item()) + 1
            else:
                assert self.dim_size > 0, "Warming up is needed before capturing the model into a CUDA graph"
                warn(f"CUDA graph capture will lock the batch to the current number of samples ({self.dim_size}). "
                     "Changing this will result in a crash")
            num_systems = self.dim_size
        return int(num_systems)

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


def _exp_cutoff(d, rc: float):
    """short-range damping of the Coulomb term (reference output_modules.py:321-330, from aimnet)"""
    return torch.exp(-1.0 / (1.0 - (d / rc).clamp(0, 1.0 - 1e-6).pow(2))) / 0.36787944117144233


def _triu_indices(x, batch):
    """all i < j pairs inside a molecule (reference :333-341)"""
    N = x.shape[0]
    i, j = torch.triu_indices(N, N, 1, device=x.device, dtype=torch.long).unbind(0)
    mask = batch[i] == batch[j]
    return torch.stack([i[mask], j[mask]])


class ScalarPlusWeightedCoulomb(OutputModel):
    """Scalar head + damped Coulomb energy of the charge channels predicted by TensorNet2, weighted mean over the channels
    (reference output_modules.py:344-609).  Input features: [N, hidden_channels + (num_layers + 1) q_dim].
    coulomb_cutoff None: all pairs of a molecule (no PBC); else a half neighbour list inside the cutoff with the
    reaction-field form (works with PBC when the box is larger than twice the cutoff)."""

    def __init__(self, hidden_channels, activation="silu", allow_prior_model=True, reduce_op="sum", dtype=torch.float,
                 static_shapes=False, cutoff=None, **kwargs):
        super().__init__(allow_prior_model=allow_prior_model, reduce_op=reduce_op, static_shapes=static_shapes)
        self.hidden_channels = hidden_channels
        self.output_network = MLP(hidden_channels, 1, hidden_channels // 2, activation,
                                  num_hidden_layers=kwargs.get("num_hidden_layers", 0), dtype=dtype)
        self.q_dim = kwargs["q_dim"]
        self.num_interaction_layers = kwargs["num_layers"]
        self.layer_weights = kwargs["q_weights"]
        self.cutoff = kwargs["coulomb_cutoff"]
        assert len(self.layer_weights) == self.num_interaction_layers + 1
        w = torch.zeros((self.num_interaction_layers + 1, self.q_dim), dtype=dtype)
        for i in range(self.num_interaction_layers + 1):
            w[i, :] = torch.tensor(self.layer_weights[i], dtype=dtype)
        self.register_buffer("qweights", w.flatten())
        self._factor = 0.5 * 27.211386024367243 * 0.5291772105638411      # half Hartree x Bohr (eV, Angstrom)
        if self.cutoff is None:
            self.mode, self.distance = "all_to_all", None
        else:
            self.mode = "cutoff"
            max_nb = kwargs.get("coulomb_max_num_neighbors", None)
            if max_nb is None:      # uniform water-like density (reference :425-430)
                max_nb = int(0.1 * 4 / 3 * math.pi * self.cutoff ** 3)
            self.distance = OptimizedDistance(cutoff_lower=0.0, cutoff_upper=self.cutoff, max_num_pairs=-max_nb, return_vecs=False,
                                              resize_to_fit=not self.static_shapes, include_transpose=False, loop=False,
                                              strategy=kwargs.get("coulomb_neighbor_strategy", "brute"))
            self.epsilon_solvent = kwargs.get("coulomb_epsilon_solvent", 78.3)
        self.edge_index = torch.empty(2, 0)
        self.reset_parameters()

    def reset_parameters(self):
        self.output_network.reset_parameters()

    def coulomb_per_atom(self, charges, z, pos, batch, box: Optional[torch.Tensor] = None):
        """e_i [N]: every pair energy is added to both of its atoms, as in the reference (:487-606)."""
        n = charges.shape[0]
        if self.mode == "all_to_all":
            if box is not None:
                raise ValueError("PBC is not supported with coulomb_cutoff = None")
            capturing = pos.is_cuda and torch.cuda.is_current_stream_capturing()
            if not capturing:
                self.edge_index = _triu_indices(pos, batch)
            else:
                assert self.edge_index.shape[1] > 0, "Warming up is needed before capturing the model into a CUDA graph"
            ei = self.edge_index
            d_ij = torch.linalg.norm(pos[ei[0]] - pos[ei[1]], dim=-1)
            shape = 1.0 / d_ij
        else:
            ei, d_ij, _ = self.distance(pos, batch, box=box)
            if self.static_shapes:      # -1 padding -> an extra dummy atom with zero charge (:531-550)
                mask = ei[0] < 0
                ei = ei.masked_fill(mask.unsqueeze(0).expand_as(ei), n)
                d_ij = d_ij.masked_fill(mask, 0)
                charges = torch.cat([charges, torch.zeros(1, charges.shape[1], device=charges.device, dtype=charges.dtype)], dim=0)
            eps = self.epsilon_solvent
            k_rf = (1.0 / self.cutoff ** 3) * (eps - 1.0) / (2.0 * eps + 1.0)
            c_rf = (1.0 / self.cutoff) * (3.0 * eps) / (2.0 * eps + 1.0)
            shape = 1.0 / d_ij + k_rf * d_ij ** 2 - c_rf
        fc = 1.0 - _exp_cutoff(d_ij, 4.6)
        e_ij = self._factor * (fc * shape).unsqueeze(-1) * charges[ei[0]] * charges[ei[1]]
        e_ij = torch.sum(e_ij * self.qweights.unsqueeze(0), dim=-1) / torch.sum(self.qweights)
        e_i = torch.zeros(charges.shape[0], device=pos.device, dtype=pos.dtype)
        e_i = e_i.index_add(0, ei[0], e_ij).index_add(0, ei[1], e_ij)
        return e_i[:n]

    def pre_reduce(self, x, v: Optional[torch.Tensor], z, pos, batch, box: Optional[torch.Tensor] = None):
        charges = x[:, self.hidden_channels:]           # the charges were appended to the node features
        if not self.static_shapes:
            assert charges.shape[1] == self.qweights.shape[0]
        x = self.output_network(x[:, : self.hidden_channels])
        return x + self.coulomb_per_atom(charges, z, pos, batch, box).unsqueeze(-1)
