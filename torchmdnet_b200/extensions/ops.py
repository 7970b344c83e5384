"""Drop-in for `torchmdnet.extensions.ops` (reference: torchmdnet/extensions/ops.py:14-106).

Same name, arguments, return order and error behaviour as the reference operator; the work is
done by hand-written sm_100a kernels behind the C ABI (include/tmdnet_b200.h) instead of NVIDIA
Warp.  Both strategies ("brute", "cell") resolve to the same deterministic kernel: the output
is sorted by (i, j) rather than in atomic-counter order (the reference documents the list as
"without any enforced ordering").
"""
from typing import Tuple

import torch
from torch import Tensor

from .neighbors import neighbor_pairs

__all__ = ["get_neighbor_pairs_kernel"]


def get_neighbor_pairs_kernel(
    strategy: str,
    positions: Tensor,
    batch: Tensor,
    box_vectors: Tensor,
    use_periodic: bool,
    cutoff_lower: float,
    cutoff_upper: float,
    max_num_pairs: int,
    loop: bool,
    include_transpose: bool,
    num_cells: int,
) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Returns (neighbors[2,P] int64 padded with -1, deltas[P,3], distances[P], num_pairs[1] int32).

    Capacity overflow is not an exception here: pairs past `max_num_pairs` are dropped and the true
    count is returned in `num_pairs`, exactly like the reference (warp_ops/neighbors.py:69-80); the
    caller asserts on it (models/utils.py:297-300).
    """
    if strategy not in ("brute", "cell"):
        raise ValueError(f"Unknown strategy: {strategy}")  # warp_ops/neighbors.py:454-455
    if positions.dim() != 2 or positions.shape[1] != 3:
        raise ValueError('Expected "positions" to have shape (N, 3)')
    if positions.dtype not in (torch.float32, torch.float64):
        raise ValueError("Unsupported dtype for positions: " + str(positions.dtype))
    return neighbor_pairs(positions, batch, box_vectors, use_periodic, cutoff_lower, cutoff_upper, max_num_pairs,
                          loop, include_transpose)
