"""Neighbour-list ops over the C ABI (tmd_neighbor_*).

Two faces:
  * `neighbor_graph(...)`   -- what the B200 model path uses: a deterministic CSR (sorted by (i, j)),
    int32 indices, capacity-sized buffers + a device-side pair counter, the reverse-edge map, and a
    position gradient that is an atomic-free CSR gather.
  * `get_neighbor_pairs_kernel(...)` (in ops.py) -- the reference's operator signature
    (torchmdnet/extensions/ops.py:14-106), built on the same kernels.
"""
from typing import NamedTuple, Optional

import torch
from torch import Tensor

from .. import _lib


class NeighborGraph(NamedTuple):
    indptr: Tensor      # [N+1] int32, CSR row pointer (rows = edge_index[0])
    nbr_i: Tensor       # [P] int32  edge_index[0]   (slots >= num_pairs hold -1)
    nbr_j: Tensor       # [P] int32  edge_index[1]
    rev: Tensor         # [P] int32  slot of the transposed pair (j, i)
    num_pairs: Tensor   # [1] int32  pairs found (may exceed P: overflow, checked by the caller)
    vec: Tensor         # [P,3]  pos_i - pos_j (minimum image)
    dist: Tensor        # [P]
    order: Optional[Tensor] = None       # [N] int32  atoms grouped by search cell (spatially adjacent rows are consecutive) ...
    order_flag: Optional[Tensor] = None  # [1] int32  ... valid iff non-zero (the cell grid was used); else use 0..N-1


def _suffix(dtype):
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise ValueError(f"neighbor ops support float32/float64 positions, got {dtype}")  # neighbors.py:426-427


def _prep_box(box: Optional[Tensor], use_periodic: bool, pos: Tensor):
    if not use_periodic:
        return None, 0
    if box is None:
        raise ValueError("use_periodic=True needs box vectors")
    if box.dim() == 2:
        box = box.unsqueeze(0)
    elif box.dim() != 3:
        raise ValueError('Expected "box_vectors" to have shape (n_batch, 3, 3)')  # neighbors.py:58-61
    box = box.to(device=pos.device, dtype=pos.dtype).contiguous()
    return box, (0 if box.shape[0] == 1 else 9)


def _build(pos, batch, box, use_periodic, lo, hi, loop, transpose, max_pairs, want_rev):
    _lib.require_cuda(pos, batch)
    n = pos.shape[0]
    dev = pos.device
    pos_c = pos.detach().contiguous()
    batch_c = batch.detach().to(torch.long).contiguous()
    box_c, stride = _prep_box(box, use_periodic, pos)
    i32 = dict(device=dev, dtype=torch.int32)
    indptr = torch.empty(n + 1, **i32)
    nbr_i = torch.full((max_pairs,), -1, **i32)
    nbr_j = torch.full((max_pairs,), -1, **i32)
    rev = torch.full((max_pairs,), -1, **i32) if want_rev else None
    vec = torch.zeros(max_pairs, 3, device=dev, dtype=pos.dtype)
    dist = torch.zeros(max_pairs, device=dev, dtype=pos.dtype)
    num_pairs = torch.zeros(1, **i32)
    if n == 0:
        indptr.zero_()
        return indptr, nbr_i, nbr_j, rev, num_pairs, vec, dist, None, None
    # cell grids, one per periodic rectangular box (decided on the device from the box tensor); otherwise segment scan
    n_boxes = 0 if box_c is None else int(box_c.shape[0])
    max_cells = min(4 * n + 64, 1 << 20) if (use_periodic and n >= 512 and 1 <= n_boxes <= 4096) else 0
    ws = torch.empty(5 * n + 8 + 3 * (max_cells + 1) + 4 * max(n_boxes, 1), **i32)
    _lib.call("tmd_neighbor_csr_" + _suffix(pos.dtype), _lib.ptr(pos_c), _lib.ptr(batch_c), _lib.ptr(box_c), stride,
              int(use_periodic), n, float(lo), float(hi), int(loop), int(transpose), int(max_pairs), max_cells, n_boxes, _lib.ptr(indptr),
              _lib.ptr(nbr_i), _lib.ptr(nbr_j), _lib.ptr(vec), _lib.ptr(dist), _lib.ptr(num_pairs), _lib.ptr(rev),
              _lib.ptr(ws), device=dev)
    # workspace layout (csrc/neighbors.cu): seg_lo[n] seg_hi[n] deg[n] flag[1] | hdr[4] cell_of[n] cell_atoms[n] ...;
    # hdr[3] != 0 <=> the cell grid was used and cell_atoms lists the atoms cell by cell
    order, order_flag = ws[4 * n + 5: 5 * n + 5], ws[3 * n + 4: 3 * n + 5]
    return indptr, nbr_i, nbr_j, rev, num_pairs, vec, dist, order, order_flag


def _grad_pos_torch(nbr_i, nbr_j, vec, dist, g_vec, g_dist, n):
    """Differentiable restatement of the position gradient (used only under create_graph=True, i.e.
    training with a force loss; reference: extensions/neighbor_utils.py:11-46)."""
    valid = (nbr_i >= 0) & (dist != 0)
    i = nbr_i.clamp(min=0).long()
    j = nbr_j.clamp(min=0).long()
    g = torch.zeros_like(vec)
    if g_vec is not None:
        g = g + g_vec
    if g_dist is not None:
        g = g + (g_dist / dist.masked_fill(~valid, 1)).unsqueeze(-1) * vec
    g = g * valid.unsqueeze(-1)
    out = torch.zeros(n, 3, device=vec.device, dtype=vec.dtype)
    return out.index_add(0, i, g).index_add(0, j, -g)


class _NeighborGraphFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, batch, box, use_periodic, lo, hi, loop, max_pairs):
        indptr, nbr_i, nbr_j, rev, num_pairs, vec, dist, order, order_flag = _build(pos, batch, box, use_periodic, lo, hi, loop,
                                                                                     True, max_pairs, True)
        ctx.save_for_backward(indptr, nbr_i, nbr_j, rev, num_pairs, vec, dist)
        ctx.n = pos.shape[0]
        if order is None:
            order = torch.zeros(0, device=pos.device, dtype=torch.int32)
            order_flag = torch.zeros(1, device=pos.device, dtype=torch.int32)
        ctx.mark_non_differentiable(indptr, nbr_i, nbr_j, rev, num_pairs, order, order_flag)
        return indptr, nbr_i, nbr_j, rev, num_pairs, vec, dist, order, order_flag

    @staticmethod
    def backward(ctx, _a, _b, _c, _d, _e, g_vec, g_dist, _o=None, _of=None):
        indptr, nbr_i, nbr_j, rev, num_pairs, vec, dist = ctx.saved_tensors
        n = ctx.n
        if torch.is_grad_enabled() and ((g_vec is not None and g_vec.requires_grad) or (g_dist is not None and g_dist.requires_grad)):
            gpos = _grad_pos_torch(nbr_i, nbr_j, vec, dist, g_vec, g_dist, n)
        else:
            gpos = torch.empty(n, 3, device=vec.device, dtype=vec.dtype)
            gv = None if g_vec is None else g_vec.contiguous()
            gd = None if g_dist is None else g_dist.contiguous()
            if n:
                _lib.call("tmd_neighbor_grad_csr_" + _suffix(vec.dtype), _lib.ptr(indptr), _lib.ptr(rev),
                          _lib.ptr(num_pairs), vec.shape[0], n, _lib.ptr(vec), _lib.ptr(dist), _lib.ptr(gv), _lib.ptr(gd),
                          _lib.ptr(gpos), device=vec.device)
        return gpos, None, None, None, None, None, None, None


def neighbor_graph(pos: Tensor, batch: Tensor, box: Optional[Tensor], use_periodic: bool, cutoff_lower: float,
                   cutoff_upper: float, loop: bool, max_pairs: int) -> NeighborGraph:
    """Full (transposed) neighbour list as a sorted CSR; differentiable in `pos` through vec/dist."""
    return NeighborGraph(*_NeighborGraphFn.apply(pos, batch, box, use_periodic, float(cutoff_lower), float(cutoff_upper),
                                                 bool(loop), int(max_pairs)))


class _NeighborPairsFn(torch.autograd.Function):
    """Reference-shaped outputs: neighbors int64 [2,P] (-1 padded), deltas, distances, num_pairs."""

    @staticmethod
    def forward(ctx, pos, batch, box, use_periodic, lo, hi, loop, transpose, max_pairs):
        _, nbr_i, nbr_j, _, num_pairs, vec, dist, _, _ = _build(pos, batch, box, use_periodic, lo, hi, loop, transpose,
                                                                max_pairs, False)
        neighbors = torch.stack([nbr_i, nbr_j]).to(torch.long)
        ctx.save_for_backward(neighbors, vec, dist)
        ctx.n = pos.shape[0]
        ctx.mark_non_differentiable(neighbors, num_pairs)
        return neighbors, vec, dist, num_pairs

    @staticmethod
    def backward(ctx, _n, g_vec, g_dist, _c):
        neighbors, vec, dist = ctx.saved_tensors
        n = ctx.n
        if torch.is_grad_enabled() and ((g_vec is not None and g_vec.requires_grad) or (g_dist is not None and g_dist.requires_grad)):
            gpos = _grad_pos_torch(neighbors[0], neighbors[1], vec, dist, g_vec, g_dist, n)
        else:
            gpos = torch.zeros(n, 3, device=vec.device, dtype=vec.dtype)
            gv = None if g_vec is None else g_vec.contiguous()
            gd = None if g_dist is None else g_dist.contiguous()
            if n and vec.shape[0]:
                _lib.call("tmd_neighbor_grad_coo_" + _suffix(vec.dtype), _lib.ptr(neighbors), vec.shape[0], _lib.ptr(vec),
                          _lib.ptr(dist), _lib.ptr(gv), _lib.ptr(gd), _lib.ptr(gpos), device=vec.device)
        return gpos, None, None, None, None, None, None, None, None


def neighbor_pairs(pos, batch, box, use_periodic, cutoff_lower, cutoff_upper, max_num_pairs, loop, include_transpose):
    return _NeighborPairsFn.apply(pos, batch, box, bool(use_periodic), float(cutoff_lower), float(cutoff_upper),
                                  bool(loop), bool(include_transpose), int(max_num_pairs))
