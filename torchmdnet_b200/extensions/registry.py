"""Op-registry seam: the reference's `torch.library` op names, backed by the sm_100a kernels.

The reference registers 25 custom ops in the namespaces `torchmdnet::`, `nvtnet::` and `tensornet::`
(torchmdnet/extensions/warp_ops/*.py) so that TorchScript (`extensions/ops.py:66-92`),
`torch.compile(fullgraph=True)` (`calculators.py:297-303`) and `torch.export` can resolve them.  Importing this module
registers the forward and backward primitives under the SAME names and signatures (`mutates_args=()`, CUDA only --
there is no CPU path), with `register_fake` shape functions and `register_autograd` chaining fwd -> bwd, plus the seven
`*_bwd_bwd_primitive` ops (double backward: training with a force loss) with the reference's argument and result order; they
evaluate the pull-back of the differentiable backward of tensornet_ops.py (kernel re-entry for the (bi)linear ops, torch
restatements for the quadratic ones).

Do not import it next to the reference's own `torchmdnet.extensions.warp_ops` (same op names).
"""
from typing import List

import torch
from torch import Tensor

from . import tensornet_ops as T
from .neighbors import _build, _suffix
from .. import _lib

p = _lib.ptr
_CUDA = ["cuda"]
REGISTERED: List[str] = []


def _op(name):
    REGISTERED.append(name)
    return torch.library.custom_op(name, mutates_args=(), device_types=_CUDA)


# ------------------------------------------------------------------------------------------ neighbour list
def _neighbor_fwd(positions, batch, box_vectors, use_periodic, cutoff_lower, cutoff_upper, max_num_pairs, loop,
                  include_transpose) -> List[Tensor]:
    _, nbr_i, nbr_j, _, num_pairs, vec, dist, _, _ = _build(positions, batch, box_vectors if use_periodic else None, use_periodic,
                                                      cutoff_lower, cutoff_upper, loop, include_transpose, max_num_pairs, False)
    return [torch.stack([nbr_i, nbr_j]).to(torch.long), vec, dist, num_pairs]


@_op("torchmdnet::warp_neighbor_brute_fwd")
def warp_neighbor_brute_fwd(positions: Tensor, batch: Tensor, box_vectors: Tensor, use_periodic: bool, cutoff_lower: float,
                            cutoff_upper: float, max_num_pairs: int, loop: bool, include_transpose: bool) -> List[Tensor]:
    return _neighbor_fwd(positions, batch, box_vectors, use_periodic, cutoff_lower, cutoff_upper, max_num_pairs, loop, include_transpose)


@_op("torchmdnet::warp_neighbor_cell_fwd")
def warp_neighbor_cell_fwd(positions: Tensor, batch: Tensor, box_vectors: Tensor, use_periodic: bool, cutoff_lower: float,
                           cutoff_upper: float, max_num_pairs: int, loop: bool, include_transpose: bool,
                           num_cells: int) -> List[Tensor]:
    return _neighbor_fwd(positions, batch, box_vectors, use_periodic, cutoff_lower, cutoff_upper, max_num_pairs, loop, include_transpose)


def _neighbor_fake(positions, batch, box_vectors, use_periodic, cutoff_lower, cutoff_upper, max_num_pairs, loop,
                   include_transpose, num_cells=0):
    dev, dt = positions.device, positions.dtype
    return [torch.empty(2, max_num_pairs, device=dev, dtype=torch.long), torch.empty(max_num_pairs, 3, device=dev, dtype=dt),
            torch.empty(max_num_pairs, device=dev, dtype=dt), torch.empty(1, device=dev, dtype=torch.int32)]


def _neighbor_setup(ctx, inputs, output):
    neighbors, deltas, distances, _ = output
    ctx.save_for_backward(neighbors, deltas, distances)
    ctx.n, ctx.nargs = inputs[0].shape[0], len(inputs)


def _neighbor_bwd(ctx, grads):
    neighbors, vec, dist = ctx.saved_tensors
    g_vec, g_dist = grads[1], grads[2]
    gpos = torch.zeros(ctx.n, 3, device=vec.device, dtype=vec.dtype)
    if ctx.n and vec.shape[0]:
        gv = None if g_vec is None else g_vec.contiguous()     # bound to locals: the copies must outlive the launch
        gd = None if g_dist is None else g_dist.contiguous()
        _lib.call("tmd_neighbor_grad_coo_" + _suffix(vec.dtype), p(neighbors), vec.shape[0], p(vec), p(dist), p(gv), p(gd), p(gpos),
                  device=vec.device)
    return (gpos,) + (None,) * (ctx.nargs - 1)


for _name in ("torchmdnet::warp_neighbor_brute_fwd", "torchmdnet::warp_neighbor_cell_fwd"):
    torch.library.register_fake(_name)(_neighbor_fake)
    torch.library.register_autograd(_name, _neighbor_bwd, setup_context=_neighbor_setup)


# ------------------------------------------------------------------------------------------ graph transform
@_op("nvtnet::count_row_col_primitive")
def count_row_col_primitive(edge_index: Tensor, num_nodes: int) -> List[Tensor]:
    ei = edge_index.long()
    valid = (ei[0] >= 0) & (ei[1] >= 0)
    rc = torch.zeros(num_nodes + 1, dtype=torch.int32, device=ei.device)
    cc = torch.zeros(num_nodes + 1, dtype=torch.int32, device=ei.device)
    one = torch.ones(int(ei.shape[1]), dtype=torch.int32, device=ei.device) * valid.to(torch.int32)
    rc.index_add_(0, ei[0].clamp(min=0) + 1, one)
    cc.index_add_(0, ei[1].clamp(min=0) + 1, one)
    return [rc, cc]


@torch.library.register_fake("nvtnet::count_row_col_primitive")
def _(edge_index, num_nodes):
    z = torch.empty(num_nodes + 1, dtype=torch.int32, device=edge_index.device)
    return [z, torch.empty_like(z)]


@_op("nvtnet::convert_to_sparse_primitive")
def convert_to_sparse_primitive(edge_index: Tensor, row_count: Tensor, col_count: Tensor, row_indptr: Tensor,
                                col_indptr: Tensor) -> List[Tensor]:
    rd, ri, _, cd, ci, _ = T.graph_transform(edge_index.to(torch.int32), row_indptr.shape[0] - 1)
    return [ri, ci, rd, cd]


@torch.library.register_fake("nvtnet::convert_to_sparse_primitive")
def _(edge_index, row_count, col_count, row_indptr, col_indptr):
    return [torch.empty(edge_index.shape[1], dtype=torch.int32, device=edge_index.device) for _ in range(4)]


# ------------------------------------------------------------------------------------------ TensorNet ops
def _pair(fwd_name, bwd_name, fwd_impl, bwd_impl, fwd_fake, bwd_fake, n_in, single_out):
    """register fwd + bwd primitives and chain them with register_autograd (inputs are saved, like the reference)."""
    torch.library.register_fake(fwd_name)(fwd_fake)
    torch.library.register_fake(bwd_name)(bwd_fake)

    def setup(ctx, inputs, output):
        ctx.save_for_backward(*[t for t in inputs if isinstance(t, Tensor)])

    def bwd(ctx, *grads):
        g = grads if single_out else grads[0]
        g = [x.contiguous() for x in (g if isinstance(g, (list, tuple)) else [g])]
        out = getattr(getattr(torch.ops, bwd_name.split("::")[0]), bwd_name.split("::")[1])(*g, *ctx.saved_tensors)
        out = list(out) if isinstance(out, (list, tuple)) else [out]
        return tuple(out + [None] * (n_in - len(out)))

    torch.library.register_autograd(fwd_name, bwd, setup_context=setup)


def _t(N, F, ref):
    return T._triple(N, F, ref)


@_op("tensornet::compose_tensor_fwd_primitive")
def compose_fwd(x: Tensor, y: Tensor, z: Tensor) -> Tensor:
    return T._Compose.forward(None, x, y, z)


@_op("tensornet::compose_tensor_bwd_primitive")
def compose_bwd(grad_output: Tensor, x: Tensor, y: Tensor, z: Tensor) -> List[Tensor]:
    return list(T._Compose.backward(None, grad_output))


_pair("tensornet::compose_tensor_fwd_primitive", "tensornet::compose_tensor_bwd_primitive", None, None,
      lambda x, y, z: x.new_empty(x.shape[0], 3, 3, x.shape[-1]),
      lambda g, x, y, z: [torch.empty_like(x), torch.empty_like(y), torch.empty_like(z)], 3, True)


@_op("tensornet::decompose_tensor_fwd_primitive")
def decompose_fwd(x: Tensor) -> List[Tensor]:
    return list(T._Decompose.forward(None, x))


@_op("tensornet::decompose_tensor_bwd_primitive")
def decompose_bwd(grad_output_i: Tensor, grad_output_a: Tensor, grad_output_s: Tensor, x: Tensor) -> Tensor:
    return T._Decompose.backward(None, grad_output_i, grad_output_a, grad_output_s)


_pair("tensornet::decompose_tensor_fwd_primitive", "tensornet::decompose_tensor_bwd_primitive", None, None,
      lambda x: [x.new_empty(x.shape[0], k, x.shape[-1]) for k in (1, 3, 5)], lambda gi, ga, gs, x: torch.empty_like(x), 1, False)


def _register_matmul(kind):
    fwd_name, bwd_name = f"tensornet::tensor_matmul_{kind}_3x3_fwd_primitive", f"tensornet::tensor_matmul_{kind}_3x3_bwd_primitive"

    @_op(fwd_name)
    def mm_fwd(x: Tensor, y: Tensor) -> Tensor:
        if x.ndim != 4 or y.ndim != 4 or x.shape[1] != 3 or x.shape[2] != 3 or y.shape[1] != 3 or y.shape[2] != 3:
            raise ValueError("x and y must be [N,3,3,F] tensors")
        x, y = T._c(x), T._c(y)
        z = torch.empty_like(x)
        _lib.call(f"tmd_matmul_{kind}_fwd", p(x), p(y), p(z), x.shape[0], x.shape[-1], device=x.device)
        return z

    @_op(bwd_name)
    def mm_bwd(grad_output: Tensor, x: Tensor, y: Tensor) -> List[Tensor]:
        x, y, g = T._c(x), T._c(y), T._c(grad_output)
        dx, dy = torch.empty_like(x), torch.empty_like(y)
        _lib.call(f"tmd_matmul_{kind}_bwd", p(x), p(y), p(g), p(dx), p(dy), x.shape[0], x.shape[-1], device=x.device)
        return [dx, dy]

    _pair(fwd_name, bwd_name, None, None, lambda x, y: torch.empty_like(x),
          lambda g, x, y: [torch.empty_like(x), torch.empty_like(y)], 2, True)


_register_matmul("o3")
_register_matmul("so3")


@_op("tensornet::tensor_norm3_fwd_primitive")
def norm3_fwd(x: Tensor) -> Tensor:
    x = T._c(x)
    out = x.new_empty(x.shape[0], 3 * x.shape[-1])
    _lib.call("tmd_norm3_fwd", p(x), p(out), x.shape[0], x.shape[-1], device=x.device)
    return out


@_op("tensornet::tensor_norm3_bwd_primitive")
def norm3_bwd(grad_output: Tensor, x: Tensor) -> List[Tensor]:
    x, g = T._c(x), T._c(grad_output)
    dx = torch.empty_like(x)
    _lib.call("tmd_norm3_bwd", p(g), p(x), p(dx), x.shape[0], x.shape[-1], device=x.device)
    return [dx]


_pair("tensornet::tensor_norm3_fwd_primitive", "tensornet::tensor_norm3_bwd_primitive", None, None,
      lambda x: x.new_empty(x.shape[0], 3 * x.shape[-1]), lambda g, x: [torch.empty_like(x)], 1, True)


@_op("tensornet::message_passing_fwd_primitive")
def mp_fwd(x: Tensor, y: Tensor, z: Tensor, edge_attr: Tensor, row_data: Tensor, row_indices: Tensor, row_indptr: Tensor,
           col_data: Tensor, col_indices: Tensor, col_indptr: Tensor) -> List[Tensor]:
    x, y, z, w = T._c(x), T._c(y), T._c(z), T._c(edge_attr)
    N, F = x.shape[0], x.shape[-1]
    o = _t(N, F, x)
    _lib.call("tmd_mp_fwd", p(x), p(y), p(z), p(w), p(T._ci(row_data)), p(T._ci(row_indices)), p(T._ci(row_indptr)), N, F,
              p(o[0]), p(o[1]), p(o[2]), device=x.device)
    return list(o)


@_op("tensornet::message_passing_bwd_primitive")
def mp_bwd(grad_output_x: Tensor, grad_output_y: Tensor, grad_output_z: Tensor, x: Tensor, y: Tensor, z: Tensor,
           edge_attr: Tensor, row_data: Tensor, row_indices: Tensor, row_indptr: Tensor, col_data: Tensor,
           col_indices: Tensor, col_indptr: Tensor) -> List[Tensor]:
    x, y, z, w = T._c(x), T._c(y), T._c(z), T._c(edge_attr)
    N, F = x.shape[0], x.shape[-1]
    g = _t(N, F, x)
    dw = torch.zeros_like(w)
    _lib.call("tmd_mp_bwd", p(x), p(y), p(z), p(w), p(T._c(grad_output_x)), p(T._c(grad_output_y)), p(T._c(grad_output_z)),
              p(T._ci(col_data)), p(T._ci(col_indices)), p(T._ci(col_indptr)), N, F, p(g[0]), p(g[1]), p(g[2]), p(dw), device=x.device)
    return [g[0], g[1], g[2], dw]


_pair("tensornet::message_passing_fwd_primitive", "tensornet::message_passing_bwd_primitive", None, None,
      lambda x, y, z, w, *a: [torch.empty_like(x), torch.empty_like(y), torch.empty_like(z)],
      lambda gx, gy, gz, x, y, z, w, *a: [torch.empty_like(x), torch.empty_like(y), torch.empty_like(z), torch.empty_like(w)], 10, False)


@_op("tensornet::radial_message_passing_fwd_primitive")
def radial_fwd(edge_vec_norm: Tensor, edge_attr: Tensor, row_data: Tensor, row_indptr: Tensor) -> List[Tensor]:
    r, w = T._c(edge_vec_norm), T._c(edge_attr)
    N, F = row_indptr.shape[0] - 1, w.shape[-1]
    o = _t(N, F, w)
    _lib.call("tmd_radial_mp_fwd", p(r), p(w), p(T._ci(row_data)), p(T._ci(row_indptr)), N, F, p(o[0]), p(o[1]), p(o[2]), device=w.device)
    return list(o)


@_op("tensornet::radial_message_passing_bwd_primitive")
def radial_bwd(grad_output_I: Tensor, grad_output_A: Tensor, grad_output_S: Tensor, edge_vec_norm: Tensor, edge_attr: Tensor,
               row_data: Tensor, row_indptr: Tensor) -> List[Tensor]:
    r, w = T._c(edge_vec_norm), T._c(edge_attr)
    N, F = row_indptr.shape[0] - 1, w.shape[-1]
    dw, dr = torch.zeros_like(w), torch.zeros_like(r)
    _lib.call("tmd_radial_mp_bwd", p(r), p(w), p(T._ci(row_data)), p(T._ci(row_indptr)), N, F, p(T._c(grad_output_I)),
              p(T._c(grad_output_A)), p(T._c(grad_output_S)), p(dw), p(dr), device=w.device)
    return [dr, dw]


_pair("tensornet::radial_message_passing_fwd_primitive", "tensornet::radial_message_passing_bwd_primitive", None, None,
      lambda r, w, rd, rp: [w.new_empty(rp.shape[0] - 1, k, w.shape[-1]) for k in (1, 3, 5)],
      lambda gi, ga, gs, r, w, rd, rp: [torch.empty_like(r), torch.empty_like(w)], 4, False)


# ------------------------------------------------------------------------------------------ double backward primitives
# Registered as CompositeImplicitAutograd implementations (torch.library.Library.define / impl): they are compositions of the
# differentiable backward of tensornet_ops.py (kernel re-entry for the (bi)linear ops, torch restatements for the quadratic ones)
# and must run ABOVE the autograd dispatch key -- inside a backend `custom_op` body autograd is excluded and nothing would record.
_LIB = torch.library.Library("tensornet", "FRAGMENT")


def _composite(name, schema, fn):
    _LIB.define(f"{name}{schema}")
    _LIB.impl(name, fn, "CompositeImplicitAutograd")
    REGISTERED.append(f"tensornet::{name}")


def _pullback(bwd_fn, grad_outs, inputs, cotangents):
    """the reference's `*_bwd_bwd` contract: given the cotangents of the BACKWARD op's results, return (d/d grad_outs, d/d inputs).
    `bwd_fn(grad_outs, inputs)` evaluates the backward op with differentiable operations (tensornet_ops.py)."""
    with torch.enable_grad():
        go = [t.detach().requires_grad_(True) for t in grad_outs]
        xi = [t.detach().requires_grad_(True) for t in inputs]
        res = list(bwd_fn(go, xi))
        keep = [(r, c.detach()) for r, c in zip(res, cotangents) if r is not None and r.requires_grad]
        g = torch.autograd.grad([r for r, _ in keep], go + xi, [c for _, c in keep], allow_unused=True) if keep else [None] * (len(go) + len(xi))
    g = [torch.zeros_like(t) if gi is None else gi for gi, t in zip(g, go + xi)]
    return g[: len(go)], g[len(go):]


def _compose_bwd_bwd(grad_output, grad_grad_x, grad_grad_y, grad_grad_z, x, y, z):
    dgo, _ = _pullback(lambda go, xi: T._ComposeT.apply(go[0]), [grad_output], [], [grad_grad_x, grad_grad_y, grad_grad_z])
    return [dgo[0], torch.zeros_like(x), torch.zeros_like(y), torch.zeros_like(z)]     # linear op: no dependence on (x, y, z)


_composite("compose_tensor_bwd_bwd_primitive",
           "(Tensor grad_output, Tensor grad_grad_x, Tensor grad_grad_y, Tensor grad_grad_z, Tensor x, Tensor y, Tensor z) -> Tensor[]", _compose_bwd_bwd)


def _decompose_bwd_bwd(grad_output_i, grad_output_a, grad_output_s, grad_grad_x, x):
    dgo, _ = _pullback(lambda go, xi: [T._DecomposeT.apply(*go)], [grad_output_i, grad_output_a, grad_output_s], [], [grad_grad_x])
    return [dgo[0], dgo[1], dgo[2], torch.zeros_like(x)]


_composite("decompose_tensor_bwd_bwd_primitive",
           "(Tensor grad_output_i, Tensor grad_output_a, Tensor grad_output_s, Tensor grad_grad_x, Tensor x) -> Tensor[]", _decompose_bwd_bwd)


def _register_matmul_bwd_bwd(kind):
    fn = T._MatMulO3 if kind == "o3" else T._MatMulSO3

    def mm_bwd_bwd(grad_output, grad_grad_x, grad_grad_y, x, y):
        if x.ndim != 4 or y.ndim != 4 or x.shape[1] != 3 or x.shape[2] != 3 or y.shape[1] != 3 or y.shape[2] != 3:
            raise ValueError("x and y must be [N,3,3,F] tensors")
        dgo, dx = _pullback(lambda go, xi: [fn.apply(go[0], xi[1].transpose(1, 2)), fn.apply(xi[0].transpose(1, 2), go[0])],
                            [grad_output], [x, y], [grad_grad_x, grad_grad_y])
        return [dgo[0], dx[0], dx[1]]

    _composite(f"tensor_matmul_{kind}_3x3_bwd_bwd_primitive",
               "(Tensor grad_output, Tensor grad_grad_x, Tensor grad_grad_y, Tensor x, Tensor y) -> Tensor[]", mm_bwd_bwd)


_register_matmul_bwd_bwd("o3")
_register_matmul_bwd_bwd("so3")


def _norm3_bwd_bwd(grad_grad_x, x, grad_output):
    def bwd(go, xi):
        n = torch.cat(T._t_norm3(*T._Decompose.apply(xi[0])), -1)
        return torch.autograd.grad([n], [xi[0]], [go[0]], create_graph=True)
    dgo, dx = _pullback(bwd, [grad_output], [x], [grad_grad_x])
    return [dgo[0], dx[0]]


_composite("tensor_norm3_bwd_bwd_primitive", "(Tensor grad_grad_x, Tensor x, Tensor grad_output) -> Tensor[]", _norm3_bwd_bwd)


def _mp_bwd_bwd(grad_output_x, grad_output_y, grad_output_z, grad_grad_x, grad_grad_y, grad_grad_z, grad_grad_edge_attr, x, y, z, edge_attr,
                row_data, row_indices, row_indptr, col_data, col_indices, col_indptr):
    graph = (row_data, row_indices, row_indptr, col_data, col_indices, col_indptr)
    dgo, dx = _pullback(lambda go, xi: T._MPBwd.apply(xi[0], xi[1], xi[2], xi[3], go[0], go[1], go[2], *graph),
                        [grad_output_x, grad_output_y, grad_output_z], [x, y, z, edge_attr],
                        [grad_grad_x, grad_grad_y, grad_grad_z, grad_grad_edge_attr])
    return [dgo[0], dgo[1], dgo[2], dx[0], dx[1], dx[2], dx[3]]


_composite("message_passing_bwd_bwd_primitive",
           "(Tensor grad_output_x, Tensor grad_output_y, Tensor grad_output_z, Tensor grad_grad_x, Tensor grad_grad_y, Tensor grad_grad_z, "
           "Tensor grad_grad_edge_attr, Tensor x, Tensor y, Tensor z, Tensor edge_attr, Tensor? row_data, Tensor row_indices, Tensor row_indptr, "
           "Tensor? col_data, Tensor col_indices, Tensor col_indptr) -> Tensor[]", _mp_bwd_bwd)


def _radial_bwd_bwd(grad_output_I, grad_output_A, grad_output_S, grad_grad_edge_vec_norm, grad_grad_edge_attr, edge_vec_norm, edge_attr,
                    row_data, row_indptr):
    def bwd(go, xi):
        out = T._t_radial_mp(xi[0], xi[1], row_data, row_indptr)
        return torch.autograd.grad(list(out), [xi[0], xi[1]], list(go), create_graph=True)
    dgo, dx = _pullback(bwd, [grad_output_I, grad_output_A, grad_output_S], [edge_vec_norm, edge_attr],
                        [grad_grad_edge_vec_norm, grad_grad_edge_attr])
    return [dgo[0], dgo[1], dgo[2], dx[0], dx[1]]


_composite("radial_message_passing_bwd_bwd_primitive",
           "(Tensor grad_output_I, Tensor grad_output_A, Tensor grad_output_S, Tensor grad_grad_edge_vec_norm, Tensor grad_grad_edge_attr, "
           "Tensor edge_vec_norm, Tensor edge_attr, Tensor? row_data, Tensor row_indptr) -> Tensor[]", _radial_bwd_bwd)
