"""Drop-in for `torchmdnet.extensions.warp_ops` (reference: warp_ops/__init__.py:46-56) plus the fused
compact-irreps stages the B200 model path uses.

Function names, argument order and tensor layouts of the 8 reference handles are kept
(`graph_transform`, `fn_radial_message_passing`, `fn_message_passing`, `fn_compose_tensor`,
`fn_decompose_tensor`, `fn_tensor_matmul_o3_3x3`, `fn_tensor_matmul_so3_3x3`, `fn_tensor_norm3`), so
`models/tensornet.py:54-66` of the reference can import them unchanged.  Each is a
torch.autograd.Function whose forward and backward are hand-written sm_100a kernels behind the C ABI.
fp32 only (the reference also instantiates fp16/fp64; not needed by any config in scope).

Sign convention: A follows the reference's *PyTorch* path (`vector_to_skewtensor`,
models/tensornet.py:84-103), i.e. the transpose of what the Warp radial kernel emits
("Sign convention trap"); identical energies for O(3) models.
"""
from typing import List, Tuple

import torch
from torch import Tensor
from torch.autograd.function import once_differentiable

from .. import _lib

__all__ = ["graph_transform", "fn_radial_message_passing", "fn_message_passing", "fn_compose_tensor",
           "fn_decompose_tensor", "fn_tensor_matmul_o3_3x3", "fn_tensor_matmul_so3_3x3", "fn_tensor_norm3"]

p = _lib.ptr


def _c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        raise ValueError(f"torchmdnet_b200 TensorNet ops are fp32-only, got {t.dtype}")
    _lib.require_cuda(t)
    return t.contiguous()


def _ci(t):
    if t is None:
        return None
    if t.dtype != torch.int32:
        raise ValueError("graph index arrays must be int32")
    return t.contiguous()


def _triple(N, F, ref):
    return (torch.empty(N, 1, F, device=ref.device, dtype=ref.dtype), torch.empty(N, 3, F, device=ref.device, dtype=ref.dtype),
            torch.empty(N, 5, F, device=ref.device, dtype=ref.dtype))


# ------------------------------------------------------------------------------------------ graph
def graph_transform(edge_index: Tensor, num_nodes: int):
    """COO int32 [2,E] (negative entries = padding) -> (row_data,row_indices,row_indptr,col_data,col_indices,col_indptr),
    all int32 (reference: warp_ops/graph_transform.py:160-179).  Slots inside a row are ordered by edge id."""
    _lib.require_cuda(edge_index)
    if edge_index.dtype != torch.int32:
        raise ValueError("graph_transform expects an int32 edge_index")
    ei = edge_index.contiguous()
    E, n, dev = ei.shape[1], int(num_nodes), ei.device
    i32 = dict(device=dev, dtype=torch.int32)
    row_data, row_idx, col_data, col_idx = (torch.zeros(E, **i32) for _ in range(4))
    row_ptr, col_ptr = torch.empty(n + 1, **i32), torch.empty(n + 1, **i32)
    ws = torch.empty(2 * (n + 1), **i32)
    _lib.call("tmd_graph_transform", p(ei), E, n, p(row_data), p(row_idx), p(row_ptr), p(col_data), p(col_idx), p(col_ptr),
              p(ws), device=dev)
    return row_data, row_idx, row_ptr, col_data, col_idx, col_ptr


# ------------------------------------------------------------------------------------------ granular ops
class _Compose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, A, S):
        I, A, S = _c(I), _c(A), _c(S)
        N, F = I.shape[0], I.shape[-1]
        X = torch.empty(N, 3, 3, F, device=I.device, dtype=I.dtype)
        _lib.call("tmd_compose_fwd", p(I), p(A), p(S), p(X), N, F, device=I.device)
        return X

    @staticmethod
    @once_differentiable
    def backward(ctx, dX):
        dX = _c(dX)
        N, F = dX.shape[0], dX.shape[-1]
        dI, dA, dS = _triple(N, F, dX)
        _lib.call("tmd_compose_bwd", p(dX), p(dI), p(dA), p(dS), N, F, device=dX.device)
        return dI, dA, dS


class _Decompose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X):
        X = _c(X)
        N, F = X.shape[0], X.shape[-1]
        I, A, S = _triple(N, F, X)
        _lib.call("tmd_decompose_fwd", p(X), p(I), p(A), p(S), N, F, device=X.device)
        return I, A, S

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        dI, dA, dS = _c(dI), _c(dA), _c(dS)
        N, F = dI.shape[0], dI.shape[-1]
        dX = torch.empty(N, 3, 3, F, device=dI.device, dtype=dI.dtype)
        _lib.call("tmd_decompose_bwd", p(dI), p(dA), p(dS), p(dX), N, F, device=dI.device)
        return dX


def _make_matmul(kind):
    class _MatMul(torch.autograd.Function):
        @staticmethod
        def forward(ctx, X, Y):
            X, Y = _c(X), _c(Y)
            Z = torch.empty_like(X)
            _lib.call(f"tmd_matmul_{kind}_fwd", p(X), p(Y), p(Z), X.shape[0], X.shape[-1], device=X.device)
            ctx.save_for_backward(X, Y)
            return Z

        @staticmethod
        @once_differentiable
        def backward(ctx, dZ):
            X, Y = ctx.saved_tensors
            dZ = _c(dZ)
            dX, dY = torch.empty_like(X), torch.empty_like(Y)
            _lib.call(f"tmd_matmul_{kind}_bwd", p(X), p(Y), p(dZ), p(dX), p(dY), X.shape[0], X.shape[-1], device=X.device)
            return dX, dY

    return _MatMul


_MatMulO3, _MatMulSO3 = _make_matmul("o3"), _make_matmul("so3")


class _Norm3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X):
        X = _c(X)
        N, F = X.shape[0], X.shape[-1]
        out = torch.empty(N, 3 * F, device=X.device, dtype=X.dtype)
        _lib.call("tmd_norm3_fwd", p(X), p(out), N, F, device=X.device)
        ctx.save_for_backward(X)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (X,) = ctx.saved_tensors
        dout = _c(dout)
        dX = torch.empty_like(X)
        _lib.call("tmd_norm3_bwd", p(dout), p(X), p(dX), X.shape[0], X.shape[-1], device=X.device)
        return dX


class _RadialMP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rhat, w, row_data, row_indptr):
        rhat, w = _c(rhat), _c(w)
        row_data, row_indptr = _ci(row_data), _ci(row_indptr)
        N, F = row_indptr.shape[0] - 1, w.shape[-1]
        I, A, S = _triple(N, F, w)
        _lib.call("tmd_radial_mp_fwd", p(rhat), p(w), p(row_data), p(row_indptr), N, F, p(I), p(A), p(S), device=w.device)
        ctx.save_for_backward(rhat, w, row_data, row_indptr)
        ctx.has_data = row_data is not None
        return I, A, S

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        rhat, w, row_data, row_indptr = ctx.saved_tensors
        dI, dA, dS = _c(dI), _c(dA), _c(dS)
        N, F = row_indptr.shape[0] - 1, w.shape[-1]
        dw, dr = torch.zeros_like(w), torch.zeros_like(rhat)  # padded edge slots keep zero gradient
        _lib.call("tmd_radial_mp_bwd", p(rhat), p(w), p(row_data), p(row_indptr), N, F, p(dI), p(dA), p(dS), p(dw), p(dr),
                  device=w.device)
        return dr, dw, None, None


class _MP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, A, S, w, row_data, row_indices, row_indptr, col_data, col_indices, col_indptr):
        I, A, S, w = _c(I), _c(A), _c(S), _c(w)
        row_data, row_indices, row_indptr = _ci(row_data), _ci(row_indices), _ci(row_indptr)
        N, F = I.shape[0], I.shape[-1]
        oI, oA, oS = _triple(N, F, I)
        _lib.call("tmd_mp_fwd", p(I), p(A), p(S), p(w), p(row_data), p(row_indices), p(row_indptr), N, F, p(oI), p(oA),
                  p(oS), device=I.device)
        ctx.save_for_backward(I, A, S, w, _ci(col_data), _ci(col_indices), _ci(col_indptr))
        return oI, oA, oS

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        I, A, S, w, col_data, col_indices, col_indptr = ctx.saved_tensors
        dI, dA, dS = _c(dI), _c(dA), _c(dS)
        N, F = I.shape[0], I.shape[-1]
        gI, gA, gS = _triple(N, F, I)
        need_w = ctx.needs_input_grad[3]
        dw = torch.zeros_like(w) if need_w else None
        _lib.call("tmd_mp_bwd", p(I), p(A), p(S), p(w), p(dI), p(dA), p(dS), p(col_data), p(col_indices), p(col_indptr), N, F,
                  p(gI), p(gA), p(gS), p(dw), device=I.device)
        return gI, gA, gS, dw, None, None, None, None, None, None


def fn_compose_tensor(I: Tensor, A: Tensor, S: Tensor) -> Tensor:
    """(I[N,1,F], A[N,3,F], S[N,5,F]) -> X[N,3,3,F]   (warp_ops/compose_tensor.py)."""
    return _Compose.apply(I, A, S)


def fn_decompose_tensor(X: Tensor) -> List[Tensor]:
    """X[N,3,3,F] -> [I[N,1,F], A[N,3,F], S[N,5,F]]   (warp_ops/decompose_tensor.py)."""
    return list(_Decompose.apply(X))


def fn_tensor_matmul_o3_3x3(X: Tensor, Y: Tensor) -> Tensor:
    """X Y + Y X per (atom, channel)   (warp_ops/equivariant_o3_matmul.py)."""
    return _MatMulO3.apply(X, Y)


def fn_tensor_matmul_so3_3x3(X: Tensor, Y: Tensor) -> Tensor:
    """X Y per (atom, channel)   (warp_ops/equivariant_so3_matmul.py)."""
    return _MatMulSO3.apply(X, Y)


def fn_tensor_norm3(X: Tensor) -> Tensor:
    """X[N,3,3,F] -> [N,3F] = (tr^2/3, |A|^2, |S|^2)   (warp_ops/tensor_norm3.py)."""
    return _Norm3.apply(X)


def fn_radial_message_passing(edge_vec_norm: Tensor, edge_attr: Tensor, row_data: Tensor, row_indptr: Tensor) -> List[Tensor]:
    """Embedding message passing (warp_ops/tensornet_radial_mp.py:37-118): edge_vec_norm[E,3],
    edge_attr[E,3,F], CSR (row_data = edge ids, row_indptr[N+1]) -> [I[N,1,F], A[N,3,F], S[N,5,F]]."""
    return list(_RadialMP.apply(edge_vec_norm, edge_attr, row_data, row_indptr))


def fn_message_passing(I, A, S, edge_attr, row_data, row_indices, row_indptr, col_data, col_indices, col_indptr) -> List[Tensor]:
    """Interaction message passing (warp_ops/tensornet_mp.py:37-111): weighted gather-sum over the CSR
    rows; the CSC arrays drive the atomic-free backward."""
    return list(_MP.apply(I, A, S, edge_attr, row_data, row_indices, row_indptr, col_data, col_indices, col_indptr))


# ------------------------------------------------------------------------------------------ fused compact stages
def _flat3(I, A, S):
    return _c(I), _c(A), _c(S)


class _Norm2C(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, A, S):
        I, A, S = _flat3(I, A, S)
        N, F = I.shape[0], I.shape[-1]
        out = torch.empty(N, F, device=I.device, dtype=I.dtype)
        _lib.call("tmd_norm2c_fwd", p(I), p(A), p(S), p(out), N, F, device=I.device)
        ctx.save_for_backward(I, A, S)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        I, A, S = ctx.saved_tensors
        N, F = I.shape[0], I.shape[-1]
        g = _triple(N, F, I)
        _lib.call("tmd_norm2c_bwd", p(I), p(A), p(S), p(_c(dout)), p(g[0]), p(g[1]), p(g[2]), N, F, device=I.device)
        return g


class _Norm3C(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, A, S):
        I, A, S = _flat3(I, A, S)
        N, F = I.shape[0], I.shape[-1]
        out = torch.empty(N, 3 * F, device=I.device, dtype=I.dtype)
        _lib.call("tmd_norm3c_fwd", p(I), p(A), p(S), p(out), N, F, device=I.device)
        ctx.save_for_backward(I, A, S)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        I, A, S = ctx.saved_tensors
        N, F = I.shape[0], I.shape[-1]
        g = _triple(N, F, I)
        _lib.call("tmd_norm3c_bwd", p(I), p(A), p(S), p(_c(dout)), p(g[0]), p(g[1]), p(g[2]), N, F, device=I.device)
        return g


class _Gate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, A, S, g):
        I, A, S = _flat3(I, A, S)
        g = _c(g)
        N, F = I.shape[0], I.shape[-1]
        o = _triple(N, F, I)
        _lib.call("tmd_gate_fwd", p(I), p(A), p(S), p(g), p(o[0]), p(o[1]), p(o[2]), N, F, device=I.device)
        ctx.save_for_backward(I, A, S, g)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        I, A, S, g = ctx.saved_tensors
        N, F = I.shape[0], I.shape[-1]
        o = _triple(N, F, I)
        dg = torch.empty_like(g)
        _lib.call("tmd_gate_bwd", p(I), p(A), p(S), p(g), p(_c(dI)), p(_c(dA)), p(_c(dS)), p(o[0]), p(o[1]), p(o[2]), p(dg),
                  N, F, device=I.device)
        return o[0], o[1], o[2], dg


class _P1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, A, S):
        I, A, S = _flat3(I, A, S)
        N, F = I.shape[0], I.shape[-1]
        o = _triple(N, F, I)
        _lib.call("tmd_p1_fwd", p(I), p(A), p(S), p(o[0]), p(o[1]), p(o[2]), N, F, device=I.device)
        ctx.save_for_backward(I, A, S)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        I, A, S = ctx.saved_tensors
        N, F = I.shape[0], I.shape[-1]
        g = _triple(N, F, I)
        _lib.call("tmd_p1_bwd", p(I), p(A), p(S), p(_c(dI)), p(_c(dA)), p(_c(dS)), p(g[0]), p(g[1]), p(g[2]), N, F,
                  device=I.device)
        return g


class _P2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, YI, YA, YS, MI, MA, MS, qf, o3):
        YI, YA, YS = _flat3(YI, YA, YS)
        MI, MA, MS = _flat3(MI, MA, MS)
        qf = _c(qf)
        N, F = YI.shape[0], YI.shape[-1]
        o = _triple(N, F, YI)
        _lib.call("tmd_p2_fwd", p(YI), p(YA), p(YS), p(MI), p(MA), p(MS), p(qf), int(o3), p(o[0]), p(o[1]), p(o[2]), N, F,
                  device=YI.device)
        ctx.save_for_backward(YI, YA, YS, MI, MA, MS, qf)
        ctx.o3 = int(o3)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        YI, YA, YS, MI, MA, MS, qf = ctx.saved_tensors
        N, F = YI.shape[0], YI.shape[-1]
        gy, gm = _triple(N, F, YI), _triple(N, F, YI)
        _lib.call("tmd_p2_bwd", p(YI), p(YA), p(YS), p(MI), p(MA), p(MS), p(qf), ctx.o3, p(_c(dI)), p(_c(dA)), p(_c(dS)),
                  p(gy[0]), p(gy[1]), p(gy[2]), p(gm[0]), p(gm[1]), p(gm[2]), N, F, device=YI.device)
        return gy[0], gy[1], gy[2], gm[0], gm[1], gm[2], None, None


class _P3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, XI, XA, XS, DI, DA, DS, qf):
        XI, XA, XS = _flat3(XI, XA, XS)
        DI, DA, DS = _flat3(DI, DA, DS)
        qf = _c(qf)
        N, F = XI.shape[0], XI.shape[-1]
        o = _triple(N, F, XI)
        _lib.call("tmd_p3_fwd", p(XI), p(XA), p(XS), p(DI), p(DA), p(DS), p(qf), p(o[0]), p(o[1]), p(o[2]), N, F,
                  device=XI.device)
        ctx.save_for_backward(DI, DA, DS, qf)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, dI, dA, dS):
        DI, DA, DS, qf = ctx.saved_tensors
        N, F = DI.shape[0], DI.shape[-1]
        gx, gd = _triple(N, F, DI), _triple(N, F, DI)
        _lib.call("tmd_p3_bwd", p(DI), p(DA), p(DS), p(qf), p(_c(dI)), p(_c(dA)), p(_c(dS)), p(gx[0]), p(gx[1]), p(gx[2]),
                  p(gd[0]), p(gd[1]), p(gd[2]), N, F, device=DI.device)
        return gx[0], gx[1], gx[2], gd[0], gd[1], gd[2], None


def norm2_compact(I, A, S) -> Tensor:
    """|X|^2 per channel from compact irreps -> [N,F]   (tensor_norm of the composed tensor, tensornet.py:144-146)."""
    return _Norm2C.apply(I, A, S)


def norm3_compact(I, A, S) -> Tensor:
    """(3 I^2, |A|^2, |S|^2) -> [N,3F]   (tensornet.py:384-398 tail)."""
    return _Norm3C.apply(I, A, S)


def gate_compact(I, A, S, g) -> Tuple[Tensor, Tensor, Tensor]:
    """(I*g0, A*g1, S*g2), g [N,3,F]   (tensornet.py:607-615)."""
    return _Gate.apply(I, A, S, g)


def interaction_normalize(I, A, S):
    """P1: X/(|X|^2+1) in compact form   (tensornet.py:745)."""
    return _P1.apply(I, A, S)


def interaction_product(YI, YA, YS, MI, MA, MS, qf, o3: bool):
    """P2: decompose(C)/(|C|^2+1) with C = qf (Y M + M Y) [O(3)] or 2 Y M [SO(3)]   (tensornet.py:786-806)."""
    return _P2.apply(YI, YA, YS, MI, MA, MS, qf, o3)


def interaction_update(XI, XA, XS, DI, DA, DS, qf):
    """P3: X + dX + qf dX.dX in compact form   (tensornet.py:812)."""
    return _P3.apply(XI, XA, XS, DI, DA, DS, qf)
