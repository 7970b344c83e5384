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


# ------------------------------------------------------------------------------------------ double backward
# Training with a force loss differentiates THROUGH the force evaluation (`autograd.grad(..., create_graph=True)`,
# models/model.py:618-628 of the reference, whose ops chain fwd -> bwd -> bwd_bwd primitives).  Every Function below therefore
# has two backward branches: grad mode off (inference, the hand-written backward kernel) and grad mode on (create_graph): the
# gradient is produced by DIFFERENTIABLE operations -- either the same kernels re-entered through autograd Functions when the op is
# (bi)linear (compose / decompose / 3x3 products / message passing: `_MPBwd`; any order), or a torch restatement of the
# per-(atom, channel) formula whose second derivatives torch builds itself (`_DDFn`: second order, what a force loss needs).
def _needs_graph(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class _DDFn(torch.autograd.Function):
    """H(inputs, grads) = d<ref_fn(inputs), grads>/d inputs as ONE autograd node.

    The torch restatement is evaluated on detached leaf copies, so H holds the PARTIAL derivatives w.r.t. each input: several
    stages receive inputs that depend on one another (P2 gets (Y, M) with M = message_passing(Y); P3 gets (X, dX) with dX a function
    of X), and differentiating the restatement directly on the original tensors would let autograd walk from one input into the
    history of the other and add that path a second time.  The backward of this node is the vector-Jacobian product of H w.r.t.
    (inputs, grads) taken through the graph built in the forward (second order; a third derivative raises)."""

    @staticmethod
    def forward(ctx, ref_fn, n_in, *args):
        ins = [a.detach().requires_grad_(True) for a in args[:n_in]]
        gs = [a.detach().requires_grad_(True) for a in args[n_in:]]
        with torch.enable_grad():
            outs = ref_fn(*ins)
            outs = list(outs) if isinstance(outs, (tuple, list)) else [outs]
            g = torch.autograd.grad(outs, ins, gs, create_graph=True, allow_unused=True)
        ctx.leaves, ctx.g = ins + gs, g
        ctx.set_materialize_grads(False)
        return tuple(torch.zeros_like(t) if x is None else x.detach() for x, t in zip(g, ins))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gg):
        pairs = [(x, y) for x, y in zip(ctx.g, gg) if x is not None and y is not None and x.requires_grad]
        if not pairs:
            return (None, None) + (None,) * len(ctx.leaves)
        with torch.enable_grad():
            res = torch.autograd.grad([x for x, _ in pairs], ctx.leaves, [y for _, y in pairs], allow_unused=True, retain_graph=True)
        return (None, None) + tuple(res)


def _dd(ref_fn, inputs, grads):
    """create_graph branch of the per-(atom, channel) stages: the first-order gradient as a differentiable function of the
    Function's original inputs and of the incoming gradients (see _DDFn); None for inputs that need no gradient."""
    g = _DDFn.apply(ref_fn, len(inputs), *inputs, *grads)
    return tuple(gi if t.requires_grad else None for gi, t in zip(g, inputs))


def _t_compose(I, A, S):
    """torch restatement of tn::compose (csrc/tn_math.h): compact irreps -> X [N,3,3,F]"""
    i = I[:, 0]
    a01, a02, a12 = A.unbind(1)
    s00, s01, s02, s11, s12 = S.unbind(1)
    rows = [i + s00, a01 + s01, a02 + s02, s01 - a01, i + s11, a12 + s12, s02 - a02, s12 - a12, i - (s00 + s11)]
    return torch.stack(rows, 1).view(I.shape[0], 3, 3, I.shape[-1])


def _t_decompose(X):
    tr3 = (X[:, 0, 0] + X[:, 1, 1] + X[:, 2, 2]) / 3
    I = tr3.unsqueeze(1)
    A = torch.stack([(X[:, 0, 1] - X[:, 1, 0]) * 0.5, (X[:, 0, 2] - X[:, 2, 0]) * 0.5, (X[:, 1, 2] - X[:, 2, 1]) * 0.5], 1)
    S = torch.stack([X[:, 0, 0] - tr3, (X[:, 0, 1] + X[:, 1, 0]) * 0.5, (X[:, 0, 2] + X[:, 2, 0]) * 0.5, X[:, 1, 1] - tr3,
                     (X[:, 1, 2] + X[:, 2, 1]) * 0.5], 1)
    return I, A, S


def _t_norm3(I, A, S):
    s22 = S[:, 0] + S[:, 3]
    nS = S[:, 0] ** 2 + S[:, 3] ** 2 + s22 ** 2 + 2 * (S[:, 1] ** 2 + S[:, 2] ** 2 + S[:, 4] ** 2)
    return 3 * I[:, 0] ** 2, 2 * (A ** 2).sum(1), nS


def _t_norm2(I, A, S):
    a, b, c = _t_norm3(I, A, S)
    return a + b + c


def _t_mm(X, Y):
    return torch.einsum("nikf,nkjf->nijf", X, Y)


def _t_p1(I, A, S):
    s = 1.0 / (_t_norm2(I, A, S) + 1.0)
    return I * s.unsqueeze(1), A * s.unsqueeze(1), S * s.unsqueeze(1)


def _t_p2(YI, YA, YS, MI, MA, MS, qf, o3):
    Y, M = _t_compose(YI, YA, YS), _t_compose(MI, MA, MS)
    C = (qf.view(-1, 1, 1, 1) * (_t_mm(Y, M) + _t_mm(M, Y))) if o3 else 2 * _t_mm(Y, M)
    return _t_p1(*_t_decompose(C))


def _t_p3(XI, XA, XS, DI, DA, DS, qf):
    D = _t_compose(DI, DA, DS)
    QI, QA, QS = _t_decompose(_t_mm(D, D))
    q = qf.view(-1, 1, 1)
    return XI + DI + q * QI, XA + DA + q * QA, XS + DS + q * QS


def _row_of_slot(indptr, E):
    """row id of every CSR slot (slots past indptr[-1] -> clamped to the last row, their weights are zero-padded)"""
    k = torch.arange(E, device=indptr.device, dtype=indptr.dtype)
    return torch.bucketize(k, indptr[1:], right=True).clamp_(max=indptr.shape[0] - 2).long()


def _t_radial_mp(rhat, w, row_data, row_indptr):
    """torch restatement of the embedding message passing (tensornet.py:405-445, PyTorch sign convention)"""
    N, F = row_indptr.shape[0] - 1, w.shape[-1]
    nslot = int(row_indptr[-1].item()) if row_data is None else row_data.shape[0]
    nslot = min(nslot, w.shape[0]) if row_data is None else nslot
    row = _row_of_slot(row_indptr, nslot)
    e = torch.arange(nslot, device=w.device) if row_data is None else row_data.long()
    valid = (torch.arange(nslot, device=w.device) < row_indptr[-1]).to(w.dtype)
    r, ww = rhat.index_select(0, e), w.index_select(0, e) * valid.view(-1, 1, 1)
    a = torch.stack([-r[:, 2], r[:, 1], -r[:, 0]], 1)
    m = (r * r).sum(1) / 3
    sb = torch.stack([r[:, 0] ** 2 - m, r[:, 0] * r[:, 1], r[:, 0] * r[:, 2], r[:, 1] ** 2 - m, r[:, 1] * r[:, 2]], 1)
    I = torch.zeros(N, 1, F, device=w.device, dtype=w.dtype).index_add(0, row, ww[:, 0:1])
    A = torch.zeros(N, 3, F, device=w.device, dtype=w.dtype).index_add(0, row, ww[:, 1:2] * a.unsqueeze(-1))
    S = torch.zeros(N, 5, F, device=w.device, dtype=w.dtype).index_add(0, row, ww[:, 2:3] * sb.unsqueeze(-1))
    return I, A, S


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
    def backward(ctx, dX):
        if _needs_graph(dX):
            return _ComposeT.apply(dX)       # the transpose map, itself differentiable (its backward is _Compose)
        dX = _c(dX)
        N, F = dX.shape[0], dX.shape[-1]
        dI, dA, dS = _triple(N, F, dX)
        _lib.call("tmd_compose_bwd", p(dX), p(dI), p(dA), p(dS), N, F, device=dX.device)
        return dI, dA, dS


class _ComposeT(torch.autograd.Function):
    """(dI, dA, dS) = compose^T dX  (kernel tmd_compose_bwd); backward = compose."""

    @staticmethod
    def forward(ctx, dX):
        dX = _c(dX)
        N, F = dX.shape[0], dX.shape[-1]
        dI, dA, dS = _triple(N, F, dX)
        _lib.call("tmd_compose_bwd", p(dX), p(dI), p(dA), p(dS), N, F, device=dX.device)
        return dI, dA, dS

    @staticmethod
    def backward(ctx, gI, gA, gS):
        return _Compose.apply(gI, gA, gS)


class _Decompose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X):
        X = _c(X)
        N, F = X.shape[0], X.shape[-1]
        I, A, S = _triple(N, F, X)
        _lib.call("tmd_decompose_fwd", p(X), p(I), p(A), p(S), N, F, device=X.device)
        return I, A, S

    @staticmethod
    def backward(ctx, dI, dA, dS):
        if _needs_graph(dI, dA, dS):
            return _DecomposeT.apply(dI, dA, dS)
        dI, dA, dS = _c(dI), _c(dA), _c(dS)
        N, F = dI.shape[0], dI.shape[-1]
        dX = torch.empty(N, 3, 3, F, device=dI.device, dtype=dI.dtype)
        _lib.call("tmd_decompose_bwd", p(dI), p(dA), p(dS), p(dX), N, F, device=dI.device)
        return dX


class _DecomposeT(torch.autograd.Function):
    """dX = decompose^T (dI, dA, dS)  (kernel tmd_decompose_bwd); backward = decompose."""

    @staticmethod
    def forward(ctx, dI, dA, dS):
        dI, dA, dS = _c(dI), _c(dA), _c(dS)
        N, F = dI.shape[0], dI.shape[-1]
        dX = torch.empty(N, 3, 3, F, device=dI.device, dtype=dI.dtype)
        _lib.call("tmd_decompose_bwd", p(dI), p(dA), p(dS), p(dX), N, F, device=dI.device)
        return dX

    @staticmethod
    def backward(ctx, gX):
        return _Decompose.apply(gX)


def _make_matmul(kind):
    class _MatMul(torch.autograd.Function):
        @staticmethod
        def forward(ctx, X, Y):
            ctx.save_for_backward(X, Y)
            X, Y = _c(X), _c(Y)
            Z = torch.empty_like(X)
            _lib.call(f"tmd_matmul_{kind}_fwd", p(X), p(Y), p(Z), X.shape[0], X.shape[-1], device=X.device)
            return Z

        @staticmethod
        def backward(ctx, dZ):
            X, Y = ctx.saved_tensors
            if _needs_graph(dZ, X, Y):
                # o3: Z = XY + YX -> dX = dZ Y^T + Y^T dZ = o3(dZ, Y^T), dY = o3(X^T, dZ);  so3: Z = XY -> dX = so3(dZ, Y^T),
                # dY = so3(X^T, dZ): the forward kernel re-entered through this Function (transposes are torch views)
                return _MatMul.apply(dZ, Y.transpose(1, 2)), _MatMul.apply(X.transpose(1, 2), dZ)
            X, Y, dZ = _c(X), _c(Y), _c(dZ)
            dX, dY = torch.empty_like(X), torch.empty_like(Y)
            _lib.call(f"tmd_matmul_{kind}_bwd", p(X), p(Y), p(dZ), p(dX), p(dY), X.shape[0], X.shape[-1], device=X.device)
            return dX, dY

    return _MatMul


_MatMulO3, _MatMulSO3 = _make_matmul("o3"), _make_matmul("so3")


class _Norm3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X):
        ctx.save_for_backward(X)
        X = _c(X)
        N, F = X.shape[0], X.shape[-1]
        out = torch.empty(N, 3 * F, device=X.device, dtype=X.dtype)
        _lib.call("tmd_norm3_fwd", p(X), p(out), N, F, device=X.device)
        return out

    @staticmethod
    def backward(ctx, dout):
        (X,) = ctx.saved_tensors
        if _needs_graph(dout, X):
            # the decomposition re-enters the kernel-backed (closed) _Decompose; the three quadratic forms are torch elementwise
            return _dd(lambda x: torch.cat(_t_norm3(*_Decompose.apply(x)), -1), (X,), (dout,))[0]
        X, dout = _c(X), _c(dout)
        dX = torch.empty_like(X)
        _lib.call("tmd_norm3_bwd", p(dout), p(X), p(dX), X.shape[0], X.shape[-1], device=X.device)
        return dX


class _RadialMP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rhat, w, row_data, row_indptr):
        ctx.save_for_backward(rhat, w, row_data, row_indptr)
        rhat, w = _c(rhat), _c(w)
        row_data, row_indptr = _ci(row_data), _ci(row_indptr)
        N, F = row_indptr.shape[0] - 1, w.shape[-1]
        I, A, S = _triple(N, F, w)
        _lib.call("tmd_radial_mp_fwd", p(rhat), p(w), p(row_data), p(row_indptr), N, F, p(I), p(A), p(S), device=w.device)
        return I, A, S

    @staticmethod
    def backward(ctx, dI, dA, dS):
        rhat, w, row_data, row_indptr = ctx.saved_tensors
        if _needs_graph(dI, dA, dS, rhat, w):
            g = _dd(lambda r, ww: _t_radial_mp(r, ww, row_data, row_indptr), (rhat, w), (dI, dA, dS))
            return g[0], g[1], None, None
        rhat, w = _c(rhat), _c(w)
        row_data, row_indptr = _ci(row_data), _ci(row_indptr)
        dI, dA, dS = _c(dI), _c(dA), _c(dS)
        N, F = row_indptr.shape[0] - 1, w.shape[-1]
        dw, dr = torch.zeros_like(w), torch.zeros_like(rhat)  # padded edge slots keep zero gradient
        _lib.call("tmd_radial_mp_bwd", p(rhat), p(w), p(row_data), p(row_indptr), N, F, p(dI), p(dA), p(dS), p(dw), p(dr),
                  device=w.device)
        return dr, dw, None, None


def _mp_fwd_call(I, A, S, w, data, indices, indptr):
    I, A, S, w = _c(I), _c(A), _c(S), _c(w)
    N, F = I.shape[0], I.shape[-1]
    oI, oA, oS = _triple(N, F, I)
    _lib.call("tmd_mp_fwd", p(I), p(A), p(S), p(w), p(_ci(data)), p(_ci(indices)), p(_ci(indptr)), N, F, p(oI), p(oA), p(oS),
              device=I.device)
    return oI, oA, oS


def _mp_bwd_call(I, A, S, w, dI, dA, dS, col_data, col_indices, col_indptr, need_w=True):
    I, A, S, w, dI, dA, dS = _c(I), _c(A), _c(S), _c(w), _c(dI), _c(dA), _c(dS)
    N, F = I.shape[0], I.shape[-1]
    gI, gA, gS = _triple(N, F, I)
    dw = torch.zeros_like(w) if need_w else None
    _lib.call("tmd_mp_bwd", p(I), p(A), p(S), p(w), p(dI), p(dA), p(dS), p(_ci(col_data)), p(_ci(col_indices)), p(_ci(col_indptr)), N, F,
              p(gI), p(gA), p(gS), p(dw), device=I.device)
    return gI, gA, gS, dw


class _MP(torch.autograd.Function):
    """out = G(T, w): out[b] = sum_{k in row b} w[e_k] * T[j_k]   (bilinear in the node tensors T and the edge weights w)"""

    @staticmethod
    def forward(ctx, I, A, S, w, row_data, row_indices, row_indptr, col_data, col_indices, col_indptr):
        ctx.save_for_backward(I, A, S, w, row_data, row_indices, row_indptr, col_data, col_indices, col_indptr)
        return _mp_fwd_call(I, A, S, w, row_data, row_indices, row_indptr)

    @staticmethod
    def backward(ctx, dI, dA, dS):
        I, A, S, w, row_data, row_indices, row_indptr, col_data, col_indices, col_indptr = ctx.saved_tensors
        if _needs_graph(dI, dA, dS, I, A, S, w):
            g = _MPBwd.apply(I, A, S, w, dI, dA, dS, row_data, row_indices, row_indptr, col_data, col_indices, col_indptr)
            return g[0], g[1], g[2], g[3], None, None, None, None, None, None
        gI, gA, gS, dw = _mp_bwd_call(I, A, S, w, dI, dA, dS, col_data, col_indices, col_indptr, ctx.needs_input_grad[3])
        return gI, gA, gS, dw, None, None, None, None, None, None


class _MPBwd(torch.autograd.Function):
    r"""(gT, gw) = B(T, w, dOut): gT[j] = sum_{e: j} w_e dOut[i_e] (the transposed gather), gw_e = <dOut[i_e], T[j_e]> per irrep
    (kernel tmd_mp_bwd).  Its own backward -- the reference's message_passing_bwd_bwd (warp_kernels/tensornet_mp.py:133-245) --
    only needs the same two kernels again, re-entered through _MP / _MPBwd:
        d/dT     = transposed gather of ggw (x) dOut          \ one tmd_mp_bwd launch with (T, w) := (ggT, ggw)
        d/dw     = <dOut[i], ggT[j]>                           /
        d/ddOut  = G(ggT, w) + G(T, ggw)                        two tmd_mp_fwd launches"""

    @staticmethod
    def forward(ctx, I, A, S, w, dI, dA, dS, row_data, row_indices, row_indptr, col_data, col_indices, col_indptr):
        ctx.save_for_backward(I, A, S, w, dI, dA, dS, row_data, row_indices, row_indptr, col_data, col_indices, col_indptr)
        return _mp_bwd_call(I, A, S, w, dI, dA, dS, col_data, col_indices, col_indptr, True)

    @staticmethod
    def backward(ctx, ggI, ggA, ggS, ggw):
        I, A, S, w, dI, dA, dS, row_data, row_indices, row_indptr, col_data, col_indices, col_indptr = ctx.saved_tensors
        z = lambda g, like: torch.zeros_like(like) if g is None else g
        ggI, ggA, ggS, ggw = z(ggI, I), z(ggA, A), z(ggS, S), z(ggw, w)
        graph = (row_data, row_indices, row_indptr, col_data, col_indices, col_indptr)
        if _needs_graph(ggI, ggA, ggS, ggw, I, A, S, w, dI, dA, dS):      # third order and beyond: stay inside autograd Functions
            dT = _MPBwd.apply(ggI, ggA, ggS, ggw, dI, dA, dS, *graph)
            o1, o2 = _MP.apply(ggI, ggA, ggS, w, *graph), _MP.apply(I, A, S, ggw, *graph)
        else:
            dT = _mp_bwd_call(ggI, ggA, ggS, ggw, dI, dA, dS, col_data, col_indices, col_indptr, True)
            o1, o2 = _mp_fwd_call(ggI, ggA, ggS, w, row_data, row_indices, row_indptr), _mp_fwd_call(I, A, S, ggw, row_data, row_indices, row_indptr)
        return (dT[0], dT[1], dT[2], dT[3], o1[0] + o2[0], o1[1] + o2[1], o1[2] + o2[2], None, None, None, None, None, None)


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
        ctx.save_for_backward(I, A, S)
        I, A, S = _flat3(I, A, S)
        N, F = I.shape[0], I.shape[-1]
        out = torch.empty(N, F, device=I.device, dtype=I.dtype)
        _lib.call("tmd_norm2c_fwd", p(I), p(A), p(S), p(out), N, F, device=I.device)
        return out

    @staticmethod
    def backward(ctx, dout):
        I, A, S = ctx.saved_tensors
        if _needs_graph(dout, I, A, S):
            return _dd(_t_norm2, (I, A, S), (dout,))
        I, A, S = _flat3(I, A, S)
        N, F = I.shape[0], I.shape[-1]
        g = _triple(N, F, I)
        _lib.call("tmd_norm2c_bwd", p(I), p(A), p(S), p(_c(dout)), p(g[0]), p(g[1]), p(g[2]), N, F, device=I.device)
        return g


class _Norm3C(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, A, S):
        ctx.save_for_backward(I, A, S)
        I, A, S = _flat3(I, A, S)
        N, F = I.shape[0], I.shape[-1]
        out = torch.empty(N, 3 * F, device=I.device, dtype=I.dtype)
        _lib.call("tmd_norm3c_fwd", p(I), p(A), p(S), p(out), N, F, device=I.device)
        return out

    @staticmethod
    def backward(ctx, dout):
        I, A, S = ctx.saved_tensors
        if _needs_graph(dout, I, A, S):
            return _dd(lambda i, a, s_: torch.cat(_t_norm3(i, a, s_), -1), (I, A, S), (dout,))
        I, A, S = _flat3(I, A, S)
        N, F = I.shape[0], I.shape[-1]
        g = _triple(N, F, I)
        _lib.call("tmd_norm3c_bwd", p(I), p(A), p(S), p(_c(dout)), p(g[0]), p(g[1]), p(g[2]), N, F, device=I.device)
        return g


class _Gate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, A, S, g):
        ctx.save_for_backward(I, A, S, g)
        I, A, S = _flat3(I, A, S)
        g = _c(g)
        N, F = I.shape[0], I.shape[-1]
        o = _triple(N, F, I)
        _lib.call("tmd_gate_fwd", p(I), p(A), p(S), p(g), p(o[0]), p(o[1]), p(o[2]), N, F, device=I.device)
        return o

    @staticmethod
    def backward(ctx, dI, dA, dS):
        I, A, S, g = ctx.saved_tensors
        if _needs_graph(dI, dA, dS, I, A, S, g):
            return _dd(lambda i, a, s_, gg: (i * gg[:, 0:1], a * gg[:, 1:2], s_ * gg[:, 2:3]), (I, A, S, g), (dI, dA, dS))
        I, A, S = _flat3(I, A, S)
        g = _c(g)
        N, F = I.shape[0], I.shape[-1]
        o = _triple(N, F, I)
        dg = torch.empty_like(g)
        _lib.call("tmd_gate_bwd", p(I), p(A), p(S), p(g), p(_c(dI)), p(_c(dA)), p(_c(dS)), p(o[0]), p(o[1]), p(o[2]), p(dg),
                  N, F, device=I.device)
        return o[0], o[1], o[2], dg


class _P1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, A, S):
        ctx.save_for_backward(I, A, S)
        I, A, S = _flat3(I, A, S)
        N, F = I.shape[0], I.shape[-1]
        o = _triple(N, F, I)
        _lib.call("tmd_p1_fwd", p(I), p(A), p(S), p(o[0]), p(o[1]), p(o[2]), N, F, device=I.device)
        return o

    @staticmethod
    def backward(ctx, dI, dA, dS):
        I, A, S = ctx.saved_tensors
        if _needs_graph(dI, dA, dS, I, A, S):
            return _dd(_t_p1, (I, A, S), (dI, dA, dS))
        I, A, S = _flat3(I, A, S)
        N, F = I.shape[0], I.shape[-1]
        g = _triple(N, F, I)
        _lib.call("tmd_p1_bwd", p(I), p(A), p(S), p(_c(dI)), p(_c(dA)), p(_c(dS)), p(g[0]), p(g[1]), p(g[2]), N, F,
                  device=I.device)
        return g


class _P2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, YI, YA, YS, MI, MA, MS, qf, o3):
        ctx.save_for_backward(YI, YA, YS, MI, MA, MS, qf)
        YI, YA, YS = _flat3(YI, YA, YS)
        MI, MA, MS = _flat3(MI, MA, MS)
        qf = _c(qf)
        N, F = YI.shape[0], YI.shape[-1]
        o = _triple(N, F, YI)
        _lib.call("tmd_p2_fwd", p(YI), p(YA), p(YS), p(MI), p(MA), p(MS), p(qf), int(o3), p(o[0]), p(o[1]), p(o[2]), N, F,
                  device=YI.device)
        ctx.o3 = int(o3)
        return o

    @staticmethod
    def backward(ctx, dI, dA, dS):
        YI, YA, YS, MI, MA, MS, qf = ctx.saved_tensors
        if _needs_graph(dI, dA, dS, YI, YA, YS, MI, MA, MS, qf):
            g = _dd(lambda *t: _t_p2(*t, bool(ctx.o3)), (YI, YA, YS, MI, MA, MS, qf), (dI, dA, dS))
            return g + (None,)
        YI, YA, YS = _flat3(YI, YA, YS)
        MI, MA, MS = _flat3(MI, MA, MS)
        qf = _c(qf)
        N, F = YI.shape[0], YI.shape[-1]
        gy, gm = _triple(N, F, YI), _triple(N, F, YI)
        _lib.call("tmd_p2_bwd", p(YI), p(YA), p(YS), p(MI), p(MA), p(MS), p(qf), ctx.o3, p(_c(dI)), p(_c(dA)), p(_c(dS)),
                  p(gy[0]), p(gy[1]), p(gy[2]), p(gm[0]), p(gm[1]), p(gm[2]), N, F, device=YI.device)
        return gy[0], gy[1], gy[2], gm[0], gm[1], gm[2], None, None


class _P3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, XI, XA, XS, DI, DA, DS, qf):
        ctx.save_for_backward(XI, XA, XS, DI, DA, DS, qf)
        XI, XA, XS = _flat3(XI, XA, XS)
        DI, DA, DS = _flat3(DI, DA, DS)
        qf = _c(qf)
        N, F = XI.shape[0], XI.shape[-1]
        o = _triple(N, F, XI)
        _lib.call("tmd_p3_fwd", p(XI), p(XA), p(XS), p(DI), p(DA), p(DS), p(qf), p(o[0]), p(o[1]), p(o[2]), N, F,
                  device=XI.device)
        return o

    @staticmethod
    def backward(ctx, dI, dA, dS):
        XI, XA, XS, DI, DA, DS, qf = ctx.saved_tensors
        if _needs_graph(dI, dA, dS, XI, XA, XS, DI, DA, DS, qf):
            return _dd(_t_p3, (XI, XA, XS, DI, DA, DS, qf), (dI, dA, dS))
        DI, DA, DS = _flat3(DI, DA, DS)
        qf = _c(qf)
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
