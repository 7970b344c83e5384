"""CPU: the torch restatements that back the create_graph (double-backward) branches of extensions/tensornet_ops.py agree with
the oracle's formulas (oracle/tensornet_oracle.py, pinned to the reference) -- value, first and second derivatives."""
import torch

import tensornet_oracle as orc
from torchmdnet_b200.extensions import tensornet_ops as T

torch.manual_seed(0)
N, F = 7, 6


def compact():
    return (torch.randn(N, 1, F, dtype=torch.float64, requires_grad=True), torch.randn(N, 3, F, dtype=torch.float64, requires_grad=True),
            torch.randn(N, 5, F, dtype=torch.float64, requires_grad=True))


def oracle_X(I, A, S):
    """compact irreps -> the oracle's full tensors: I [N,F], A / S [N,3,3,F]"""
    X = T._t_compose(I, A, S)
    return orc.decompose(X), X


def close(a, b, tol=1e-11):
    assert (a - b).abs().max() <= tol * (1 + b.abs().max()), (a - b).abs().max()


def test_compose_decompose_norms():
    I, A, S = compact()
    (Io, Ao, So), X = oracle_X(I, A, S)
    close(orc.compose(Io, Ao, So), X)
    close(T._t_norm2(I, A, S), orc.tnorm(X))
    n = T._t_norm3(I, A, S)
    close(n[0], 3 * Io ** 2); close(n[1], orc.tnorm(Ao)); close(n[2], orc.tnorm(So))
    I2, A2, S2 = T._t_decompose(X)
    close(I2, I); close(A2, A); close(S2, S)


def test_p_stages_value_and_second_derivative():
    Y, M = compact(), compact()
    qf = torch.rand(N, dtype=torch.float64) + 0.5
    for o3 in (True, False):
        out = T._t_p2(*Y, *M, qf, o3)
        Yf, Mf = T._t_compose(*Y), T._t_compose(*M)
        C = qf.view(-1, 1, 1, 1) * (orc.mm3(Mf, Yf) + orc.mm3(Yf, Mf)) if o3 else 2 * orc.mm3(Yf, Mf)
        Io, Ao, So = orc.decompose(C)
        den = (orc.tnorm(C) + 1)
        ref = T._t_decompose(orc.compose(Io / den, Ao / den[:, None, None, :], So / den[:, None, None, :]))
        for a, b in zip(out, ref):
            close(a, b)
        # second derivative of a scalar of the outputs w.r.t. the inputs: same graph both ways, checked through autograd
        w = [torch.randn_like(t) for t in out]
        g1 = torch.autograd.grad(sum((a * b).sum() for a, b in zip(out, w)), Y + M, create_graph=True)
        g2 = torch.autograd.grad(sum((a * b).sum() for a, b in zip(ref, w)), Y + M, create_graph=True)
        h1 = torch.autograd.grad(sum((g ** 2).sum() for g in g1), Y + M)
        h2 = torch.autograd.grad(sum((g ** 2).sum() for g in g2), Y + M)
        for a, b in zip(h1, h2):
            close(a, b, 1e-9)
    X, D = compact(), compact()
    out = T._t_p3(*X, *D, qf)
    Xf, Df = T._t_compose(*X), T._t_compose(*D)
    ref = T._t_decompose(Xf + Df + qf.view(-1, 1, 1, 1) * orc.mm3(Df, Df))
    for a, b in zip(out, ref):
        close(a, b)
    out = T._t_p1(*X)
    ref = T._t_decompose(Xf / (orc.tnorm(Xf) + 1)[:, None, None, :])
    for a, b in zip(out, ref):
        close(a, b)


def test_radial_message_passing_restatement():
    """against the oracle's embedding accumulation (tensornet.py:405-445: I = sum w1, A = skew(sum w2 rhat), S = symtraceless(sum w3 rhat rhat^T))"""
    n, E = 5, 17
    indptr = torch.tensor([0, 4, 4, 9, 13, 17], dtype=torch.int32)
    rhat = torch.randn(E, 3, dtype=torch.float64)
    w = torch.randn(E, 3, F, dtype=torch.float64)
    I, A, S = T._t_radial_mp(rhat, w, None, indptr)
    row = torch.repeat_interleave(torch.arange(n), (indptr[1:] - indptr[:-1]).long())
    Io = torch.zeros(n, F, dtype=torch.float64).index_add(0, row, w[:, 0])
    Av = torch.zeros(n, 3, F, dtype=torch.float64).index_add(0, row, w[:, 1, None, :] * rhat[:, :, None])
    outer = rhat[:, :, None] * rhat[:, None, :]
    Sf = torch.zeros(n, 3, 3, F, dtype=torch.float64).index_add(0, row, w[:, 2, None, None, :] * outer[..., None])
    Ao = orc.skew(Av)
    So = 0.5 * (Sf + Sf.transpose(1, 2)) - Sf.diagonal(dim1=1, dim2=2).mean(-1)[:, None, None, :] * torch.eye(3, dtype=torch.float64)[None, :, :, None]
    I2, A2, S2 = T._t_decompose(orc.compose(Io, Ao, So))
    close(I, I2); close(A, A2); close(S, S2)
