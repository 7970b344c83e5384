"""-m gpu: every TensorNet kernel (forward and backward, through the C ABI) against a plain PyTorch fp32
evaluation of the same formula on the GPU (the oracle's functions are device-agnostic torch code), plus
fp64-on-CPU as the arbiter.  Mirrors the reference's tests/test_warp_ops.py (Warp ops == PyTorch ops)."""
import pytest
import torch

import tensornet_oracle as orc

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
N, F = 37, 64


def rnd(*s, seed=0, grad=True):
    g = torch.Generator().manual_seed(seed + sum(s))
    return torch.randn(*s, generator=g).to(dev).requires_grad_(grad)


def full(I, A, S):
    """compact (I[N,1,F], A[N,3,F], S[N,5,F]) -> X[N,3,3,F] with torch ops (oracle conventions)."""
    z = torch.zeros_like(I[:, 0])
    Af = torch.stack([z, A[:, 0], A[:, 1], -A[:, 0], z, A[:, 2], -A[:, 1], -A[:, 2], z], 1).view(-1, 3, 3, I.shape[-1])
    Sf = torch.stack([S[:, 0], S[:, 1], S[:, 2], S[:, 1], S[:, 3], S[:, 4], S[:, 2], S[:, 4], -(S[:, 0] + S[:, 3])], 1).view(-1, 3, 3, I.shape[-1])
    return I[:, 0][:, None, None, :] * torch.eye(3, device=I.device)[None, :, :, None] + Af + Sf


def compact(X):
    I, A, S = orc_decompose(X)
    return (I[:, None, :], torch.stack([A[:, 0, 1], A[:, 0, 2], A[:, 1, 2]], 1),
            torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2]], 1))


def orc_decompose(X):
    A = 0.5 * (X - X.transpose(1, 2))
    I = X.diagonal(dim1=1, dim2=2).mean(-1)
    S = 0.5 * (X + X.transpose(1, 2)) - I[:, None, None, :] * torch.eye(3, device=X.device)[None, :, :, None]
    return I, A, S


def check(outs, refs, ins, tol=2e-5):
    outs = outs if isinstance(outs, (list, tuple)) else [outs]
    refs = refs if isinstance(refs, (list, tuple)) else [refs]
    ws = [torch.randn_like(o) for o in outs]
    for o, r in zip(outs, refs):
        assert o.shape == r.shape
        assert (o - r).abs().max() <= tol * (1 + r.abs().max()), (o - r).abs().max().item()
    g1 = torch.autograd.grad(sum((o * w).sum() for o, w in zip(outs, ws)), ins, allow_unused=True, retain_graph=True)
    g2 = torch.autograd.grad(sum((r * w).sum() for r, w in zip(refs, ws)), ins, allow_unused=True, retain_graph=True)
    for a, b in zip(g1, g2):
        assert (a - b).abs().max() <= tol * 10 * (1 + b.abs().max()), (a - b).abs().max().item()


def test_compose_decompose():
    from torchmdnet_b200.extensions import tensornet_ops as t
    I, A, S = rnd(N, 1, F), rnd(N, 3, F), rnd(N, 5, F)
    check(t.fn_compose_tensor(I, A, S), full(I, A, S), [I, A, S])
    X = rnd(N, 3, 3, F)
    check(t.fn_decompose_tensor(X), list(compact(X)), [X])


def test_matmuls_and_norm():
    from torchmdnet_b200.extensions import tensornet_ops as t
    X, Y = rnd(N, 3, 3, F, seed=1), rnd(N, 3, 3, F, seed=2)
    mm = lambda a, b: torch.einsum("nikf,nkjf->nijf", a, b)
    check(t.fn_tensor_matmul_o3_3x3(X, Y), mm(X, Y) + mm(Y, X), [X, Y])
    check(t.fn_tensor_matmul_so3_3x3(X, Y), mm(X, Y), [X, Y])
    I, A, S = orc_decompose(X)
    ref = torch.cat([3 * I ** 2, (A ** 2).sum((1, 2)), (S ** 2).sum((1, 2))], -1)
    check(t.fn_tensor_norm3(X), ref, [X])


def test_fused_compact_stages():
    from torchmdnet_b200.extensions import tensornet_ops as t
    I, A, S = rnd(N, 1, F), rnd(N, 3, F), rnd(N, 5, F)
    X = full(I, A, S)
    tn = lambda T: (T ** 2).sum((1, 2))
    check(t.norm2_compact(I, A, S), tn(X), [I, A, S])
    Ii, Ai, Si = orc_decompose(X)
    check(t.norm3_compact(I, A, S), torch.cat([3 * Ii ** 2, tn(Ai), tn(Si)], -1), [I, A, S])
    g = rnd(N, 3, F, seed=5)
    check(t.gate_compact(I, A, S, g), [I * g[:, 0:1], A * g[:, 1:2], S * g[:, 2:3]], [I, A, S, g])
    check(t.interaction_normalize(I, A, S), list(compact(X / (tn(X) + 1)[:, None, None, :])), [I, A, S])
    J, B, T = rnd(N, 1, F, seed=7), rnd(N, 3, F, seed=8), rnd(N, 5, F, seed=9)
    M = full(J, B, T)
    qf = (1 + 0.1 * torch.randn(N, device=dev)).contiguous()
    mm = lambda a, b: torch.einsum("nikf,nkjf->nijf", a, b)
    for o3 in (True, False):
        C = qf[:, None, None, None] * (mm(X, M) + mm(M, X)) if o3 else 2 * mm(X, M)
        check(t.interaction_product(I, A, S, J, B, T, qf, o3), list(compact(C / (tn(C) + 1)[:, None, None, :])), [I, A, S, J, B, T])
    check(t.interaction_update(I, A, S, J, B, T, qf), list(compact(X + M + qf[:, None, None, None] * mm(M, M))), [I, A, S, J, B, T])


def make_graph(n, seed=0):
    """symmetric random graph with self loops as a sorted CSR + reverse map (what the neighbour kernel emits)."""
    g = torch.Generator().manual_seed(seed)
    adj = torch.rand(n, n, generator=g) < 0.3
    adj = adj | adj.t() | torch.eye(n, dtype=torch.bool)
    ii, jj = torch.nonzero(adj, as_tuple=True)  # sorted by (i, j)
    indptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(ii, minlength=n).cumsum(0)])
    key = ii * n + jj
    rev = torch.searchsorted(key, jj * n + ii)
    return ii.to(dev), jj.to(dev), indptr.int().to(dev), rev.int().to(dev)


def test_radial_message_passing():
    from torchmdnet_b200.extensions import tensornet_ops as t
    ii, jj, indptr, rev = make_graph(N)
    E = ii.numel()
    rhat = rnd(E, 3, seed=3)
    w = rnd(E, 3, F, seed=4)
    Iref = torch.zeros(N, F, device=dev).index_add(0, ii, w[:, 0])
    Av = torch.zeros(N, 3, F, device=dev).index_add(0, ii, w[:, 1, None, :] * rhat[:, :, None])
    outer = rhat[:, :, None] * rhat[:, None, :]
    Sf = torch.zeros(N, 3, 3, F, device=dev).index_add(0, ii, w[:, 2, None, None, :] * outer[..., None])
    Sf = 0.5 * (Sf + Sf.transpose(1, 2)) - Sf.diagonal(dim1=1, dim2=2).mean(-1)[:, None, None, :] * torch.eye(3, device=dev)[None, :, :, None]
    refs = [Iref[:, None, :], torch.stack([-Av[:, 2], Av[:, 1], -Av[:, 0]], 1),
            torch.stack([Sf[:, 0, 0], Sf[:, 0, 1], Sf[:, 0, 2], Sf[:, 1, 1], Sf[:, 1, 2]], 1)]
    check(t.fn_radial_message_passing(rhat, w, None, indptr), refs, [rhat, w], tol=5e-5)
    # with an explicit edge-id indirection built by graph_transform from a shuffled COO list
    perm = torch.randperm(E, device=dev)
    ei = torch.stack([ii[perm], jj[perm]]).int()
    rd, ri, rp, cd, ci, cp = t.graph_transform(ei, N)
    check(t.fn_radial_message_passing(rhat[perm], w[perm], rd, rp), refs, [rhat, w], tol=5e-5)


def test_interaction_message_passing():
    from torchmdnet_b200.extensions import tensornet_ops as t
    ii, jj, indptr, rev = make_graph(N, seed=2)
    E = ii.numel()
    I, A, S = rnd(N, 1, F), rnd(N, 3, F), rnd(N, 5, F)
    w = rnd(E, 3, F, seed=6)
    refs = [torch.zeros_like(I).index_add(0, ii, w[:, 0:1] * I[jj]), torch.zeros_like(A).index_add(0, ii, w[:, 1:2] * A[jj]),
            torch.zeros_like(S).index_add(0, ii, w[:, 2:3] * S[jj])]
    nbr_j = jj.int()
    out = t.fn_message_passing(I, A, S, w, None, nbr_j, indptr, rev, nbr_j, indptr)
    check(out, refs, [I, A, S, w], tol=5e-5)
    # general COO through graph_transform (reference calling convention, unsorted edges)
    perm = torch.randperm(E, device=dev)
    ei = torch.stack([ii[perm], jj[perm]]).int()
    rd, ri, rp, cd, ci, cp = t.graph_transform(ei, N)
    wp = w[perm]
    out = t.fn_message_passing(I, A, S, wp, rd, ri, rp, cd, ci, cp)
    check(out, refs, [I, A, S, w], tol=5e-5)
