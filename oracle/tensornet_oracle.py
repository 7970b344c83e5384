"""CPU oracle for the TensorNet / Equivariant-Transformer energy+force path.

TEST INFRASTRUCTURE ONLY.  Nothing under torchmdnet_b200/ imports this file; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may call it, and only as the checker / CPU baseline -- never as the product.

It is a functional (state-dict in, tensors out) restatement in plain PyTorch-CPU of
the reference algorithm, every function citing the reference file:line it follows
(paths relative to torchmd/torchmd-net @ 833c0770).  Forces come from torch.autograd,
exactly like the reference (torchmdnet/models/model.py:618-628).  It works in fp32
and fp64 (fp64 is the arbiter when two fp32 paths disagree at 1e-5).

Parity is PINNED: oracle/make_golden.py (run in the build container, where the
reference is importable with the two stubs in oracle/ref_import.py) checks this file
against the unmodified reference's own `create_model/load_model/TorchMD_Net` on the
caffeine known-answer test (tests/test_calculator.py:102-141), tests/expected.pkl
(tests/test_model.py:277-329) and seeded random systems, and tests/test_oracle.py
re-checks it against the committed fixtures in tests/golden/ on every run.
"""
from __future__ import annotations

import math

import numpy as np
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor
import torch.nn.functional as Fn

StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# a1  neighbour list
# --------------------------------------------------------------------------------------
def round_half_away(x: Tensor) -> Tensor:
    """C `round()` (ties away from zero) -- what wp.round does in
    extensions/warp_kernels/neighbors_brute.py:124-135; torch.round is ties-to-even."""
    r = torch.round(x)
    t = torch.trunc(x)
    tie = (x - t).abs() == 0.5
    return torch.where(tie, t + torch.sign(x), r)


def neighbor_list(
    pos: Tensor,
    batch: Tensor,
    box: Optional[Tensor],
    cutoff_lower: float,
    cutoff_upper: float,
    loop: bool = True,
    include_transpose: bool = True,
    block: int = 512,
) -> Tuple[Tensor, Tensor, Tensor]:
    """All pairs with cutoff_lower <= r < cutoff_upper inside one `batch` id.

    Follows extensions/warp_kernels/neighbors_brute.py:98-143: the canonical pair is
    (i, j) with i > j and delta = pos_i - pos_j; triclinic minimum image by sequential
    rounding z -> y -> x on the *updated* components; the transposed pair gets -delta;
    self pairs (loop) bypass the lower cutoff.  Cut-offs are compared squared.
    Every product/sum is a separate rounding (no FMA) so that the CUDA kernel, written
    with __fmul_rn/__fsub_rn, is bit-identical.

    Returns (edge_index[2,E] int64, vec[E,3], dist[E]) sorted lexicographically by
    (edge_index[0], edge_index[1]) -- the reference's order is nondeterministic
    (atomic slot counter), its own tests compare after np.lexsort
    (tests/test_neighbors.py:14-16).
    """
    n = pos.shape[0]
    dt = pos.dtype
    # the reference passes python floats cutoff**2 into a float-typed kernel argument
    lo2 = torch.tensor(float(cutoff_lower) ** 2, dtype=dt)
    hi2 = torch.tensor(float(cutoff_upper) ** 2, dtype=dt)
    periodic = box is not None
    if periodic:
        bx = box.to(dt)
        if bx.dim() == 2:
            bx = bx.unsqueeze(0)
    out_i, out_j, out_v, out_d = [], [], [], []
    ar = torch.arange(n)
    for s in range(0, n, block):
        rows = ar[s : s + block]
        ii = rows[:, None].expand(-1, n)
        jj = ar[None, :].expand(rows.numel(), -1)
        keep = (batch[ii] == batch[jj]) & (jj < ii)  # canonical orientation i > j
        ii, jj = ii[keep], jj[keep]
        dx = pos[ii, 0] - pos[jj, 0]
        dy = pos[ii, 1] - pos[jj, 1]
        dz = pos[ii, 2] - pos[jj, 2]
        if periodic:
            b = bx[batch[ii]] if bx.shape[0] > 1 else bx.expand(ii.numel(), 3, 3)
            s3 = round_half_away(dz / b[:, 2, 2])
            dx = dx - s3 * b[:, 2, 0]
            dy = dy - s3 * b[:, 2, 1]
            dz = dz - s3 * b[:, 2, 2]
            s2 = round_half_away(dy / b[:, 1, 1])
            dx = dx - s2 * b[:, 1, 0]
            dy = dy - s2 * b[:, 1, 1]
            s1 = round_half_away(dx / b[:, 0, 0])
            dx = dx - s1 * b[:, 0, 0]
        d2 = (dx * dx + dy * dy) + dz * dz
        ok = (d2 < hi2) & (d2 >= lo2)
        ii, jj = ii[ok], jj[ok]
        v = torch.stack([dx[ok], dy[ok], dz[ok]], dim=1)
        # numpy's sqrt is the IEEE correctly-rounded one (torch's vectorised CPU sqrt is off by 1 ulp on some
        # inputs); CUDA's sqrtf / __fsqrt_rn -- what the reference's Warp kernel and ours use -- is IEEE too
        d = torch.from_numpy(np.sqrt(d2[ok].numpy()))
        out_i.append(ii); out_j.append(jj); out_v.append(v); out_d.append(d)
        if include_transpose:
            out_i.append(jj); out_j.append(ii); out_v.append(-v); out_d.append(d)
    if loop:
        out_i.append(ar); out_j.append(ar)
        out_v.append(torch.zeros(n, 3, dtype=dt)); out_d.append(torch.zeros(n, dtype=dt))
    if not out_i:
        return torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, 3, dtype=dt), torch.zeros(0, dtype=dt)
    ei, ej = torch.cat(out_i), torch.cat(out_j)
    vec, dist = torch.cat(out_v), torch.cat(out_d)
    order = torch.argsort(ei * n + ej, stable=True)
    return torch.stack([ei[order], ej[order]]), vec[order], dist[order]


def neighbor_list_from_pos(pos, batch, box, hp, loop=True, graph=None):
    """Differentiable wrt pos: indices from `neighbor_list`, vec/dist recomputed with
    autograd-tracked ops (the reference's backward is extensions/neighbor_utils.py:11-46,
    which is exactly d vec/d pos = +-1 and d dist/d vec = vec/dist, masked at dist==0)."""
    with torch.no_grad():
        if graph is not None:   # prebuilt (ei, vec0, dist0) on pos.device: lets a caller keep the host-side O(N^2) search out of a timed region
            ei, vec0, dist0 = graph
        else:
            ei, vec0, dist0 = neighbor_list(pos.detach().cpu(), batch.cpu(), None if box is None else box.cpu(),
                                            float(hp["cutoff_lower"]), float(hp["cutoff_upper"]), loop=loop)
            ei, vec0, dist0 = ei.to(pos.device), vec0.to(pos.device), dist0.to(pos.device)
    i, j = ei[0], ei[1]
    shift = vec0 - (pos.detach()[i] - pos.detach()[j])  # integer box-vector combination
    vec = pos[i] - pos[j] + shift
    self_mask = i == j
    safe = vec.detach().clone()
    safe[self_mask] = 1.0
    dist = torch.where(self_mask, torch.zeros_like(dist0),
                       torch.sqrt(((torch.where(self_mask[:, None], safe, vec)) ** 2).sum(-1)))
    return ei, vec, dist


# --------------------------------------------------------------------------------------
# a4  radial basis and cutoff
# --------------------------------------------------------------------------------------
def cosine_cutoff(d: Tensor, lo: float, hi: float) -> Tensor:
    """models/utils.py:500-528."""
    if lo > 0:
        c = 0.5 * (torch.cos(math.pi * (2 * (d - lo) / (hi - lo) + 1.0)) + 1.0)
        return c * (d < hi) * (d > lo)
    c = 0.5 * (torch.cos(d * math.pi / hi) + 1.0)
    return c * (d < hi)


def expnorm_rbf(d: Tensor, means: Tensor, betas: Tensor, lo: float, hi: float) -> Tensor:
    """models/utils.py:356-407 (ExpNormalSmearing); alpha = 5/(hi-lo); the envelope is
    CosineCutoff(0, hi) regardless of cutoff_lower (utils.py:371)."""
    alpha = 5.0 / (hi - lo)
    d = d.unsqueeze(-1)
    return cosine_cutoff(d, 0.0, hi) * torch.exp(-betas * (torch.exp(alpha * (-d + lo)) - means) ** 2)


def gauss_rbf(d: Tensor, offset: Tensor, coeff: Tensor) -> Tensor:
    """models/utils.py:316-353 (GaussianSmearing)."""
    return torch.exp(coeff * (d.unsqueeze(-1) - offset) ** 2)


def expnorm_initial_params(lo: float, hi: float, num_rbf: int, dtype=torch.float32):
    """models/utils.py:382-395."""
    start = torch.exp(torch.scalar_tensor(-hi + lo, dtype=dtype))
    means = torch.linspace(start, 1, num_rbf, dtype=dtype)
    betas = torch.tensor([(2 / num_rbf * (1 - start)) ** -2] * num_rbf, dtype=dtype)
    return means, betas


# --------------------------------------------------------------------------------------
# a8  3x3 irreducible-tensor algebra, [N,3,3,F] layout (models/tensornet.py:18-146)
# --------------------------------------------------------------------------------------
def _eye(x):
    return torch.eye(3, dtype=x.dtype, device=x.device)[None, :, :, None]


def skew(v: Tensor) -> Tensor:
    """tensornet.py:84-103 -- X01=-vz, X02=vy, X12=-vx (PyTorch convention; the Warp
    kernel's transpose convention is NOT used)."""
    z = torch.zeros_like(v[:, 0])
    return torch.stack([z, -v[:, 2], v[:, 1], v[:, 2], z, -v[:, 0], -v[:, 1], v[:, 0], z], dim=1).view(
        v.shape[0], 3, 3, v.shape[-1])


def decompose(X: Tensor):
    """tensornet.py:18-27."""
    A = 0.5 * (X - X.transpose(1, 2))
    I = X.diagonal(dim1=1, dim2=2).mean(-1)  # [N,F]
    S = 0.5 * (X + X.transpose(1, 2)) - I[:, None, None, :] * _eye(X)
    return I, A, S


def compose(I, A, S):
    """tensornet.py:30-39."""
    return I[:, None, None, :] * _eye(A) + A + S


def tnorm(X: Tensor) -> Tensor:
    """tensornet.py:144-146 -- squared Frobenius norm per channel (no sqrt)."""
    return (X ** 2).sum((1, 2))


def mm3(Xa: Tensor, Xb: Tensor) -> Tensor:
    """per-(atom,channel) 3x3 product, tensornet.py:42-51."""
    return torch.einsum("nikf,nkjf->nijf", Xa, Xb)


def _act(name):
    return {"silu": Fn.silu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "mish": Fn.mish,
            "swish": Fn.silu, "ssp": lambda x: Fn.softplus(x) - math.log(2.0)}[name]


def _lin(sd, key, x, bias=True):
    return Fn.linear(x, sd[key + ".weight"], sd[key + ".bias"] if bias else None)


# --------------------------------------------------------------------------------------
# a5-a10  TensorNet representation
# --------------------------------------------------------------------------------------
def _edge_sum(fn, E: int, edge_chunks: int, *edge_tensors):
    """sum over edge chunks of fn(*[t[chunk] for t in edge_tensors]) (a tuple of per-atom tensors).  edge_chunks <= 1: one
    call on all edges (the plain restatement).  edge_chunks > 1: the same arithmetic chunk by chunk under activation
    checkpointing, so that a 10k-atom box fits in host memory in fp64 (per-edge intermediates [E,3,3,F] are the footprint);
    only the order of the per-atom partial sums changes."""
    if edge_chunks <= 1 or E == 0:
        return fn(*edge_tensors)
    from torch.utils.checkpoint import checkpoint
    step = -(-E // edge_chunks)
    total = None
    for s in range(0, E, step):
        part = checkpoint(fn, *[t[s:s + step] for t in edge_tensors], use_reentrant=False)
        total = part if total is None else tuple(a + b for a, b in zip(total, part))
    return total


def _tensor_embedding(sd: StateDict, te: str, F_: int, act, z, i, j, rhat, rbf, C, n: int, edge_chunks: int = 0) -> Tensor:
    """models/tensornet.py:526-619 (TensorEmbedding.forward) -> X [N,3,3,F]; shared by TensorNet and TensorNet2."""
    Z = sd[te + "emb.weight"][z]

    def edges(i, j, rhat, rbf, C):
        Zij = _lin(sd, te + "emb2", torch.cat([Z[i], Z[j]], dim=-1))
        w = (C[:, None] * Zij)[:, None, :] * torch.stack(
            [_lin(sd, te + "distance_proj1", rbf), _lin(sd, te + "distance_proj2", rbf), _lin(sd, te + "distance_proj3", rbf)],
            dim=1)  # [E,3,F]
        dt = w.dtype
        dev = w.device
        I = torch.zeros(n, F_, dtype=dt, device=dev).index_add(0, i, w[:, 0])
        Av = torch.zeros(n, 3, F_, dtype=dt, device=dev).index_add(0, i, w[:, 1, None, :] * rhat[:, :, None])
        outer = rhat[:, :, None] * rhat[:, None, :]
        Sf = torch.zeros(n, 3, 3, F_, dtype=dt, device=dev).index_add(0, i, w[:, 2, None, None, :] * outer[..., None])
        return I, Av, Sf

    I, Av, Sf = _edge_sum(edges, i.shape[0], edge_chunks, i, j, rhat, rbf, C)
    return _tensor_embedding_nodes(sd, te, F_, act, I, Av, Sf)


def _tensor_embedding_nodes(sd: StateDict, te: str, F_: int, act, I, Av, Sf) -> Tensor:
    """per-atom half of TensorEmbedding.forward (tensornet.py:587-617)."""
    A = skew(Av)
    S = 0.5 * (Sf + Sf.transpose(1, 2)) - Sf.diagonal(dim1=1, dim2=2).mean(-1)[:, None, None, :] * _eye(Sf)
    X = compose(I, A, S)
    nrm = Fn.layer_norm(tnorm(X), (F_,), sd[te + "init_norm.weight"], sd[te + "init_norm.bias"])
    nrm = act(_lin(sd, te + "linears_scalar.0", nrm))
    nrm = act(_lin(sd, te + "linears_scalar.1", nrm)).reshape(-1, 3, F_)
    I = Fn.linear(I, sd[te + "linears_tensor.0.weight"]) * nrm[:, 0]
    A = Fn.linear(A, sd[te + "linears_tensor.1.weight"]) * nrm[:, 1, None, None, :]
    S = Fn.linear(S, sd[te + "linears_tensor.2.weight"]) * nrm[:, 2, None, None, :]
    return compose(I, A, S)


def tensornet_representation(sd: StateDict, hp: dict, z, ei, vec, dist, q_atom, prefix="representation_model.",
                             edge_chunks: int = 0):
    """models/tensornet.py:308-402 (TensorNet.forward), :543-619 (TensorEmbedding.forward),
    :729-814 (Interaction.forward), with the pure-PyTorch message passing :405-445, :622-679."""
    F_ = int(hp["embedding_dimension"])
    lo, hi = float(hp["cutoff_lower"]), float(hp["cutoff_upper"])
    act = _act(hp["activation"])
    group = hp.get("equivariance_invariance_group", "O(3)")
    n = z.shape[0]
    i, j = ei[0], ei[1]
    p = prefix
    if hp["rbf_type"] == "expnorm":
        rbf = expnorm_rbf(dist, sd[p + "distance_expansion.means"], sd[p + "distance_expansion.betas"], lo, hi)
    else:
        rbf = gauss_rbf(dist, sd[p + "distance_expansion.offset"], sd[p + "distance_expansion.coeff"])
    self_mask = i == j
    rhat = vec / dist.masked_fill(self_mask, 1).unsqueeze(1)  # tensornet.py:363-366
    C = cosine_cutoff(dist, lo, hi)

    X = _tensor_embedding(sd, p + "tensor_embedding.", F_, act, z, i, j, rhat, rbf, C, n, edge_chunks)
    dt = X.dtype

    # ---- interaction layers (tensornet.py:729-814)
    qf = (1 + 0.1 * q_atom.to(dt))[:, None, None, None]
    for l in range(int(hp["num_layers"])):
        lp = p + f"layers.{l}."
        X = X / (tnorm(X) + 1)[:, None, None, :]
        I, A, S = decompose(X)
        I = Fn.linear(I, sd[lp + "linears_tensor.0.weight"])
        A = Fn.linear(A, sd[lp + "linears_tensor.1.weight"])
        S = Fn.linear(S, sd[lp + "linears_tensor.2.weight"])
        Y = compose(I, A, S)

        def edges(i, j, rbf, C, I=I, A=A, S=S, lp=lp):
            f = rbf
            for k in range(3):
                f = act(_lin(sd, lp + f"linears_scalar.{k}", f))
            f = (f * C[:, None]).reshape(-1, 3, F_)
            Im = torch.zeros_like(I).index_add(0, i, f[:, 0] * I[j])
            Am = torch.zeros_like(A).index_add(0, i, f[:, 1, None, None, :] * A[j])
            Sm = torch.zeros_like(S).index_add(0, i, f[:, 2, None, None, :] * S[j])
            return Im, Am, Sm

        Im, Am, Sm = _edge_sum(edges, i.shape[0], edge_chunks, i, j, rbf, C)
        M = compose(Im, Am, Sm)
        if group == "O(3)":
            Cn = qf * (mm3(M, Y) + mm3(Y, M))
        else:
            Cn = 2 * mm3(Y, M)
        I, A, S = decompose(Cn)
        np1 = (tnorm(Cn) + 1)
        I = Fn.linear(I / np1, sd[lp + "linears_tensor.3.weight"])
        A = Fn.linear(A / np1[:, None, None, :], sd[lp + "linears_tensor.4.weight"])
        S = Fn.linear(S / np1[:, None, None, :], sd[lp + "linears_tensor.5.weight"])
        dX = compose(I, A, S)
        X = X + dX + qf * mm3(dX, dX)

    # ---- tail (tensornet.py:384-402)
    I, A, S = decompose(X)
    x = torch.cat([3 * I ** 2, tnorm(A), tnorm(S)], dim=-1)
    x = Fn.layer_norm(x, (3 * F_,), sd[p + "out_norm.weight"], sd[p + "out_norm.bias"])
    return act(_lin(sd, p + "linear", x))


# --------------------------------------------------------------------------------------
# TensorNet2 (neutral charge equilibration) + ScalarPlusWeightedCoulomb
# (oracle first: the GPU path for this row is not built yet; these functions and their goldens pin the arithmetic)
# --------------------------------------------------------------------------------------
def _mol_sum(x: Tensor, batch: Tensor, m: int) -> Tensor:
    """tensornet2.py:65-100 (ChargePredict.mol_sum): per-molecule sum."""
    return torch.zeros(m, x.shape[-1], dtype=x.dtype, device=x.device).index_add(0, batch, x)


def charge_predict(sd: StateDict, key: str, X: Tensor, batch: Tensor, Q_atom: Tensor, q_dim: int, act, m: int) -> Tensor:
    """tensornet2.py:142-157 (ChargePredict.forward, pure-PyTorch branch) and :102-140 (qeq):
    features (I, |A|^2, |S|^2) -> LayerNorm -> MLP(3F -> F -> F -> 2 q_dim) -> (charges, f);
    new_charges = charges + f^2 / (sum_mol f^2 + 1e-6) * (Q - sum_mol charges)."""
    I, A, S = decompose(X)
    x = torch.cat([I, tnorm(A), tnorm(S)], dim=-1)
    F3 = x.shape[-1]
    x = Fn.layer_norm(x, (F3,), sd[key + "q_norm.weight"], sd[key + "q_norm.bias"])
    x = act(_lin(sd, key + "q_mlp.layers.0", x))
    x = act(_lin(sd, key + "q_mlp.layers.2", x))
    x = _lin(sd, key + "q_mlp.layers.4", x)
    charges, f = x[:, :q_dim], x[:, q_dim:]
    f_u = f ** 2
    F_u = _mol_sum(f_u, batch, m) + 1.0e-6
    Q_u = _mol_sum(charges, batch, m)
    dQ = Q_atom.to(x.dtype).unsqueeze(-1) - Q_u[batch]
    return charges + f_u / F_u[batch] * dQ


def tensornet2_representation(sd: StateDict, hp: dict, z, ei, vec, dist, batch, q_mol, prefix="representation_model."):
    """models/tensornet2.py:343-462 (TensorNet2.forward, pure-PyTorch branch) and :519-626 (Interaction.forward): the TensorNet
    embedding (tensornet.py:543-619), then per layer an edge MLP on cat(rbf, q_i, q_j) (NOT a function of r alone and not
    symmetric in (i, j)), the TensorNet tensor update without the (1 + 0.1 q) factor, and a charge re-equilibration."""
    F_ = int(hp["embedding_dimension"])
    q_dim = int(hp["q_dim"])
    lo, hi = float(hp["cutoff_lower"]), float(hp["cutoff_upper"])
    act = _act(hp["activation"])
    group = hp.get("equivariance_invariance_group", "O(3)")
    n = z.shape[0]
    m = int(batch.max()) + 1
    i, j = ei[0], ei[1]
    p = prefix
    rbf = expnorm_rbf(dist, sd[p + "distance_expansion.means"], sd[p + "distance_expansion.betas"], lo, hi) \
        if hp["rbf_type"] == "expnorm" else gauss_rbf(dist, sd[p + "distance_expansion.offset"], sd[p + "distance_expansion.coeff"])
    rhat = vec / dist.masked_fill(i == j, 1).unsqueeze(1)
    C = cosine_cutoff(dist, lo, hi)
    Q_atom = torch.zeros(n, dtype=dist.dtype, device=dist.device) if q_mol is None else q_mol[batch].to(dist.dtype)
    X = _tensor_embedding(sd, p + "tensor_embedding.", F_, act, z, i, j, rhat, rbf, C, n)
    charges = [charge_predict(sd, p + "charge_predict_0.", X, batch, Q_atom, q_dim, act, m)]
    for l in range(int(hp["num_layers"])):
        lp = p + f"layers.{l}."
        f = torch.cat([rbf, charges[-1][i], charges[-1][j]], dim=-1)
        for k in range(3):
            f = act(_lin(sd, lp + f"linears_scalar.{k}", f))
        f = (f * C[:, None]).reshape(-1, 3, F_)
        X = X / (tnorm(X) + 1)[:, None, None, :]
        I, A, S = decompose(X)
        I = Fn.linear(I, sd[lp + "linears_tensor.0.weight"])
        A = Fn.linear(A, sd[lp + "linears_tensor.1.weight"])
        S = Fn.linear(S, sd[lp + "linears_tensor.2.weight"])
        Y = compose(I, A, S)
        Im = torch.zeros_like(I).index_add(0, i, f[:, 0] * I[j])
        Am = torch.zeros_like(A).index_add(0, i, f[:, 1, None, None, :] * A[j])
        Sm = torch.zeros_like(S).index_add(0, i, f[:, 2, None, None, :] * S[j])
        M = compose(Im, Am, Sm)
        Cn = (mm3(M, Y) + mm3(Y, M)) if group == "O(3)" else 2 * mm3(Y, M)
        I, A, S = decompose(Cn)
        np1 = tnorm(Cn) + 1
        I = Fn.linear(I / np1, sd[lp + "linears_tensor.3.weight"])
        A = Fn.linear(A / np1[:, None, None, :], sd[lp + "linears_tensor.4.weight"])
        S = Fn.linear(S / np1[:, None, None, :], sd[lp + "linears_tensor.5.weight"])
        dX = compose(I, A, S)
        X = X + dX + mm3(dX, dX)
        charges.append(charge_predict(sd, p + f"charge_predicts.{l}.", X, batch, Q_atom, q_dim, act, m))
    I, A, S = decompose(X)
    x = torch.cat([3 * I ** 2, tnorm(A), tnorm(S)], dim=-1)
    x = Fn.layer_norm(x, (3 * F_,), sd[p + "out_norm.weight"], sd[p + "out_norm.bias"])
    x = act(_lin(sd, p + "linear", x))
    return x, torch.cat(charges, dim=-1)


def exp_cutoff(d: Tensor, rc: float) -> Tensor:
    """output_modules.py:321-330 (_exp_cutoff): short-range damping taken from aimnet."""
    return torch.exp(-1.0 / (1.0 - (d / rc).clamp(0, 1.0 - 1e-6).pow(2))) / 0.36787944117144233


def weighted_coulomb_per_atom(hp: dict, charges: Tensor, pos: Tensor, batch: Tensor, box: Optional[Tensor]) -> Tensor:
    """output_modules.py:344-609 (ScalarPlusWeightedCoulomb.pre_reduce): damped Coulomb energy of the predicted charge channels,
    weighted mean over channels, split equally between the two atoms of a pair.  coulomb_cutoff None: all pairs of a molecule;
    else pairs within the cutoff (half list, minimum image) with the reaction-field form."""
    q_dim, L = int(hp["q_dim"]), int(hp["num_layers"])
    w = torch.tensor([float(hp["q_weights"][l]) for l in range(L + 1)], dtype=charges.dtype, device=charges.device).repeat_interleave(q_dim)
    factor = 0.5 * 27.211386024367243 * 0.5291772105638411
    n = pos.shape[0]
    rc = hp.get("coulomb_cutoff", None)
    ii, jj = torch.triu_indices(n, n, 1).unbind(0)
    keep = batch[ii] == batch[jj]
    ii, jj = ii[keep], jj[keep]
    d = pos[ii] - pos[jj]
    if rc is None:
        assert box is None, "PBC is not supported with coulomb_cutoff = None"
        r = torch.linalg.norm(d, dim=-1)
        shape = 1.0 / r
    else:
        if box is not None:
            b = box if box.dim() == 3 else box[None].expand(int(batch.max()) + 1, 3, 3)
            bb = b[batch[ii]].to(d.dtype)
            # triclinic minimum image in the reference's order z, y, x
            s3 = round_half_away(d[:, 2] / bb[:, 2, 2]); d = d - s3[:, None] * bb[:, 2]
            s2 = round_half_away(d[:, 1] / bb[:, 1, 1]); d = d - s2[:, None] * bb[:, 1]
            s1 = round_half_away(d[:, 0] / bb[:, 0, 0]); d = d - s1[:, None] * bb[:, 0]
        r = torch.linalg.norm(d, dim=-1)
        inside = r < float(rc)
        ii, jj, r = ii[inside], jj[inside], r[inside]
        eps = float(hp.get("coulomb_epsilon_solvent", 78.3))
        k_rf = (1.0 / rc ** 3) * (eps - 1.0) / (2.0 * eps + 1.0)
        c_rf = (1.0 / rc) * (3.0 * eps) / (2.0 * eps + 1.0)
        shape = 1.0 / r + k_rf * r ** 2 - c_rf
    fc = 1.0 - exp_cutoff(r, 4.6)
    e_ij = factor * (fc * shape).unsqueeze(-1) * charges[ii] * charges[jj]
    e_ij = (e_ij * w[None, :]).sum(-1) / w.sum()
    e_i = torch.zeros(n, dtype=charges.dtype, device=charges.device).index_add(0, ii, e_ij).index_add(0, jj, e_ij)
    return e_i


# --------------------------------------------------------------------------------------
# a13-a15  Equivariant Transformer
# --------------------------------------------------------------------------------------
def et_representation(sd: StateDict, hp: dict, z, ei, vec, dist, prefix="representation_model."):
    """models/torchmd_et.py:188-221 (forward), :315-426 (attention layer),
    models/utils.py:83-117 (NeighborEmbedding)."""
    F_ = int(hp["embedding_dimension"]); H = int(hp["num_heads"]); hd = F_ // H
    lo, hi = float(hp["cutoff_lower"]), float(hp["cutoff_upper"])
    act = _act(hp["activation"]); attn_act = _act(hp["attn_activation"])
    p = prefix
    n = z.shape[0]
    src, dst = ei[0], ei[1]  # ET: query/target = edge_index[1], key/value/source = edge_index[0]
    x = sd[p + "embedding.weight"][z]
    if hp["rbf_type"] == "expnorm":
        rbf = expnorm_rbf(dist, sd[p + "distance_expansion.means"], sd[p + "distance_expansion.betas"], lo, hi)
    else:
        rbf = gauss_rbf(dist, sd[p + "distance_expansion.offset"], sd[p + "distance_expansion.coeff"])
    notself = src != dst
    dhat = vec / dist.masked_fill(~notself, 1).unsqueeze(1)  # torchmd_et.py:205-206 (self loops keep 0)
    C = cosine_cutoff(dist, lo, hi)
    if hp.get("neighbor_embedding", True):
        ne = p + "neighbor_embedding."
        W = _lin(sd, ne + "distance_proj", rbf) * C[:, None]
        msg = (W * sd[ne + "embedding.weight"][z][dst])[notself]
        xn = torch.zeros(n, F_, dtype=x.dtype, device=x.device).index_add(0, src[notself], msg)
        x = _lin(sd, ne + "combine", torch.cat([x, xn], dim=1))
    v3 = torch.zeros(n, 3, F_, dtype=x.dtype, device=x.device)
    di = hp.get("distance_influence", "both")
    vcut = bool(hp.get("vector_cutoff", False))
    for l in range(int(hp["num_layers"])):
        lp = p + f"attention_layers.{l}."
        xl = Fn.layer_norm(x, (F_,), sd[lp + "layernorm.weight"], sd[lp + "layernorm.bias"])
        qh = _lin(sd, lp + "q_proj", xl).reshape(-1, H, hd)
        kh = _lin(sd, lp + "k_proj", xl).reshape(-1, H, hd)
        vh = _lin(sd, lp + "v_proj", xl).reshape(-1, H, 3 * hd)
        vp = Fn.linear(v3, sd[lp + "vec_proj.weight"])
        vec1, vec2, vec3 = torch.split(vp, F_, dim=-1)
        vec_dot = (vec1 * vec2).sum(dim=1)
        dk = act(_lin(sd, lp + "dk_proj", rbf)).reshape(-1, H, hd) if di in ("keys", "both") else None
        dv = act(_lin(sd, lp + "dv_proj", rbf)).reshape(-1, H, 3 * hd) if di in ("values", "both") else None
        q_i, k_j, v_j = qh[dst], kh[src], vh[src]
        vec_j = v3.reshape(-1, 3, H, hd)[src]
        attn = (q_i * k_j).sum(-1) if dk is None else (q_i * k_j * dk).sum(-1)
        attn = attn_act(attn)
        if vcut:
            v_j = v_j * C[:, None, None]
        else:
            attn = attn * C[:, None]
        if dv is not None:
            v_j = v_j * dv
        s, m1, m2 = torch.split(v_j, hd, dim=2)
        xs = s * attn.unsqueeze(2)
        vm = vec_j * m1.unsqueeze(1) + m2.unsqueeze(1) * dhat[:, :, None, None]
        xagg = torch.zeros(n, H, hd, dtype=x.dtype, device=x.device).index_add(0, dst, xs).reshape(-1, F_)
        vagg = torch.zeros(n, 3, H, hd, dtype=x.dtype, device=x.device).index_add(0, dst, vm).reshape(-1, 3, F_)
        o1, o2, o3 = torch.split(_lin(sd, lp + "o_proj", xagg), F_, dim=1)
        x = x + vec_dot * o2 + o3
        v3 = v3 + vec3 * o1.unsqueeze(1) + vagg
    x = Fn.layer_norm(x, (F_,), sd[p + "out_norm.weight"], sd[p + "out_norm.bias"])
    return x, v3


def gated_equivariant_block(sd, key, x, v, out_channels, act, scalar_activation):
    """models/utils.py:626-655."""
    vec1 = torch.norm(Fn.linear(v, sd[key + "vec1_proj.weight"]), dim=-2)
    # zero rows are masked out of the norm in the reference (utils.py:629-645); same value (0)
    vec2 = Fn.linear(v, sd[key + "vec2_proj.weight"])
    h = torch.cat([x, vec1], dim=-1)
    h = _lin(sd, key + "update_net.layers.2", act(_lin(sd, key + "update_net.layers.0", h)))
    xo, vo = torch.split(h, out_channels, dim=-1)
    vo = vo.unsqueeze(1) * vec2
    if scalar_activation:
        xo = act(xo)
    return xo, vo


# --------------------------------------------------------------------------------------
# a11-a12  heads, reduction, forces
# --------------------------------------------------------------------------------------
def energy_and_forces(sd: StateDict, hp: dict, z: Tensor, pos: Tensor, batch: Optional[Tensor] = None,
                      box: Optional[Tensor] = None, q: Optional[Tensor] = None, derivative: bool = True,
                      edge_chunks: int = 0, graph=None, create_graph: bool = False):
    """models/model.py:530-631 (TorchMD_Net.forward): representation -> output head
    (output_modules.py:79-163) -> *std -> (Atomref prior, priors/atomref.py) -> scatter-sum
    per molecule (output_modules.py:43-73, utils.py:671-701) -> +mean -> F = -dE/dpos."""
    batch = torch.zeros_like(z) if batch is None else batch
    pos = pos.detach().clone().requires_grad_(derivative)
    ei, vec, dist = neighbor_list_from_pos(pos, batch, box, hp, loop=True, graph=graph)
    act = _act(hp["activation"])
    model = hp["model"]
    if model == "tensornet":
        q_atom = torch.zeros(z.shape[0], dtype=pos.dtype, device=pos.device) if q is None else q[batch]
        x = tensornet_representation(sd, hp, z, ei, vec, dist, q_atom, edge_chunks=edge_chunks)
        e = _lin(sd, "output_model.output_network.layers.2", act(_lin(sd, "output_model.output_network.layers.0", x)))
    elif model == "tensornet2":
        x, charges = tensornet2_representation(sd, hp, z, ei, vec, dist, batch, q)
        e = _lin(sd, "output_model.output_network.layers.2", act(_lin(sd, "output_model.output_network.layers.0", x)))
        if "Coul" in hp["output_model"]:
            e = e + weighted_coulomb_per_atom(hp, charges, pos, batch, box).unsqueeze(-1)
    elif model == "equivariant-transformer":
        x, v = et_representation(sd, hp, z, ei, vec, dist)
        F_ = int(hp["embedding_dimension"])
        x, v = gated_equivariant_block(sd, "output_model.output_network.0.", x, v, F_ // 2, act, True)
        x, v = gated_equivariant_block(sd, "output_model.output_network.1.", x, v, 1, act, False)
        e = x + v.sum() * 0
    else:
        raise ValueError(model)
    e = e * sd["std"]
    if "prior_model.0.atomref.weight" in sd:
        e = e + sd["prior_model.0.atomref.weight"][z]
    m = int(batch.max()) + 1
    y = torch.zeros(m, 1, dtype=e.dtype, device=e.device).index_add(0, batch, e)
    if hp.get("reduce_op", "add") == "mean":
        y = y / torch.bincount(batch, minlength=m).to(y.dtype)[:, None]
    y = y + sd["mean"]
    if not derivative:
        return y.detach(), None
    if create_graph:   # training with a force loss (models/model.py:618-628: create_graph=self.training): outputs stay differentiable
        (dy,) = torch.autograd.grad([y], [pos], grad_outputs=[torch.ones_like(y)], create_graph=True)
        return y, -dy
    (dy,) = torch.autograd.grad([y], [pos], grad_outputs=[torch.ones_like(y)])
    return y.detach(), -dy


def cast_state_dict(sd: StateDict, dtype) -> StateDict:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
