"""-m gpu: the fused pair pipeline (tmd_pair_index / tmd_pair_chain / tmd_mp_pair_bwd / tmd_embed_*) against plain
PyTorch fp32 evaluations of the reference formulas (models/tensornet.py:554-585, :738-743, :622-679)."""
import pytest
import torch
import torch.nn.functional as Fn

import tensornet_oracle as orc
import systems

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def graph_for(n=150, seed=0, cap_mult=96):
    from torchmdnet_b200.extensions.neighbors import neighbor_graph
    z, pos, batch, box = systems.random_box(n, seed=seed)
    pos = pos.to(dev).requires_grad_(True)
    g = neighbor_graph(pos, batch.to(dev), box.to(dev), True, 0.0, 5.0, True, n * cap_mult)
    return z.to(dev), pos, g


def test_pair_index():
    from torchmdnet_b200.extensions import fused
    z, pos, g = graph_for()
    pidx = fused.build_pair_index(g)
    E, P = int(g.num_pairs), int(pidx.num_unique)
    i, j = g.nbr_i[:E].long(), g.nbr_j[:E].long()
    n = z.numel()
    assert P == int((j <= i).sum()) == (E - n) // 2 + n
    pid = pidx.pid[:E].long()
    assert pid.min() == 0 and pid.max() == P - 1
    assert torch.equal(pid, pid[g.rev[:E].long()])                       # both directions share the pair
    assert torch.equal(torch.bincount(pid, minlength=P), torch.where(i[pidx.pair_edge[:P].long()] == j[pidx.pair_edge[:P].long()], 1, 2))
    ce = pidx.pair_edge[:P].long()
    assert (j[ce] <= i[ce]).all() and torch.equal(pid[ce], torch.arange(P, device=dev))
    assert torch.equal(pidx.pair_r[:P], g.dist[:E][ce])


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("cfg", ["interaction", "embedding", "single_act"])
def test_pair_chain_value_and_tangent(cfg, tc, monkeypatch):
    """tc=True: tcgen05 3xTF32 GEMM chain (tmd_pair_chain_tc); tc=False: fp32 CUDA-core chain (tmd_pair_chain)."""
    from torchmdnet_b200.extensions import fused
    monkeypatch.setattr(fused, "USE_TENSOR_CORES", tc)
    from torchmdnet_b200.models.utils import ExpNormalSmearing, CosineCutoff
    torch.manual_seed(0)
    R, F = 32, 128
    rbf = ExpNormalSmearing(0.0, 5.0, R, trainable=False).to(dev)
    if cfg == "interaction":
        lins = [torch.nn.Linear(R, F), torch.nn.Linear(F, 2 * F), torch.nn.Linear(2 * F, 3 * F)]
        chain = fused.RadialChain([[l] for l in lins], act=True, mul_cutoff=True)
    elif cfg == "embedding":
        lins = [torch.nn.Linear(R, F) for _ in range(3)]
        chain = fused.RadialChain([lins], act=False, mul_cutoff=False)
    else:
        lins = [torch.nn.Linear(R, 3 * F)]
        chain = fused.RadialChain([lins], act=True, mul_cutoff=False)
    for l in lins:
        l.to(dev)
    assert chain.supported()
    P = 1000 + 7  # not a multiple of the 32-pair tile
    r = (torch.rand(P, device=dev) * 5.5).requires_grad_(True)
    r.data[:5] = torch.tensor([0.0, 4.9999, 5.0, 5.2, 1e-3], device=dev)
    nu = torch.tensor([P], device=dev, dtype=torch.int32)
    f, fp = fused.run_chain(chain, r.detach(), nu, rbf, 0.0, 5.0, True)
    x = rbf(r)
    if cfg == "interaction":
        for l in lins:
            x = Fn.silu(l(x))
        ref = x * CosineCutoff(0.0, 5.0)(r).unsqueeze(-1)
    elif cfg == "embedding":
        ref = torch.cat([l(x) for l in lins], -1)
    else:
        ref = Fn.silu(lins[0](x))
    assert (f - ref).abs().max() < 2e-5 * (1 + ref.abs().max())
    w = torch.randn_like(ref)
    (gr,) = torch.autograd.grad((ref * w).sum(), r)
    mine = (fp * w).sum(-1)
    assert (mine - gr).abs().max() < 1e-4 * (1 + gr.abs().max()), (mine - gr).abs().max()
    f2, none = fused.run_chain(chain, r.detach(), nu, rbf, 0.0, 5.0, False)
    assert none is None and torch.equal(f2, f)
    # only the first num_unique rows are produced
    nu2 = torch.tensor([100], device=dev, dtype=torch.int32)
    f3, _ = fused.run_chain(chain, r.detach(), nu2, rbf, 0.0, 5.0, False)
    assert torch.equal(f3[:100], f[:100])


def test_fused_interaction_message_passing():
    from torchmdnet_b200.extensions import fused
    from torchmdnet_b200.models.utils import ExpNormalSmearing, CosineCutoff
    torch.manual_seed(1)
    R, F = 32, 64
    z, pos, g = graph_for(n=120, seed=2)
    n = z.numel()
    pidx = fused.build_pair_index(g)
    E, P = int(g.num_pairs), int(pidx.num_unique)
    rbf = ExpNormalSmearing(0.0, 5.0, R, trainable=False).to(dev)
    lins = [torch.nn.Linear(R, F).to(dev), torch.nn.Linear(F, 2 * F).to(dev), torch.nn.Linear(2 * F, 3 * F).to(dev)]
    chain = fused.RadialChain([[l] for l in lins], act=True, mul_cutoff=True)
    Y = [torch.randn(n, k, F, device=dev, requires_grad=True) for k in (1, 3, 5)]
    dist = g.dist[:E]
    pair_r = fused.pair_distances(dist, pidx, P)
    M = fused.interaction_message_passing(Y, pair_r, chain, rbf, 0.0, 5.0, g, pidx)
    # reference: per directed edge
    x = rbf(dist)
    for l in lins:
        x = Fn.silu(l(x))
    fe = (x * CosineCutoff(0.0, 5.0)(dist).unsqueeze(-1)).view(-1, 3, F)
    i, j = g.nbr_i[:E].long(), g.nbr_j[:E].long()
    ref = [torch.zeros_like(Y[k]).index_add(0, i, fe[:, k:k + 1] * Y[k][j]) for k in range(3)]
    ws = [torch.randn_like(m) for m in M]
    for a, b in zip(M, ref):
        assert (a - b).abs().max() < 5e-5 * (1 + b.abs().max())
    g1 = torch.autograd.grad(sum((a * w).sum() for a, w in zip(M, ws)), Y + [pos], retain_graph=True)
    g2 = torch.autograd.grad(sum((a * w).sum() for a, w in zip(ref, ws)), Y + [pos], retain_graph=True)
    for a, b in zip(g1, g2):
        assert (a - b).abs().max() < 2e-4 * (1 + b.abs().max()), (a - b).abs().max()


def test_fused_embedding():
    from torchmdnet_b200.extensions import fused
    from torchmdnet_b200.extensions import tensornet_ops as ops
    from torchmdnet_b200.models.utils import ExpNormalSmearing, CosineCutoff
    torch.manual_seed(2)
    R, F = 32, 64
    z, pos, g = graph_for(n=100, seed=4)
    n = z.numel()
    pidx = fused.build_pair_index(g)
    E, P = int(g.num_pairs), int(pidx.num_unique)
    rbf = ExpNormalSmearing(0.0, 5.0, R, trainable=False).to(dev)
    lins = [torch.nn.Linear(R, F).to(dev) for _ in range(3)]
    chain = fused.RadialChain([lins], act=False, mul_cutoff=False)
    U, V = torch.randn(n, F, device=dev), torch.randn(n, F, device=dev)
    dist, vec = g.dist[:E], g.vec[:E]
    rhat = vec / dist.masked_fill(dist == 0, 1).unsqueeze(1)
    pair_r = fused.pair_distances(dist, pidx, P)
    out = fused.embedding_message_passing(dist, rhat, U, V, pair_r, chain, rbf, 0.0, 5.0, g, pidx)
    i, j = g.nbr_i[:E].long(), g.nbr_j[:E].long()
    x = rbf(dist)
    w = (CosineCutoff(0.0, 5.0)(dist).unsqueeze(-1) * (U[i] + V[j])).unsqueeze(1) * torch.stack([l(x) for l in lins], 1)
    ref = ops.fn_radial_message_passing(rhat, w, None, g.indptr)
    ws = [torch.randn_like(m) for m in out]
    for a, b in zip(out, ref):
        assert (a - b).abs().max() < 5e-5 * (1 + b.abs().max())
    (g1,) = torch.autograd.grad(sum((a * w_).sum() for a, w_ in zip(out, ws)), [pos], retain_graph=True)
    (g2,) = torch.autograd.grad(sum((a * w_).sum() for a, w_ in zip(ref, ws)), [pos], retain_graph=True)
    assert (g1 - g2).abs().max() < 2e-4 * (1 + g2.abs().max()), (g1 - g2).abs().max()
