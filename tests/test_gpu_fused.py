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


@pytest.mark.parametrize("tc", ["table", "f16", "tf32", False])
@pytest.mark.parametrize("cfg", ["interaction", "embedding", "single_act"])
def test_pair_chain_value_and_tangent(cfg, tc, monkeypatch):
    """tc='f16'/'tf32': tcgen05 GEMM chain with fp16 / tf32 split terms (tmd_pair_chain_tc); False: fp32 CUDA-core chain;
    'table': the validated per-model table of the frozen chain (tmd_radial_table) -- all against torch autograd."""
    from torchmdnet_b200.extensions import fused
    monkeypatch.setattr(fused, "USE_TENSOR_CORES", bool(tc))
    monkeypatch.setattr(fused, "USE_F16_SPLIT", tc in ("f16", "table"))
    monkeypatch.setattr(fused, "USE_RADIAL_TABLE", tc == "table")
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
    if tc == "table":
        rep = chain.table_report   # the gate accepted a table: error vs fp64 below 2e-6, vs the exact GPU chain below 1e-5
        print(cfg, rep)
        assert chain.table(rbf, 0.0, 5.0) is not None and rep["err_f_vs_fp64"] < 2e-6 and rep["err_fp_vs_fp64"] < 2e-6
        assert rep["err_f_vs_exact"] < 1e-5 and rep["err_fp_vs_exact"] < 1e-5
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
    M = fused.interaction_message_passing(Y, dist, pair_r.detach(), chain, rbf, 0.0, 5.0, g, pidx)
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
    out = fused.embedding_message_passing(dist, vec, U, V, pair_r, chain, rbf, 0.0, 5.0, g, pidx)
    # the same through per-element tables: U_i = T_u[z_i], V_j = T_v[z_j]
    zz = torch.randint(0, 7, (n,), device=dev)
    TU, TV = torch.randn(7, F, device=dev), torch.randn(7, F, device=dev)
    out_t = fused.embedding_message_passing(dist, vec, TU, TV, pair_r, chain, rbf, 0.0, 5.0, g, pidx, zidx=zz)
    out_r = fused.embedding_message_passing(dist, vec, TU[zz], TV[zz], pair_r, chain, rbf, 0.0, 5.0, g, pidx)
    for a, b in zip(out_t, out_r):
        assert torch.equal(a, b)
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


@pytest.mark.parametrize("R,F,nz,lo", [(32, 128, 2, 0.0), (32, 64, 7, 0.0), (50, 96, 16, 0.5), (64, 256, 3, 0.0)])
def test_embedding_by_moments(R, F, nz, lo):
    """csrc/embed_moments.cu (neighbour sums of C phi(rhat) rbf_k first, the radial projections per atom after) vs the composed
    reference formula (models/tensornet.py:554-585) in torch: outputs and position gradients"""
    from torchmdnet_b200.extensions import fused
    from torchmdnet_b200.extensions import tensornet_ops as ops
    from torchmdnet_b200.models.utils import ExpNormalSmearing, CosineCutoff
    torch.manual_seed(5)
    z, pos, g = graph_for(n=150, seed=8)
    n = z.numel()
    E = int(g.num_pairs)
    rbf = ExpNormalSmearing(lo, 5.0, R, trainable=False).to(dev)
    lins = [torch.nn.Linear(R, F).to(dev) for _ in range(3)]
    for l in lins:
        for t in (l.weight, l.bias):
            t.requires_grad_(False)
    elements = torch.randperm(40, device=dev)[:nz].sort().values
    zz = elements[torch.randint(0, nz, (n,), device=dev)]
    TU, TV = torch.randn(41, F, device=dev), torch.randn(41, F, device=dev)
    dist, vec = g.dist[:E], g.vec[:E]
    slots = fused.SpeciesSlots().get(zz)
    assert slots is not None and slots[1].numel() <= nz
    mom = fused.EmbedMomentWeights(lins)
    assert mom.supported(rbf)
    Wk, Wh = mom.tensors()
    Uz, Vz = mom.species_rows(TU, TV, slots[1])
    out = fused.embedding_moments(dist, vec, slots[0], Uz, Vz, Wk, Wh, rbf, lo, 5.0, g)
    i, j = g.nbr_i[:E].long(), g.nbr_j[:E].long()
    rhat = vec / dist.masked_fill(dist == 0, 1).unsqueeze(1)
    x = rbf(dist)
    w = (CosineCutoff(lo, 5.0)(dist).unsqueeze(-1) * (TU[zz[i]] + TV[zz[j]])).unsqueeze(1) * torch.stack([l(x) for l in lins], 1)
    ref = ops._t_radial_mp(rhat, w, None, g.indptr)
    for a, b in zip(out, ref):
        assert (a - b).abs().max() < 2e-5 * (1 + b.abs().max()), float((a - b).abs().max() / (1 + b.abs().max()))
    ws = [torch.randn_like(m) for m in out]
    (g1,) = torch.autograd.grad(sum((a * w_).sum() for a, w_ in zip(out, ws)), [pos], retain_graph=True)
    (g2,) = torch.autograd.grad(sum((a * w_).sum() for a, w_ in zip(ref, ws)), [pos], retain_graph=True)
    assert (g1 - g2).abs().max() < 1e-4 * (1 + g2.abs().max()), float((g1 - g2).abs().max() / (1 + g2.abs().max()))


@pytest.mark.parametrize("vcut", [True, False])
def test_et_attention_message(vcut):
    """fused ET message kernel (tmd_et_msg_fwd/bwd) vs the reference formulas (models/torchmd_et.py:355-426) in torch."""
    from torchmdnet_b200.extensions import fused
    from torchmdnet_b200.models.utils import ExpNormalSmearing, CosineCutoff
    torch.manual_seed(3)
    R, F, H = 32, 64, 4
    hd = F // H
    z, pos, g = graph_for(n=90, seed=6)
    n = z.numel()
    pidx = fused.build_pair_index(g)
    E, P = int(g.num_pairs), int(pidx.num_unique)
    rbf = ExpNormalSmearing(0.0, 5.0, R, trainable=False).to(dev)
    dk_proj, dv_proj = torch.nn.Linear(R, F).to(dev), torch.nn.Linear(R, 3 * F).to(dev)
    chain = fused.RadialChain([[dk_proj, dv_proj]], act=True, mul_cutoff=False)
    q, k = (torch.randn(n, F, device=dev, requires_grad=True) for _ in range(2))
    v = torch.randn(n, 3 * F, device=dev, requires_grad=True)
    vec = torch.randn(n, 3, F, device=dev, requires_grad=True)
    dist, dvec = g.dist[:E], g.vec[:E]
    rhat = dvec / dist.masked_fill(dist == 0, 1).unsqueeze(1)
    pair_r = fused.pair_distances(dist, pidx, P)
    xa, va = fused.et_message(q, k, v, vec, dist, rhat, pair_r, chain, rbf, 0.0, 5.0, vcut, hd, g, pidx)
    # reference: edge (e0 = source, e1 = target); our slot (row t, nbr s) is the edge (s -> t) with d_ij = -rhat
    t_idx, s_idx = g.nbr_i[:E].long(), g.nbr_j[:E].long()
    f_ij = rbf(dist)
    cutoff = CosineCutoff(0.0, 5.0)(dist)
    dk = torch.nn.functional.silu(dk_proj(f_ij)).view(-1, H, hd)
    dv = torch.nn.functional.silu(dv_proj(f_ij)).view(-1, H, 3 * hd)
    q_i, k_j, v_j = q.view(-1, H, hd)[t_idx], k.view(-1, H, hd)[s_idx], v.view(-1, H, 3 * hd)[s_idx]
    vec_j = vec.view(-1, 3, H, hd)[s_idx]
    attn = torch.nn.functional.silu((q_i * k_j * dk).sum(-1))
    if vcut:
        v_j = v_j * cutoff.view(-1, 1, 1)
    else:
        attn = attn * cutoff.unsqueeze(1)
    v_j = v_j * dv
    xs, m1, m2 = torch.split(v_j, hd, dim=2)
    xs = xs * attn.unsqueeze(2)
    vm = vec_j * m1.unsqueeze(1) + m2.unsqueeze(1) * (-rhat).unsqueeze(2).unsqueeze(3)
    xref = torch.zeros(n, H, hd, device=dev).index_add(0, t_idx, xs).view(n, F)
    vref = torch.zeros(n, 3, H, hd, device=dev).index_add(0, t_idx, vm).view(n, 3, F)
    assert (xa - xref).abs().max() < 5e-5 * (1 + xref.abs().max())
    assert (va - vref).abs().max() < 5e-5 * (1 + vref.abs().max())
    wx, wv = torch.randn_like(xref), torch.randn_like(vref)
    g1 = torch.autograd.grad((xa * wx).sum() + (va * wv).sum(), [q, k, v, vec, pos], retain_graph=True)
    g2 = torch.autograd.grad((xref * wx).sum() + (vref * wv).sum(), [q, k, v, vec, pos], retain_graph=True)
    for name, a, b in zip("q k v vec pos".split(), g1, g2):
        assert (a - b).abs().max() < 2e-4 * (1 + b.abs().max()), (name, (a - b).abs().max().item())


@pytest.mark.parametrize("F", [64, 128, 256])
def test_irreps_linear_tensor_cores(F):
    """grouped 3xTF32 node linears (tmd_irreps_linear_tc) vs three fp64 nn.Linear, forward and input gradients."""
    from torchmdnet_b200.extensions import fused
    torch.manual_seed(0)
    n = 300 + 37  # segments not multiples of 128
    lins = [torch.nn.Linear(F, F, bias=False).to(dev) for _ in range(3)]
    w = fused.IrrepsLinearWeights(lins)
    assert w.supported()
    X = [torch.randn(n, k, F, device=dev, requires_grad=True) for k in (1, 3, 5)]
    out = fused.irreps_linear(X, w)
    ref = [torch.nn.functional.linear(x.double(), l.weight.double()) for x, l in zip(X, lins)]
    for a, b in zip(out, ref):
        assert a.shape == b.shape and (a.double() - b).abs().max() < 3e-6 * (1 + b.abs().max())
    ws = [torch.randn_like(o) for o in out]
    g1 = torch.autograd.grad(sum((o * w_).sum() for o, w_ in zip(out, ws)), X)
    g2 = torch.autograd.grad(sum((r * w_.double()).sum() for r, w_ in zip(ref, ws)), X)
    for a, b in zip(g1, g2):
        assert (a.double() - b.double()).abs().max() < 3e-6 * (1 + b.abs().max())


@pytest.mark.parametrize("case", ["water_cells", "water_cells_static", "random_box_no_cells", "molecules", "F64", "tail_group"])
def test_cooperative_message_passing_bit_exact(case, monkeypatch):
    """csrc/mp_coop.cu (group-staged neighbour rows, async radial-row rings) == the warp-per-atom kernels, bit for bit: every row
    accumulates its edges in the same ascending-j order with the same FMA sequence.  Cases: cell-ordered groups (periodic water
    box, exact-size and capacity-padded buffers), identity order (small random box: segment scan, no cell grid), a non-periodic
    molecule batch, F = 64 (half the lanes idle), and an atom count that leaves a partial last group."""
    from torchmdnet_b200.extensions import fused
    from torchmdnet_b200.extensions.neighbors import neighbor_graph
    torch.manual_seed(0)
    F = 64 if case == "F64" else 128
    if case in ("water_cells", "water_cells_static", "F64"):
        z, pos, batch, box = systems.water_box(n_side=6, L=46.42 * 6 / 15, seed=1)
        cap = z.numel() * (64 if case != "water_cells_static" else 96)
    elif case == "tail_group":
        z, pos, batch, box = systems.random_box(1003, seed=2)
        cap = z.numel() * 96
    elif case == "random_box_no_cells":
        z, pos, batch, box = systems.random_box(300, seed=3)
        cap = z.numel() * 96
    else:
        z, pos, batch, box = systems.molecule_batch(n_mol=12, seed=4)
        cap = z.numel() * 48
    n = z.numel()
    g = neighbor_graph(pos.to(dev), batch.to(dev), None if box is None else box.to(dev), box is not None, 0.0, 5.0, True, cap)
    assert int(g.num_pairs) <= cap
    uses_cells = bool(g.order_flag.item()) if g.order_flag is not None else False
    assert uses_cells == (case in ("water_cells", "water_cells_static", "F64", "tail_group"))
    pidx = fused.build_pair_index(g)
    P = int(pidx.num_unique)
    Y = [torch.randn(n, k, F, device=dev) for k in (1, 3, 5)]
    dM = [torch.randn(n, k, F, device=dev) for k in (1, 3, 5)]
    f, fp = torch.randn(P, 3, F, device=dev), torch.randn(P, 3, F, device=dev)
    monkeypatch.setattr(fused, "USE_MP_COOP", False)
    ref = fused.MessagePassingPlan(g, pidx, F)
    assert not ref.coop
    M0 = ref.forward(Y, f)
    dY0, gd0 = ref.backward(Y, dM, f, fp)
    dY0n, _ = ref.backward(Y, dM, f, None)
    monkeypatch.setattr(fused, "USE_MP_COOP", True)
    plan = fused.MessagePassingPlan(g, pidx, F)
    assert plan.coop
    M1 = plan.forward(Y, f)
    dY1, gd1 = plan.backward(Y, dM, f, fp)
    dY1n, gdn = plan.backward(Y, dM, f, None)
    torch.cuda.synchronize()
    E = int(g.num_pairs)
    # the plan itself: groups cover every row once; union positions point at the edge's neighbour
    rows = plan.rows.long()
    assert torch.equal(torch.sort(rows).values, torch.arange(n, device=dev))
    gcnt = plan.gcnt.long()
    assert (gcnt >= 0).all(), "no group may overflow the staged path on these inputs"
    grp_of_row = torch.empty(n, dtype=torch.long, device=dev); grp_of_row[rows] = torch.arange(n, device=dev) // plan.group
    ge = grp_of_row[g.nbr_i[:E].long()]
    up = plan.upos[:E].to(torch.int32) & 0xFFFF
    t = (up & 0x7FFF).long()
    assert (t < gcnt[ge]).all()
    assert torch.equal(plan.union_j[(plan.goff.long()[ge] + t)].long(), g.nbr_j[:E].long())
    assert torch.equal((up & 0x8000) != 0, g.nbr_i[:E] == g.nbr_j[:E])
    print(f"{case}: n={n} E={E} groups={gcnt.numel()} mean union={gcnt.float().mean():.1f} edges/union={E / gcnt.sum().item():.2f}")
    for a, b in zip(M0 + dY0 + dY0n, M1 + dY1 + dY1n):
        assert torch.equal(a, b)
    assert torch.equal(gd0[:E], gd1[:E]) and gdn is None


def test_cooperative_message_passing_overflow_fallback():
    """a group whose rows hold more than 2048 distinct neighbours is flagged (gcnt = -1) and gathered directly -- same bits."""
    from torchmdnet_b200.extensions import fused
    from torchmdnet_b200.extensions.neighbors import neighbor_graph
    torch.manual_seed(1)
    n, F = 1000, 64
    g0 = torch.Generator().manual_seed(5)
    pos = torch.rand(n, 3, generator=g0) * 6.0          # everything within the 12 A cutoff of everything: rows of 1000 neighbours
    batch = torch.zeros(n, dtype=torch.long)
    g = neighbor_graph(pos.to(dev), batch.to(dev), None, False, 0.0, 12.0, True, n * n)
    pidx = fused.build_pair_index(g)
    P = int(pidx.num_unique)
    Y = [torch.randn(n, k, F, device=dev) for k in (1, 3, 5)]
    dM = [torch.randn(n, k, F, device=dev) for k in (1, 3, 5)]
    f, fp = torch.randn(P, 3, F, device=dev), torch.randn(P, 3, F, device=dev)
    fused_coop = fused.USE_MP_COOP
    try:
        fused.USE_MP_COOP = False
        ref = fused.MessagePassingPlan(g, pidx, F)
        M0 = ref.forward(Y, f); dY0, gd0 = ref.backward(Y, dM, f, fp)
        fused.USE_MP_COOP = True
        plan = fused.MessagePassingPlan(g, pidx, F)
        assert (plan.gcnt[:-1] == -1).all()      # (the partial last group may fit the staged path: 4 rows x 1000 edges)
        M1 = plan.forward(Y, f); dY1, gd1 = plan.backward(Y, dM, f, fp)
    finally:
        fused.USE_MP_COOP = fused_coop
    for a, b in zip(M0 + dY0, M1 + dY1):
        assert torch.equal(a, b)
    assert torch.equal(gd0[: int(g.num_pairs)], gd1[: int(g.num_pairs)])


def test_neighbor_embedding_fused_vs_module():
    """row a13: fused NeighborEmbedding sum (per-pair radial factor + tmd_nbr_embed_fwd/bwd) == the module's torch-op forward
    (reference models/utils.py:83-117), value and position gradient, through tables and through the exact chain"""
    from torchmdnet_b200.extensions import fused
    from torchmdnet_b200.models.utils import NeighborEmbedding, ExpNormalSmearing
    torch.manual_seed(5)
    R, F = 32, 64
    z, pos, g = graph_for(n=120, seed=7)
    n = z.numel()
    ne = NeighborEmbedding(F, R, 0.0, 5.0, max_z=16).to(dev)
    for p_ in ne.parameters():
        p_.requires_grad_(False)
    rbf = ExpNormalSmearing(0.0, 5.0, R, trainable=False).to(dev)
    x = torch.randn(n, F, device=dev)
    pidx = fused.build_pair_index(g)
    E, P = int(g.num_pairs), int(pidx.num_unique)
    dist = g.dist[:E]
    ei = torch.stack([g.nbr_i[:E].long(), g.nbr_j[:E].long()])
    ref = ne(z, x, ei, dist, rbf(dist))
    w = torch.randn_like(ref)
    (g_ref,) = torch.autograd.grad((ref * w).sum(), [pos], retain_graph=True)
    for table in (True, False):
        fused_tab = fused.USE_RADIAL_TABLE
        fused.USE_RADIAL_TABLE = table
        try:
            out = ne.forward_fused(z, x, dist, pidx.pair_r[:P], g, pidx, rbf, None)
            (g_out,) = torch.autograd.grad((out * w).sum(), [pos], retain_graph=True)
        finally:
            fused.USE_RADIAL_TABLE = fused_tab
        assert (out - ref).abs().max() < 2e-5 * (1 + ref.abs().max())
        assert (g_out - g_ref).abs().max() < 2e-4 * (1 + g_ref.abs().max())
