"""CPU checks of the host-side logic added in round 2 (no kernel launches):
the algebra behind the moment form of the embedding (csrc/embed_moments.cu), element slots, the cooperative-group size chosen through
the C ABI, the whole-frame op's fake implementation (torch.compile support) and the fall-backs of the frozen-Linear wrapper."""
import math

import torch

import systems  # noqa: F401  (puts the repo on the path through conftest)


def _cos_cut(d, lo, hi):
    if lo > 0:
        return 0.5 * (torch.cos(math.pi * (2 * (d - lo) / (hi - lo) + 1.0)) + 1.0) * (d < hi) * (d > lo)
    return 0.5 * (torch.cos(d * math.pi / hi) + 1.0) * (d < hi)


def test_embedding_moment_factorisation_identity():
    """sum_e C (U_i + V_{z_j}) (W_g rbf_e + b_g) phi_c(rhat_e)  ==  sum_v (U_i + V_v) sum_k W~_g[:,k] m[v][k][c]  with the moments
    m[v][k][c] = sum_{e: z_j = v} C phi_c rho_k (rho_R = 1): the exact reordering csrc/embed_moments.cu evaluates, in fp64"""
    torch.manual_seed(0)
    E, R, F, Z = 57, 16, 24, 3
    lo, hi = 0.0, 5.0
    d = torch.rand(E, dtype=torch.float64) * 4.8 + 0.1
    rhat = torch.nn.functional.normalize(torch.randn(E, 3, dtype=torch.float64), dim=1)
    zj = torch.randint(0, Z, (E,))
    U, V = torch.randn(F, dtype=torch.float64), torch.randn(Z, F, dtype=torch.float64)
    W, b = torch.randn(3, F, R, dtype=torch.float64), torch.randn(3, F, dtype=torch.float64)
    means, betas = torch.linspace(0.05, 1.0, R, dtype=torch.float64), torch.full((R,), 30.0, dtype=torch.float64)
    alpha = 5.0 / (hi - lo)
    rbf = _cos_cut(d, 0.0, hi)[:, None] * torch.exp(-betas * (torch.exp(alpha * (lo - d))[:, None] - means) ** 2)    # [E,R]
    C = _cos_cut(d, lo, hi)
    a = torch.stack([-rhat[:, 2], rhat[:, 1], -rhat[:, 0]], 1)
    m3 = (rhat * rhat).sum(1) / 3
    sb = torch.stack([rhat[:, 0] ** 2 - m3, rhat[:, 0] * rhat[:, 1], rhat[:, 0] * rhat[:, 2], rhat[:, 1] ** 2 - m3, rhat[:, 1] * rhat[:, 2]], 1)
    phi = torch.cat([torch.ones(E, 1, dtype=torch.float64), a, sb], 1)                                               # [E,9]
    grp = [0, 1, 1, 1, 2, 2, 2, 2, 2]
    # direct: per edge weights, then the neighbour sum
    direct = torch.zeros(9, F, dtype=torch.float64)
    for c in range(9):
        w = C[:, None] * (U[None] + V[zj]) * (rbf @ W[grp[c]].t() + b[grp[c]])                                      # [E,F]
        direct[c] = (w * phi[:, c:c + 1]).sum(0)
    # moments first
    rho = torch.cat([rbf, torch.ones(E, 1, dtype=torch.float64)], 1)                                                  # [E,R+1]
    mom = torch.zeros(Z, R + 1, 9, dtype=torch.float64)
    mom.index_add_(0, zj, C[:, None, None] * rho[:, :, None] * phi[:, None, :])
    Wt = torch.cat([W, b[:, :, None]], 2)                                                                             # [3,F,R+1]
    fact = torch.zeros(9, F, dtype=torch.float64)
    for c in range(9):
        for v in range(Z):
            fact[c] += (U + V[v]) * (Wt[grp[c]] @ mom[v, :, c])
    assert (direct - fact).abs().max() < 1e-11 * direct.abs().max()


def test_species_slots_and_moment_weight_layouts():
    from torchmdnet_b200.extensions import fused
    z = torch.tensor([8, 1, 1, 6, 8, 1, 17])
    slots = fused.SpeciesSlots()
    zslot, species = slots.get(z)
    assert species.tolist() == [1, 6, 8, 17] and zslot.dtype == torch.int32
    assert torch.equal(species[zslot.long()], z)
    assert slots.get(z)[0] is zslot                                   # cached against the identity of z
    z2 = z.clone()
    assert slots.get(z2)[0] is not zslot                              # another tensor: recomputed
    assert fused.SpeciesSlots().get(torch.arange(1, fused.MAX_SPECIES + 2)) is None     # too many elements -> chain path
    lins = [torch.nn.Linear(8, 12) for _ in range(3)]
    mom = fused.EmbedMomentWeights(lins)
    Wk, Wh = mom.tensors()
    assert Wk.shape == (3, 9, 12) and Wh.shape == (3, 3, 8, 4)
    for g, lin in enumerate(lins):
        assert torch.equal(Wk[g, :8], lin.weight.detach().t()) and torch.equal(Wk[g, 8], lin.bias.detach())
        for h in range(12):
            assert torch.equal(Wh[g, h // 4, :, h % 4], lin.weight.detach()[h])
    with torch.no_grad():
        lins[0].weight.mul_(2.0)                                       # in-place update -> layouts follow the new version
    assert torch.equal(mom.tensors()[0][0, :8], lins[0].weight.detach().t())


def test_mp_group_size_through_the_c_abi():
    """tmd_mp_pick_group: rows per cooperative group in [G-3, G] minimising rounds x rows on 148 one-CTA slots"""
    import __graft_entry__ as g
    g.build()
    from torchmdnet_b200 import _lib
    lib = _lib.library()
    G = lib.tmd_mp_group_size()
    assert G == 15
    for n in (1, 100, 1000, 10125, 40000, 123457):
        gs = lib.tmd_mp_pick_group(n)
        assert G - 3 <= gs <= G
        cost = lambda k: -(-(-(-n // k)) // 148) * k
        assert cost(gs) == min(cost(k) for k in range(G - 3, G + 1))
    assert lib.tmd_mp_pick_group(10125) == 14                          # the C4 frame: 724 groups = 4.89 rounds


def test_frame_op_fake_and_rows_linear_fallbacks():
    from torch._subclasses.fake_tensor import FakeTensorMode
    from torchmdnet_b200.extensions import frame_op, fused  # noqa: F401  (registers torchmdnet_b200::energy_forces)
    from torchmdnet_b200.models.model import create_model
    with FakeTensorMode():
        z = torch.zeros(10, dtype=torch.long, device="cuda")
        pos = torch.zeros(10, 3, device="cuda")
        y, f = torch.ops.torchmdnet_b200.energy_forces(1, z, pos, z, None, None, 2)
        assert y.shape == (2, 1) and f.shape == (10, 3) and y.device.type == "cuda"
    import copy
    m = create_model(dict(systems.tensornet_hparams(num_layers=1, F=32, R=16, max_nb=32)))
    m2 = copy.deepcopy(m)
    assert m._frame_op_handle != m2._frame_op_handle and frame_op._models[m2._frame_op_handle] is m2
    assert "_frame_op_handle" not in m.state_dict()
    # the frozen-Linear wrapper falls back to the module itself whenever the GEMM launcher does not apply (here: CPU tensors)
    lin = torch.nn.Linear(32, 64)
    x = torch.randn(5, 32)
    assert torch.equal(fused.rows_linear(x, lin), lin(x))


def test_torchscript_in_process_wrapper():
    """torch.jit.script(model) (reference tests/test_model.py:207-262) compiles: the scripted graph is one call of the whole-frame op
    with TorchMD_Net.forward's signature; the model is pinned so the handle stays valid.  (Execution needs a GPU; the op body is the one
    `test_torch_compile_fullgraph` runs.)"""
    import io
    from torchmdnet_b200.extensions import frame_op
    from torchmdnet_b200.models.model import create_model
    m = create_model(dict(systems.tensornet_hparams(num_layers=1, F=32, R=16, max_nb=32)))
    sm = torch.jit.script(m)
    assert "torchmdnet_b200::energy_forces" in str(sm.graph)
    names = [a.name for a in sm.forward.schema.arguments]
    assert names == ["self", "z", "pos", "batch", "box", "q", "s", "extra_args", "num_systems"]
    assert frame_op._pinned[m._frame_op_handle] is m
    buf = io.BytesIO()
    torch.jit.save(sm, buf)                       # serialisable; loadable where this package (and the model) lives
    buf.seek(0)
    sm2 = torch.jit.load(buf)
    assert "torchmdnet_b200::energy_forces" in str(sm2.graph)
