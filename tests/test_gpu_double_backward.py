"""-m gpu: double backward: `create_graph=True` through every kernel-backed op, and one force-matching
training step of the whole model against the CPU oracle's parameter gradients.

Reference: `TorchMD_Net.forward` calls autograd.grad(y, pos, create_graph=self.training) (models/model.py:618-628); the loss on
the forces is then back-propagated to the weights through the `*_bwd_bwd_primitive` kernels (warp_kernels/*.py); its own test is
tests/test_model.py:331-352 (gradcheck of the whole model)."""
import random

import numpy as np
import pytest
import torch

import systems
import tensornet_oracle as orc

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def second_order(fn, inputs, seed=0):
    """(first-order grads with create_graph, grads of a scalar of them w.r.t. inputs and the upstream weights)"""
    g = torch.Generator(device="cpu").manual_seed(seed)
    outs = fn(*inputs)
    outs = list(outs) if isinstance(outs, (tuple, list)) else [outs]
    ws = [torch.randn(o.shape, generator=g).to(o.device).requires_grad_(True) for o in outs]
    need = [t for t in inputs if t.requires_grad]
    g1 = torch.autograd.grad(sum((o * w).sum() for o, w in zip(outs, ws)), need, create_graph=True)
    vs = [torch.randn(t.shape, generator=g).to(t.device) for t in g1]
    g2 = torch.autograd.grad(sum((a * v).sum() for a, v in zip(g1, vs)), need + ws, allow_unused=True)
    return g1, g2


def compare(fn_kernel, fn_torch, make_inputs, tol=2e-4):
    a = second_order(fn_kernel, make_inputs())
    b = second_order(fn_torch, make_inputs())
    for ga, gb in zip(a[0] + a[1], b[0] + b[1]):
        if gb is None:
            assert ga is None or float(ga.abs().max()) == 0.0
            continue
        assert ga is not None
        assert (ga - gb).abs().max() <= tol * (1e-6 + gb.abs().max()), float((ga - gb).abs().max() / gb.abs().max())


N, F = 40, 32


def leaf(*shape, seed):
    g = torch.Generator().manual_seed(seed)
    return (0.8 * torch.randn(*shape, generator=g)).to(dev).requires_grad_(True)


def test_granular_ops_double_backward():
    from torchmdnet_b200.extensions import tensornet_ops as T
    X = lambda s: leaf(N, 3, 3, F, seed=s)
    compare(T.fn_compose_tensor, T._t_compose, lambda: [leaf(N, 1, F, seed=1), leaf(N, 3, F, seed=2), leaf(N, 5, F, seed=3)])
    compare(T.fn_decompose_tensor, T._t_decompose, lambda: [X(4)])
    compare(T.fn_tensor_matmul_o3_3x3, lambda a, b: T._t_mm(a, b) + T._t_mm(b, a), lambda: [X(5), X(6)])
    compare(T.fn_tensor_matmul_so3_3x3, T._t_mm, lambda: [X(7), X(8)])
    compare(T.fn_tensor_norm3, lambda x: torch.cat(T._t_norm3(*T._t_decompose(x)), -1), lambda: [X(9)])


def test_compact_stages_double_backward():
    from torchmdnet_b200.extensions import tensornet_ops as T
    c = lambda s: [leaf(N, 1, F, seed=s), leaf(N, 3, F, seed=s + 1), leaf(N, 5, F, seed=s + 2)]
    qf = (torch.rand(N, generator=torch.Generator().manual_seed(3)) + 0.5).to(dev)
    compare(T.interaction_normalize, T._t_p1, lambda: c(10))
    for o3 in (True, False):
        compare(lambda *t: T.interaction_product(*t, qf, o3), lambda *t: T._t_p2(*t, qf, o3), lambda: c(20) + c(30))
    compare(lambda *t: T.interaction_update(*t, qf), lambda *t: T._t_p3(*t, qf), lambda: c(40) + c(50))
    compare(T.norm2_compact, T._t_norm2, lambda: c(60))
    compare(T.norm3_compact, lambda *t: torch.cat(T._t_norm3(*t), -1), lambda: c(70))
    compare(T.gate_compact, lambda i, a, s, g: (i * g[:, 0:1], a * g[:, 1:2], s * g[:, 2:3]), lambda: c(80) + [leaf(N, 3, F, seed=83)])


def test_dependent_inputs_double_backward():
    """stage inputs that are functions of one another (the model's P2 gets (Y, M(Y)), P3 gets (X, dX(X))): the create_graph branch
    must return PARTIAL derivatives per input, autograd adds the chain through the other input itself"""
    from torchmdnet_b200.extensions import tensornet_ops as T
    c = lambda s: [leaf(N, 1, F, seed=s), leaf(N, 3, F, seed=s + 1), leaf(N, 5, F, seed=s + 2)]
    qf = (torch.rand(N, generator=torch.Generator().manual_seed(5)) + 0.5).to(dev)

    def chain(p1, p2, p3):
        def fn(I, A, S):
            Y = p1(I, A, S)
            M = p1(Y[0] * 1.5, Y[1] - 0.3 * A, Y[2] + S)          # M depends on Y (stand-in for the message passing)
            C = p2(Y[0], Y[1], Y[2], M[0], M[1], M[2], qf, True)
            D = (0.7 * C[0] + Y[0], C[1] * 1.1, C[2] - Y[2])       # dX depends on X = Y
            return p3(Y[0], Y[1], Y[2], D[0], D[1], D[2], qf)
        return fn

    compare(chain(T.interaction_normalize, T.interaction_product, T.interaction_update), chain(T._t_p1, T._t_p2, T._t_p3), lambda: c(90),
            tol=5e-4)


def test_message_passing_double_backward():
    """interaction and embedding message passing on a real neighbour graph: the interaction op re-enters its own kernels
    (tmd_mp_fwd / tmd_mp_bwd) in the double backward"""
    from torchmdnet_b200.extensions import tensornet_ops as T
    from torchmdnet_b200.extensions.neighbors import neighbor_graph
    z, pos, batch, box = systems.random_box(60, seed=1)
    g = neighbor_graph(pos.to(dev), batch.to(dev), box.to(dev), True, 0.0, 5.0, True, 60 * 96)
    E, n = int(g.num_pairs), 60
    i, j = g.nbr_i[:E].long(), g.nbr_j[:E].long()
    idx, rev, ptr = g.nbr_j[:E].contiguous(), g.rev[:E].contiguous(), g.indptr

    def mp_torch(I, A, S, w):
        o = lambda T_, k: torch.zeros_like(T_).index_add(0, i, w[:, k:k + 1] * T_[j])
        return o(I, 0), o(A, 1), o(S, 2)

    mk = lambda: [leaf(n, 1, F, seed=1), leaf(n, 3, F, seed=2), leaf(n, 5, F, seed=3), leaf(E, 3, F, seed=4)]
    compare(lambda I, A, S, w: T.fn_message_passing(I, A, S, w, None, idx, ptr, rev, idx, ptr), mp_torch, mk)
    rhat = torch.nn.functional.normalize(g.vec[:E].detach() + 1e-3, dim=1)
    mk2 = lambda: [rhat.clone().requires_grad_(True), leaf(E, 3, F, seed=5)]
    compare(lambda r, w: T.fn_radial_message_passing(r, w, None, ptr), lambda r, w: T._t_radial_mp(r, w, None, ptr), mk2)


def test_bwd_bwd_primitives_registered_and_consistent():
    """the reference's `tensornet::*_bwd_bwd_primitive` names resolve and agree with autograd's double backward of the same op"""
    import torchmdnet_b200.extensions.registry  # noqa: F401
    from torchmdnet_b200.extensions import tensornet_ops as T
    x, y = leaf(N, 3, 3, F, seed=1), leaf(N, 3, 3, F, seed=2)
    go, ggx, ggy = leaf(N, 3, 3, F, seed=3), leaf(N, 3, 3, F, seed=4), leaf(N, 3, 3, F, seed=5)
    ggo, dx, dy = torch.ops.tensornet.tensor_matmul_o3_3x3_bwd_bwd_primitive(go.detach(), ggx.detach(), ggy.detach(), x.detach(), y.detach())
    z = T._t_mm(x, y) + T._t_mm(y, x)
    gx, gy = torch.autograd.grad(z, [x, y], go, create_graph=True)
    r = torch.autograd.grad([gx, gy], [go, x, y], [ggx.detach(), ggy.detach()])
    for a, b in zip((ggo, dx, dy), r):
        assert (a - b).abs().max() <= 2e-4 * b.abs().max()
    ggo, dxx = torch.ops.tensornet.tensor_norm3_bwd_bwd_primitive(ggx.detach(), x.detach(), leaf(N, 3 * F, seed=6).detach())
    assert ggo.shape == (N, 3 * F) and dxx.shape == x.shape


@pytest.mark.parametrize("group", ["O(3)", "SO(3)"])
def test_force_matching_training_step_vs_oracle(group):
    """loss = sum E^2 + sum |F|^2 in training mode: parameter gradients of the CUDA model == the oracle's (fp64), 5e-3 of max|grad| per tensor"""
    from torchmdnet_b200.models.model import create_model
    hp = systems.tensornet_hparams(num_layers=2, F=32, R=16, max_nb=64, group=group)
    random.seed(4); np.random.seed(4); torch.manual_seed(4)
    m = create_model(dict(hp))
    sd64 = {k: v.detach().double().requires_grad_(v.is_floating_point() and v.dim() > 0) for k, v in m.state_dict().items()}
    z, pos, batch, _ = systems.molecule_batch(n_mol=3, lo=8, hi=12, spread=3.5, seed=2)
    q = torch.tensor([0.0, 1.0, -1.0])
    y, f = orc.energy_and_forces(sd64, hp, z, pos.double(), batch, None, q.double(), create_graph=True)
    loss_ref = (y ** 2).sum() + (f ** 2).sum()
    names = [k for k, v in sd64.items() if v.requires_grad]
    g_ref = dict(zip(names, torch.autograd.grad(loss_ref, [sd64[k] for k in names], allow_unused=True)))
    m = m.to(dev).train()
    y, f = m(z.to(dev), pos.to(dev), batch.to(dev), q=q.to(dev))
    assert f.requires_grad, "forces must stay attached to the graph in training mode"
    loss = (y ** 2).sum() + (f ** 2).sum()
    assert abs(loss.item() - loss_ref.item()) <= 5e-4 * abs(loss_ref.item())   # fp32 forces enter squared
    loss.backward()
    params = dict(m.named_parameters())
    checked = 0
    for k, gr in g_ref.items():
        if k not in params or gr is None:
            continue
        gp = params[k].grad
        assert gp is not None, f"{k}: no gradient (detached in the double backward?)"
        scale = gr.abs().max().item()
        if scale < 1e-12:
            continue
        err = (gp.double().cpu() - gr).abs().max().item() / scale
        assert err < 5e-3, (k, err)   # second derivatives in fp32 vs the fp64 oracle
        checked += 1
    assert checked >= 20
