"""-m gpu: the reference's torch.library op names resolve to the sm_100a kernels (op-registry seam)."""
import pytest
import torch

import systems
import tensornet_oracle as orc

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def test_registered_names():
    from torchmdnet_b200.extensions import registry
    assert len(registry.REGISTERED) == 25   # the reference registers 25 op names
    for n in registry.REGISTERED:
        ns, op = n.split("::")
        assert hasattr(getattr(torch.ops, ns), op)


def test_neighbor_op_through_torch_ops():
    from torchmdnet_b200.extensions import registry  # noqa: F401
    z, pos, batch, box = systems.random_box(120, seed=3)
    ei, v, d = orc.neighbor_list(pos, batch, box, 0.0, 4.0, True, True)
    p = pos.to(dev).requires_grad_(True)
    nb, vec, dist, npairs = torch.ops.torchmdnet.warp_neighbor_brute_fwd(p, batch.to(dev), box.to(dev), True, 0.0, 4.0, ei.shape[1] + 5, True, True)
    n = int(npairs)
    assert torch.equal(nb[:, :n].cpu(), ei) and torch.equal(dist[:n].cpu(), d)
    w = torch.randn(n, device=dev)
    (dist[:n] * w).sum().backward()
    pc = pos.clone().requires_grad_(True)
    _, _, dd = orc.neighbor_list_from_pos(pc, batch, box, dict(cutoff_lower=0.0, cutoff_upper=4.0), loop=True)
    (dd * w.cpu()).sum().backward()
    assert torch.allclose(p.grad.cpu(), pc.grad, atol=1e-4)
    out = torch.ops.torchmdnet.warp_neighbor_cell_fwd(pos.to(dev), batch.to(dev), box.to(dev), True, 0.0, 4.0, ei.shape[1] + 5, True, True, 64)
    assert torch.equal(out[0][:, :n].cpu(), ei)


def test_tensornet_ops_through_torch_ops():
    from torchmdnet_b200.extensions import registry  # noqa: F401
    from torchmdnet_b200.extensions import tensornet_ops as T
    N, F = 33, 32
    I, A, S = (torch.randn(N, k, F, device=dev, requires_grad=True) for k in (1, 3, 5))
    X = torch.ops.tensornet.compose_tensor_fwd_primitive(I, A, S)
    assert torch.allclose(X, T.fn_compose_tensor(I, A, S))
    I2, A2, S2 = torch.ops.tensornet.decompose_tensor_fwd_primitive(X)
    assert torch.allclose(I2, I, atol=1e-6) and torch.allclose(A2, A, atol=1e-6) and torch.allclose(S2, S, atol=1e-6)
    Y = torch.randn(N, 3, 3, F, device=dev, requires_grad=True)
    Z = torch.ops.tensornet.tensor_matmul_o3_3x3_fwd_primitive(X, Y)
    nrm = torch.ops.tensornet.tensor_norm3_fwd_primitive(Z)
    g1 = torch.autograd.grad(nrm.sum(), [I, A, S, Y], retain_graph=True)
    ref = T.fn_tensor_norm3(T.fn_tensor_matmul_o3_3x3(T.fn_compose_tensor(I, A, S), Y))
    g2 = torch.autograd.grad(ref.sum(), [I, A, S, Y])
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5)
    # message passing through the registry, graph from graph_transform primitives
    ii, jj = torch.nonzero(torch.rand(N, N, device=dev) < 0.3, as_tuple=True)
    ei = torch.stack([ii, jj]).int()
    rc, cc = torch.ops.nvtnet.count_row_col_primitive(ei, N)
    rp, cp = torch.cumsum(rc, 0, dtype=torch.int32), torch.cumsum(cc, 0, dtype=torch.int32)
    ri, ci, rd, cd = torch.ops.nvtnet.convert_to_sparse_primitive(ei, rc, cc, rp, cp)
    w = torch.randn(ei.shape[1], 3, F, device=dev, requires_grad=True)
    out = torch.ops.tensornet.message_passing_fwd_primitive(I, A, S, w, rd, ri, rp, cd, ci, cp)
    ref = [torch.zeros_like(t).index_add(0, ii, w[:, k:k + 1] * t[jj]) for k, t in enumerate((I, A, S))]
    for a, b in zip(out, ref):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)
    g1 = torch.autograd.grad(sum(o.sum() for o in out), [I, w], retain_graph=True)
    g2 = torch.autograd.grad(sum(o.sum() for o in ref), [I, w])
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)
