"""-m gpu: fused TensorNet tail + Scalar head + per-molecule reduction (csrc/head.cu, rows a10 + a11) against the torch-op chain
it replaces: norm3 -> LayerNorm -> Linear -> SiLU -> MLP -> * std -> + atomref -> scatter-sum -> + mean, value and gradient
w.r.t. the node tensors (the seed of the force pass)."""
import random

import numpy as np
import pytest
import torch

import systems

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def seeded_model(hp, seed, prior=False):
    from torchmdnet_b200.models.model import create_model
    from torchmdnet_b200 import priors
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    pm = None
    if prior:
        pm = priors.Atomref(max_z=hp["max_z"])
        pm.atomref.weight.data.copy_(torch.randn(hp["max_z"], 1))
    m = create_model(dict(hp), prior_model=pm, mean=torch.tensor(0.37), std=torch.tensor(1.9))
    return m.eval().to(dev)


@pytest.mark.parametrize("rows_gemm", [True, False])
@pytest.mark.parametrize("F,n_mol,prior", [(128, 1, False), (64, 5, True), (192, 3, False), (256, 2, True)])
def test_head_kernels_vs_torch_ops(F, n_mol, prior, rows_gemm, monkeypatch):
    """both forms of rows a10 + a11: weight-tiled GEMM launches (csrc/rows_gemm.cu, default) and the one-launch-per-32-atom-tile
    kernels (csrc/head.cu, TMD_ROWS_GEMM=0)"""
    from torchmdnet_b200.extensions import fused, tensornet_ops as ops
    monkeypatch.setattr(fused, "USE_ROWS_GEMM", rows_gemm)
    from torchmdnet_b200.models.utils import scatter
    hp = systems.tensornet_hparams(num_layers=1, F=F, R=16, max_nb=64)
    m = seeded_model(hp, 3, prior)
    for p_ in m.parameters():
        p_.requires_grad_(False)
    rep, out = m.representation_model, m.output_model
    torch.manual_seed(0)
    N = 37 * n_mol + 5
    X = [(0.7 * torch.randn(N, k, F, device=dev)).requires_grad_(True) for k in (1, 3, 5)]
    z = torch.randint(1, 9, (N,), device=dev)
    batch = torch.sort(torch.randint(0, n_mol, (N,), device=dev)).values
    hw = rep.head_weights(out.output_network)
    assert hw.supported() and hw.frozen()
    atomref = m.prior_model[0].atomref.weight if prior else None
    y = fused.tail_head_energy(X, z, batch, n_mol, hw, m.std, m.mean, atomref)
    gy = torch.randn(n_mol, 1, device=dev)
    g = torch.autograd.grad(y, X, gy)
    # torch-op chain (what TorchMD_Net.forward runs with fused_head_path = False)
    X2 = [t.detach().clone().requires_grad_(True) for t in X]
    x = rep.act(rep.linear(rep.out_norm(ops.norm3_compact(*X2))))
    e = out.pre_reduce(x, None, z, None, batch) * m.std
    if prior:
        e = m.prior_model[0].pre_reduce(e, z, None, batch)
    y2 = scatter(e, batch, dim=0, dim_size=n_mol, reduce="sum") + m.mean
    g2 = torch.autograd.grad(y2, X2, gy)
    assert (y - y2).abs().max() <= 2e-5 * (1 + y2.abs().max())
    for a, b in zip(g, g2):
        assert (a - b).abs().max() <= 2e-5 * (1e-6 + b.abs().max())
    # unsorted batch ids: the reduction falls back to the scan path, same sums
    perm = torch.randperm(N, device=dev)
    y3 = fused.tail_head_energy([t.detach()[perm] for t in X], z[perm], batch[perm], n_mol, hw, m.std, m.mean, atomref)
    assert (y3 - y2).abs().max() <= 2e-5 * (1 + y2.abs().max())


def test_model_selects_fused_head_and_matches_torch_head():
    hp = systems.tensornet_hparams(num_layers=2, F=128, R=32, max_nb=96)
    m = seeded_model(hp, 5)
    z, pos, batch, box = systems.random_box(300, seed=1)
    args = (z.to(dev), pos.to(dev), batch.to(dev))
    assert m._fused_head() is None                 # parameters require grad: torch-op head (weight gradients stay available)
    for p_ in m.parameters():
        p_.requires_grad_(False)
    assert m._fused_head() is not None
    y1, f1 = m(*args, box=box.to(dev))
    m.fused_head_path = False
    y0, f0 = m(*args, box=box.to(dev))
    assert (y1 - y0).abs().max() <= 1e-5 * (1 + y0.abs().max()) and (f1 - f0).abs().max() <= 2e-5 * f0.abs().max()


@pytest.mark.parametrize("F,prior", [(128, False), (64, True)])
def test_et_head_kernels_vs_torch_head(F, prior):
    """row a15: the two gated equivariant blocks of EquivariantScalar + std / atomref / per-molecule sum / mean as one autograd node
    (fused.et_head_energy: GEMM launches + the geb_* kernels, hand-sequenced backward) == the torch-op head, energies and forces"""
    from torchmdnet_b200.models.model import create_model
    from torchmdnet_b200 import priors
    hp = systems.et_hparams(num_layers=2, F=F, R=32, H=4, cutoff=5.0, max_nb=64)   # output_model "Scalar" -> EquivariantScalar for the ET
    random.seed(7); np.random.seed(7); torch.manual_seed(7)
    pm = None
    if prior:
        pm = priors.Atomref(max_z=hp["max_z"])
        pm.atomref.weight.data.copy_(torch.randn(hp["max_z"], 1))
    m = create_model(dict(hp), prior_model=pm, mean=torch.tensor(-0.21), std=torch.tensor(1.7)).eval().to(dev)
    z, pos, batch, _ = systems.molecule_batch(n_mol=5, lo=9, hi=15, spread=3.0, seed=3)
    args = (z.to(dev), pos.to(dev), batch.to(dev))
    assert type(m.output_model).__name__ == "EquivariantScalar"
    assert m._fused_et_head() is None               # parameters require grad: torch-op head
    for p_ in m.parameters():
        p_.requires_grad_(False)
    assert m._fused_et_head() is not None
    y1, f1 = m(*args)
    m.fused_head_path = False
    y0, f0 = m(*args)
    assert (y1 - y0).abs().max() <= 2e-5 * (1 + y0.abs().max()), float((y1 - y0).abs().max())
    assert (f1 - f0).abs().max() <= 2e-5 * (1e-6 + f0.abs().max()), float((f1 - f0).abs().max() / f0.abs().max())
