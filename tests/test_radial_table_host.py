"""-m "not gpu": the math of the radial tables on the host.

The table construction of `RadialChain.table` (fp64 nodes at i*2^-k, rows (f_i, f_{i+1}-f_i, h f'_i, h f'_{i+1})) is repeated
here on the CPU and evaluated with the fp32 arithmetic of `radial_table_kernel` (csrc/radial_table.cu) restated in numpy; it must
reproduce the fp64 chain -- the same function the oracle evaluates per edge (oracle/tensornet_oracle.py) -- to the gate's 2e-6
for the value and the derivative.  The GPU tests check the kernel itself against this construction (test_gpu_fused.py, 'table')."""
import math

import numpy as np
import pytest
import torch

from torchmdnet_b200.extensions import fused
from torchmdnet_b200.models.utils import ExpNormalSmearing


def hermite_f32(table, h, r):
    """numpy float32 restatement of radial_table_kernel: t = r/h - i is exact because h is a power of two."""
    f32 = np.float32
    n_int = table.shape[0]
    x = np.maximum(r.astype(f32), f32(0)) * f32(1.0 / h)
    i = np.minimum(x.astype(np.int64), n_int - 1)
    t = np.minimum(x - i.astype(f32), f32(1))[:, None]
    f0, D, m0, m1 = (table[i, k] for k in range(4))
    u = f32(1) - t
    a, b, g = t * t * (f32(3) - f32(2) * t), t * u * u, -t * t * u
    val = a * D + (b * m0 + (g * m1 + f0))
    inv_h = f32(1.0 / h)
    da, db, dg = f32(6) * t * u * inv_h, u * (f32(1) - f32(3) * t) * inv_h, t * (f32(3) * t - f32(2)) * inv_h
    der = da * D + (db * m0 + dg * m1)
    return val.astype(f32), der.astype(f32)


@pytest.mark.parametrize("cfg", ["interaction", "embedding"])
def test_table_reproduces_the_fp64_chain(cfg):
    torch.manual_seed(0)
    R, F, lo, hi = 32, 64, 0.0, 5.0
    rbf = ExpNormalSmearing(lo, hi, R, trainable=False)
    if cfg == "interaction":
        lins = [torch.nn.Linear(R, F), torch.nn.Linear(F, 2 * F), torch.nn.Linear(2 * F, 3 * F)]
        chain = fused.RadialChain([[l] for l in lins], act=True, mul_cutoff=True)
    else:
        lins = [torch.nn.Linear(R, F) for _ in range(3)]
        chain = fused.RadialChain([lins], act=False, mul_cutoff=False)
    accepted = None
    for k in (9, 10, 11, 12):
        h = 2.0 ** -k
        n_int = int(math.ceil(hi / h))
        nodes = torch.arange(n_int + 1, dtype=torch.float64) * h
        assert torch.equal(nodes.float().double(), nodes)            # node positions are exact in fp32
        f, fp = chain._eval_fp64(rbf, lo, hi, nodes)
        m = fp * h
        table = torch.stack([f[:-1], f[1:] - f[:-1], m[:-1], m[1:]], 1).float().numpy()
        rt = ((torch.arange(n_int, dtype=torch.float64).unsqueeze(1) + torch.tensor([0.25, 0.5, 0.75], dtype=torch.float64)) * h).reshape(-1)
        rt = rt[rt < hi]
        f64, fp64 = chain._eval_fp64(rbf, lo, hi, rt)
        val, der = hermite_f32(table, h, rt.numpy())
        e_f = np.abs(val - f64.numpy()).max() / f64.abs().max().item()
        e_fp = np.abs(der - fp64.numpy()).max() / fp64.abs().max().item()
        if e_f < fused.TABLE_TOL and e_fp < fused.TABLE_TOL:
            accepted = (k, e_f, e_fp)
            break
    assert accepted is not None, "no table spacing down to 2^-12 met the gate"
    print(cfg, "k=%d err_f=%.1e err_fp=%.1e" % accepted)


def test_fp64_chain_matches_torch_autograd():
    """`_eval_fp64` (forward-mode) is the function the reference evaluates per edge: compare with reverse-mode autograd of the
    module composition (ExpNormalSmearing -> Linear/SiLU x3 -> CosineCutoff, reference models/tensornet.py:738-743)."""
    from torchmdnet_b200.models.utils import CosineCutoff
    torch.manual_seed(1)
    R, F, lo, hi = 16, 32, 0.0, 4.5
    rbf = ExpNormalSmearing(lo, hi, R, trainable=False, dtype=torch.float64)
    lins = [torch.nn.Linear(R, F).double(), torch.nn.Linear(F, 2 * F).double(), torch.nn.Linear(2 * F, 3 * F).double()]
    chain = fused.RadialChain([[l] for l in lins], act=True, mul_cutoff=True)
    r = (torch.rand(200, dtype=torch.float64) * hi).requires_grad_(True)
    x = rbf(r)
    for l in lins:
        x = torch.nn.functional.silu(l(x))
    ref = x * CosineCutoff(lo, hi)(r).unsqueeze(-1)
    f, fp = chain._eval_fp64(rbf, lo, hi, r.detach())
    assert (f - ref).abs().max() < 1e-13
    w = torch.randn_like(ref)
    (g,) = torch.autograd.grad((ref * w).sum(), r)
    assert ((fp * w).sum(-1) - g).abs().max() < 1e-11
