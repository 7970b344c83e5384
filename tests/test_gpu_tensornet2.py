"""-m gpu: TensorNet2 + ScalarPlusWeightedCoulomb assembled from the TensorNet kernels (models/tensornet2.py)
against the reference outputs in tests/golden/next/ (oracle/make_golden_tn2.py: the unmodified reference in fp64 and fp32).
The oracle side of the row is pinned (tests/test_oracle.py::test_oracle_tensornet2_pinned) and the host side is covered on the
CPU (tests/test_tensornet2_host.py)."""
import os

import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("static_shapes", [False, True])
@pytest.mark.parametrize("name", ["tn2_mol4_alltoall", "tn2_box120_rf"])
def test_tensornet2_golden_energy_forces(name, static_shapes):
    from torchmdnet_b200.models.model import create_model
    dev = torch.device("cuda:0")
    fx = torch.load(os.path.join(GOLDEN, "next", name + ".pt"), weights_only=False)
    hp = dict(fx["hparams"]); hp["static_shapes"] = static_shapes
    m = create_model(hp)
    m.load_state_dict(fx["state_dict"])
    m = m.eval().to(dev)
    g = lambda t: None if t is None else t.to(dev)
    y, f = m(fx["z"].to(dev), fx["pos"].to(dev).clone(), fx["batch"].to(dev), box=g(fx["box"]), q=g(fx["q"]))
    eE = ((y.double().cpu() - fx["ref_E64"]).abs().max() / max(1.0, fx["ref_E64"].abs().max().item())).item()
    eF = ((f.double().cpu() - fx["ref_F64"]).abs().max() / fx["ref_F64"].abs().max()).item()
    print(f"{name} static={static_shapes}: relE={eE:.2e} relF={eF:.2e}")
    assert eE < TOL and eF < TOL
