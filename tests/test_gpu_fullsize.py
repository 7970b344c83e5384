"""-m gpu: parity at BASELINE.json's FULL sizes: the timed C4 frame (10 125 atoms: cell-grid neighbour
search, fp16 operand range bound, radial-table gate on that model, capacity max_num_neighbors=64), C3 (ET-SPICE, 256 molecules) and
C5 per GPU (8 x 5000-atom boxes) against fp64 oracle outputs committed under tests/golden/full (oracle/make_golden_full.py; inputs
and weights are reproduced from their seeds, the fixture stores E, F and a digest of the inputs)."""
import hashlib
import os
import random

import numpy as np
import pytest
import torch

import systems
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
TOL = 1e-4   # north_star: energies / forces within 1e-4 relative fp32


def seeded_model(hp, seed):
    from torchmdnet_b200.models.model import create_model
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    m = create_model(dict(hp)).eval().to(dev)
    for p_ in m.parameters():
        p_.requires_grad_(False)
    return m


def load(name, z, pos, batch):
    path = os.path.join(GOLDEN, "full", name + ".pt")
    if not os.path.exists(path):
        pytest.skip(f"{name}: full-size golden not generated")
    fx = torch.load(path, weights_only=False)
    h = hashlib.sha256()
    for t in (z, pos, batch):
        h.update(t.detach().cpu().contiguous().numpy().tobytes())
    assert h.hexdigest() == fx["input_digest"], "the seeded generator no longer reproduces the golden's input"
    return fx


def errs(y, f, fx):
    eE = ((y.double().cpu() - fx["ref_E64"]).abs().max() / max(1.0, fx["ref_E64"].abs().max().item())).item()
    eF = ((f.double().cpu() - fx["ref_F64"]).abs().max() / fx["ref_F64"].abs().max()).item()
    return eE, eF


@pytest.mark.parametrize("path", ["tables", "exact", "graph"])
def test_c4_full_frame(path, monkeypatch):
    """the bench workload itself: tables (default), exact tcgen05 chain, and the CUDA-graph replay that bench.py times"""
    from torchmdnet_b200.extensions import fused
    monkeypatch.setattr(fused, "USE_RADIAL_TABLE", path != "exact")
    hp = systems.tensornet_hparams(num_layers=3, F=128, R=32, cutoff=5.0, max_nb=64, static_shapes=(path == "graph"))
    z, pos, batch, box = systems.water_box(n_side=15, L=46.42, seed=0)
    fx = load("c4_water10125_tensornet_L3", z, pos, batch)
    m = seeded_model(hp, fx["weight_seed"])
    assert m.representation_model._fused_path_ok()
    if path == "graph":
        from torchmdnet_b200.calculators import CUDAGraphForceField
        ff = CUDAGraphForceField(m, z.to(dev), batch.to(dev), box=box.to(dev), num_systems=1).capture(pos.to(dev))
        y, f = ff(pos.to(dev))
    else:
        y, f = m(z.to(dev), pos.to(dev), batch.to(dev), box=box.to(dev), num_systems=1)
    eE, eF = errs(y, f, fx)
    print(f"C4 full [{path}]: relE={eE:.2e} relF={eF:.2e}")
    assert eE < TOL and eF < TOL


def test_c3_et_spice_256_molecules():
    hp = systems.et_hparams()
    z, pos, batch, _ = systems.molecule_batch(n_mol=256, seed=0)
    fx = load("c3_et_spice_256mol", z, pos, batch)
    m = seeded_model(hp, fx["weight_seed"])
    y, f = m(z.to(dev), pos.to(dev), batch.to(dev), num_systems=256)
    eE, eF = errs(y, f, fx)
    print(f"C3 full: relE={eE:.2e} relF={eF:.2e}")
    assert eE < TOL and eF < TOL


def test_c5_eight_boxes_of_5000():
    hp = systems.tensornet_hparams(num_layers=2, F=128, R=32, cutoff=5.0, max_nb=128)
    z, pos, batch, box = systems.random_box(5000, seed=0, n_boxes=8)
    fx = load("c5_8x5000_tensornet_L2", z, pos, batch)
    m = seeded_model(hp, fx["weight_seed"])
    y, f = m(z.to(dev), pos.to(dev), batch.to(dev), box=box.to(dev), num_systems=8)
    eE, eF = errs(y, f, fx)
    print(f"C5 per GPU: relE={eE:.2e} relF={eF:.2e}")
    assert eE < TOL and eF < TOL


def test_tn2_256_charged_molecules():
    """SURVEY 8(f) item 3 at bench size: TensorNet2 + ScalarPlusWeightedCoulomb (all-pairs Coulomb, total charges -1/0/+1) on the
    `bench.py --workload tn2` input against the fp64 oracle (pinned to the unmodified reference by oracle/make_golden_tn2.py)"""
    import bench
    hp, z, pos, batch, box, _, q = bench.workload("tn2", seed=0)
    fx = load("tn2_256mol_charged", z, pos, batch)
    m = seeded_model(hp, fx["weight_seed"])
    y, f = m(z.to(dev), pos.to(dev), batch.to(dev), q=q.to(dev), num_systems=256)
    eE, eF = errs(y, f, fx)
    print(f"TN2 256 molecules: relE={eE:.2e} relF={eF:.2e}")
    assert eE < TOL and eF < TOL
