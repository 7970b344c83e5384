"""-m "not gpu": the host side of TensorNet2 + ScalarPlusWeightedCoulomb that runs without the CUDA kernels:
construction (state-dict keys and seeded weights identical to the reference's), the charge head (LayerNorm -> MLP -> charge
equilibration) and the all-pairs Coulomb term, each against the pinned oracle.  The assembled GPU forward is checked by
tests/test_gpu_tensornet2.py (not yet validated on a B200: opt-in)."""
import os
import random

import numpy as np
import pytest
import torch

import tensornet_oracle as orc
from conftest import GOLDEN

NEXT = os.path.join(GOLDEN, "next")
CASES = ["tn2_mol4_alltoall", "tn2_box120_rf"]


def seeded(hp, seed):
    from torchmdnet_b200.models.model import create_model
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    return create_model(dict(hp))


@pytest.mark.parametrize("name", CASES)
def test_seeded_weights_and_keys_match_reference(name):
    """create_model under the fixture's seed == the reference's create_model: same keys, bit-identical tensors
    (same parameter creation order and RNG consumption as tensornet2.py:226-341, output_modules.py:355-460)."""
    fx = torch.load(os.path.join(NEXT, name + ".pt"), weights_only=False)
    sd, ref = seeded(fx["hparams"], fx["weight_seed"]).state_dict(), fx["state_dict"]
    assert set(sd) == set(ref)
    assert all(torch.equal(sd[k], ref[k]) for k in ref)


def test_charge_head_matches_oracle():
    fx = torch.load(os.path.join(NEXT, "tn2_mol4_alltoall.pt"), weights_only=False)
    hp = fx["hparams"]
    m = seeded(hp, 0).double()
    m.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in fx["state_dict"].items()})
    sd = orc.cast_state_dict(fx["state_dict"], torch.float64)
    g = torch.Generator().manual_seed(5)
    n, F_ = fx["z"].numel(), hp["embedding_dimension"]
    X = torch.randn(n, 3, 3, F_, generator=g, dtype=torch.float64)
    batch, Q = fx["batch"], fx["q"].double()[fx["batch"]]
    nsys = int(batch.max()) + 1
    I, A, S = orc.decompose(X)
    feat = torch.cat([I, orc.tnorm(A), orc.tnorm(S)], -1)
    for key, mod in [("charge_predict_0.", m.representation_model.charge_predict_0),
                     ("charge_predicts.1.", m.representation_model.charge_predicts[1])]:
        want = orc.charge_predict(sd, "representation_model." + key, X, batch, Q, hp["q_dim"], orc._act(hp["activation"]), nsys)
        got = mod.from_features(feat, batch, Q)
        assert (got - want).abs().max() < 1e-12
        tot = torch.zeros(nsys, hp["q_dim"], dtype=torch.float64).index_add(0, batch, got)
        assert (tot - fx["q"].double()[:, None]).abs().max() < 1e-3      # equilibrated to the molecule's total charge


def test_all_pairs_coulomb_matches_oracle():
    fx = torch.load(os.path.join(NEXT, "tn2_mol4_alltoall.pt"), weights_only=False)
    hp = fx["hparams"]
    head = seeded(hp, 0).output_model.double()
    g = torch.Generator().manual_seed(6)
    n = fx["z"].numel()
    charges = 0.3 * torch.randn(n, (hp["num_layers"] + 1) * hp["q_dim"], generator=g, dtype=torch.float64)
    pos = fx["pos"].double().requires_grad_(True)
    got = head.coulomb_per_atom(charges, fx["z"], pos, fx["batch"], None)
    want = orc.weighted_coulomb_per_atom(hp, charges, pos, fx["batch"], None)
    assert (got - want).abs().max() < 1e-12 * max(1.0, want.abs().max().item())
    (gg,) = torch.autograd.grad(got.sum(), pos, retain_graph=True)
    (gw,) = torch.autograd.grad(want.sum(), pos)
    assert (gg - gw).abs().max() < 1e-11 * max(1.0, gw.abs().max().item())
    with pytest.raises(ValueError):
        head.coulomb_per_atom(charges, fx["z"], pos, fx["batch"], torch.eye(3, dtype=torch.float64) * 30)


def test_load_model_roundtrip_and_legacy_name(tmp_path):
    """reference model.py:208-374: a Lightning-style checkpoint ('model.' key prefix, hyper_parameters) loads back bit-identically;
    the development name of the architecture ('tensornet-nqe', model.py:244-247) maps to tensornet2."""
    from torchmdnet_b200.models.model import load_model
    fx = torch.load(os.path.join(NEXT, "tn2_box120_rf.pt"), weights_only=False)
    hp = dict(fx["hparams"], model="tensornet-nqe")
    path = str(tmp_path / "tn2.ckpt")
    torch.save({"hyper_parameters": hp, "state_dict": {"model." + k: v for k, v in fx["state_dict"].items()}}, path)
    m = load_model(path, derivative=True)
    assert type(m.representation_model).__name__ == "TensorNet2" and type(m.output_model).__name__ == "ScalarPlusWeightedCoulomb"
    sd = m.state_dict()
    assert all(torch.equal(sd[k], v) for k, v in fx["state_dict"].items())
    assert m.output_model.mode == "cutoff" and m.representation_model.output_charges
