"""Generates tests/golden/next/tn2_*.pt from the UNMODIFIED reference (build container only) -- SURVEY 8(f) item 3, oracle first.

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/make_golden_tn2.py
TensorNet2 (models/tensornet2.py) + ScalarPlusWeightedCoulomb (models/output_modules.py:344-609) are evaluated through the
reference's own create_model / TorchMD_Net in fp32 and fp64 on seeded inputs; oracle/tensornet_oracle.py must reproduce them
(1e-11 relative in fp64), which pins `tensornet2_representation`, `charge_predict` and `weighted_coulomb_per_atom` before any GPU
code for this row exists.  The GPU path of this row is not built yet: these fixtures are consumed by tests/test_oracle.py only.
"""
import os
import random
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.simplefilter("ignore")

from ref_import import import_reference  # noqa: E402
import tensornet_oracle as orc  # noqa: E402
import systems  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "next")   # not globbed by the product parity tests
os.makedirs(OUT, exist_ok=True)
import_reference()
from torchmdnet.models.model import create_model  # noqa: E402  (the reference)


def tn2_hparams(F=64, L=2, R=32, q_dim=16, cutoff=5.0, coulomb_cutoff=None, q_weights=(1.0, 0.5, 2.0), group="O(3)"):
    hp = systems.tensornet_hparams(num_layers=L, F=F, R=R, cutoff=cutoff, max_nb=128, group=group)
    hp.update(model="tensornet2", output_model="ScalarPlusWeightedCoulomb", q_dim=q_dim, q_weights=list(q_weights),
              coulomb_cutoff=coulomb_cutoff, coulomb_max_num_neighbors=256, reduce_op="sum", static_shapes=False)
    return hp


def run_case(name, hp, z, pos, batch, box, q, seed):
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    model = create_model(dict(hp)).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    out = {"hparams": hp, "weight_seed": seed, "z": z, "pos": pos.float(), "batch": batch, "box": box, "q": q, "state_dict": sd}
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        m = create_model(dict(hp, precision=32 if dt == torch.float32 else 64)).eval()
        m.load_state_dict({k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd.items()})
        y, f = m(z, pos.float().to(dt).clone(), batch, box=None if box is None else box.to(dt), q=None if q is None else q.to(dt))
        yo, fo = orc.energy_and_forces(orc.cast_state_dict(sd, dt), hp, z, pos.float().to(dt), batch,
                                       None if box is None else box.to(dt), None if q is None else q.to(dt))
        eE = ((y.detach() - yo).abs().max() / max(1.0, y.detach().abs().max().item())).item()
        eF = ((f.detach() - fo).abs().max() / f.detach().abs().max()).item()
        print(f"[{name}] fp{tag}: E {y.detach().flatten().tolist()[:2]} oracle-vs-reference relE {eE:.2e} relF {eF:.2e}")
        tol = 1e-11 if dt == torch.float64 else 2e-5
        assert eE < tol and eF < tol, (name, tag, eE, eF)
        out["ref_E" + tag], out["ref_F" + tag] = y.detach(), f.detach()
    torch.save(out, os.path.join(OUT, name + ".pt"))


if __name__ == "__main__":
    # (1) molecules in vacuum, total charges 0/+1/-1/0, all-pairs Coulomb (coulomb_cutoff None)
    z, pos, batch, _ = systems.molecule_batch(n_mol=4, lo=8, hi=14, seed=11)
    run_case("tn2_mol4_alltoall", tn2_hparams(), z, pos, batch, None, torch.tensor([0.0, 1.0, -1.0, 0.0]), seed=21)
    # (2) periodic box, reaction-field Coulomb inside 5 A (half list, minimum image), SO(3) variant
    z, pos, batch, box = systems.random_box(120, density=0.1, seed=12)
    run_case("tn2_box120_rf", tn2_hparams(coulomb_cutoff=5.0, group="SO(3)"), z, pos, batch, box, None, seed=22)
