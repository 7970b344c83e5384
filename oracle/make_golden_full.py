"""Full-size goldens for the bench workloads (build container only): C4 (10 125-atom water box, the timed frame), C3 (ET-SPICE,
256 molecules) and C5 per GPU (8 x 5000-atom boxes, TensorNet-ANI2X config).

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/make_golden_full.py [c4] [c3] [c5]

The fp64 evaluation of a 10k-atom box does not fit in this container's 62 GB with the plain restatement (per-edge [E,3,3,F]
intermediates), so `energy_and_forces(..., edge_chunks=K)` evaluates the same arithmetic chunk by chunk under activation
checkpointing.  This script first PINS that mode: on the water375 fixture it must reproduce the unmodified reference's fp64
energies/forces (tests/golden/water375_tensornet_L3.pt: ref_E64 / ref_F64, written by make_golden.py from the reference itself)
to 1e-11 / 1e-10, and on a 1029-atom slice of the C4 generator the chunked and plain oracles must agree to 1e-12.  Only then the
full-size fixtures are written: inputs are reproduced from the seed (oracle/systems.py), weights from the seed
(create_model under seed 1), so a fixture holds only E [M,1] and F [N,3] in fp64 (+ digests).
"""
import hashlib
import os
import random
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import systems  # noqa: E402
import tensornet_oracle as orc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
OUT = os.path.join(GOLDEN, "full")   # not globbed by the small-fixture tests (different schema: no inputs stored)


def seeded_sd(hp, seed=1):
    from torchmdnet_b200.models.model import create_model   # same RNG consumption as the reference (tests/test_oracle.py digest)
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    return {k: v.detach().clone() for k, v in create_model(dict(hp)).state_dict().items()}


def digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def pin_chunked_mode():
    fx = torch.load(os.path.join(GOLDEN, "water375_tensornet_L3.pt"), weights_only=False)
    sd = orc.cast_state_dict(seeded_sd(fx["hparams"], fx["weight_seed"]), torch.float64)
    y, f = orc.energy_and_forces(sd, fx["hparams"], fx["z"], fx["pos"].double(), fx["batch"], fx["box"].double(), edge_chunks=7)
    eE, eF = (y - fx["ref_E64"]).abs().max().item(), (f - fx["ref_F64"]).abs().max().item()
    print(f"[pin] chunked oracle vs the reference's fp64 (water375): |dE|={eE:.2e} |dF|={eF:.2e}")
    assert eE < 1e-11 * max(1.0, fx["ref_E64"].abs().max().item()) and eF < 1e-10
    hp = systems.tensornet_hparams(num_layers=3, max_nb=64)
    z, pos, batch, box = systems.water_box(n_side=7, L=46.42 * 7 / 15, seed=0)
    sd = orc.cast_state_dict(seeded_sd(hp, 1), torch.float64)
    y0, f0 = orc.energy_and_forces(sd, hp, z, pos.double(), batch, box.double())
    y1, f1 = orc.energy_and_forces(sd, hp, z, pos.double(), batch, box.double(), edge_chunks=16)
    print(f"[pin] chunked vs plain oracle (1029 atoms): |dE|={(y0 - y1).abs().max():.2e} |dF|={(f0 - f1).abs().max():.2e}")
    assert (y0 - y1).abs().max() < 1e-12 * max(1.0, y0.abs().max().item()) and (f0 - f1).abs().max() < 1e-12


def write(name, hp, z, pos, batch, box, gen, chunks, seed=1):
    t0 = time.time()
    sd = orc.cast_state_dict(seeded_sd(hp, seed), torch.float64)
    y, f = orc.energy_and_forces(sd, hp, z, pos.double(), batch, None if box is None else box.double(), edge_chunks=chunks)
    fx = dict(name=name, hparams=hp, generator=gen, weight_seed=seed, n_atoms=int(z.numel()), ref_E64=y, ref_F64=f,
              input_digest=digest(z, pos, batch), note="fp64 oracle (edge_chunks mode, pinned by make_golden_full.pin_chunked_mode)")
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(f"[{name}] N={z.numel()} E={y.flatten()[:3].tolist()} max|F|={f.abs().max():.4f}  ({time.time() - t0:.0f} s)")


def write_tn2():
    """bench.py --workload tn2: TensorNet2 + ScalarPlusWeightedCoulomb on 256 charged molecules (the oracle functions of this row are
    pinned to the unmodified reference by make_golden_tn2.py)"""
    import bench
    hp, z, pos, batch, box, desc, q = bench.workload("tn2", seed=0)
    t0 = time.time()
    sd = orc.cast_state_dict(seeded_sd(hp, 1), torch.float64)
    y, f = orc.energy_and_forces(sd, hp, z, pos.double(), batch, None, q.double())
    fx = dict(name="tn2_256mol_charged", hparams=hp, generator="bench.workload('tn2', seed=0)", weight_seed=1, n_atoms=int(z.numel()),
              ref_E64=y, ref_F64=f, input_digest=digest(z, pos, batch), note="fp64 oracle (plain mode)")
    torch.save(fx, os.path.join(OUT, "tn2_256mol_charged.pt"))
    print(f"[tn2] N={z.numel()} E={y.flatten()[:3].tolist()} max|F|={f.abs().max():.4f}  ({time.time() - t0:.0f} s)")


def main():
    which = set(sys.argv[1:]) or {"c4", "c3", "c5"}
    if "tn2" in which:
        torch.set_num_threads(os.cpu_count() or 8)
        os.makedirs(OUT, exist_ok=True)
        write_tn2()
        which.discard("tn2")
        if not which:
            return
    torch.set_num_threads(os.cpu_count() or 8)
    os.makedirs(OUT, exist_ok=True)
    pin_chunked_mode()
    if "c4" in which:   # bench.py --workload c4, rank 0
        hp = systems.tensornet_hparams(num_layers=3, F=128, R=32, cutoff=5.0, max_nb=64)
        z, pos, batch, box = systems.water_box(n_side=15, L=46.42, seed=0)
        write("c4_water10125_tensornet_L3", hp, z, pos, batch, box, "systems.water_box(n_side=15, L=46.42, seed=0)", 96)
    if "c3" in which:   # bench.py --workload c3
        hp = systems.et_hparams()
        z, pos, batch, box = systems.molecule_batch(n_mol=256, seed=0)
        write("c3_et_spice_256mol", hp, z, pos, batch, None, "systems.molecule_batch(n_mol=256, seed=0)", 0)
    if "c5" in which:   # bench.py --workload c5 (per GPU)
        hp = systems.tensornet_hparams(num_layers=2, F=128, R=32, cutoff=5.0, max_nb=128)
        z, pos, batch, box = systems.random_box(5000, seed=0, n_boxes=8)
        write("c5_8x5000_tensornet_L2", hp, z, pos, batch, box, "systems.random_box(5000, seed=0, n_boxes=8)", 256)


if __name__ == "__main__":
    main()
