"""Generates tests/golden/*.pt from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/make_golden.py
It (1) imports the reference through oracle/ref_import.py, (2) evaluates the reference's
own create_model/load_model/TorchMD_Net on seeded inputs in fp32 and fp64, (3) asserts
that oracle/tensornet_oracle.py reproduces them (this is what pins the oracle), and
(4) writes small fixtures that travel to the GPU box, where /root/reference is absent.
"""
import hashlib
import os
import pickle
import random
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.simplefilter("ignore")

from ref_import import import_reference, REFERENCE_ROOT  # noqa: E402
import tensornet_oracle as orc  # noqa: E402
import systems  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
import_reference()
from torchmdnet.models.model import create_model, load_model  # noqa: E402  (the reference)


def seed_everything(seed):  # what pl.seed_everything does
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode()); h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def run_ref(model, z, pos, batch, box=None, q=None, dtype=torch.float32):
    model = model.to(dtype)
    y, f = model(z, pos.to(dtype).clone(), batch, box=None if box is None else box.to(dtype),
                 q=None if q is None else q.to(dtype))
    return y.detach(), f.detach()


def check(name, a, b, tol):
    err = (a - b).abs().max().item()
    scale = b.abs().max().item() + 1e-30
    print(f"  {name}: max|d|={err:.3e} rel={err / scale:.3e}")
    assert err <= tol * max(1.0, scale), (name, err, scale)


def parse_pdb(path):
    sym = {"H": 1, "C": 6, "N": 7, "O": 8}
    z, p = [], []
    for line in open(path):
        if line.startswith(("HETATM", "ATOM")):
            p.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
            z.append(sym[line[76:78].strip()])
    return torch.tensor(z), torch.tensor(p, dtype=torch.float64)


def oracle_pair(sd, hp, z, pos, batch, box, q):
    out = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        sdd = orc.cast_state_dict(sd, dt)
        y, f = orc.energy_and_forces(sdd, hp, z, pos.to(dt), batch, None if box is None else box.to(dt),
                                     None if q is None else q.to(dt))
        out[tag] = (y, f)
    return out


def case_from_model(name, model, hp, z, pos, batch, box=None, q=None, store_sd=False, extra=None):
    print(f"[{name}] N={z.numel()}")
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    yr32, fr32 = run_ref(model, z, pos, batch, box, q, torch.float32)
    yr64, fr64 = run_ref(model.double(), z, pos.double(), batch, box, q, torch.float64)
    model.float()
    o = oracle_pair(sd, hp, z, pos, batch, box, q)
    check("oracle64 vs ref64 E", o["f64"][0], yr64, 1e-11)
    check("oracle64 vs ref64 F", o["f64"][1], fr64, 1e-10)
    check("oracle32 vs ref32 E", o["f32"][0], yr32, 2e-5)
    check("oracle32 vs ref32 F", o["f32"][1], fr32, 2e-5)
    check("ref32 vs ref64 E", yr32.double(), yr64, 1e-4)
    check("ref32 vs ref64 F", fr32.double(), fr64, 1e-4)
    fx = dict(name=name, hparams=hp, z=z, pos=pos, batch=batch, box=box, q=q, ref_E32=yr32, ref_F32=fr32,
              ref_E64=yr64, ref_F64=fr64, sd_digest=sd_digest(sd))
    if store_sd:
        fx["state_dict"] = sd
    if extra:
        fx.update(extra)
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    return fx


def main():
    # ---- C1a: caffeine known-answer test (tests/test_calculator.py:102-141) -------------
    ck = torch.load(os.path.join(REFERENCE_ROOT, "tests", "example_tensornet.ckpt"), map_location="cpu", weights_only=False)
    slim = {"state_dict": ck["state_dict"], "hyper_parameters": ck["hyper_parameters"]}
    torch.save(slim, os.path.join(OUT, "example_tensornet_slim.ckpt"))
    z, pos = parse_pdb(os.path.join(REFERENCE_ROOT, "tests", "caffeine.pdb"))
    model = load_model(os.path.join(OUT, "example_tensornet_slim.ckpt"), derivative=True, max_num_neighbors=64,
                       static_shapes=False)
    hp = dict(torch.load(os.path.join(OUT, "example_tensornet_slim.ckpt"), weights_only=False)["hyper_parameters"])
    hp.update(derivative=True, max_num_neighbors=64, static_shapes=False)
    batch = torch.zeros_like(z)
    kat_E = -113.6652
    kat_F0 = torch.tensor([-7.76999056e-01, -6.89736724e-01, -9.31625906e-03])
    kat_F23 = torch.tensor([7.02065229e-02, 7.10679889e-02, 2.37307906e-01])
    fx = case_from_model("caffeine_tensornet", model, hp, z, pos.float(), batch, None, torch.tensor([0.0]),
                         store_sd=True, extra=dict(kat_E=kat_E, kat_F0=kat_F0, kat_F23=kat_F23))
    assert abs(fx["ref_E32"].item() - kat_E) < 1e-4
    assert (fx["ref_F32"][0] - kat_F0).abs().max() < 1e-4 and (fx["ref_F32"][23] - kat_F23).abs().max() < 1e-4
    # charged variant (q enters the O(3) branch, tensornet.py:789,812)
    case_from_model("caffeine_tensornet_q2", model, hp, z, pos.float(), batch, None, torch.tensor([2.0]))

    # ---- C1b: tests/expected.pkl (tests/test_model.py:277-329) ---------------------------
    import yaml
    with open(os.path.join(REFERENCE_ROOT, "tests", "expected.pkl"), "rb") as f:
        expected = pickle.load(f)
    for model_name, cfg in (("tensornet", "TensorNet-QM9.yaml"), ("equivariant-transformer", "ET-QM9.yaml")):
        args = yaml.load(open(os.path.join(REFERENCE_ROOT, "examples", cfg)), Loader=yaml.FullLoader)
        args.setdefault("precision", 32)
        args.update(model=model_name, seed=1234, prior_model=None, output_model="Scalar", derivative=True)
        args.setdefault("box_vecs", None); args.setdefault("remove_ref_energy", False)
        seed_everything(1234)
        model = create_model(args)
        zs = torch.tensor([1, 6, 7, 8, 9], dtype=torch.long)
        z = zs[torch.randint(0, len(zs), (5,))]
        pos = torch.randn(len(z), 3)
        batch = torch.zeros(len(z), dtype=torch.long); batch[len(batch) // 2:] = 1
        exp = expected[model_name]["Scalar"]
        tag = "expected_" + model_name.replace("-", "_")
        keep = {k: v for k, v in args.items() if isinstance(v, (int, float, str, bool, type(None), list))}
        fx = case_from_model(tag, model, keep, z, pos, batch, store_sd=False,  # weights are reproduced from seed 1234

                             extra=dict(pkl_pred=exp["pred"].detach(), pkl_deriv=exp["deriv"].detach(), seed=1234))
        check("ref vs expected.pkl pred", fx["ref_E32"], exp["pred"].detach(), 1e-5)
        check("ref vs expected.pkl deriv", fx["ref_F32"], exp["deriv"].detach(), 1e-5)

    # ---- C2-like periodic boxes, seeded weights (weights reproduced from the seed) --------
    def seeded(hp, seed):
        seed_everything(seed)
        return create_model(dict(hp))

    hp = systems.tensornet_hparams(num_layers=2, max_nb=128)
    z, pos, batch, box = systems.random_box(300, seed=0)
    case_from_model("box300_tensornet_L2", seeded(hp, 1), hp, z, pos, batch, box, extra=dict(weight_seed=1))
    z, pos, batch, box = systems.random_box(1000, seed=0)
    case_from_model("box1000_tensornet_L2", seeded(hp, 1), hp, z, pos, batch, box, extra=dict(weight_seed=1))
    # SO(3) group + charges + two boxes with different per-sample triclinic cells
    hp2 = systems.tensornet_hparams(num_layers=1, F=64, R=16, max_nb=96, group="SO(3)")
    z, pos, batch, box = systems.random_box(150, seed=3, n_boxes=2)
    box = box.clone(); box[1] = torch.tensor([[13.0, 0, 0], [1.5, 12.0, 0], [-2.0, 1.0, 14.0]])
    q = torch.tensor([1.0, -1.0])
    case_from_model("box2x150_tensornet_so3_tri", seeded(hp2, 2), hp2, z, pos, batch, box, q, extra=dict(weight_seed=2))
    hp3 = systems.tensornet_hparams(num_layers=1, F=64, R=16, max_nb=96, group="O(3)")
    case_from_model("box2x150_tensornet_o3_tri", seeded(hp3, 2), hp3, z, pos, batch, box, q, extra=dict(weight_seed=2))
    # water box slice (C4 generator at n_side=5 -> 375 atoms, L=15.47)
    hp4 = systems.tensornet_hparams(num_layers=3, max_nb=64)
    z, pos, batch, box = systems.water_box(n_side=5, L=46.42 / 3, seed=0)
    case_from_model("water375_tensornet_L3", seeded(hp4, 1), hp4, z, pos, batch, box, extra=dict(weight_seed=1))

    # ---- C3-like ET-SPICE molecule batch ---------------------------------------------------
    hpe = systems.et_hparams()
    z, pos, batch, _ = systems.molecule_batch(n_mol=6, seed=0)
    case_from_model("mol6_et_spice", seeded(hpe, 1), hpe, z, pos, batch, None, extra=dict(weight_seed=1))
    hpe2 = systems.et_hparams(num_layers=2, vector_cutoff=False, cutoff=5.0, R=32)
    z, pos, batch, box = systems.random_box(200, seed=5)
    case_from_model("box200_et", seeded(hpe2, 4), hpe2, z, pos, batch, box, extra=dict(weight_seed=4))

    # ---- neighbour-list vectors from the reference's own numpy checker ----------------------
    # (tests/test_neighbors.py:19-71 restated with numpy: tril_indices + sequential rounding + '< cutoff')
    nl = []
    rng = np.random.RandomState(7)
    boxes = {"none": None, "rect": np.diag([10.0, 11.0, 12.0]), "tri": np.array([[10.0, 0, 0], [2.0, 9.0, 0], [1.0, -2.0, 11.0]])}
    for bname, bx in boxes.items():
        for nb in (1, 3):
            for cutoff in (1.0, 4.0):
                for loop in (False, True):
                    for tr in (False, True):
                        n_per = rng.randint(3, 40, size=nb)
                        batch_np = np.repeat(np.arange(nb), n_per)
                        lbox = 10.0
                        p = (rng.rand(batch_np.size, 3) * lbox - 0.5 * lbox).astype(np.float32)
                        cum = np.concatenate([[0], np.cumsum(n_per)])
                        pairs = np.concatenate([np.array(np.tril_indices(int(n_per[i]), -1)) + cum[i] for i in range(nb)], axis=1)
                        if tr:
                            pairs = np.concatenate([pairs, np.flip(pairs, axis=0)], axis=1)
                        if loop:
                            il = np.arange(cum[-1]); pairs = np.concatenate([pairs, np.stack([il, il])], axis=1)
                        d = p[pairs[0]] - p[pairs[1]]
                        if bx is not None:
                            b32 = bx.astype(np.float32)
                            d -= np.outer(np.round(d[:, 2] / b32[2, 2]), b32[2])
                            d -= np.outer(np.round(d[:, 1] / b32[1, 1]), b32[1])
                            d -= np.outer(np.round(d[:, 0] / b32[0, 0]), b32[0])
                        r = np.linalg.norm(d, axis=-1)
                        m = r < cutoff
                        pairs, d, r = pairs[:, m], d[m], r[m]
                        o = np.lexsort(pairs)  # sorts by pairs[1] then pairs[0] (last key primary)
                        o = np.lexsort((pairs[1], pairs[0]))
                        nl.append(dict(box=bname, box_vectors=None if bx is None else torch.tensor(bx, dtype=torch.float32),
                                       pos=torch.tensor(p), batch=torch.tensor(batch_np), cutoff=cutoff, loop=loop,
                                       include_transpose=tr, neighbors=torch.tensor(pairs[:, o].copy()),
                                       deltas=torch.tensor(d[o].copy()), distances=torch.tensor(r[o].copy())))
    # check the oracle neighbour list against the numpy reference
    for c in nl:
        ei, vec, dist = orc.neighbor_list(c["pos"], c["batch"], c["box_vectors"], 0.0, c["cutoff"], c["loop"], c["include_transpose"])
        assert torch.equal(ei, c["neighbors"]), "oracle neighbour indices differ from the numpy reference"
        assert torch.allclose(vec, c["deltas"], atol=1e-5) and torch.allclose(dist, c["distances"], atol=1e-5)
    print(f"[neighbors] {len(nl)} cases: oracle indices bit-equal to the numpy reference")
    torch.save(nl, os.path.join(OUT, "neighbor_cases.pt"))
    print("done; fixtures in", OUT)


if __name__ == "__main__":
    main()
