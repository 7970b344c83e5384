"""Bisect the double backward of the training path: per-parameter error of d(sum E^2 + sum |F|^2)/d theta vs the fp64 oracle, with
each kernel-backed op swapped for its torch restatement in turn (development aid, GPU)."""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p_)
import systems
import tensornet_oracle as orc
from torchmdnet_b200.extensions import tensornet_ops as T
from torchmdnet_b200.models.model import create_model

dev = torch.device("cuda:0")


def run(group="O(3)", patches=(), verbose=False):
    hp = systems.tensornet_hparams(num_layers=2, F=32, R=16, max_nb=64, group=group)
    random.seed(4); np.random.seed(4); torch.manual_seed(4)
    m = create_model(dict(hp))
    sd64 = {k: v.detach().double().requires_grad_(v.is_floating_point() and v.dim() > 0) for k, v in m.state_dict().items()}
    z, pos, batch, _ = systems.molecule_batch(n_mol=3, lo=8, hi=12, spread=3.5, seed=2)
    q = torch.tensor([0.0, 1.0, -1.0])
    y, f = orc.energy_and_forces(sd64, hp, z, pos.double(), batch, None, q.double(), create_graph=True)
    res = {}
    for term in ("E", "F", "EF"):
        loss_ref = ((y ** 2).sum() if "E" in term else 0) + ((f ** 2).sum() if "F" in term else 0)
        names = [k for k, v in sd64.items() if v.requires_grad]
        g_ref = dict(zip(names, torch.autograd.grad(loss_ref, [sd64[k] for k in names], allow_unused=True, retain_graph=True)))
        saved = {}
        for name, fn in patches:
            saved[name] = getattr(T, name)
            setattr(T, name, fn)
        try:
            mm = m.to(dev).train()
            mm.zero_grad()
            yy, ff = mm(z.to(dev), pos.to(dev), batch.to(dev), q=q.to(dev))
            loss = ((yy ** 2).sum() if "E" in term else 0) + ((ff ** 2).sum() if "F" in term else 0)
            loss.backward()
        finally:
            for name, fn in saved.items():
                setattr(T, name, fn)
        params = dict(mm.named_parameters())
        worst = ("", 0.0)
        for k, gr in g_ref.items():
            if k not in params or gr is None:
                continue
            gp = params[k].grad
            scale = gr.abs().max().item()
            if scale < 1e-12:
                continue
            err = float("nan") if gp is None else (gp.double().cpu() - gr).abs().max().item() / scale
            if verbose:
                print(f"    {term} {k:70s} {err:.2e}")
            if not (err <= worst[1]):
                worst = (k, err)
        res[term] = worst
    return res


def mp_torch(I, A, S, w, row_data, row_indices, row_indptr, col_data, col_indices, col_indptr):
    E = w.shape[0]
    row = T._row_of_slot(row_indptr, E)
    valid = (torch.arange(E, device=w.device) < row_indptr[-1]).to(w.dtype).view(-1, 1, 1)
    j = row_indices[:E].clamp(min=0).long()
    ww = w * valid
    o = lambda X, k: torch.zeros_like(X).index_add(0, row, ww[:, k:k + 1] * X[j])
    return [o(I, 0), o(A, 1), o(S, 2)]


PATCHES = {
    "mp": [("fn_message_passing", mp_torch)],
    "radial_mp": [("fn_radial_message_passing", lambda r, w, rd, rp: list(T._t_radial_mp(r, w, rd, rp)))],
    "p1": [("interaction_normalize", T._t_p1)],
    "p2": [("interaction_product", T._t_p2)],
    "p3": [("interaction_update", T._t_p3)],
    "norm2": [("norm2_compact", T._t_norm2)],
    "norm3": [("norm3_compact", lambda i, a, s: torch.cat(T._t_norm3(i, a, s), -1))],
    "gate": [("gate_compact", lambda i, a, s, g: (i * g[:, 0:1], a * g[:, 1:2], s * g[:, 2:3]))],
}

if __name__ == "__main__":
    for group in ("O(3)", "SO(3)"):
        print(group, "no patch", run(group, verbose=(group == "O(3)")))
        allp = [x for v in PATCHES.values() for x in v]
        print(group, "all torch", run(group, allp))
        for name, pt in PATCHES.items():
            print(group, "torch", name, run(group, pt))
        for name in PATCHES:
            rest = [x for k, v in PATCHES.items() if k != name for x in v]
            print(group, "kernel only", name, run(group, rest))
