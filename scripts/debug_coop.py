"""Debug driver for the cooperative message-passing kernels: plan / forward / backward with a synchronize after each step
(run under compute-sanitizer on the GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import systems
from torchmdnet_b200.extensions import fused
from torchmdnet_b200.extensions.neighbors import neighbor_graph

dev = torch.device("cuda:0")
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 6
z, pos, batch, box = systems.water_box(n_side=n_side, L=46.42 * n_side / 15, seed=1)
n, F = z.numel(), 128
g = neighbor_graph(pos.to(dev), batch.to(dev), box.to(dev), True, 0.0, 5.0, True, n * 64)
pidx = fused.build_pair_index(g)
P = int(pidx.num_unique)
torch.manual_seed(0)
Y = [torch.randn(n, k, F, device=dev) for k in (1, 3, 5)]
dM = [torch.randn(n, k, F, device=dev) for k in (1, 3, 5)]
f, fp = torch.randn(P, 3, F, device=dev), torch.randn(P, 3, F, device=dev)
torch.cuda.synchronize(); print("graph ok, n =", n, "E =", int(g.num_pairs), "cells:", int(g.order_flag), flush=True)
plan = fused.MessagePassingPlan(g, pidx, F)
torch.cuda.synchronize(); print("plan ok: groups", plan.gcnt.numel(), "gcnt min/max", int(plan.gcnt.min()), int(plan.gcnt.max()),
                                "ucount max", int(plan.ucount[: int(plan.cursor)].max()), "cursor", int(plan.cursor), flush=True)
M = plan.forward(Y, f)
torch.cuda.synchronize(); print("fwd ok", float(M[0].abs().sum()), flush=True)
dY, gd = plan.backward(Y, dM, f, fp)
torch.cuda.synchronize(); print("bwd ok", float(dY[0].abs().sum()), float(gd.abs().sum()), flush=True)
fused.USE_MP_COOP = False
ref = fused.MessagePassingPlan(g, pidx, F)
M0 = ref.forward(Y, f); dY0, gd0 = ref.backward(Y, dM, f, fp)
E = int(g.num_pairs)
print("bit-exact:", all(torch.equal(a, b) for a, b in zip(M0 + dY0, M + dY)), torch.equal(gd0[:E], gd[:E]))
