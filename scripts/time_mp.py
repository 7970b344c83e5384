"""times the message-passing entry points of one C4 layer (the frame's pid form) -- tuning helper"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import systems
from torchmdnet_b200.extensions import fused
from torchmdnet_b200.extensions.neighbors import neighbor_graph
dev = torch.device("cuda:0")
z, pos, batch, box = systems.water_box(n_side=15, L=46.42, seed=0)
n, F = z.numel(), 128
g = neighbor_graph(pos.to(dev), batch.to(dev), box.to(dev), True, 0.0, 5.0, True, n * 64)
pidx = fused.build_pair_index(g)
P = int(pidx.num_unique)
Y = [torch.randn(n, k, F, device=dev) for k in (1, 3, 5)]
dM = [torch.randn(n, k, F, device=dev) for k in (1, 3, 5)]
f, fp = torch.randn(P, 3, F, device=dev), torch.randn(P, 3, F, device=dev)
plan = fused.MessagePassingPlan(g, pidx, F)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
print(f"coop={plan.coop} fwd {t(lambda: plan.forward(Y, f)):.1f} us  bwd {t(lambda: plan.backward(Y, dM, f, fp)):.1f} us  plan {t(lambda: fused.MessagePassingPlan(g, pidx, F)):.1f} us")
