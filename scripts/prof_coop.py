"""per-warp cycle breakdown of mp_coop_fwd_kernel (needs a library built with -DTMD_COOP_PROF)"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import systems
from torchmdnet_b200 import _lib
from torchmdnet_b200.extensions import fused
from torchmdnet_b200.extensions.neighbors import neighbor_graph
dev = torch.device("cuda:0")
z, pos, batch, box = systems.water_box(n_side=15, L=46.42, seed=0)
n, F = z.numel(), 128
g = neighbor_graph(pos.to(dev), batch.to(dev), box.to(dev), True, 0.0, 5.0, True, n * 64)
pidx = fused.build_pair_index(g)
P = int(pidx.num_unique)
Y = [torch.randn(n, k, F, device=dev) for k in (1, 3, 5)]
f = torch.randn(P, 3, F, device=dev)
plan = fused.MessagePassingPlan(g, pidx, F)
for _ in range(3): plan.forward(Y, f)
ng = plan.gcnt.numel()
prof = torch.zeros(ng, 16, 8, dtype=torch.int64, device=dev)
lib = _lib.library()
lib.tmd_mp_coop_set_profile.argtypes = [ctypes.c_void_p]
lib.tmd_mp_coop_set_profile(prof.data_ptr())
plan.forward(Y, f); torch.cuda.synchronize()
p = prof.cpu().double()
w = p[:, :15]
valid = w[..., 5] > 0
tot = w[..., 0][valid]
print("groups", ng, "rows", int(valid.sum()))
for k, name in ((0, "total"), (1, "issued spin"), (2, "full wait"), (3, "f wait"), (4, "lane0 tail")):
    x = w[..., k][valid]
    print(f"{name:12s} mean {x.mean():9.0f} cyc  ({100 * x.sum() / tot.sum():5.1f} %)  per edge {x.sum() / w[..., 5][valid].sum():7.1f}")
gs = torch.where(valid, w[..., 6], torch.full_like(w[..., 6], 1e30)).min(1).values
ge = w[..., 7].max(1).values
dur = (ge - gs)
print("group duration cycles: mean %.0f min %.0f max %.0f" % (dur.mean(), dur.min(), dur.max()))
slow = w[..., 0].max(1).values
print("slowest-row total per group mean %.0f ; mean edges/row %.1f ; compute+other per edge %.1f" % (slow.mean(), w[..., 5][valid].mean(),
      ((tot.sum() - w[..., 1][valid].sum() - w[..., 2][valid].sum() - w[..., 3][valid].sum() - w[..., 4][valid].sum()) / w[..., 5][valid].sum())))
