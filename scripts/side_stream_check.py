"""development aid: eager frame, then CUDAGraphForceField warm-up + capture + replay of the same C4 model (faulthandler on)"""
import faulthandler, sys, os
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, bench
from torchmdnet_b200.calculators import CUDAGraphForceField
dev = torch.device("cuda:0")
hp, z, pos, batch, box, desc, q_sys = bench.workload("c4", seed=0)
hp["static_shapes"] = True
model = bench.seeded_model(hp).to(dev)
for p_ in model.parameters(): p_.requires_grad_(False)
zd, pd, bd, xd = z.to(dev), pos.to(dev), batch.to(dev), box.to(dev)
y, f = model(zd, pd.clone(), bd, box=xd, num_systems=1); torch.cuda.synchronize(); print("eager ok", float(y.detach()), flush=True)
ff = CUDAGraphForceField(model, zd, bd, box=xd, num_systems=1)
print("ff built", flush=True)
ff.capture(pd)
print("captured", flush=True)
y, f = ff(pd); torch.cuda.synchronize(); print("graph ok", float(y.detach()))
