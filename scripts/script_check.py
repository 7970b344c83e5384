"""development aid: torch.jit.script(model) (in-process wrapper around the whole-frame op) against the eager frame on the GPU"""
import os, random, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import systems
from torchmdnet_b200.models.model import create_model
dev = torch.device("cuda:0")
hp = systems.tensornet_hparams(num_layers=2, F=64, R=32, max_nb=96)
random.seed(3); np.random.seed(3); torch.manual_seed(3)
m = create_model(dict(hp)).eval().to(dev)
for p_ in m.parameters():
    p_.requires_grad_(False)
z, pos, batch, box = systems.random_box(300, seed=5)
zd, pd, bd, xd = z.to(dev), pos.to(dev), batch.to(dev), box.to(dev)
y0, f0 = m(zd, pd.clone(), bd, box=xd)
sm = torch.jit.script(m)
y1, f1 = sm(zd, pd.clone(), bd, box=xd)
y2, f2 = sm(zd, pd.clone(), bd, xd, None, None, None, 1)
print("scripted vs eager: dE", float((y1 - y0).abs().max()), "dF", float((f1 - f0).abs().max()), "| with num_systems: dE",
      float((y2 - y0).abs().max()), "dF", float((f2 - f0).abs().max()), "| max|F|", float(f0.abs().max()))
