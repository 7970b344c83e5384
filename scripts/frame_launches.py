"""One eager C4 frame inside a cudaProfilerStart/Stop range (the table build and warm-ups stay outside):

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python scripts/frame_launches.py [--workload c4] [--exact]

`scripts/launch_table.py gpurun_out/launches.csv` turns the csv into the per-kernel share table committed under profiles/."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--exact", action="store_true", help="radial tables off")
ap.add_argument("--frames", type=int, default=1)
args = ap.parse_args()

import __graft_entry__ as ge  # noqa: E402
ge.build()
from torchmdnet_b200.extensions import fused  # noqa: E402

if args.exact:
    fused.USE_RADIAL_TABLE = False
dev = torch.device("cuda:0")
hp, z, pos, batch, box, desc, q_sys = bench.workload(args.workload, seed=0)
hp["static_shapes"] = hp["model"] == "tensornet"
model = bench.seeded_model(hp).to(dev)
for p_ in model.parameters():
    p_.requires_grad_(False)
zd, pd, bd = z.to(dev), pos.to(dev), batch.to(dev)
xd = None if box is None else box.to(dev)
n_sys = int(batch.max()) + 1
qd = None if q_sys is None else q_sys.to(dev)
for _ in range(3):
    model(zd, pd.clone(), bd, box=xd, q=qd, num_systems=n_sys)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.frames):
    y, f = model(zd, pd.clone(), bd, box=xd, q=qd, num_systems=n_sys)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(desc, float(y.sum()))
