"""Dev helper (GPU): where does the tcgen05 radial chain spend its time?  Times the 3-GEMM chain at the C4 pair count with
parts of the kernel disabled through TMD_TC_DBG (1 epilogue off, 2 MMA off, 3 no global stores, 4 A tile re-read from row 0)
and with different n-blocks (TMD_TC_BN is latched at first use, so BN is varied by running the script again)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchmdnet_b200.extensions import fused
from torchmdnet_b200.models.utils import ExpNormalSmearing

dev = torch.device("cuda:0")
torch.manual_seed(0)
R, F = 32, 128
P = int(sys.argv[1]) if len(sys.argv) > 1 else 276446
rbf = ExpNormalSmearing(0.0, 5.0, R, trainable=False).to(dev)
lins = [torch.nn.Linear(R, F).to(dev), torch.nn.Linear(F, 2 * F).to(dev), torch.nn.Linear(2 * F, 3 * F).to(dev)]
r = torch.rand(P, device=dev) * 5.0
nu = torch.tensor([P], device=dev, dtype=torch.int32)
for nl in (3, 2, 1):
    chain = fused.RadialChain([[l] for l in lins[:nl]], act=True, mul_cutoff=True)
    for dbg in (0, 1, 2, 3, 4):
        os.environ["TMD_TC_DBG"] = str(dbg)
        for _ in range(3):
            fused.run_chain(chain, r, nu, rbf, 0.0, 5.0, True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            fused.run_chain(chain, r, nu, rbf, 0.0, 5.0, True)
        e1.record()
        torch.cuda.synchronize()
        print(f"layers={nl} dbg={dbg}: {e0.elapsed_time(e1) / 10:.3f} ms per chain call (incl. rbf split)", flush=True)
