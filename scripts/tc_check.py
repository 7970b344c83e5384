"""Dev helper (GPU): tcgen05 chain vs the fp32 CUDA-core chain vs torch, value and tangent. Run under `timeout`."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as Fn
from torchmdnet_b200.extensions import fused
from torchmdnet_b200.models.utils import ExpNormalSmearing, CosineCutoff

dev = torch.device("cuda:0")
torch.manual_seed(0)
R, F = 32, 128
rbf = ExpNormalSmearing(0.0, 5.0, R, trainable=False).to(dev)
cfgs = {
    "interaction": ([torch.nn.Linear(R, F), torch.nn.Linear(F, 2 * F), torch.nn.Linear(2 * F, 3 * F)], True, True, False),
    "embedding": ([torch.nn.Linear(R, F) for _ in range(3)], False, False, True),
}
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for name, (lins, act, cut, concat) in cfgs.items():
    for l in lins:
        l.to(dev)
    chain = fused.RadialChain([lins] if concat else [[l] for l in lins], act=act, mul_cutoff=cut)
    r = (torch.rand(P, device=dev) * 5.2).double()
    r[:4] = torch.tensor([0.0, 4.9999, 5.0, 1e-3], device=dev, dtype=torch.float64)
    nu = torch.tensor([P], device=dev, dtype=torch.int32)
    # fp64 truth
    rd = r.clone().requires_grad_(True)
    rbf64 = ExpNormalSmearing(0.0, 5.0, R, trainable=False, dtype=torch.float64).to(dev).double()
    x = rbf64(rd)
    if concat:
        ref = torch.cat([Fn.linear(x, l.weight.double(), l.bias.double()) for l in lins], -1)
    else:
        for l in lins:
            x = Fn.silu(Fn.linear(x, l.weight.double(), l.bias.double()))
        ref = x * CosineCutoff(0.0, 5.0)(rd).unsqueeze(-1)
    w = torch.randn_like(ref)
    (gref,) = torch.autograd.grad((ref * w).sum(), rd)
    for tc, f16 in ((False, False), (True, False), (True, True)):
        fused.USE_TENSOR_CORES = tc
        fused.USE_F16_SPLIT = f16
        f, fp = fused.run_chain(chain, r.float(), nu, rbf, 0.0, 5.0, True)
        torch.cuda.synchronize()
        ev = (f.double() - ref).abs().max().item() / ref.abs().max().item()
        et = ((fp.double() * w).sum(-1) - gref).abs().max().item() / gref.abs().max().item()
        t0 = time.perf_counter()
        for _ in range(5):
            fused.run_chain(chain, r.float(), nu, rbf, 0.0, 5.0, True)
        torch.cuda.synchronize()
        print(f"{name:12s} tc={tc} f16={f16}: rel err value {ev:.2e} tangent {et:.2e}  time {1e3 * (time.perf_counter() - t0) / 5:.3f} ms (P={P})", flush=True)
