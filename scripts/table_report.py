"""Dev helper (GPU): the radial tables of a model -- gate report per chain, and golden parity with tables on / off."""
import os, sys, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
from torchmdnet_b200.extensions import fused
import test_gpu_model as T

dev = torch.device("cuda:0")
for name in T.CASES:
    fx = torch.load(os.path.join(T.GOLDEN, name + ".pt"), weights_only=False)
    if fx["hparams"]["model"] != "tensornet":
        continue
    out = {}
    for on in (True, False):
        fused.USE_RADIAL_TABLE = on
        m = T.build(fx)
        y, f = T.run(m, fx)
        out[on] = (T.rel(y, fx["ref_E64"]), T.relF(f, fx["ref_F64"]))
        if on:
            rep = m.representation_model
            chains = [rep.tensor_embedding.radial_chain()] + [l.radial_chain() for l in rep.layers]
            reps = [getattr(c, "table_report", None) for c in chains]
    print(f"{name:32s} table: relE {out[True][0]:.2e} relF {out[True][1]:.2e} | exact: relE {out[False][0]:.2e} relF {out[False][1]:.2e}")
    for r in reps:
        print("     ", None if r is None else {k: (f"{v:.1e}" if isinstance(v, float) else v) for k, v in r.items()})
