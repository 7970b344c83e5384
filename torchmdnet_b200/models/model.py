"""Model API kept from the reference: `create_model`, `load_model`, `load_ensemble`, `TorchMD_Net`,
`Ensemble`, `create_prior_models` (reference `torchmdnet/models/model.py:21-680`).

`torchmd-train` (LNNP, module.py:86-89) and `TMDNETCalculator` (calculators.py:213-225) only touch these
entry points, the `hyper_parameters` dict and the state-dict keys -- all preserved -- so they run on the
B200 kernels unchanged.  Architectures in scope: "tensornet", "equivariant-transformer"  and
"tensornet2" (+ ScalarPlusWeightedCoulomb); the deprecated graph-network / transformer models raise.
"""
import os
import re
import tempfile
import warnings
import zipfile
from glob import glob
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import output_modules
from .utils import dtype_mapping
from .. import priors


def create_model(args, prior_model=None, mean=None, std=None):
    """Build a TorchMD_Net from the reference's hyper-parameter dict (model.py:21-164)."""
    dtype = dtype_mapping[args["precision"]]
    args.setdefault("box_vecs", None)
    args.setdefault("static_shapes", False)
    args.setdefault("vector_cutoff", False)
    shared = dict(
        hidden_channels=args["embedding_dimension"], num_layers=args["num_layers"], num_rbf=args["num_rbf"],
        rbf_type=args["rbf_type"], trainable_rbf=args["trainable_rbf"], activation=args["activation"],
        cutoff_lower=float(args["cutoff_lower"]), cutoff_upper=float(args["cutoff_upper"]), max_z=args["max_z"],
        max_num_neighbors=args["max_num_neighbors"],
        box_vecs=torch.tensor(args["box_vecs"], dtype=dtype) if args["box_vecs"] is not None else None, dtype=dtype)
    kind = args["model"]
    if kind == "tensornet":
        from .tensornet import TensorNet
        is_equivariant = False  # TensorNet feeds the plain Scalar head
        representation = TensorNet(equivariance_invariance_group=args["equivariance_invariance_group"],
                                   static_shapes=args["static_shapes"], **shared)
    elif kind == "equivariant-transformer":
        from .torchmd_et import TorchMD_ET
        is_equivariant = True
        representation = TorchMD_ET(attn_activation=args["attn_activation"], num_heads=args["num_heads"],
                                    distance_influence=args["distance_influence"],
                                    neighbor_embedding=args["neighbor_embedding"], vector_cutoff=args["vector_cutoff"],
                                    static_shapes=args["static_shapes"], **shared)
    elif kind == "tensornet2":
        from .tensornet2 import TensorNet2   # host structure + composed forward (see the module docstring)
        is_equivariant = False
        representation = TensorNet2(equivariance_invariance_group=args["equivariance_invariance_group"],
                                    static_shapes=args["static_shapes"], q_dim=args.get("q_dim", 0),
                                    output_charges="Coul" in args["output_model"], **shared)
    elif kind in ("graph-network", "transformer"):
        raise ValueError(f'architecture "{kind}" is outside the B200 hot-path scope (tensornet, equivariant-transformer)')
    else:
        raise ValueError(f'Unknown architecture: {kind}')

    if args.get("atom_filter", -1) > -1:
        raise ValueError("atom_filter is not supported by torchmdnet_b200 (incompatible with forces in the reference too)")
    if args["prior_model"] and prior_model is None:
        prior_model = create_prior_models(args)

    head_name = ("Equivariant" if is_equivariant else "") + args["output_model"]
    if not hasattr(output_modules, head_name):
        raise ValueError(f'output model "{head_name}" is outside the B200 hot-path scope (Scalar / EquivariantScalar)')
    output_model = getattr(output_modules, head_name)(
        args["embedding_dimension"], activation=args["activation"], reduce_op=args["reduce_op"], dtype=dtype,
        static_shapes=args.get("static_shapes", False), num_hidden_layers=args.get("output_mlp_num_layers", 0),
        num_layers=args["num_layers"] if args["output_model"] == "ScalarPlusWeightedCoulomb" else 0,   # reference issue 343
        q_dim=args.get("q_dim", 0), q_weights=args.get("q_weights", []), coulomb_cutoff=args.get("coulomb_cutoff", None),
        coulomb_max_num_neighbors=args.get("coulomb_max_num_neighbors", None),
        coulomb_neighbor_strategy=args.get("coulomb_neighbor_strategy", "brute"))
    return TorchMD_Net(representation, output_model, prior_model=prior_model, mean=mean, std=std,
                       derivative=args["derivative"], dtype=dtype)


def load_ensemble(filepath, args=None, device="cpu", return_std=False, **kwargs):
    """Ensemble from a list of checkpoints or a zip of *.ckpt (model.py:167-205)."""
    if isinstance(filepath, (list, tuple)):
        assert all(isinstance(f, str) for f in filepath), "Invalid filepath list."
        models = [load_model(f, args=args, device=device, **kwargs) for f in filepath]
    elif filepath.endswith(".zip"):
        with tempfile.TemporaryDirectory() as tmp:
            with zipfile.ZipFile(filepath, "r") as zf:
                zf.extractall(tmp)
            ckpts = glob(os.path.join(tmp, "*.ckpt"))
            assert len(ckpts) > 0, "No checkpoint files found in zip file."
            models = [load_model(f, args=args, device=device, **kwargs) for f in ckpts]
    else:
        raise ValueError("Invalid filepath. Must be a list of paths or a path to a zip file.")
    return Ensemble(models, return_std=return_std)


def _remix_rows(weight, bias):
    """Old [N,F,3,3] checkpoints store the 3F output rows channel-major (f*3+k); the new layout is
    irrep-major (k*F+f)  (model.py:321-331)."""
    a, b = weight.shape
    return (weight.view(a // 3, 3, b).transpose(0, 1).reshape(a, b), bias.view(a // 3, 3).transpose(0, 1).reshape(a))


def load_model(filepath, args=None, device="cpu", return_std=False, **kwargs):
    """Load a Lightning checkpoint written by torchmd-train (model.py:208-374): hyper_parameters -> kwargs
    override -> create_model -> key patches -> old-layout remap -> load_state_dict."""
    if isinstance(filepath, (list, tuple)) or filepath.endswith(".zip"):
        return load_ensemble(filepath, args=args, device=device, return_std=return_std, **kwargs)
    assert isinstance(filepath, str)
    ckpt = torch.load(filepath, map_location="cpu", weights_only=False)
    if args is None:
        args = ckpt["hyper_parameters"]
    delta_learning = args.get("remove_ref_energy", False)
    for key, value in kwargs.items():
        if key not in args:
            warnings.warn(f"Unknown hyperparameter: {key}={value}")
        args[key] = value
    if args["model"] in ("tensornetv2_alt", "tensornet-nqe"):
        args["model"] = "tensornet2"
    model = create_model(args)
    if delta_learning and "remove_ref_energy" in kwargs and not kwargs["remove_ref_energy"]:
        assert len(model.prior_model) > 0, "Atomref prior must be added during training (with enable=False) for total energy prediction."
        assert isinstance(model.prior_model[-1], priors.Atomref), "I expected the last prior to be Atomref."
        model.prior_model[-1].enable = True

    sd = {re.sub(r"^model\.", "", k): v for k, v in ckpt["state_dict"].items()}
    for pat, rep in ((r"output_model.output_network.(\d+).update_net.(\d+).", r"output_model.output_network.\1.update_net.layers.\2."),
                     (r"output_model.output_network.([02]).(weight|bias)", r"output_model.output_network.layers.\1.\2")):
        sd = {re.sub(pat, rep, k): v for k, v in sd.items()}
    sd.setdefault("representation_model.distance.box", torch.zeros(3, 3))
    if getattr(model.output_model, "distance", None) is not None:   # Coulomb head with a cutoff (reference model.py:286-291)
        sd.setdefault("output_model.distance.box", torch.zeros(3, 3))

    old_layout = "check_errors" in ckpt.get("hyper_parameters", {})
    if kwargs.get("compatibility_load", old_layout) and args["model"] in ("tensornet", "tensornet2"):
        if old_layout and "compatibility_load" not in kwargs:
            warnings.warn("Old-format checkpoint detected ('check_errors' found in hyper_parameters). "
                          "Automatically applying compatibility_load to remap linear-layer weights. "
                          "Pass compatibility_load=False to suppress this.")
        keys = ["representation_model.tensor_embedding.linears_scalar.1"]
        if args["model"] == "tensornet":   # TensorNet2 interaction layers keep their layout (reference model.py:299)
            keys += [f"representation_model.layers.{l}.linears_scalar.2" for l in range(args["num_layers"])]
        for k in keys:
            sd[k + ".weight"], sd[k + ".bias"] = _remix_rows(sd[k + ".weight"], sd[k + ".bias"])
    model.load_state_dict(sd)
    return model.to(device)


def create_prior_models(args, dataset=None):
    """Parse `prior_model` / `prior_args` (model.py:377-448); only Atomref is in scope."""
    out = []
    if not args["prior_model"]:
        return out
    spec = args["prior_model"] if isinstance(args["prior_model"], list) else [args["prior_model"]]
    names, pargs = [], []
    for item in spec:
        if isinstance(item, dict):
            for k, v in item.items():
                names.append(k)
                pargs.append({} if v is None else v)
        else:
            names.append(item)
            pargs.append({})
    if args.get("prior_args") is not None:
        pargs = args["prior_args"] if isinstance(args["prior_args"], list) else [args["prior_args"]]
    for name, a in zip(names, pargs):
        assert hasattr(priors, name), f"Unknown prior model {name}. Available models are {', '.join(priors.__all__)}"
        out.append(getattr(priors, name)(dataset=dataset, **a))
    return out


class TorchMD_Net(nn.Module):
    """representation -> output head -> *std -> priors -> per-molecule reduce -> +mean -> forces
    (reference model.py:451-631; same constructor and forward signature)."""

    def __init__(self, representation_model, output_model, prior_model=None, mean=None, std=None, derivative=False,
                 dtype=torch.float32):
        super().__init__()
        self.representation_model = representation_model.to(dtype=dtype)
        self.output_model = output_model.to(dtype=dtype)
        if not output_model.allow_prior_model and prior_model is not None:
            prior_model = None
            warnings.warn("Prior model was given but the output model does not allow prior models. Dropping the prior model.")
        if isinstance(prior_model, priors.BasePrior):
            prior_model = [prior_model]
        self.prior_model = None if prior_model is None else nn.ModuleList(prior_model).to(dtype=dtype)
        self.derivative = derivative
        self.register_buffer("mean", (torch.scalar_tensor(0) if mean is None else mean).to(dtype=dtype))
        self.register_buffer("std", (torch.scalar_tensor(1) if std is None else std).to(dtype=dtype))
        self.reset_parameters()
        from ..extensions import frame_op      # handle of this module for the whole-frame op (torch.compile support); set here, outside
        self._frame_op_handle = frame_op.handle_of(self)   # any traced region: Dynamo cannot trace the weak registry itself

    def __prepare_scriptable__(self):
        """torch.jit.script(model): the scripting compiler gets a module whose graph is one call of the whole-frame op
        (extensions/frame_op.py); in-process only, like the reference's Python-registered ops (tests/test_model.py:207-262)"""
        from ..extensions import frame_op
        return frame_op.scriptable(self)

    def __setstate__(self, state):
        super().__setstate__(state)            # copies / unpickled modules get a handle of their own
        from ..extensions import frame_op
        self._frame_op_handle = frame_op.handle_of(self)

    def reset_parameters(self):
        self.representation_model.reset_parameters()
        self.output_model.reset_parameters()
        if self.prior_model is not None:
            for prior in self.prior_model:
                prior.reset_parameters()

    fused_head_path = None   # None: automatic; False: always the torch-op tail / head

    def _fused_head(self):
        """(HeadWeights, atomref weight or None) when the frame ends in the configuration csrc/head.cu implements -- TensorNet on its
        fused inference path, `Scalar` head with the default 2-layer SiLU MLP, sum reduction, no prior or Atomref only, frozen
        weights -- else None (torch-op tail and head, which also provide weight gradients)."""
        from .tensornet import TensorNet
        rep, out = self.representation_model, self.output_model
        if self.fused_head_path is False or type(rep) is not TensorNet or type(out) is not output_modules.Scalar:
            return None
        if rep.activation != "silu" or out.reduce_op not in ("add", "sum") or self.mean is None or self.std is None:
            return None
        layers = out.output_network.layers
        if len(layers) != 3 or not isinstance(layers[1], nn.SiLU) or self.mean.dtype != torch.float32:
            return None
        atomref = None
        if self.prior_model is not None and len(self.prior_model) > 0:
            if len(self.prior_model) != 1 or type(self.prior_model[0]) is not priors.Atomref or self.prior_model[0].atomref.weight.requires_grad:
                return None
            if self.prior_model[0].enable:
                atomref = self.prior_model[0].atomref.weight
        hw = rep.head_weights(out.output_network)
        if not (hw.supported() and hw.frozen() and rep._fused_path_ok()):
            return None
        return hw, atomref

    def _fused_et_head(self):
        """(block weights, atomref weight or None) when the frame ends in EquivariantScalar on a frozen ET with a sum reduction: the
        two gated equivariant blocks, * std, + atomref, per-molecule sum and + mean run as one autograd node (fused.et_head_energy,
        row a15); else None."""
        from .torchmd_et import TorchMD_ET
        rep, out = self.representation_model, self.output_model
        if self.fused_head_path is False or type(rep) is not TorchMD_ET or type(out) is not output_modules.EquivariantScalar:
            return None
        if out.reduce_op not in ("add", "sum") or self.mean is None or self.std is None or self.mean.dtype != torch.float32:
            return None
        atomref = None
        if self.prior_model is not None and len(self.prior_model) > 0:
            if len(self.prior_model) != 1 or type(self.prior_model[0]) is not priors.Atomref or self.prior_model[0].atomref.weight.requires_grad:
                return None
            if self.prior_model[0].enable:
                atomref = self.prior_model[0].atomref.weight
        from ..extensions import fused
        if getattr(self, "_et_head_w", None) is None or any(w.block is not b for w, b in zip(self._et_head_w, out.output_network)):
            self._et_head_w = [fused.GatedBlockWeights(b) for b in out.output_network]
        ws = self._et_head_w
        if not all(w.supported() and w.frozen() for w in ws) or ws[-1].block.out_channels != 1:
            return None
        if any(t.requires_grad for t in rep.parameters()):
            return None
        return ws, atomref

    def forward(self, z: Tensor, pos: Tensor, batch: Optional[Tensor] = None, box: Optional[Tensor] = None,
                q: Optional[Tensor] = None, s: Optional[Tensor] = None, extra_args: Optional[Dict[str, Tensor]] = None,
                num_systems: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        if torch.compiler.is_compiling():
            # torch.compile(model, fullgraph=True): the frame is one opaque op with a fake implementation (extensions/frame_op.py)
            from ..extensions import frame_op
            if not self.derivative:
                raise RuntimeError("torch.compile of torchmdnet_b200 models needs derivative=True (the compiled frame returns the forces "
                                   "itself); for energy.backward() callers use TMDNETCalculator(compile=True) / CUDAGraphForceField")
            if num_systems is None:
                raise RuntimeError("torch.compile of torchmdnet_b200 models needs num_systems (the output shape must be known at trace time)")
            batch = torch.zeros_like(z) if batch is None else batch
            return torch.ops.torchmdnet_b200.energy_forces(self._frame_op_handle, z, pos, batch, box, q, num_systems)
        return self._forward_impl(z, pos, batch, box=box, q=q, s=s, extra_args=extra_args, num_systems=num_systems)

    def _forward_impl(self, z: Tensor, pos: Tensor, batch: Optional[Tensor] = None, box: Optional[Tensor] = None,
                      q: Optional[Tensor] = None, s: Optional[Tensor] = None, extra_args: Optional[Dict[str, Tensor]] = None,
                      num_systems: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        assert z.dim() == 1 and z.dtype == torch.long
        if not pos.is_cuda:
            raise RuntimeError("torchmdnet_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        batch = torch.zeros_like(z) if batch is None else batch
        if self.derivative:
            pos.requires_grad_(True)
        fused_head = self._fused_head()
        if fused_head is not None:
            # rows a10 + a11 as three launches (csrc/head.cu): norm3 -> LayerNorm -> Linear -> SiLU -> MLP -> * std -> + atomref ->
            # per-molecule sum -> + mean, with a hand-written backward to the node tensors
            from ..extensions import fused
            hw, atomref = fused_head
            X = self.representation_model.forward_irreps(z, pos, batch, box=box, q=q, s=s)
            m = self.output_model.resolve_num_systems(batch, num_systems)
            y = fused.tail_head_energy(X, z, batch, m, hw, self.std, self.mean, atomref)
        elif self._fused_et_head() is not None:
            from ..extensions import fused
            ws, atomref = self._fused_et_head()
            x, v, z, pos, batch = self.representation_model(z, pos, batch, box=box, q=q, s=s)
            m = self.output_model.resolve_num_systems(batch, num_systems)
            y = fused.et_head_energy(x, v, z, batch, m, ws, self.std, self.mean, atomref)
        else:
            x, v, z, pos, batch = self.representation_model(z, pos, batch, box=box, q=q, s=s)
            x = self.output_model.pre_reduce(x, v, z, pos, batch, box=box)
            if self.std is not None:
                x = x * self.std
            if self.prior_model is not None:
                for prior in self.prior_model:
                    x = prior.pre_reduce(x, z, pos, batch, extra_args)
            x = self.output_model.reduce(x, batch, num_systems=num_systems)
            if self.mean is not None:
                x = x + self.mean
            y = self.output_model.post_reduce(x)
            if self.prior_model is not None:
                for prior in self.prior_model:
                    y = prior.post_reduce(y, z, pos, batch, box, extra_args)
        if self.derivative:
            dy = torch.autograd.grad([y], [pos], grad_outputs=[torch.ones_like(y)], create_graph=self.training,
                                     retain_graph=self.training)[0]
            assert dy is not None, "Autograd returned None for the force prediction."
            return y, -dy
        return y, torch.empty(0)


class Ensemble(nn.ModuleList):
    """Mean (and optionally std) over several TorchMD_Net models (model.py:634-680)."""

    def __init__(self, modules: List[nn.Module], return_std: bool = False):
        for m in modules:
            assert isinstance(m, TorchMD_Net)
        super().__init__(modules)
        self.return_std = return_std

    def forward(self, *args, **kwargs):
        ys, fs = zip(*[m(*args, **kwargs)[:2] for m in self])
        y, f = torch.stack(ys), torch.stack(fs)
        if self.return_std:
            return y.mean(0), f.mean(0), y.std(0), f.std(0)
        return y.mean(0), f.mean(0)
