"""Inference adapter: whole-frame CUDA-graph replay of a TorchMD_Net model.

The reference's TorchMD adapter `External` captures the model into a CUDA graph when asked to
(torchmdnet/calculators.py:117-169, "dramatically increase performance during inference",
docs/source/usage.rst:132).  `CUDAGraphForceField` is that capability for the B200 path: the
model must be built with `static_shapes=True` (capacity-sized neighbour buffers + device-side counters:
nothing in the frame depends on a host-visible shape, so ~250 launches replay with zero CPU work).
"""
import warnings
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

try:  # optional, exactly like the reference's ASE adapter (calculators.py:180)
    import ase.calculators.calculator as ase_calc
except ImportError:  # ASE is not a dependency of the hot path: a minimal stand-in keeps the adapter usable and testable
    ase_calc = None

# preset output transforms of the reference adapter (torchmdnet/calculators.py:9-23; the misspelt name is the reference's)
tranforms = {
    "eV/A -> kcal/mol/A": lambda energy, forces: (energy * 23.0609, forces * 23.0609),
    "Hartree/Bohr -> kcal/mol/A": lambda energy, forces: (energy * 627.509, forces * 627.509 / 0.529177),
    "Hartree/A -> kcal/mol/A": lambda energy, forces: (energy * 627.509, forces * 627.509),
}


class CUDAGraphForceField:
    """energy, forces = ff(pos) for a fixed (z, batch, box, q); positions may change every call.

    Raises at the next synchronisation if a frame overflows the neighbour capacity
    (`max_num_neighbors * n_atoms`, checked on the device exactly like the eager path)."""

    def __init__(self, model, z: Tensor, batch: Optional[Tensor] = None, box: Optional[Tensor] = None,
                 q: Optional[Tensor] = None, num_systems: Optional[int] = None, warmup: int = 3, allreduce_total: bool = False,
                 group=None):
        rep = model.representation_model
        if not getattr(rep, "static_shapes", False):
            raise ValueError("CUDA-graph capture needs a representation model with static_shapes=True "
                             f"({type(rep).__name__} has none or it is off)")
        if not z.is_cuda:
            raise RuntimeError("CUDAGraphForceField needs CUDA tensors")
        self.model = model
        dev = z.device
        self.z = z.clone()
        self.batch = torch.zeros_like(z) if batch is None else batch.clone()
        self.box = None if box is None else box.clone()
        self.q = None if q is None else q.clone()
        self.num_systems = int(self.batch.max().item()) + 1 if num_systems is None else int(num_systems)
        self.pos = torch.zeros(z.shape[0], 3, device=dev, dtype=torch.float32)
        self.graph = None
        self._warmup = warmup
        self._derivative = bool(model.derivative)
        # multi-GPU data parallelism: the ONE collective of the path -- an all-reduce of the scalar total energy over the
        # process group -- is recorded inside the captured frame (NCCL kernels are graph-capturable), so a replay needs no host-side
        # launch per step.  Every rank must construct / capture / replay collectively.
        self._allreduce, self._group = bool(allreduce_total), group
        self.total_energy = None

    def _frame(self):
        pos = self.pos
        if not self._derivative:
            pos = pos.detach().requires_grad_(True)
        y, f = self.model(self.z, pos, self.batch, box=self.box, q=self.q, num_systems=self.num_systems)
        if not self._derivative:  # ASE-style: forces from the returned energy (calculators.py:315)
            (g,) = torch.autograd.grad([y.sum()], [pos])
            f = -g
        if self._allreduce:
            import torch.distributed as dist
            self.total_energy = y.detach().sum().reshape(1)
            dist.all_reduce(self.total_energy, op=dist.ReduceOp.SUM, group=self._group)
        return y, f

    def capture(self, pos: Tensor):
        with torch.no_grad():
            self.pos.copy_(pos)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                self._frame()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.energy, self.forces = self._frame()
        return self

    def __call__(self, pos: Tensor) -> Tuple[Tensor, Tensor]:
        if self.graph is None:
            self.capture(pos)
        with torch.no_grad():
            self.pos.copy_(pos, non_blocking=True)
        self.graph.replay()
        return self.energy, self.forces


class External:
    """TorchMD adapter with the reference's interface (torchmdnet/calculators.py:26-177):
    `External(netfile, embeddings, device, output_transform, use_cuda_graph, cuda_graph_warmup_steps, dtype, **kwargs)` and
    `calculate(pos, box) -> (energy [n_replicas,1], forces [n_replicas,n_atoms,3])`.

    `embeddings` is [n_replicas, n_atoms] atomic numbers; replicas are independent systems (batch ids 0..n_replicas-1).
    With `use_cuda_graph=True` the whole frame is captured once (static shapes) and replayed: positions and box are copied into
    the graph's static buffers every call.  This package has no CPU path: `device` must be a CUDA device."""

    def __init__(self, netfile, embeddings, device="cuda", output_transform=None, use_cuda_graph=False,
                 cuda_graph_warmup_steps=12, dtype=torch.float32, **kwargs):
        from .models.model import load_model
        if not str(device).startswith("cuda"):
            raise ValueError("torchmdnet_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        if isinstance(netfile, str):
            extra_args = dict(kwargs)
            if use_cuda_graph:
                warnings.warn("CUDA graphs are enabled, setting static_shapes=True")
                extra_args["static_shapes"] = True
            self.model = load_model(netfile, device=device, derivative=True, **extra_args)
        elif isinstance(netfile, torch.nn.Module):
            if kwargs:
                warnings.warn("Warning: extra arguments are being ignored when passing a torch.nn.Module")
            self.model = netfile
        else:
            raise ValueError(f"Expected a path to a checkpoint file or a torch.nn.Module, got {type(netfile)}")
        self.device = device
        self.n_atoms = embeddings.size(1)
        self.n_replicas = embeddings.size(0)
        self.embeddings = embeddings.reshape(-1).to(device)
        self.batch = torch.arange(embeddings.size(0), device=device).repeat_interleave(embeddings.size(1))
        self.model.eval()
        for p_ in self.model.parameters():   # inference adapter: weights are constants (this is what selects the fused pair pipeline)
            p_.requires_grad_(False)
        if not output_transform:
            self.output_transformer = lambda energy, forces: (energy, forces)
        elif output_transform in tranforms:
            self.output_transformer = tranforms[output_transform]
        elif callable(output_transform):
            self.output_transformer = output_transform
        else:
            raise ValueError(f"Unknown output_transform {output_transform!r}: a key of `tranforms` or a callable")
        self.use_cuda_graph = use_cuda_graph
        self.cuda_graph_warmup_steps = cuda_graph_warmup_steps
        self._ff = None
        if isinstance(dtype, str):
            try:
                dtype = getattr(torch, dtype)
            except AttributeError:
                raise ValueError(f"Unknown torch dtype {dtype}")
        if dtype != torch.float32:
            raise ValueError("the B200 TensorNet / ET kernels compute in float32")
        self.dtype = dtype

    def calculate(self, pos, box=None):
        pos = pos.to(self.device).to(self.dtype).reshape(-1, 3)
        if box is not None:
            box = box.to(self.device).to(self.dtype)
        if self.use_cuda_graph:
            if self._ff is None:
                self._ff = CUDAGraphForceField(self.model, self.embeddings, self.batch, box=box, num_systems=self.n_replicas,
                                               warmup=self.cuda_graph_warmup_steps)
            elif box is not None:
                if self._ff.box is None:
                    raise ValueError("the frame was captured without a box; pass the box on the first call when use_cuda_graph=True")
                with torch.no_grad():
                    self._ff.box.copy_(box)
            energy, forces = self._ff(pos.detach())
        else:
            energy, forces = self.model(self.embeddings, pos, self.batch, box=box, num_systems=self.n_replicas)
        assert forces is not None, "The model is not returning forces"
        assert energy is not None, "The model is not returning energy"
        return self.output_transformer(energy.clone().detach(), forces.clone().reshape(-1, self.n_atoms, 3).detach())


class _MiniCalculator:
    """The part of ase.calculators.calculator.Calculator the adapter relies on (used only when ASE is not installed)."""
    implemented_properties = []

    def __init__(self):
        self.atoms, self.results = None, {}

    def calculate(self, atoms=None, properties=None, system_changes=None):
        if atoms is not None:
            self.atoms = atoms.copy() if hasattr(atoms, "copy") else atoms

    def check_state(self, atoms, tol: float = 1e-15):
        return ["positions"]

    def get_potential_energy(self, atoms=None):
        self.calculate(atoms, ["energy"], all_changes)
        return self.results["energy"]

    def get_forces(self, atoms=None):
        self.calculate(atoms, ["forces"], all_changes)
        return self.results["forces"]


all_changes = list(ase_calc.all_changes) if ase_calc is not None else ["positions", "numbers", "cell", "pbc", "initial_charges",
                                                                        "initial_magmoms"]
_CalculatorBase = ase_calc.Calculator if ase_calc is not None else _MiniCalculator


class TMDNETCalculator(_CalculatorBase):
    """ASE adapter with the reference's interface (torchmdnet/calculators.py:183-320):
    `TMDNETCalculator(model_file, device, dtype, compile, **kwargs)`; `calculate(atoms, properties, system_changes)` fills
    `results["energy"]` (float, eV) and `results["forces"]` (numpy [N,3]).

    As in the reference the model is loaded with `derivative=False`, its parameters are frozen, the total charge comes from
    `atoms.info["charge"]`, the box from `atoms.cell` when any direction is periodic, and the forces are obtained by
    `energy.backward()` on the returned energy (calculators.py:315) -- plain first-order autograd through the hand-written
    backward kernels.  `compile=True` is the reference's "make MD fast" switch (torch.compile(mode="reduce-overhead"), i.e. CUDA
    graphs behind a tracing compiler); here it captures the whole frame into a CUDA graph directly (`CUDAGraphForceField`,
    static shapes) and re-captures when the atomic numbers change, like the reference recompiles.  `device` must be a CUDA
    device: this build has no CPU path (`device="cpu"`, the reference's default, raises)."""

    implemented_properties = ["energy", "forces"]

    def __init__(self, model_file, device="cuda", dtype=torch.float32, compile=False, **kwargs):
        from .models.model import load_model
        _CalculatorBase.__init__(self)
        if not str(device).startswith("cuda"):
            raise ValueError("torchmdnet_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        if isinstance(dtype, str):
            dtype = getattr(torch, dtype)
        if dtype != torch.float32:
            raise ValueError("the B200 TensorNet / ET kernels compute in float32")
        self.device = device
        self.remove_ref_energy = kwargs.pop("remove_ref_energy", True)
        self.max_num_neighbors = kwargs.pop("max_num_neighbors", 64)
        self.static_shapes = bool(compile)
        self.model_file, self.model_kwargs = model_file, kwargs
        if isinstance(model_file, torch.nn.Module):   # convenience the reference does not have: an already-built TorchMD_Net
            self.model = model_file
            self.model.derivative = False
        else:
            self.model = load_model(model_file, derivative=False, remove_ref_energy=self.remove_ref_energy,
                                    max_num_neighbors=self.max_num_neighbors, static_shapes=self.static_shapes, **kwargs)
        for parameter in self.model.parameters():
            parameter.requires_grad = False
        self.model.to(device=self.device, dtype=dtype).eval()
        self.compile, self.compiled = bool(compile), False
        self._ff = None
        self.evals = 0

    def check_state(self, atoms, tol: float = 1e-15) -> list:
        # the total charge lives in atoms.info, which ASE does not watch (reference calculators.py:240-251)
        state = _CalculatorBase.check_state(self, atoms, tol=tol)
        if (not state) and (self.atoms is not None) and (self.atoms.info != atoms.info):
            state.append("info")
        return state

    def calculate(self, atoms=None, properties=None, system_changes=all_changes):
        if not properties:
            properties = ["energy"]
        _CalculatorBase.calculate(self, atoms, properties, system_changes)
        dev = self.device
        numbers = torch.tensor(np.asarray(atoms.numbers), device=dev, dtype=torch.long)
        total_charge = torch.tensor([atoms.info.get("charge", 0)], device=dev, dtype=torch.float32)
        box = torch.tensor(np.asarray(atoms.cell.array), device=dev, dtype=torch.float32) if np.any(atoms.pbc) else None
        batch = torch.zeros(numbers.shape[0], device=dev, dtype=torch.long)
        if self.compile:
            if self._ff is None or "numbers" in system_changes and not torch.equal(self._ff.z, numbers):
                self._ff = CUDAGraphForceField(self.model, numbers, batch, box=box, q=total_charge, num_systems=1)
                self.compiled = True
            else:
                with torch.no_grad():
                    self._ff.q.copy_(total_charge)
                    if box is not None:
                        self._ff.box.copy_(box)
            pos = torch.tensor(np.asarray(atoms.positions), device=dev, dtype=torch.float32)
            energy, forces = self._ff(pos)
        else:
            positions = torch.tensor(np.asarray(atoms.positions), device=dev, dtype=torch.float32, requires_grad=True)
            energy, _ = self.model(numbers, positions, batch=batch, q=total_charge, box=box)
            energy.backward()
            forces = -positions.grad
        self.results["energy"] = energy.detach().cpu().item()
        self.results["forces"] = forces.detach().cpu().numpy()
        self.evals += 1
