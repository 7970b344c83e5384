"""Inference adapter: whole-frame CUDA-graph replay of a TorchMD_Net model.

The reference's TorchMD adapter `External` captures the model into a CUDA graph when asked to
(torchmdnet/calculators.py:117-169, "dramatically increase performance during inference",
docs/source/usage.rst:132).  `CUDAGraphForceField` is that capability for the B200 path: the
model must be built with `static_shapes=True` (capacity-sized neighbour buffers + device-side counters:
nothing in the frame depends on a host-visible shape, so ~250 launches replay with zero CPU work).
"""
from typing import Optional, Tuple

import torch
from torch import Tensor


class CUDAGraphForceField:
    """energy, forces = ff(pos) for a fixed (z, batch, box, q); positions may change every call.

    Raises at the next synchronisation if a frame overflows the neighbour capacity
    (`max_num_neighbors * n_atoms`, checked on the device exactly like the eager path)."""

    def __init__(self, model, z: Tensor, batch: Optional[Tensor] = None, box: Optional[Tensor] = None,
                 q: Optional[Tensor] = None, num_systems: Optional[int] = None, warmup: int = 3):
        rep = model.representation_model
        if not getattr(rep, "static_shapes", True):
            raise ValueError("CUDA-graph capture needs a model created with static_shapes=True")
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

    def _frame(self):
        pos = self.pos
        if not self._derivative:
            pos = pos.detach().requires_grad_(True)
        y, f = self.model(self.z, pos, self.batch, box=self.box, q=self.q, num_systems=self.num_systems)
        if not self._derivative:  # ASE-style: forces from the returned energy (calculators.py:315)
            (g,) = torch.autograd.grad([y.sum()], [pos])
            f = -g
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
