"""`torch.compile(model, fullgraph=True)` support (reference calculators.py:285-309).

The reference makes MD fast with `torch.compile(model, backend="inductor", dynamic=False, fullgraph=True, mode="reduce-overhead")`;
that works there because every Warp kernel sits behind a `torch.library` op with a fake (shape) implementation.  Here the frozen
energy + force frame is a fixed sequence of ctypes launches that Dynamo cannot trace, so the WHOLE frame is exposed as one opaque
op with a fake implementation,

    torchmdnet_b200::energy_forces(int model, Tensor z, Tensor pos, Tensor batch, Tensor? box, Tensor? q, int num_systems)
        -> (Tensor energy [M,1], Tensor forces [N,3])

and `TorchMD_Net.forward` dispatches to it while `torch.compiler.is_compiling()`.  The compiled graph then contains exactly one
node; with `mode="reduce-overhead"` Inductor captures it into a CUDA graph, which every kernel of the frame allows (static shapes,
device-side counters).  Only the force-returning form (`derivative=True`) is exposed: the op computes the forces itself, so no
autograd formula is registered.  `model` is a handle into a weak registry (custom ops take no module arguments)."""
import concurrent.futures
import weakref
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

_models = weakref.WeakValueDictionary()
_pool = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="tmd-frame")
_next = [1]


def handle_of(model) -> int:
    """register `model` (weakly) and return its integer handle; called from TorchMD_Net.__init__, never from traced code"""
    h = _next[0]
    _next[0] += 1
    _models[h] = model
    return h


@torch.library.custom_op("torchmdnet_b200::energy_forces", mutates_args=(), device_types=["cuda"])
def energy_forces(model: int, z: Tensor, pos: Tensor, batch: Tensor, box: Optional[Tensor], q: Optional[Tensor],
                  num_systems: int) -> Tuple[Tensor, Tensor]:
    m = _models.get(model)
    if m is None:
        raise RuntimeError("torchmdnet_b200::energy_forces: the model behind this handle no longer exists")
    # The body of a custom op runs below the autograd dispatch keys of the calling thread, so the force evaluation inside the frame
    # (autograd chaining the hand-written backward kernels) would record nothing there.  Dispatch state is thread-local: the frame runs
    # on a worker thread bound to the caller's device and current stream (stream order and graph capture are those of the caller).
    dev, stream = pos.device, torch.cuda.current_stream(pos.device)

    def body():
        with torch.cuda.device(dev), torch.cuda.stream(stream), torch.enable_grad():
            y, f = m._forward_impl(z, pos.detach().clone(), batch, box=box, q=q, num_systems=num_systems)
            return y.detach(), f.detach()

    return _pool.submit(body).result()


@energy_forces.register_fake
def _(model, z, pos, batch, box, q, num_systems):
    return pos.new_empty(num_systems, 1), pos.new_empty(pos.shape[0], 3)


# ------------------------------------------------------------------------------------------ TorchScript (in-process)
# The reference scripts its model in-process (tests/test_model.py:207-262; its Warp ops are Python-registered too, so a scripted
# module only runs where those registrations exist).  `torch.jit.script(model)` on a TorchMD_Net of this build goes through
# `__prepare_scriptable__`, which hands the compiler this module: the scripted graph is ONE call of torchmdnet_b200::energy_forces.
_pinned = {}     # handle -> model: a scripted wrapper refers to its model by handle only, so the model is kept alive here


class ScriptedFrame(torch.nn.Module):
    """forward(z, pos, batch, box, q, s, extra_args, num_systems) -> (energy, forces) with TorchMD_Net.forward's signature"""

    def __init__(self, handle: int):
        super().__init__()
        self.handle = handle

    def forward(self, z: Tensor, pos: Tensor, batch: Optional[Tensor] = None, box: Optional[Tensor] = None, q: Optional[Tensor] = None,
                s: Optional[Tensor] = None, extra_args: Optional[Dict[str, Tensor]] = None,
                num_systems: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        if batch is None:
            batch = torch.zeros_like(z)
        n = 0
        if num_systems is not None:
            n = num_systems
        else:
            n = int(batch.max().item()) + 1
        y, f = torch.ops.torchmdnet_b200.energy_forces(self.handle, z, pos, batch, box, q, n)
        return y, f


def scriptable(model) -> ScriptedFrame:
    _pinned[model._frame_op_handle] = model
    return ScriptedFrame(model._frame_op_handle)
