"""Multi-GPU data parallelism over independent molecules / boxes.

The path shards naturally: edges exist only between atoms with equal `batch` id (reference
warp_kernels/neighbors_brute.py:107-110) and the energy reduction is per `batch`
(models/output_modules.py:43-73), so whole systems are assigned to ranks (contiguous ranges balanced by
atom count), weights are replicated, forces stay on the owning rank and the only exchange is ONE
all-reduce of the scalar total energy / loss (or an all-gather of the per-system energies when the
caller wants them all).  One process per GPU; the process group is torch.distributed (NCCL over NVLink on
the GPU box, gloo in the CPU tests).  A single box is not domain-decomposed (neither does the reference).
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def partition_systems(atoms_per_system: Tensor, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) system ranges per rank, balanced by atom count (greedy prefix split)."""
    n = int(atoms_per_system.numel())
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    csum = torch.cumsum(atoms_per_system.to(torch.float64), 0)
    total = float(csum[-1]) if n else 0.0
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(torch.searchsorted(csum, torch.tensor(target, dtype=torch.float64)).item()) if n else 0
        # choose the closer of the two neighbouring cut points, keep monotone
        if 0 < k < n and abs(float(csum[k - 1]) - target) <= abs(float(csum[k]) - target):
            k = k - 1
        k = min(max(k + 1 if n else 0, bounds[-1]), n)
        bounds.append(k)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def shard_batch(z: Tensor, pos: Tensor, batch: Tensor, box: Optional[Tensor], q: Optional[Tensor], rank: int,
                world_size: int):
    """Local slice of a (sorted-`batch`) multi-system input for `rank`; `batch` ids are re-based to start at 0.
    Returns (z, pos, batch, box, q, (first_system, last_system_exclusive), atom_slice)."""
    n_sys = int(batch.max().item()) + 1 if batch.numel() else 0
    counts = torch.bincount(batch, minlength=n_sys)
    s0, s1 = partition_systems(counts, world_size)[rank]
    offs = torch.cat([torch.zeros(1, dtype=torch.long, device=counts.device), torch.cumsum(counts, 0)])
    a0, a1 = int(offs[s0]), int(offs[s1])
    sl = slice(a0, a1)
    lbox = box
    if box is not None and box.dim() == 3 and box.shape[0] > 1:
        lbox = box[s0:s1]
    lq = None if q is None else q[s0:s1]
    return z[sl], pos[sl], batch[sl] - s0, lbox, lq, (s0, s1), sl


def allreduce_total_energy(local_energies: Tensor, group=None) -> Tensor:
    """THE collective of the path: sum over ranks of the local total energy (a 4-byte all-reduce)."""
    tot = local_energies.sum().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    return tot


def gather_system_energies(local_energies: Tensor, ranges: List[Tuple[int, int]], group=None) -> Tensor:
    """Optional: per-system energies of ALL systems on every rank (all-gather of ragged [M_local,1] blocks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_energies
    world = dist.get_world_size(group)
    width = max(e - s for s, e in ranges)
    buf = torch.zeros(width, *local_energies.shape[1:], dtype=local_energies.dtype, device=local_energies.device)
    buf[: local_energies.shape[0]] = local_energies
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[: e - s] for o, (s, e) in zip(out, ranges)], 0)


def sharded_energy_forces(model, z, pos, batch, box=None, q=None, group=None):
    """Evaluate a multi-system input with systems sharded over the process group.
    Returns (total_energy [1] identical on all ranks, local per-system energies, local forces, system range)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lz, lpos, lbatch, lbox, lq, rng, _ = shard_batch(z, pos, batch, box, q, rank, world)
    n_local = rng[1] - rng[0]
    if lz.numel() == 0:
        y = torch.zeros(0, 1, device=pos.device, dtype=pos.dtype)
        f = torch.zeros(0, 3, device=pos.device, dtype=pos.dtype)
    else:
        y, f = model(lz, lpos, lbatch, box=lbox, q=lq, num_systems=n_local)
    return allreduce_total_energy(y.detach(), group), y, f, rng
