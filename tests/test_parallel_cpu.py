"""world_size-2 gloo tests of the host-side multi-GPU logic: system partitioning, local
re-basing of `batch`, and the single scalar all-reduce -- with a stand-in energy model (sum of |pos|^2 per system)
since the CUDA kernels cannot run here."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torchmdnet_b200 import parallel


def test_partition_balanced_and_contiguous():
    counts = torch.tensor([30, 20, 40, 10, 25, 35, 22, 18])
    for w in (1, 2, 3, 4, 8):
        r = parallel.partition_systems(counts, w)
        assert r[0][0] == 0 and r[-1][1] == 8 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        loads = [int(counts[s:e].sum()) for s, e in r]
        assert max(loads) - min(loads) <= 45
    assert parallel.partition_systems(torch.tensor([5, 5]), 4)[-1][1] == 2  # more ranks than systems: empty shards allowed


class ToyModel(torch.nn.Module):
    def forward(self, z, pos, batch, box=None, q=None, num_systems=None):
        pos = pos.clone().requires_grad_(True)
        e = (pos ** 2).sum(-1, keepdim=True) * z.unsqueeze(-1).to(pos.dtype)
        y = torch.zeros(num_systems, 1, dtype=pos.dtype).index_add(0, batch, e)
        (g,) = torch.autograd.grad(y.sum(), pos)
        return y.detach(), -g


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    sizes = torch.tensor([7, 3, 9, 4, 6])
    batch = torch.arange(5).repeat_interleave(sizes)
    pos = torch.randn(int(sizes.sum()), 3, generator=g)
    z = torch.randint(1, 9, (pos.shape[0],), generator=g)
    box = torch.eye(3).unsqueeze(0).repeat(5, 1, 1) * torch.arange(1, 6).view(5, 1, 1)
    tot, y, f, rng = parallel.sharded_energy_forces(ToyModel(), z, pos, batch, box=box)
    yall = parallel.gather_system_energies(y, parallel.partition_systems(sizes, world))
    lz, lpos, lbatch, lbox, _, rng2, sl = parallel.shard_batch(z, pos, batch, box, None, rank, world)
    # plain lists, not tensors: a tensor in an mp.Queue travels as a shared-memory handle that dies with this process
    q.put((rank, tot.item(), yall.tolist(), f.tolist(), rng, (sl.start, sl.stop), int(lbatch.min()) if lbatch.numel() else 0, lbox.shape[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_energy_two_ranks():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    g = torch.Generator().manual_seed(0)
    sizes = torch.tensor([7, 3, 9, 4, 6])
    batch = torch.arange(5).repeat_interleave(sizes)
    pos = torch.randn(int(sizes.sum()), 3, generator=g)
    z = torch.randint(1, 9, (pos.shape[0],), generator=g)
    yref, fref = ToyModel()(z, pos, batch, num_systems=5)
    for rank, tot, yall, f, rng, (a0, a1), bmin, nbox in res:
        yall, f = torch.tensor(yall), torch.tensor(f).reshape(-1, 3)
        assert abs(tot - yref.sum().item()) < 1e-4          # one scalar all-reduce -> same total on every rank
        assert torch.allclose(yall, yref, atol=1e-5)         # optional all-gather of per-system energies
        assert torch.allclose(f, fref[a0:a1], atol=1e-6)     # forces stay local to the owning rank
        assert bmin == 0 and nbox == rng[1] - rng[0]         # batch ids re-based, per-sample boxes sliced
    assert res[0][4][1] == res[1][4][0] and res[0][4][0] == 0 and res[1][4][1] == 5
