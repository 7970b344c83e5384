"""-m gpu: neighbour-list parity (rows a1-a3 of SURVEY.md section 8) through the reference-shaped operator
`get_neighbor_pairs_kernel` / `OptimizedDistance`, i.e. through the C ABI.

Bar: indices bit-exact (after the canonical (i,j) sort, which our kernel emits natively); deltas/distances
bit-exact against the oracle (same operation order, no FMA) and allclose against the reference's numpy checker.
Mirrors the reference's tests/test_neighbors.py:74-375,608-705 and tests/test_model_utils.py:13-88.
"""
import os

import pytest
import torch

import tensornet_oracle as orc
import systems
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def run_op(pos, batch, box, lo, hi, loop, tr, cap, strategy="brute"):
    from torchmdnet_b200.extensions.ops import get_neighbor_pairs_kernel
    nb, vec, dist, npairs = get_neighbor_pairs_kernel(strategy, pos.to(dev), batch.to(dev),
                                                     torch.zeros(3, 3, device=dev) if box is None else box.to(dev),
                                                     box is not None, lo, hi, cap, loop, tr, 0)
    n = int(npairs.item())
    assert n <= cap
    assert (nb[:, n:] == -1).all() and (vec[n:] == 0).all() and (dist[n:] == 0).all()  # padding contract
    return nb[:, :n].cpu(), vec[:n].cpu(), dist[:n].cpu(), n


@pytest.mark.parametrize("strategy", ["brute", "cell"])
def test_numpy_reference_vectors(strategy):
    cases = torch.load(os.path.join(GOLDEN, "neighbor_cases.pt"), weights_only=False)
    for c in cases:
        nb, vec, dist, n = run_op(c["pos"], c["batch"], c["box_vectors"], 0.0, c["cutoff"], c["loop"],
                                  c["include_transpose"], c["neighbors"].shape[1] + 7, strategy)
        assert torch.equal(nb, c["neighbors"]), (c["box"], c["cutoff"], c["loop"], c["include_transpose"])
        assert torch.allclose(vec, c["deltas"], atol=1e-5) and torch.allclose(dist, c["distances"], atol=1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("boxkind", ["none", "cubic", "triclinic_per_batch"])
@pytest.mark.parametrize("lo", [0.0, 1.5])
def test_bit_exact_vs_oracle(dtype, boxkind, lo):
    z, pos, batch, box = systems.random_box(160, seed=11, n_boxes=3, dtype=dtype)
    if boxkind == "none":
        box = None
    elif boxkind == "cubic":
        box = box[0]
    else:
        box = box.clone()
        box[1] = torch.tensor([[12.0, 0, 0], [1.0, 11.5, 0], [-2.0, 1.5, 12.5]], dtype=dtype)
        box[2] = torch.tensor([[13.0, 0, 0], [0.0, 14.0, 0], [0.0, 0.0, 11.7]], dtype=dtype)
    for loop in (True, False):
        for tr in (True, False):
            ei, v, d = orc.neighbor_list(pos, batch, box, lo, 5.0, loop, tr)
            nb, vec, dist, n = run_op(pos, batch, box, lo, 5.0, loop, tr, ei.shape[1] + 3)
            assert torch.equal(nb, ei)
            assert torch.equal(vec, v) and torch.equal(dist, d)  # bit-exact, not allclose


def test_unsorted_batch_and_empty():
    g = torch.Generator().manual_seed(3)
    pos = torch.rand(97, 3, generator=g) * 6
    batch = torch.randint(0, 4, (97,), generator=g)  # NOT sorted
    ei, v, d = orc.neighbor_list(pos, batch, None, 0.0, 3.0, True, True)
    nb, vec, dist, n = run_op(pos, batch, None, 0.0, 3.0, True, True, ei.shape[1] + 1)
    assert torch.equal(nb, ei) and torch.equal(vec, v)
    nb, vec, dist, n = run_op(torch.zeros(0, 3), torch.zeros(0, dtype=torch.long), None, 0.0, 3.0, True, True, 5)
    assert n == 0


def test_distance_module_semantics():
    """reference tests/test_model_utils.py:13-88."""
    from torchmdnet_b200.models.utils import OptimizedDistance
    for lo in (0.0, 2.0):
        for hi in (5.0, 10.0):
            for loop in (False, True):
                dist = OptimizedDistance(lo, hi, max_num_pairs=-100, return_vecs=True, loop=loop).to(dev)
                batch = torch.tensor([0, 0], device=dev)
                extra = 2 if loop else 0
                pos = torch.tensor([[0, 0, 0], [(lo + hi) / 2, 0, 0]], dtype=torch.float, device=dev)
                ei, ew, ev = dist(pos, batch)
                assert ei.size(1) == 2 + extra and ei.dtype == torch.long and ev is not None
                if lo > 0:
                    ei, _, _ = dist(torch.tensor([[0, 0, 0], [lo / 2, 0, 0]], dtype=torch.float, device=dev), batch)
                    assert ei.size(1) == extra
                ei, _, _ = dist(torch.tensor([[0, 0, 0], [hi + 1, 0, 0]], dtype=torch.float, device=dev), batch)
                assert ei.size(1) == extra


def test_overflow_error_isolated():
    """overflow -> RuntimeError through the async device assert (reference test_neighbor_count_error);
    a device-side assert poisons the CUDA context, so this runs in a subprocess."""
    import subprocess, sys
    code = ("import torch,sys;sys.path.insert(0,'.');from torchmdnet_b200.models.utils import OptimizedDistance;"
            "d=OptimizedDistance(0.0,5.0,max_num_pairs=-2,loop=True).cuda();"
            "\ntry:\n d(torch.rand(100,3,device='cuda'),torch.zeros(100,dtype=torch.long,device='cuda'));torch.cuda.synchronize();print('NOERR')"
            "\nexcept RuntimeError as e:\n print('RAISED')")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(GOLDEN)))
    assert "RAISED" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_position_gradients(dtype):
    """vec/dist gradients wrt positions (reference tests/test_neighbors.py:151-317) for both API faces."""
    from torchmdnet_b200.extensions.ops import get_neighbor_pairs_kernel
    from torchmdnet_b200.extensions.neighbors import neighbor_graph
    z, pos, batch, box = systems.random_box(80, seed=5, n_boxes=2, dtype=dtype)
    box = box[0]
    hp = dict(cutoff_lower=0.0, cutoff_upper=4.0)
    pc = pos.clone().requires_grad_(True)
    ei, vec, dist = orc.neighbor_list_from_pos(pc, batch, box, hp, loop=True)
    g = torch.Generator().manual_seed(0)
    wv, wd = torch.randn(vec.shape, generator=g, dtype=dtype), torch.randn(dist.shape, generator=g, dtype=dtype)
    ((vec * wv).sum() + (dist * wd).sum()).backward()
    tol = 1e-4 if dtype == torch.float32 else 1e-10
    E = ei.shape[1]
    pg = pos.to(dev).requires_grad_(True)
    nb, v2, d2, _ = get_neighbor_pairs_kernel("brute", pg, batch.to(dev), box.to(dev), True, 0.0, 4.0, E + 9, True, True, 0)
    ((v2[:E] * wv.to(dev)).sum() + (d2[:E] * wd.to(dev)).sum()).backward()
    assert torch.allclose(pg.grad.cpu(), pc.grad, atol=tol, rtol=tol)
    pg2 = pos.to(dev).requires_grad_(True)
    gr = neighbor_graph(pg2, batch.to(dev), box.to(dev), True, 0.0, 4.0, True, E + 9)
    ((gr.vec[:E] * wv.to(dev)).sum() + (gr.dist[:E] * wd.to(dev)).sum()).backward()
    assert torch.allclose(pg2.grad.cpu(), pc.grad, atol=tol, rtol=tol)
    # structure: rev is an involution and points at the transposed pair
    rev = gr.rev[:E].long()
    assert torch.equal(rev[rev], torch.arange(E, device=dev))
    assert torch.equal(gr.nbr_i[:E][rev], gr.nbr_j[:E]) and torch.equal(gr.vec[:E][rev], -gr.vec[:E])
    assert torch.equal(gr.indptr.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(ei[0], minlength=160).cumsum(0)]))


def test_large_and_many_batches():
    """3200 atoms (tests/test_neighbors.py:320-375) and 128 batches; exactness via set properties at full size."""
    z, pos, batch, box = systems.random_box(3200, seed=1)
    ei, v, d = orc.neighbor_list(pos, batch, box, 0.0, 5.0, True, True)
    nb, vec, dist, n = run_op(pos, batch, box, 0.0, 5.0, True, True, 3200 * 64)
    assert torch.equal(nb, ei) and torch.equal(dist, d)
    z, pos, batch, _ = systems.molecule_batch(n_mol=128, seed=2)
    ei, v, d = orc.neighbor_list(pos, batch, None, 0.0, 5.0, False, True)
    nb, vec, dist, n = run_op(pos, batch, None, 0.0, 5.0, False, True, ei.shape[1])
    assert torch.equal(nb, ei) and torch.equal(vec, v)


def test_per_box_cell_grids():
    """several rectangular periodic boxes of DIFFERENT sizes in one batch (per-sample box matrices): every box gets its own cell
    grid (csrc/neighbors.cu; the reference's cell kernel handles one box only, warp_ops/neighbors.py:222-229); indices, deltas and
    distances bit-exact vs the oracle, and the grid path was really taken (order flag)"""
    from torchmdnet_b200.extensions.neighbors import neighbor_graph
    g0 = torch.Generator().manual_seed(11)
    sizes, dims = [700, 500, 900], [(19.0, 17.5, 21.0), (16.0, 18.2, 17.1), (23.0, 20.0, 19.5)]
    pos, batch, boxes = [], [], []
    for b, (n, L) in enumerate(zip(sizes, dims)):
        Lt = torch.tensor(L)
        pos.append((torch.rand(n, 3, generator=g0) * 1.4 - 0.2) * Lt)      # some atoms outside the primary cell
        batch.append(torch.full((n,), b, dtype=torch.long))
        boxes.append(torch.diag(Lt))
    pos, batch, box = torch.cat(pos), torch.cat(batch), torch.stack(boxes)
    ei, v, d = orc.neighbor_list(pos, batch, box, 0.0, 5.0, True, True)
    nb, vec, dist, n = run_op(pos, batch, box, 0.0, 5.0, True, True, ei.shape[1] + 7)
    assert n == ei.shape[1]
    assert torch.equal(nb, ei) and torch.equal(vec, v) and torch.equal(dist, d)
    g = neighbor_graph(pos.to(dev), batch.to(dev), box.to(dev), True, 0.0, 5.0, True, ei.shape[1] + 7)
    assert int(g.order_flag) != 0, "the per-box cell grids were not used"
    order = g.order.cpu().long()
    assert torch.equal(torch.sort(order).values, torch.arange(pos.shape[0]))       # a permutation: atoms listed cell by cell
    assert torch.equal(batch[order], torch.sort(batch).values)                      # box-major
    # one triclinic box among them: the whole call falls back to the ordered scan, same result
    box2 = box.clone(); box2[1, 1, 0] = 2.0
    ei2, v2, d2 = orc.neighbor_list(pos, batch, box2, 0.0, 5.0, True, True)
    nb2, vec2, dist2, n2 = run_op(pos, batch, box2, 0.0, 5.0, True, True, ei2.shape[1] + 7)
    assert torch.equal(nb2, ei2) and torch.equal(dist2, d2)
    g2 = neighbor_graph(pos.to(dev), batch.to(dev), box2.to(dev), True, 0.0, 5.0, True, ei2.shape[1] + 7)
    assert int(g2.order_flag) == 0


def test_graph_transform():
    from torchmdnet_b200.extensions.tensornet_ops import graph_transform
    g = torch.Generator().manual_seed(0)
    n, E = 50, 700
    ei = torch.randint(0, n, (2, E), generator=g, dtype=torch.int32)
    ei[:, 650:] = -1  # static-shape padding
    rd, ri, rp, cd, ci, cp = [t.cpu().long() for t in graph_transform(ei.to(dev), n)]
    valid = ei[:, :650].long()
    assert torch.equal(rp, torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(valid[0], minlength=n).cumsum(0)]))
    assert torch.equal(cp, torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(valid[1], minlength=n).cumsum(0)]))
    for r in range(n):
        ids = rd[rp[r]:rp[r + 1]]
        assert torch.equal(ids, torch.nonzero(valid[0] == r).flatten()) and torch.equal(ri[rp[r]:rp[r + 1]], valid[1][ids])
        ids = cd[cp[r]:cp[r + 1]]
        assert torch.equal(ids, torch.nonzero(valid[1] == r).flatten()) and torch.equal(ci[cp[r]:cp[r + 1]], valid[0][ids])


def test_cuda_graph_capture():
    """reference tests/test_neighbors.py:451-605: the op must be capturable (no host sync inside)."""
    from torchmdnet_b200.extensions.neighbors import neighbor_graph
    z, pos, batch, box = systems.random_box(300, seed=2)
    p, b, bx = pos.to(dev), batch.to(dev), box.to(dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            neighbor_graph(p, b, bx, True, 0.0, 5.0, True, 300 * 64)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            g = neighbor_graph(p, b, bx, True, 0.0, 5.0, True, 300 * 64)
    p.copy_(torch.roll(pos, 1, 0).to(dev))
    graph.replay()
    torch.cuda.synchronize()
    ei, v, d = orc.neighbor_list(torch.roll(pos, 1, 0), batch, box, 0.0, 5.0, True, True)
    n = int(g.num_pairs.item())
    assert n == ei.shape[1] and torch.equal(g.nbr_j[:n].cpu().long(), ei[1]) and torch.equal(g.dist[:n].cpu(), d)
