"""Seeded synthetic systems shared by tests/, bench.py and oracle/make_golden.py
(SURVEY.md section 8d).  TEST/BENCH INFRASTRUCTURE -- not imported by the product.

Everything is generated on CPU with torch.Generator so that the same tensors are
obtained in the build container (where the goldens are made) and on the GPU box.
"""
import math
import torch


def tensornet_hparams(num_layers=2, F=128, R=32, cutoff=5.0, max_nb=128, group="O(3)", static_shapes=False,
                      derivative=True, prior=None, **kw):
    hp = dict(model="tensornet", embedding_dimension=F, num_layers=num_layers, num_rbf=R, rbf_type="expnorm",
              trainable_rbf=False, activation="silu", cutoff_lower=0.0, cutoff_upper=cutoff, max_z=128,
              max_num_neighbors=max_nb, equivariance_invariance_group=group, static_shapes=static_shapes,
              box_vecs=None, derivative=derivative, atom_filter=-1, prior_model=prior, output_model="Scalar",
              reduce_op="add", precision=32, aggr="add", neighbor_embedding=True, num_heads=8,
              distance_influence="both", attn_activation="silu", vector_cutoff=False)
    hp.update(kw)
    return hp


def et_hparams(num_layers=5, F=128, R=64, H=8, cutoff=10.0, max_nb=128, vector_cutoff=True, derivative=True, **kw):
    """examples/ET-SPICE.yaml hyper-parameters."""
    hp = dict(model="equivariant-transformer", embedding_dimension=F, num_layers=num_layers, num_rbf=R,
              rbf_type="expnorm", trainable_rbf=False, activation="silu", attn_activation="silu",
              neighbor_embedding=True, num_heads=H, distance_influence="both", cutoff_lower=0.0,
              cutoff_upper=cutoff, max_z=100, max_num_neighbors=max_nb, box_vecs=None, vector_cutoff=vector_cutoff,
              derivative=derivative, atom_filter=-1, prior_model=None, output_model="Scalar", reduce_op="add",
              precision=32, aggr="add", static_shapes=False, equivariance_invariance_group="O(3)")
    hp.update(kw)
    return hp


def random_box(n_atoms, density=0.1, seed=0, n_boxes=1, dtype=torch.float32):
    """C2-style: uniform random positions in a cubic periodic box at `density` atoms/A^3."""
    g = torch.Generator().manual_seed(seed)
    L = (n_atoms / density) ** (1.0 / 3.0)
    pos = torch.rand(n_boxes * n_atoms, 3, generator=g, dtype=torch.float64) * L
    z = torch.randint(1, 9, (n_boxes * n_atoms,), generator=g)
    batch = torch.arange(n_boxes).repeat_interleave(n_atoms)
    box = torch.eye(3, dtype=torch.float64) * L
    if n_boxes > 1:
        box = box.unsqueeze(0).repeat(n_boxes, 1, 1)
    return z, pos.to(dtype), batch, box.to(dtype)


def water_box(n_side=15, L=46.42, seed=0, n_boxes=1, dtype=torch.float32):
    """C4: n_side^3 H2O on a jittered lattice (n_side=15 -> 3375 molecules, 10125 atoms,
    rho = 0.101 atoms/A^3), O-H 0.96 A, H-O-H 104.5 deg, random orientation."""
    g = torch.Generator().manual_seed(seed)
    zs, ps, bs = [], [], []
    for b in range(n_boxes):
        a = L / n_side
        grid = torch.stack(torch.meshgrid(*[torch.arange(n_side, dtype=torch.float64)] * 3, indexing="ij"), -1).reshape(-1, 3)
        O = (grid + 0.5) * a + (torch.rand(grid.shape, generator=g, dtype=torch.float64) - 0.5) * 0.6 * a
        m = O.shape[0]
        # random orthonormal frames
        u = torch.randn(m, 3, generator=g, dtype=torch.float64); u = u / u.norm(dim=1, keepdim=True)
        w = torch.randn(m, 3, generator=g, dtype=torch.float64)
        w = w - (w * u).sum(1, keepdim=True) * u; w = w / w.norm(dim=1, keepdim=True)
        th = math.radians(104.5) / 2
        h1 = O + 0.96 * (math.cos(th) * u + math.sin(th) * w)
        h2 = O + 0.96 * (math.cos(th) * u - math.sin(th) * w)
        p = torch.stack([O, h1, h2], 1).reshape(-1, 3)
        zs.append(torch.tensor([8, 1, 1]).repeat(m)); ps.append(p)
        bs.append(torch.full((3 * m,), b, dtype=torch.long))
    box = torch.eye(3, dtype=torch.float64) * L
    if n_boxes > 1:
        box = box.unsqueeze(0).repeat(n_boxes, 1, 1)
    return torch.cat(zs), torch.cat(ps).to(dtype), torch.cat(bs), box.to(dtype)


def molecule_batch(n_mol=256, lo=20, hi=40, spread=2.5, seed=0, dtype=torch.float32):
    """C3: n_mol molecules with U[lo,hi] atoms, positions randn*spread, z in {1,6,7,8}."""
    g = torch.Generator().manual_seed(seed)
    sizes = torch.randint(lo, hi + 1, (n_mol,), generator=g)
    n = int(sizes.sum())
    pos = torch.randn(n, 3, generator=g, dtype=torch.float64) * spread
    z = torch.tensor([1, 6, 7, 8])[torch.randint(0, 4, (n,), generator=g)]
    batch = torch.arange(n_mol).repeat_interleave(sizes)
    return z, pos.to(dtype), batch, None
