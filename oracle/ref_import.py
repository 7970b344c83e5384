"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the *unmodified* reference (torchmd/torchmd-net, mounted read-only at
/root/reference in the build container) on CPU so that golden vectors can be
generated from the reference's own code.

The reference cannot be imported as-is here because two third-party packages
are absent: `warp` (NVIDIA Warp, used only by torchmdnet/extensions) and
`lightning_utilities`.  We inject two stubs:

  * lightning_utilities.core.rank_zero.rank_zero_warn  -> warnings.warn
  * torchmdnet.extensions.ops.get_neighbor_pairs_kernel -> a differentiable
    pure-torch O(N^2) neighbour list with the semantics of
    torchmdnet/extensions/warp_kernels/neighbors_brute.py:98-143 (pairs i>j,
    delta = pos_i - pos_j, sequential z,y,x minimum-image rounding, transpose
    negated, self loops last, -1 padding).

With `warp` absent the reference's TensorNet falls back to its pure-PyTorch
ops (torchmdnet/models/tensornet.py:67-81, OPT=False), which is the arbiter
convention for this project (SURVEY.md section 7 "Sign convention trap").

This module only works where /root/reference exists (the build container);
the GPU box never sees it -- tests there use the committed fixtures in
tests/golden/.
"""
import os
import sys
import types
import warnings

import torch

REFERENCE_ROOT = os.environ.get("TMD_REFERENCE_ROOT", "/root/reference")


def _torch_neighbor_pairs(strategy, positions, batch, box_vectors, use_periodic,
                          cutoff_lower, cutoff_upper, max_num_pairs, loop,
                          include_transpose, num_cells):
    n = positions.shape[0]
    dev, dt = positions.device, positions.dtype
    ii, jj = torch.tril_indices(n, n, -1, device=dev)
    same = batch[ii] == batch[jj]
    ii, jj = ii[same], jj[same]
    d = positions[ii] - positions[jj]
    if use_periodic:
        box = box_vectors.to(dt)
        if box.dim() == 2:
            box = box.unsqueeze(0).expand(int(batch.max()) + 1, 3, 3)
        b = box[batch[ii]]
        s3 = torch.round(d[:, 2] / b[:, 2, 2])
        d = d - s3[:, None] * b[:, 2]
        s2 = torch.round(d[:, 1] / b[:, 1, 1])
        d = d - s2[:, None] * b[:, 1]
        s1 = torch.round(d[:, 0] / b[:, 0, 0])
        d = d - s1[:, None] * b[:, 0]
    r2 = (d * d).sum(-1)
    keep = (r2 < cutoff_upper ** 2) & (r2 >= cutoff_lower ** 2)
    ii, jj, d = ii[keep], jj[keep], d[keep]
    r = d.norm(dim=-1)
    ei, ej, vec, dist = [ii], [jj], [d], [r]
    if include_transpose:
        ei.append(jj); ej.append(ii); vec.append(-d); dist.append(r)
    if loop:
        ar = torch.arange(n, device=dev)
        ei.append(ar); ej.append(ar)
        vec.append(torch.zeros(n, 3, device=dev, dtype=dt))
        dist.append(torch.zeros(n, device=dev, dtype=dt))
    ei, ej = torch.cat(ei), torch.cat(ej)
    vec, dist = torch.cat(vec), torch.cat(dist)
    npairs = ei.numel()
    neighbors = torch.full((2, max_num_pairs), -1, device=dev, dtype=torch.long)
    deltas = torch.zeros(max_num_pairs, 3, device=dev, dtype=dt)
    distances = torch.zeros(max_num_pairs, device=dev, dtype=dt)
    m = min(npairs, max_num_pairs)
    neighbors[0, :m], neighbors[1, :m] = ei[:m], ej[:m]
    deltas = torch.cat([vec[:m], deltas[m:]])
    distances = torch.cat([dist[:m], distances[m:]])
    num_pairs = torch.tensor([npairs], dtype=torch.int32, device=dev)
    return neighbors, deltas, distances, num_pairs


def import_reference():
    """Returns the reference's `torchmdnet` package (CPU, pure-PyTorch ops)."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "torchmdnet" in sys.modules and getattr(sys.modules["torchmdnet"], "_is_reference", False):
        return sys.modules["torchmdnet"]
    # the repo ships a `torchmdnet` alias package that resolves to the B200 build: never let it stand in for the reference
    for name in [m for m in sys.modules if m == "torchmdnet" or m.startswith("torchmdnet.")]:
        del sys.modules[name]
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:] = [q for q in sys.path if os.path.abspath(q or ".") != repo_root] + [repo_root]
    # stub 1: lightning_utilities
    lu = types.ModuleType("lightning_utilities")
    lu_core = types.ModuleType("lightning_utilities.core")
    lu_rz = types.ModuleType("lightning_utilities.core.rank_zero")
    lu_rz.rank_zero_warn = lambda *a, **k: warnings.warn(str(a[0]) if a else "")
    lu.core = lu_core
    lu_core.rank_zero = lu_rz
    sys.modules.setdefault("lightning_utilities", lu)
    sys.modules.setdefault("lightning_utilities.core", lu_core)
    sys.modules.setdefault("lightning_utilities.core.rank_zero", lu_rz)
    # stub 2: the extension seam (warp is absent)
    ext = types.ModuleType("torchmdnet.extensions")
    ext.__path__ = []  # mark as package; nothing else importable -> warp_ops ImportError -> OPT=False
    ops = types.ModuleType("torchmdnet.extensions.ops")
    ops.get_neighbor_pairs_kernel = _torch_neighbor_pairs
    ext.ops = ops
    sys.modules["torchmdnet.extensions"] = ext
    sys.modules["torchmdnet.extensions.ops"] = ops
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import torchmdnet  # noqa
        import torchmdnet.models.model  # noqa
    torchmdnet._is_reference = True
    return torchmdnet
