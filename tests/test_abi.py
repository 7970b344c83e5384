"""CPU checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol the header
declares; the Python operator surface mirrors the reference's names; CPU tensors are refused loudly."""
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from torchmdnet_b200 import _lib
    return _lib.library()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "tmdnet_b200.h")).read()
    return sorted(set(re.findall(r"\b(tmd_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    from torchmdnet_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tmdnet_b200.h but not exported"
    # every bound symbol is declared in the header (no private entry points)
    for s in _lib.SIGNATURES:
        assert s in syms, f"{s} bound in _lib.py but missing from the header"
    assert lib.tmd_abi_version() == 1


def test_ctypes_signatures_match_header_prototypes():
    """every entry of _lib.SIGNATURES declares exactly the argument kinds of its prototype in include/tmdnet_b200.h (pointer /
    int / double / long long, the trailing stream included).  An entry that is one argument short still *works* through ctypes:
    the extra value is converted with the default rule (a 32-bit int), which truncates a 64-bit stream handle -- harmless on the
    default stream (0), a wild pointer inside cudaLaunchKernel on any other one."""
    from torchmdnet_b200 import _lib
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "tmdnet_b200.h")).read(), flags=re.S)
    kind = {_lib.P: "P", _lib.I: "I", _lib.D: "D", _lib.LL: "L"}
    for name, (argtypes, _) in _lib.SIGNATURES.items():
        m = re.search(r"\bint\s+" + name + r"\s*\(([^;]*?)\)\s*;", txt, flags=re.S)
        assert m, f"{name}: no prototype in the header"
        params = [q.strip() for q in m.group(1).split(",") if q.strip() and q.strip() != "void"]
        want = ["P" if "*" in q else "D" if q.startswith("double") else "L" if q.startswith("long long") else "I" for q in params]
        got = [kind[a] for a in argtypes]
        assert got == want, f"{name}: _lib.py declares {''.join(got)} ({len(got)}), the header {''.join(want)} ({len(want)})"


def test_operator_surface_matches_reference():
    from torchmdnet_b200.extensions import ops, tensornet_ops
    sig = inspect.signature(ops.get_neighbor_pairs_kernel)
    assert list(sig.parameters) == ["strategy", "positions", "batch", "box_vectors", "use_periodic", "cutoff_lower",
                                    "cutoff_upper", "max_num_pairs", "loop", "include_transpose", "num_cells"]
    for name in ["graph_transform", "fn_radial_message_passing", "fn_message_passing", "fn_compose_tensor",
                 "fn_decompose_tensor", "fn_tensor_matmul_o3_3x3", "fn_tensor_matmul_so3_3x3", "fn_tensor_norm3"]:
        assert hasattr(tensornet_ops, name)
    from torchmdnet_b200.models import model
    for name in ["create_model", "load_model", "load_ensemble", "TorchMD_Net", "Ensemble", "create_prior_models"]:
        assert hasattr(model, name)
    fwd = inspect.signature(model.TorchMD_Net.forward)
    assert list(fwd.parameters) == ["self", "z", "pos", "batch", "box", "q", "s", "extra_args", "num_systems"]


def test_op_registry_names():
    """the reference's torch.library names are registered by extensions/registry.py (CUDA only)."""
    import torch
    from torchmdnet_b200.extensions import registry
    want = ["torchmdnet::warp_neighbor_brute_fwd", "torchmdnet::warp_neighbor_cell_fwd", "nvtnet::count_row_col_primitive",
            "nvtnet::convert_to_sparse_primitive"]
    for op in ["compose_tensor", "decompose_tensor", "tensor_matmul_o3_3x3", "tensor_matmul_so3_3x3", "tensor_norm3",
               "message_passing", "radial_message_passing"]:
        want += [f"tensornet::{op}_fwd_primitive", f"tensornet::{op}_bwd_primitive", f"tensornet::{op}_bwd_bwd_primitive"]   # all 25
    assert sorted(want) == sorted(registry.REGISTERED)
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.tensornet.tensor_norm3_fwd_primitive(torch.rand(2, 3, 3, 8))  # no CPU kernel registered


def test_no_cpu_fallback(lib):
    from torchmdnet_b200.extensions.ops import get_neighbor_pairs_kernel
    from torchmdnet_b200.extensions import tensornet_ops as t
    pos = torch.rand(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        get_neighbor_pairs_kernel("brute", pos, torch.zeros(4, dtype=torch.long), torch.zeros(3, 3), False, 0.0, 5.0, 16,
                                  True, True, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        t.fn_tensor_norm3(torch.rand(2, 3, 3, 8))
    with pytest.raises(ValueError):
        get_neighbor_pairs_kernel("octree", pos, torch.zeros(4, dtype=torch.long), torch.zeros(3, 3), False, 0.0, 5.0, 16,
                                  True, True, 0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "torchmdnet_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "tensornet_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f
