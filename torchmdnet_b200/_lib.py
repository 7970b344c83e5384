"""ctypes binding of the C-ABI shared library `libtmdnet_b200.so` (include/tmdnet_b200.h).

There is NO CPU fallback: if the library is missing, or a tensor is not on a CUDA device, the
call fails loudly.  Pointers are raw `tensor.data_ptr()` values, the stream is the caller's
current torch CUDA stream -- every entry point is an asynchronous launch (CUDA-graph capturable),
exactly like the reference's `wp.launch(..., stream=torch stream)`
(torchmdnet/extensions/warp_kernels/utils.py:84-88).
"""
import ctypes
import os
from ctypes import c_void_p, c_int, c_double, c_longlong, c_char_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtmdnet_b200.so")
_lib = None
_launches = 0  # kernels launched through the library (bench.py reports it as gpu_launches)

P, I, D, LL = c_void_p, c_int, c_double, c_longlong

# symbol -> (argtypes, number of CUDA kernels one call launches)
SIGNATURES = {
    "tmd_neighbor_csr_f32": ([P, P, P, I, I, I, D, D, I, I, I, I, I, P, P, P, P, P, P, P, P, P], 9),
    "tmd_neighbor_csr_f64": ([P, P, P, I, I, I, D, D, I, I, I, I, I, P, P, P, P, P, P, P, P, P], 9),
    "tmd_neighbor_grad_csr_f32": ([P, P, P, I, I, P, P, P, P, P, P], 1),
    "tmd_neighbor_grad_csr_f64": ([P, P, P, I, I, P, P, P, P, P, P], 1),
    "tmd_neighbor_grad_coo_f32": ([P, LL, P, P, P, P, P, P], 1),
    "tmd_neighbor_grad_coo_f64": ([P, LL, P, P, P, P, P, P], 1),
    "tmd_graph_transform": ([P, I, I, P, P, P, P, P, P, P, P], 6),
    "tmd_radial_mp_fwd": ([P, P, P, P, I, I, P, P, P, P], 1),
    "tmd_radial_mp_bwd": ([P, P, P, P, I, I, P, P, P, P, P, P], 1),
    "tmd_mp_fwd": ([P, P, P, P, P, P, P, I, I, P, P, P, P], 1),
    "tmd_mp_bwd": ([P, P, P, P, P, P, P, P, P, P, I, I, P, P, P, P, P], 1),
    "tmd_compose_fwd": ([P, P, P, P, I, I, P], 1),
    "tmd_compose_bwd": ([P, P, P, P, I, I, P], 1),
    "tmd_decompose_fwd": ([P, P, P, P, I, I, P], 1),
    "tmd_decompose_bwd": ([P, P, P, P, I, I, P], 1),
    "tmd_matmul_o3_fwd": ([P, P, P, I, I, P], 1),
    "tmd_matmul_o3_bwd": ([P, P, P, P, P, I, I, P], 1),
    "tmd_matmul_so3_fwd": ([P, P, P, I, I, P], 1),
    "tmd_matmul_so3_bwd": ([P, P, P, P, P, I, I, P], 1),
    "tmd_norm3_fwd": ([P, P, I, I, P], 1),
    "tmd_norm3_bwd": ([P, P, P, I, I, P], 1),
    "tmd_norm3c_fwd": ([P, P, P, P, I, I, P], 1),
    "tmd_norm3c_bwd": ([P, P, P, P, P, P, P, I, I, P], 1),
    "tmd_norm2c_fwd": ([P, P, P, P, I, I, P], 1),
    "tmd_norm2c_bwd": ([P, P, P, P, P, P, P, I, I, P], 1),
    "tmd_gate_fwd": ([P, P, P, P, P, P, P, I, I, P], 1),
    "tmd_gate_bwd": ([P] * 11 + [I, I, P], 1),
    "tmd_p1_fwd": ([P] * 6 + [I, I, P], 1),
    "tmd_p1_bwd": ([P] * 9 + [I, I, P], 1),
    "tmd_p2_fwd": ([P] * 7 + [I] + [P] * 3 + [I, I, P], 1),
    "tmd_p2_bwd": ([P] * 7 + [I] + [P] * 9 + [I, I, P], 1),
    "tmd_p3_fwd": ([P] * 7 + [P] * 3 + [I, I, P], 1),
    "tmd_p3_bwd": ([P] * 4 + [P] * 3 + [P] * 6 + [I, I, P], 1),
    "tmd_pair_index": ([P, P, P, P, P, I, I, P, P, P, P, P, P, P, P], 3),
    "tmd_pair_chain": ([P, P, I, P, P, P, P, P, P, P, I, I, I, P, P, D, D, D, P, P, P], 1),
    "tmd_pair_chain_tc": ([P, P, I, P, P, P, P, I, I, I, P, P, D, D, D, P, P, P, P, P, P, P, P, I, P], 3),
    "tmd_irreps_linear_tc": ([P, P, P, I, I, P, P, P, P, P, P], 1),
    "tmd_radial_table": ([P, P, I, P, I, D, I, P, P, P], 1),
    "tmd_pair_rbf_split": ([P, P, I, I, I, P, P, D, D, D, P, P, I, P], 1),
    "tmd_mp_pair_bwd": ([P] * 11 + [I, I] + [P] * 4 + [P], 1),
    "tmd_et_msg_fwd": ([P] * 10 + [D, D, I, I, I, I, P, P, P], 1),
    "tmd_et_msg_bwd": ([P] * 13 + [D, D, I, I, I, I] + [P] * 6 + [P], 2),
    "tmd_nbr_embed_fwd": ([P] * 6 + [I, I, P, P], 1),
    "tmd_nbr_embed_bwd": ([P] * 6 + [I, I, P, P, P], 1),
    "tmd_mp_edge_bwd": ([P] * 11 + [I, I] + [P] * 4 + [P], 1),
    "tmd_mp_group_size": ([], 0),
    "tmd_mp_pick_group": ([I], 0),
    "tmd_mp_plan": ([P, P, P, I, I, I, P, P, P, P, P, P, P, P, P, P], 1),
    "tmd_mp_coop_fwd": ([P] * 8 + [I, I, I, I] + [P] * 9 + [P], 1),
    "tmd_mp_coop_bwd": ([P] * 12 + [I, I, I, I] + [P] * 13 + [P], 1),
    "tmd_head_fwd": ([P, P, P, I, I, I, P, P, D] + [P] * 9 + [I] + [P] * 4 + [P], 1),
    "tmd_segment_sum": ([P, P, I, I, P, P, P, P], 2),
    "tmd_head_bwd": ([P, P, P, I, I, I, P, D] + [P] * 6 + [I] + [P] * 6 + [P], 1),
    "tmd_embed_fwd": ([P] * 9 + [D, D, I, I] + [P] * 3 + [P], 1),
    "tmd_embed_bwd": ([P] * 10 + [D, D, I, I] + [P] * 5 + [P], 1),
    "tmd_gate_bwd_preact": ([P] * 11 + [I, I, I, P], 1),
    "tmd_rows_gemm": ([P, I, P, I, I, I, I, I, P, P, I, P, P, I, P], 1),
    "tmd_norm_ln_fwd": ([P, P, P, I, I, I, P, P, D, P, P, P], 1),
    "tmd_norm_ln_bwd": ([P, P, P, I, I, I] + [P] * 8 + [P], 1),
    "tmd_head_energy": ([P, I, I, P, P, P, P, P, I, P, P], 1),
    "tmd_head_seed": ([P, P, I, P, P, P, I, I, P, P], 1),
    "tmd_geb_cat_fwd": ([P, P, I, I, P, P], 1),
    "tmd_geb_cat_bwd": ([P, P, P, I, I, P, P], 1),
    "tmd_geb_out_fwd": ([P, P, I, I, I, P, P, P], 1),
    "tmd_geb_out_bwd": ([P, P, P, I, P, I, I, I, P, P, P], 1),
    "tmd_rows_narrow_fwd": ([P, I, P, I, I, I, P, P, P], 1),
    "tmd_rows_narrow_bwd": ([P, P, I, I, I, I, P, P, P], 1),
    "tmd_p1_bwd2": ([P] * 12 + [I, I, P], 1),
    "tmd_p3_bwd_dx": ([P] * 10 + [I, I, P], 1),
    "tmd_embed_moments_fwd": ([P] * 10 + [D, D, D, I, I, I, I] + [P] * 3 + [P], 1),
    "tmd_embed_moments_bwd": ([P] * 11 + [D, D, D, I, I, I, I] + [P] * 5 + [P], 1),
}


def library():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a). torchmdnet_b200 has no CPU/eager fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.tmd_last_error.restype = c_char_p
        lib.tmd_abi_version.restype = c_int
        for name, (argtypes, _) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = library / header mismatch: fail loudly
            fn.argtypes = argtypes
            fn.restype = c_int
        _lib = lib
    return _lib


def launches():
    return _launches


def ptr(t):
    return None if t is None else t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("torchmdnet_b200 kernels need CUDA tensors (sm_100a); there is no CPU fallback")


def call(name, *args, device=None, launches=None):
    """Launch entry point `name` on the current stream of `device`; args before the stream.  `launches` overrides the
    per-call kernel count of SIGNATURES for entry points whose launch count depends on the arguments."""
    global _launches
    lib = library()
    stream = torch.cuda.current_stream(device).cuda_stream
    rc = getattr(lib, name)(*args, stream)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.tmd_last_error().decode()}")
    _launches += SIGNATURES[name][1] if launches is None else launches
