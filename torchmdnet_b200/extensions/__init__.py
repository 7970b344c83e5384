"""B200-native replacement for `torchmdnet.extensions` (hand-written sm_100a kernels behind a C ABI)."""
from .ops import get_neighbor_pairs_kernel  # noqa: F401
from . import tensornet_ops as warp_ops  # noqa: F401  (same handle names as torchmdnet.extensions.warp_ops)
from .tensornet_ops import *  # noqa: F401,F403
