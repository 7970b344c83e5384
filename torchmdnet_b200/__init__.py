"""torchmdnet_b200 -- B200-native (sm_100a) TensorNet / Equivariant-Transformer energy+force path.

Drop-in for the hot path of torchmd/torchmd-net: `torchmdnet_b200.models.model.{create_model, load_model,
TorchMD_Net}` mirror `torchmdnet.models.model`, `torchmdnet_b200.extensions` mirrors `torchmdnet.extensions`.
All device work goes through the C-ABI library `libtmdnet_b200.so` (include/tmdnet_b200.h); there is no CPU path.
"""
__version__ = "0.1.0"
