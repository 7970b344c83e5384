"""`torchmdnet` resolved to the B200 build: the drop-in name (BASELINE.json north_star: "keeping the
torchmdnet.models.model.create_model/load_model and TorchMD_Net module API so it drops into torchmd-train and
TMDNETCalculator unchanged").

Put this directory's parent FIRST on sys.path / PYTHONPATH.  Then

    from torchmdnet.models.model import create_model, load_model, TorchMD_Net      # -> torchmdnet_b200.models.model
    from torchmdnet.extensions.ops import get_neighbor_pairs_kernel                # -> torchmdnet_b200.extensions.ops
    from torchmdnet.extensions.warp_ops import fn_message_passing, ...             # -> torchmdnet_b200.extensions.tensornet_ops
    from torchmdnet.calculators import External, TMDNETCalculator                  # -> torchmdnet_b200.calculators
    from torchmdnet import priors                                                  # -> torchmdnet_b200.priors

are the SAME module objects as their torchmdnet_b200 counterparts (aliased in sys.modules, no copies).  The hot path --
`models`, `extensions`, `calculators`, `priors` (reference torchmdnet/{models,extensions,calculators.py,priors}) -- is replaced;
everything else of the reference (the training harness `module`, `data`, `datasets`, `loss`, `utils`, `scripts`: out of scope,
SURVEY.md section 2) is NOT re-implemented: if an installed reference `torchmdnet` is found further down sys.path its directory
is appended to this package's `__path__`, so `torchmdnet.module.LNNP` / `torchmdnet.scripts.train` keep coming from the
reference and pick up the B200 `create_model` / `load_model` through their own `from torchmdnet.models.model import ...` lines
(module.py:13, scripts/train.py:17-23, calculators.py:6).
"""
import importlib
import importlib.util
import os
import pkgutil
import sys

import torchmdnet_b200 as _impl

__version__ = _impl.__version__
_is_b200_alias = True
_REPLACED = ("models", "extensions", "calculators", "priors")
_RENAMED = {"torchmdnet.extensions.warp_ops": "torchmdnet_b200.extensions.tensornet_ops"}


def _alias(alias: str, real: str):
    mod = importlib.import_module(real)
    sys.modules[alias] = mod
    parent, _, leaf = alias.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, mod) if parent != __name__ else globals().__setitem__(leaf, mod)
    return mod


def _install():
    for top in _REPLACED:
        real_top = f"{_impl.__name__}.{top}"
        mod = _alias(f"{__name__}.{top}", real_top)
        if hasattr(mod, "__path__"):
            for info in pkgutil.walk_packages(mod.__path__, real_top + "."):
                if info.name.rsplit(".", 1)[-1].startswith("_"):
                    continue
                _alias(__name__ + info.name[len(_impl.__name__):], info.name)
    for alias, real in _RENAMED.items():
        _alias(alias, real)
    # the rest of the reference package (training harness), if one is installed further down the path
    here = os.path.dirname(os.path.abspath(__file__))
    for entry in list(sys.path):
        cand = os.path.join(entry or ".", "torchmdnet")
        if os.path.isfile(os.path.join(cand, "__init__.py")) and os.path.abspath(cand) != here and cand not in __path__:
            __path__.append(cand)


_install()
