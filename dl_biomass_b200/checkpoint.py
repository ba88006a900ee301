"""Checkpoint interchange with the reference.

The reference pickles the whole ``DataParallel(Net)`` object (/root/reference/main.py:245, read back at
testing_model.py:30-37); that pickle needs torch_geometric to load.  What carries over is its ``state_dict``: the
B200 ``Net`` keeps PyG's attribute names, so the tensors map one to one once three spellings are normalised:

* ``module.`` -- the prefix ``torch_geometric.nn.DataParallel`` adds to every key (main.py:140);
* ``norms.<i>.module.<p>`` -- PyG >= 2.1 wraps BatchNorm1d in its own ``BatchNorm`` (``norms.<i>.<p>`` before);
* ``conv.local_nn`` / ``nn`` / ``mlp`` -- unchanged.

``reference_state_dict(net)`` writes the dictionary back in the reference's spelling so that a model trained here loads
into the reference's code (``model.module.load_state_dict`` there).
"""
from __future__ import annotations

import re
from collections import OrderedDict
from typing import Mapping

import torch

_NORM_MODULE = re.compile(r"(\.norms\.\d+)\.module\.")
_NORM_PLAIN = re.compile(r"(\.norms\.\d+)\.(weight|bias|running_mean|running_var|num_batches_tracked)$")


def normalise_state_dict(state_dict: Mapping[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Reference spelling -> this repository's spelling (idempotent)."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in state_dict.items():
        while k.startswith("module."):
            k = k[len("module."):]
        k = _NORM_MODULE.sub(r"\1.", "." + k)[1:]
        out[k] = v
    return out


def load_reference_state_dict(net: torch.nn.Module, state_dict: Mapping[str, torch.Tensor], strict: bool = True):
    """Load a ``state_dict`` saved from the reference's model (any of the spellings above) into ``net``."""
    return net.load_state_dict(normalise_state_dict(state_dict), strict=strict)


def reference_state_dict(net: torch.nn.Module, pyg_norm_wrapper: bool = True, data_parallel: bool = False
                         ) -> "OrderedDict[str, torch.Tensor]":
    """``net.state_dict()`` in the reference's spelling: ``pyg_norm_wrapper`` for PyG >= 2.1 (``norms.i.module.*``),
    ``data_parallel`` for a model that is loaded through the ``DataParallel`` wrapper (``module.`` prefix)."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in net.state_dict().items():
        if pyg_norm_wrapper:
            k = _NORM_PLAIN.sub(r"\1.module.\2", "." + k)[1:]
        if data_parallel:
            k = "module." + k
        out[k] = v
    return out
