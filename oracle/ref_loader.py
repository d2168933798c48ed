"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference modules.

Used by ``oracle/gen_golden.py`` (in the build container, where ``/root/reference``
is mounted) to produce the golden vectors under ``tests/golden/`` and by
``bench.py --impl reference`` / the ``cpu_baseline`` leg when a copy of the
reference travels to the GPU box under ``baseline/_ref``.  Nothing in the product
package (``vqa_attention_networks_b200``) may import this file.

The reference targets torch ~0.4 and needs exactly one shim on torch 2.x
(SURVEY.md "seven facts" #3): ``Tensor.view`` on a permuted tensor raises at
``mfb.py:105`` and ``mhb_coAtt.py:107``; falling back to ``reshape`` there is
value-identical because the flattened vector is only L2-normalised and viewed
back.  The shim is installed only while a reference ``forward`` runs.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types
import warnings

import torch

_CANDIDATES = ("/root/reference",
               os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref"))


def reference_root() -> str | None:
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "mhb_coAtt.py")):
            return c
    return None


def available() -> bool:
    return reference_root() is not None


def load(name: str) -> types.ModuleType:
    """Import one reference module (``mfb``, ``mhb_coAtt``, ``hieCoAtten``, ``modules``, ``networks``)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference sources not found (looked in %s)" % (_CANDIDATES,))
    key = "_vqa_ref_" + name
    if key in sys.modules:
        return sys.modules[key]
    sys.path.insert(0, root)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            # the reference's ``networks.py`` does ``from modules import ...``; keep that resolvable
            mod = importlib.import_module(name)
    finally:
        sys.path.remove(root)
    sys.modules[key] = mod
    return mod


@contextlib.contextmanager
def view_shim():
    """``Tensor.view`` -> ``reshape`` fallback for the two sites that need it on torch 2.x."""
    orig = torch.Tensor.view

    def view(self, *shape, **kw):
        try:
            return orig(self, *shape, **kw)
        except RuntimeError as e:  # "view size is not compatible with input tensor's size and stride"
            if "view size is not compatible" in str(e):
                return self.reshape(*shape)
            raise

    torch.Tensor.view = view
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            yield
    finally:
        torch.Tensor.view = orig


@contextlib.contextmanager
def injected_dropout(masks):
    """Replace ``F.dropout`` (the always-on functional dropouts of hieCoAtten.py:26-46) by
    multiplication with the next pre-scaled mask from ``masks`` (an iterator of tensors)."""
    import torch.nn.functional as F
    orig = F.dropout
    it = iter(masks)

    def dropout(x, p=0.5, training=True, inplace=False):
        m = next(it)
        assert m.shape == x.shape, (m.shape, x.shape)
        return x * m.to(x.dtype)

    F.dropout = dropout
    try:
        yield
    finally:
        F.dropout = orig


def xavier_init_(model: torch.nn.Module, seed: int = 0) -> None:
    """The reference's own init recipe (train_models.py:54-56): xavier_uniform_ on every
    parameter whose name lacks 'bias'."""
    torch.manual_seed(seed)
    for name, p in model.named_parameters():
        if name.find("bias") == -1:
            torch.nn.init.xavier_uniform_(p)
