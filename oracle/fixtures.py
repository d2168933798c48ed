"""TEST INFRASTRUCTURE ONLY -- deterministic parameter / input recipes shared by
``oracle/gen_golden.py`` (which runs the real reference) and the tests (which rebuild the very
same tensors from the seed stored in each fixture, guarded by a checksum).

Keeping tensors out of the fixtures keeps ``tests/golden/`` small: a fixture holds the case
description (shapes, seeds), a checksum of the regenerated tensors, and the reference's
outputs / gradients.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, Mapping, Sequence

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_params(shapes: Mapping[str, Sequence[int]], seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Xavier-uniform-like weights and small non-zero biases, one generator per name so that the
    result does not depend on dict order."""
    out = {}
    for i, name in enumerate(sorted(shapes)):
        shape = tuple(shapes[name])
        g = torch.Generator().manual_seed(seed * 1000003 + i)
        if len(shape) >= 2:
            rf = 1
            for s in shape[2:]:
                rf *= s
            fan_in, fan_out = shape[1] * rf, shape[0] * rf
            a = math.sqrt(6.0 / (fan_in + fan_out))
        else:
            a = 0.1
        out[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * a).to(dtype)
    return out


def make_inputs(case: Mapping, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """relu(N(0,1)) features and uniform token ids."""
    g = torch.Generator().manual_seed(int(case["input_seed"]))
    out = {}
    for name, spec in case["inputs"].items():
        kind, shape = spec[0], tuple(spec[1])
        if kind == "relu_randn":
            out[name] = torch.relu(torch.randn(shape, generator=g, dtype=torch.float64)).to(dtype)
        elif kind == "randn":
            out[name] = torch.randn(shape, generator=g, dtype=torch.float64).to(dtype)
        elif kind == "randint":
            out[name] = torch.randint(0, int(spec[2]), shape, generator=g)
        elif kind == "mask":          # pre-scaled dropout mask, keep-prob 1-p
            p = float(spec[2])
            out[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) >= p).to(dtype) / (1.0 - p))
        else:
            raise ValueError(kind)
    return out


def checksum(tensors: Mapping[str, torch.Tensor]) -> float:
    s = 0.0
    for i, k in enumerate(sorted(tensors)):
        t = tensors[k].double().reshape(-1)
        w = torch.arange(1, t.numel() + 1, dtype=torch.float64) % 97 + 1
        s += float((t * w).sum()) * (i + 1)
    return s


def subsample(t: torch.Tensor, limit: int = 4096):
    """Large gradients are stored as a strided subsample plus their norm."""
    flat = t.detach().double().reshape(-1)
    stride = max(1, flat.numel() // limit)
    return {"stride": stride, "values": flat[::stride].float().clone(), "norm": float(flat.norm()),
            "numel": flat.numel()}


def compare_subsample(t: torch.Tensor, rec, floor: float = 1e-6) -> float:
    """Relative error of ``t`` against a stored subsample (max of value-based and norm-based).
    ``floor`` is the absolute norm below which a gradient counts as numerically zero (biases in
    front of a softmax get pure rounding noise in the reference)."""
    flat = t.detach().double().reshape(-1)
    assert flat.numel() == rec["numel"], (flat.numel(), rec["numel"])
    v = flat[:: rec["stride"]]
    ref = rec["values"].double()
    e1 = float((v - ref).norm()) / max(float(ref.norm()), floor)
    e2 = abs(float(flat.norm()) - rec["norm"]) / max(rec["norm"], floor)
    return max(e1, e2)


def save_fixture(name: str, obj) -> str:
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".pt")
    torch.save(obj, path)
    return path


def load_fixture(name: str):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), map_location="cpu", weights_only=False)


def list_fixtures():
    if not os.path.isdir(GOLDEN_DIR):
        return []
    return sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.endswith(".pt"))
