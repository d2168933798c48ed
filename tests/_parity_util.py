"""Helpers shared by the GPU parity tests: tolerances, error metrics, and the log of measured errors.

Tolerances (``north_star``): fp32 mode rel-err <= 1e-4, bf16 mode rel-err <= 2e-2 on the outputs.  For log-probability
outputs the raw figure is flattered by the common -log(A) ~ -8 offset of every entry, so the tests also bound the
CENTRED error (row mean removed: what an argmax sees).  Its bounds are the largest centred error measured on the B200
for that mode (``profiles/r02_parity_measured.json``) with 2x headroom.
"""
import json
import os

import torch

from oracle import oracle as O

DEV = "cuda:0"
OUT_TOL = {"fp32": 1e-4, "bf16": 2e-2}
GRAD_TOL = {"fp32": 2e-3, "bf16": 1e-1}
# Centred log-prob / logit error.  Measured on the B200 (profiles/r02_parity_measured.json): fp32 mode 4.4e-6 .. 6.4e-5
# over all configurations, bf16 mode 1.3e-3 .. 8.4e-3 (HieCoAtten) / <= 5.0e-3 (MFB, MHBCoAtt).  The fp32 bound is the
# north_star tolerance itself (1.6x the largest measurement); the bf16 bound is the largest measurement x 2, which is
# still inside north_star's 2e-2.
CENTRED_TOL = {"fp32": 1e-4, "bf16": 1.7e-2}
# HieCoAtten's parameter gradients in bf16 mode: five chained GEMM stages with bf16-rounded operands, each followed by a
# cancelling reduction (softmax Jacobians over 196 regions / 26 tokens, bias sums over mixed-sign rows); measured worst
# case 1.01e-1 (fc_Wv.bias, Xavier init) / 9.0e-2 (default init), 1.6e-4 in fp32 mode.  Bound = measurement x 1.5.
HIE_GRAD_TOL = {"fp32": 2e-3, "bf16": 1.5e-1}

_MEASURED = {}
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(test: str, key: str, value) -> None:
    """Remember a measured error; the session hook in conftest.py writes them to gpurun_out/parity_measured.json."""
    _MEASURED.setdefault(test, {})[key] = float(value)


def dump_measured() -> None:
    if not _MEASURED:
        return
    out = os.path.join(_ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_measured.json")
        old = {}
        if os.path.isfile(path):
            try:
                old = json.load(open(path))
            except Exception:
                old = {}
        old.update(_MEASURED)
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def centred(t):
    return t - t.mean(dim=1, keepdim=True)


def xavier_(model, seed=0):
    """train_models.py:54-56: xavier_uniform_ on every parameter whose name lacks 'bias'."""
    torch.manual_seed(seed)
    for n, p in model.named_parameters():
        if n.find("bias") == -1:
            torch.nn.init.xavier_uniform_(p)
    return model


def z_from_capture(capture, N):
    """z = sign(y) y^2 at the values the kernels stored (y1 is [N*L, 1000] -> [N, L, 1000])."""
    out = {}
    for key, y in capture.items():
        y = y.detach().double()
        z = torch.sign(y) * y * y
        out["z" + key[1:]] = z.reshape(N, -1, z.shape[-1]) if key == "y1" else z
    return out


def check_grads(model, ref_grads, tol, loose=None, tag=None):
    """Per-parameter relative L2 error of model.<p>.grad against ref_grads[name]; returns the worst one."""
    worst, worst_name = 0.0, ""
    loose = loose or {}
    for name, p in model.named_parameters():
        ref = ref_grads[name]
        got = p.grad
        if ref is None or float(ref.abs().max()) == 0.0:
            assert got is None or float(got.abs().max()) == 0.0, name + " must have an exactly-zero gradient"
            continue
        assert got is not None, name
        if float(ref.norm()) < 1e-9:        # shift-invariant biases in front of a softmax: rounding noise only
            assert float(got.norm()) < 1e-4, name
            continue
        e = O.rel_err(got, ref)
        if e > worst:
            worst, worst_name = e, name
        if tag is not None:
            record(tag, "grad:" + name, e)
        assert e < max(tol, loose.get(name, 0.0)), (name, e)
    return worst, worst_name


def mhb_masks(ops, used, N, L, p=0.1, dev=DEV):
    """The pre-scaled dropout masks MHBCoAtt's fused epilogues applied, from the seeds they drew (in call order): the grid
    MFB's [N, L, 5000] mask and the two vector blocks' [N, 5000] masks.  The fp32 path draws one seed per vector block;
    the bf16 path runs both vector blocks as one [N, 10000] launch with ONE seed (fused_block.MhbFusedBlockFn)."""
    m1 = ops.dropout_mask(N * L, 5000, p, used[0], dev).double().reshape(N, L, 5000)
    if len(used) == 2:
        m23 = ops.dropout_mask(N, 10000, p, used[1], dev).double()
        return {"m1": m1, "m2": m23[:, :5000].contiguous(), "m3": m23[:, 5000:].contiguous()}
    assert len(used) == 3, used
    return {"m1": m1, "m2": ops.dropout_mask(N, 5000, p, used[1], dev).double(),
            "m3": ops.dropout_mask(N, 5000, p, used[2], dev).double()}
