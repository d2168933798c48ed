"""Fused Adam for the drop-in modules (the optimizer step right behind the fusion block).

``FusedAdam(model.parameters(), lr)`` replaces ``optim.Adam(model.parameters(), lr)`` at solver.py:30 -- same update
rule and state (``exp_avg`` / ``exp_avg_sq`` / ``step``), one multi-tensor kernel launch per 32 parameter tensors
(``vqa_b200_adam_step``).  When the model's kernel-form weight caches are attached (``attach(model)``), the kernel also
writes the bf16 GEMM-operand copy of every updated weight, so the next forward does not re-cast the parameters.
CUDA fp32 parameters only; there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List

import torch

from . import _lib, ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("FusedAdam: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._caches: List[ops.WeightCache] = []
        self._dead = set()
        self.step_count = None          # device step counter (int64 [1]) once enable_device_step() was called

    # ---- device-side step count: the form a CUDA graph of the whole iteration needs (train.GraphedTrainStep)
    def enable_device_step(self, device) -> torch.Tensor:
        """Keep the step count on the device from now on: ``step()`` increments it with a (capturable) device op and the
        update kernel forms the bias corrections from it.  Every parameter must be at the same step count."""
        if self.step_count is None:
            steps = {int(st["step"]) for st in self.state.values() if "step" in st}
            if len(steps) > 1:
                raise RuntimeError("FusedAdam.enable_device_step: parameters are at different step counts %s" % sorted(steps))
            self.step_count = torch.full((1,), steps.pop() if steps else 0, dtype=torch.int64, device=device)
        return self.step_count

    def advance_step(self):
        """One increment per training iteration; ddp.GradientAllReducer.finish(optimizer) calls it before its per-bucket
        ``step(only=...)`` calls, a plain ``step()`` calls it itself."""
        if self.step_count is not None:
            self.step_count.add_(1)

    def sync_step_from_device(self):
        """Write the device step count back into the per-parameter state (checkpoints interchange with torch.optim.Adam)."""
        if self.step_count is not None:
            n = int(self.step_count.item())
            for st in self.state.values():
                if "step" in st:
                    st["step"] = n

    def attach(self, module: torch.nn.Module) -> "FusedAdam":
        """Register the kernel-form weight caches of `module` (and its sub-modules): their bf16 entries are refreshed
        in place by the update kernel."""
        for m in module.modules():
            c = getattr(m, "_wcache", None)
            if isinstance(c, ops.WeightCache) and all(c is not k for k in self._caches):
                self._caches.append(c)
            if hasattr(m, "dead_parameters"):
                # parameters the module declares to have exactly-zero gradients on every step (MFB's dead first stage):
                # Adam would leave them where they are (m = v = 0 => update 0), so they are not read, updated or re-cast
                self._dead.update(id(p) for p in m.dead_parameters())
        return self

    def _bf16_copy(self, p):
        for c in self._caches:
            t = c.bf16_entry(p)
            if t is not None and t.is_contiguous() and t.numel() == p.numel():
                return c, t
        return None, None

    @torch.no_grad()
    def step(self, closure=None, only=None):
        """One Adam update.  `only`: optional iterable of parameters -- update just those (each parameter keeps its
        own step count), which lets ddp.GradientAllReducer.finish() update every gradient bucket as soon as ITS
        all-reduce has landed while the later buckets are still on the wire."""
        only = None if only is None else set(only)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.load()
        dev_step = self.step_count
        if dev_step is not None and only is None:
            self.advance_step()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            # tensors that share a step count go into the same launches
            by_step = {}
            for p in group["params"]:
                if p.grad is None or (only is not None and p not in only) or id(p) in self._dead:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam handles contiguous fp32 CUDA parameters only (there is no CPU path)")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                # a state dict saved by torch.optim.Adam (solver.py:30) carries tensor steps: fold them to ints so that
                # parameters sharing a step count still share launches
                st["step"] = int(st["step"]) + 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                by_step.setdefault(0 if dev_step is not None else st["step"], []).append((p, g, st))
            for step, items in by_step.items():
                stream = ops._st()
                n = len(items)
                arr = ctypes.c_void_p * n
                copies = [self._bf16_copy(p) for p, _, _ in items]
                P = arr(*[p.data_ptr() for p, _, _ in items])
                G = arr(*[g.data_ptr() for _, g, _ in items])
                M = arr(*[st["exp_avg"].data_ptr() for _, _, st in items])
                V = arr(*[st["exp_avg_sq"].data_ptr() for _, _, st in items])
                B = arr(*[(t.data_ptr() if t is not None else None) for _, t in copies])
                numel = (ctypes.c_int64 * n)(*[p.numel() for p, _, _ in items])
                ops.LaunchStats.count += (n + 31) // 32
                timing = ops._capture is None and ops.LaunchStats.wants("adam_step")
                if timing:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                if dev_step is not None:
                    rc = L.vqa_b200_adam_step_dev(n, P, G, M, V, B, numel, float(group["lr"]), float(beta1), float(beta2),
                                                  float(group["eps"]), ctypes.c_void_p(dev_step.data_ptr()), stream)
                else:
                    rc = L.vqa_b200_adam_step(n, P, G, M, V, B, numel, float(group["lr"]), float(beta1), float(beta2),
                                              float(group["eps"]), int(step), stream)
                if timing:
                    e1.record()
                    ops.LaunchStats.events.setdefault("adam_step", []).append((e0, e1))
                _lib.check(rc, "vqa_b200_adam_step")
                for (p, _, _), (c, t) in zip(items, copies):
                    torch.autograd.graph.increment_version(p)       # p was written through its raw pointer
                    if c is not None:
                        c.refreshed(p)
        return loss
