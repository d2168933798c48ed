"""Data-parallel gradient exchange: one process per GPU, parameters resident on every GPU, gradients
all-reduced over NCCL (NVLink 5 / NVSwitch) in buckets that are launched from autograd hooks while the
rest of backward is still running.

Replaces ``nn.DataParallel`` at solver.py:34-36 (per-step parameter broadcast + gather/reduce to GPU 0).

Buckets are filled in *reverse registration order* (the order gradients become ready in backward:
classifier -> vector MFB blocks -> co-attention -> img_conv1d / ques_proj1 -> question attention ->
LSTM -> embedding), so the large early buckets are on the wire while the long img_conv1d wgrad GEMM
runs.  ``None`` gradients (hieCoAtten's dead fc_Wbq) are sent as zeros so every rank issues the same
collectives; exactly-zero gradients (MFB's dead first stage) need no special care.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter], device, dtype):
        self.params = params
        # every view starts on a 16-byte boundary: the GEMM epilogues and the fused optimizer use 128-bit accesses on
        # them (a 2-element bias in the middle of a bucket would otherwise knock every later view off alignment)
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.payload = sum(p.numel() for p in params)
        self.flat = torch.zeros(self.numel, device=device, dtype=dtype)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, params)]
        self.pending = len(params)
        self.handle = None


class GradientAllReducer:
    """Bucketed, backward-overlapped gradient averaging for a module replicated on every rank."""

    def __init__(self, module: torch.nn.Module, bucket_mb: float = 32.0, process_group=None, defer_params=None,
                 contiguous_groups=None):
        """defer_params: optional iterable of parameters; buckets that become ready are HELD until every one of these
        has received its gradient, then flushed at once (later buckets launch immediately).  For MHBCoAtt the fusion /
        co-attention parameters are deferred until the block's backward is over: the all-reduce then overlaps the
        question LSTM's backward (4 ms of tiny kernels on a mostly idle GPU) instead of stealing SMs from the
        persistent one-CTA-per-SM GEMMs, whose statically strided tiles would otherwise run in two waves."""
        self.module = module
        self._defer = set(defer_params) if defer_params is not None else set()
        self._defer_left = 0
        self._held: List[_Bucket] = []
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        # contiguous_groups: lists of parameters that must sit side by side, in the given order, inside ONE bucket: one
        # wgrad GEMM then writes all their gradients (fused_block.MhbFusedBlockFn).  Default: what the module declares.
        if contiguous_groups is None:
            groups = []
            for m in module.modules():
                if hasattr(m, "fused_param_groups"):
                    groups += m.fused_param_groups()
            contiguous_groups = groups
        group_of = {}
        for g in contiguous_groups:
            for p in g:
                group_of[id(p)] = g
        self._index = {}
        self.buckets: List[_Bucket] = []
        cur, cur_bytes = [], 0
        cap = int(bucket_mb * 1024 * 1024)
        placed = set()
        for p0 in reversed(params):
            if id(p0) in placed:
                continue
            unit = [q for q in group_of.get(id(p0), [p0]) if q.requires_grad]
            for p in unit:
                placed.add(id(p))
                cur.append(p)
                cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= cap:
                self.buckets.append(_Bucket(cur, p.device, p.dtype))
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(_Bucket(cur, cur[0].device, cur[0].dtype))
        from . import ops
        for bi, b in enumerate(self.buckets):
            for pi, p in enumerate(b.params):
                self._index[p] = (bi, pi)
                p.register_post_accumulate_grad_hook(self._hook)
                if p.dim() >= 2 and p.dtype == torch.float32:
                    ops.grad_dest[id(p)] = b.views[pi]      # wgrad kernels write weight gradients straight here
        self._use_avg = dist.is_initialized() and dist.get_backend(process_group) == "nccl"

    # ---- per-step protocol: prepare() -> loss.backward() -> finish()
    def close(self):
        """Unregister the in-place gradient destinations (call before discarding the reducer)."""
        from . import ops
        for p in self._index:
            ops.grad_dest.pop(id(p), None)
            ops.grad_dest_used.discard(id(p))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def prepare(self):
        from . import ops
        for p in self._index:
            ops.grad_dest_used.discard(id(p))
        for b in self.buckets:
            b.pending = len(b.params)
            b.handle = None
        for p in self._index:
            p.grad = None
        self._defer_left = len(self._defer)
        self._held = []

    def _launch(self, b: _Bucket):
        if self.world == 1:
            return
        op = dist.ReduceOp.AVG if self._use_avg else dist.ReduceOp.SUM
        b.handle = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    def _hook(self, p: torch.nn.Parameter):
        bi, pi = self._index[p]
        b = self.buckets[bi]
        if p.grad.data_ptr() != b.views[pi].data_ptr():      # already there when a wgrad kernel wrote it in place
            b.views[pi].copy_(p.grad)
        p.grad = b.views[pi]                 # the optimizer reads the reduced values in place
        b.pending -= 1
        if p in self._defer:
            self._defer_left -= 1
        hold = self._defer_left > 0
        if b.pending == 0:
            if hold:
                self._held.append(b)
            else:
                self._launch(b)
        if not hold and self._held:
            self._flush_held()

    def _flush_held(self):
        for hb in self._held:
            self._launch(hb)
        self._held = []

    def finish(self, optimizer=None):
        """Flush parameters that received no gradient (as zeros), wait for every bucket.

        optimizer: optional optimizer whose ``step(only=params)`` updates a subset of the parameters
        (optim.FusedAdam).  Each bucket is then updated right behind its own all-reduce -- the wait is a stream
        dependency, not a host block -- so the update of the early buckets overlaps the wire time of the late ones
        and the caller must NOT call ``optimizer.step()`` again for this iteration."""
        self._defer_left = 0
        self._flush_held()
        if optimizer is not None and hasattr(optimizer, "advance_step"):
            optimizer.advance_step()                 # device-side step count: once per iteration, not once per bucket
        for b in self.buckets:
            if b.pending > 0:
                for pi, p in enumerate(b.params):
                    if p.grad is None or p.grad.data_ptr() != b.views[pi].data_ptr():
                        if p.grad is None:
                            b.views[pi].zero_()
                        else:
                            b.views[pi].copy_(p.grad)
                        p.grad = b.views[pi]
                b.pending = 0
                self._launch(b)
        for b in self.buckets:
            if b.handle is not None:
                b.handle.wait()
                if not self._use_avg:
                    b.flat.div_(self.world)
                b.handle = None
            if optimizer is not None:
                optimizer.step(only=b.params)

    def bytes_per_step(self) -> int:
        return sum(b.payload * b.flat.element_size() for b in self.buckets)
