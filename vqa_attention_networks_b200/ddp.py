"""Data-parallel gradient exchange: one process per GPU, parameters resident on every GPU, gradients
all-reduced over NCCL (NVLink 5 / NVSwitch) in buckets that are launched from autograd hooks while the
rest of backward is still running.

Replaces ``nn.DataParallel`` at solver.py:34-36 (per-step parameter broadcast + gather/reduce to GPU 0).

Buckets are filled in *reverse registration order* (the order gradients become ready in backward:
classifier -> vector MFB blocks -> co-attention -> img_conv1d / ques_proj1 -> question attention ->
LSTM -> embedding), so the large early buckets are on the wire while the long img_conv1d wgrad GEMM
runs.  ``None`` gradients (hieCoAtten's dead fc_Wbq) are sent as zeros so every rank issues the same
collectives; exactly-zero gradients (MFB's dead first stage) need no special care.

Sharded optimizer (``shard_optimizer=FusedAdam``; SURVEY.md 8f rank 1 "reduce-scatter -> Adam on shard -> all-gather").
The weights that the kernels read only through their bf16 copies (``module.bf16_only_weights()``: 94 % of MHBCoAtt's
parameters) are exchanged differently:
    backward   gradient bucket ready  -> ``ncclReduceScatter`` (AVG, fp32, in place: each rank keeps 1/P of the bucket)
    finish()   fused Adam on the rank's shard only (fp32 master weights, exp_avg, exp_avg_sq: 1/P of the state and of the
               0.48 ms full-size update), writing the new bf16 weight copies of the shard
    next step  ``ncclAllGather`` of the bf16 copies (2 bytes per weight instead of the all-reduce's second 4), issued at
               the start of the iteration in forward-use order; each weight's first consumer waits for its bucket only
               (ops.WeightCache pending entries), so the gather overlaps the question encoder's recurrence
Wire bytes per step drop to ~0.76x of the all-reduce's; everything else (biases, embedding, small weights) keeps the
all-reduce + full-size update.  The fp32 masters of a sharded weight are current on the owning rank only;
``sync_master_weights()`` completes them everywhere (called automatically before ``state_dict()`` and on train/eval
switches); the Adam state of sharded weights lives in the reducer, sharded.
"""
from __future__ import annotations

import os
from typing import List

import torch
import torch.distributed as dist


class _Pending:
    """An async collective several cache entries wait on; the first reader makes its stream wait, later ones do not."""

    def __init__(self, handle):
        self.handle = handle

    def wait(self):
        if self.handle is not None:
            self.handle.wait()
            self.handle = None


def shard_segments(offsets, numels, lo, hi):
    """Intersections of the flat index range [lo, hi) with parameters laid out at `offsets` with `numels` elements:
    [(param index, start inside the parameter, start inside the flat buffer, length), ...]."""
    out = []
    for i, (o, n) in enumerate(zip(offsets, numels)):
        a, b = max(lo, o), min(hi, o + n)
        if a < b:
            out.append((i, a - o, a, b - a))
    return out


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter], device, dtype, sharded=False, world=1):
        self.params = params
        self.sharded = sharded
        # every view starts on a 16-byte boundary: the GEMM epilogues and the fused optimizer use 128-bit accesses on
        # them (a 2-element bias in the middle of a bucket would otherwise knock every later view off alignment)
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        if sharded:                           # equal shards whose bf16 halves stay 16-byte aligned
            q = world * 8
            off = (off + q - 1) // q * q
        self.offsets = offs
        self.numel = off
        self.payload = sum(p.numel() for p in params)
        self.flat = torch.zeros(self.numel, device=device, dtype=dtype)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, params)]
        self.pending = len(params)
        self.handle = None


class GradientAllReducer:
    """Bucketed, backward-overlapped gradient averaging for a module replicated on every rank."""

    def __init__(self, module: torch.nn.Module, bucket_mb: float = 32.0, process_group=None, defer_params=None,
                 contiguous_groups=None, shard_optimizer=None):
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
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.shard_optimizer = shard_optimizer
        shard_ids = set()
        if shard_optimizer is not None:
            for m in module.modules():
                if hasattr(m, "bf16_only_weights"):
                    shard_ids.update(id(w) for w in m.bf16_only_weights())
        # parameters with exactly-zero gradients on every rank and step (MFB's dead first stage under the reference's
        # singleton-axis softmax): nothing to exchange, nothing to update -- they stay out of the buckets
        dead = set()
        for m in module.modules():
            if hasattr(m, "dead_parameters"):
                dead.update(id(p) for p in m.dead_parameters())
        self.skipped = [p for p in module.parameters() if p.requires_grad and id(p) in dead]
        params = [p for p in module.parameters() if p.requires_grad and id(p) not in dead]
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
        cap = int(bucket_mb * 1024 * 1024)
        placed = set()
        # two bucket streams: sharded weights (reduce-scatter) and everything else (all-reduce) never share a bucket
        cur = {False: ([], 0), True: ([], 0)}
        for p0 in reversed(params):
            if id(p0) in placed:
                continue
            unit = [q for q in group_of.get(id(p0), [p0]) if q.requires_grad]
            kind = all(id(q) in shard_ids for q in unit)
            lst, nbytes = cur[kind]
            for p in unit:
                placed.add(id(p))
                lst.append(p)
                nbytes += p.numel() * p.element_size()
            cur[kind] = (lst, nbytes)
            if nbytes >= cap:
                self.buckets.append(_Bucket(lst, p.device, p.dtype, sharded=kind, world=self.world))
                cur[kind] = ([], 0)
        for kind in (True, False):
            if cur[kind][0]:
                lst = cur[kind][0]
                self.buckets.append(_Bucket(lst, lst[0].device, lst[0].dtype, sharded=kind, world=self.world))
        from . import ops
        for bi, b in enumerate(self.buckets):
            for pi, p in enumerate(b.params):
                self._index[p] = (bi, pi)
                p.register_post_accumulate_grad_hook(self._hook)
                if p.dim() >= 2 and p.dtype == torch.float32:
                    p._vqa_grad_dest = b.views[pi]           # wgrad kernels write weight gradients straight here
                    p._vqa_grad_dest_used = False
        self._use_avg = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        # gradients announced as final before autograd hands them over (ops.grads_enqueued): their buckets start early
        self._early = set()
        self._early_on = os.environ.get("VQA_B200_DDP_EARLY", "1") == "1"
        ops.grad_ready_hooks.append(self._early_ready)
        self._module = module
        self._step = 0
        self._pendings: List[_Pending] = []
        self.dry_run = False                  # True: every collective is skipped (bench.py measures the exposed time with it)
        if any(b.sharded for b in self.buckets):
            self._setup_shards(module)

    # ---- sharded optimizer: set-up
    def _caches_of(self, p):
        return [c for c, ids in self._cache_owners if id(p) in ids]

    def _setup_shards(self, module):
        """Flat bf16 weight buffers (the all-gather targets) adopted by the modules' weight caches, sharded Adam state, and
        the static pointer tables of each rank's shard update."""
        import ctypes
        from . import ops
        self._cache_owners = []
        for m in module.modules():
            c = getattr(m, "_wcache", None)
            if isinstance(c, ops.WeightCache) and hasattr(m, "bf16_only_weights"):
                self._cache_owners.append((c, {id(w) for w in m.bf16_only_weights()}))
                cbs = m.__dict__.setdefault("_mode_switch_callbacks", [])
                cbs.append(self.sync_master_weights)
        for b in self.buckets:
            if not b.sharded:
                continue
            dev = b.flat.device
            b.shard = b.numel // self.world
            lo, hi = self.rank * b.shard, (self.rank + 1) * b.shard
            b.own = b.flat[lo:hi]                                   # reduce-scatter output (in place)
            b.w16 = torch.zeros(b.numel, device=dev, dtype=torch.bfloat16)
            b.w16_own = b.w16[lo:hi]
            b.m = torch.zeros(b.shard, device=dev, dtype=torch.float32)
            b.v = torch.zeros(b.shard, device=dev, dtype=torch.float32)
            numels = [p.numel() for p in b.params]
            b.w16_views = [b.w16[o:o + n] for o, n in zip(b.offsets, numels)]
            for p, v in zip(b.params, b.w16_views):
                v.copy_(p.detach().reshape(-1))                     # fp32 -> bf16, round to nearest even
                for c in self._caches_of(p):
                    c.adopt(p, v)
            segs = shard_segments(b.offsets, numels, lo, hi)
            b.segments = segs
            arr = ctypes.c_void_p * len(segs)
            b.tab = dict(
                n=len(segs),
                P=arr(*[b.params[i].data_ptr() + 4 * s for i, s, f, n in segs]),
                G=arr(*[b.flat.data_ptr() + 4 * f for i, s, f, n in segs]),
                M=arr(*[b.m.data_ptr() + 4 * (f - lo) for i, s, f, n in segs]),
                V=arr(*[b.v.data_ptr() + 4 * (f - lo) for i, s, f, n in segs]),
                B=arr(*[b.w16.data_ptr() + 2 * f for i, s, f, n in segs]),
                numel=(ctypes.c_int64 * len(segs))(*[n for i, s, f, n in segs]))
        self._sd_hook = module.register_state_dict_pre_hook(lambda *a, **k: self.sync_master_weights())

    # ---- sharded optimizer: per-step pieces
    def begin_step(self):
        """Start of an iteration: all-gather the bf16 weight copies the shard updates of the previous iteration wrote,
        in forward-use order.  Nothing waits here: each weight's first reader waits for its own bucket."""
        if self.world == 1 or self.dry_run:
            return
        for b in reversed(self.buckets):
            if not b.sharded:
                continue
            h = dist.all_gather_into_tensor(b.w16, b.w16_own, group=self.group, async_op=True)
            pend = _Pending(h)
            self._pendings.append(pend)
            for p in b.params:
                for c in self._caches_of(p):
                    c.set_pending(p, pend)

    def _join_gathers(self):
        for pend in self._pendings:
            pend.wait()
        self._pendings = []

    def _shard_step(self, b: _Bucket):
        """Fused Adam on this rank's 1/P of the bucket (fp32 masters, sharded exp_avg / exp_avg_sq) + its bf16 copies."""
        from . import _lib, ops
        opt = self.shard_optimizer
        g = opt.param_groups[0]
        beta1, beta2 = g["betas"]
        L = _lib.load()
        t = b.tab
        ops.LaunchStats.count += (t["n"] + 31) // 32
        if opt.step_count is not None:
            rc = L.vqa_b200_adam_step_dev(t["n"], t["P"], t["G"], t["M"], t["V"], t["B"], t["numel"], float(g["lr"]),
                                          float(beta1), float(beta2), float(g["eps"]),
                                          opt.step_count.data_ptr(), ops._st())
        else:
            rc = L.vqa_b200_adam_step(t["n"], t["P"], t["G"], t["M"], t["V"], t["B"], t["numel"], float(g["lr"]),
                                      float(beta1), float(beta2), float(g["eps"]), int(self._step), ops._st())
        _lib.check(rc, "vqa_b200_adam_step (shard)")
        for i, _, _, _ in b.segments:
            torch.autograd.graph.increment_version(b.params[i])

    def sync_master_weights(self):
        """Complete the fp32 weights of the sharded parameters on every rank (each rank owns 1/P of every sharded bucket
        between synchronisations).  Collective: every rank must call it."""
        if self.world == 1 or not any(b.sharded for b in self.buckets):
            return
        self._join_gathers()
        for b in self.buckets:
            if not b.sharded:
                continue
            full = torch.empty(b.numel, device=b.flat.device, dtype=torch.float32)
            lo = self.rank * b.shard
            for i, s, f, n in b.segments:
                full[f:f + n].copy_(b.params[i].detach().reshape(-1)[s:s + n])
            dist.all_gather_into_tensor(full, full[lo:lo + b.shard], group=self.group)
            with torch.no_grad():
                for p, o in zip(b.params, b.offsets):
                    p.copy_(full[o:o + p.numel()].view_as(p))

    # ---- per-step protocol: [begin_step()] -> forward -> prepare() -> loss.backward() -> finish()
    def close(self):
        """Unregister the in-place gradient destinations (call before discarding the reducer)."""
        from . import ops
        for p in self._index:
            for attr in ("_vqa_grad_dest", "_vqa_grad_dest_used"):
                if hasattr(p, attr):
                    delattr(p, attr)
        if self._early_ready in ops.grad_ready_hooks:
            ops.grad_ready_hooks.remove(self._early_ready)
        if getattr(self, "_cache_owners", None):
            for c, _ in self._cache_owners:
                c.unpin_all()
                c.clear()
            for m in self._module.modules():
                cbs = m.__dict__.get("_mode_switch_callbacks")
                if cbs and self.sync_master_weights in cbs:
                    cbs.remove(self.sync_master_weights)
            self._sd_hook.remove()
            self._cache_owners = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def prepare(self):
        from . import ops
        self._join_gathers()                 # a weight nobody read in this forward: its gather still has to be joined
        for p in self._index:
            if hasattr(p, "_vqa_grad_dest"):
                p._vqa_grad_dest_used = False
        for b in self.buckets:
            b.pending = len(b.params)
            b.handle = None
        for p in self._index:
            p.grad = None
        for p in self.skipped:
            p.grad = None
        self._defer_left = len(self._defer)
        self._held = []
        self._early = set()

    def _early_ready(self, params):
        """ops.grads_enqueued(params): the wgrad GEMM that writes these gradients straight into their bucket views is
        already on the stream.  Count them as delivered and start the bucket if that completes it -- `defer_params`
        does not hold these back: they are announced exactly where their exchange has long GEMMs to hide behind."""
        from . import ops
        if not self._early_on or self.world == 1:
            return
        for p in params:
            if p not in self._index or id(p) in self._early or not ops.grad_dest_taken(p):
                continue                      # not ours, announced twice, or not written in place (then the hook copies it)
            bi, pi = self._index[p]
            b = self.buckets[bi]
            self._early.add(id(p))
            b.pending -= 1
            if p in self._defer:
                self._defer_left -= 1
            if b.pending == 0:
                self._launch(b)

    def _launch(self, b: _Bucket):
        if self.world == 1 or self.dry_run:
            return                            # dry_run: same copies and shard updates, no collective (bench.py: exposed time)
        op = dist.ReduceOp.AVG if self._use_avg else dist.ReduceOp.SUM
        if b.sharded:
            # in place: this rank's 1/P of the bucket receives the average, the rest of the buffer is scratch afterwards
            b.handle = dist.reduce_scatter_tensor(b.own, b.flat, op=op, group=self.group, async_op=True)
        else:
            b.handle = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    def _hook(self, p: torch.nn.Parameter):
        bi, pi = self._index[p]
        b = self.buckets[bi]
        if id(p) in self._early:             # announced, counted (and possibly already on the wire): nothing left to do
            p.grad = b.views[pi]
            if not (self._defer_left > 0) and self._held:
                self._flush_held()
            return
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
        self._step += 1
        sharded = any(b.sharded for b in self.buckets)
        if sharded and optimizer is not self.shard_optimizer:
            raise RuntimeError("GradientAllReducer(shard_optimizer=opt): call finish(opt) -- the sharded weights are "
                               "updated here, by their owning ranks, and nowhere else")
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
                    (b.own if b.sharded else b.flat).div_(self.world)
                b.handle = None
            if b.sharded and self.world > 1:
                self._shard_step(b)
            elif optimizer is not None:
                optimizer.step(only=b.params)

    def bytes_per_step(self) -> int:
        """Gradient payload exchanged per step (bytes of the reduced tensors)."""
        return sum(b.payload * b.flat.element_size() for b in self.buckets)

    def wire_bytes_per_step(self) -> int:
        """Bytes each rank sends per step with ring collectives: 2 (P-1)/P S for an all-reduce of S bytes, (P-1)/P S for a
        reduce-scatter or an all-gather."""
        f = (self.world - 1) / max(1, self.world)
        tot = 0.0
        for b in self.buckets:
            if b.sharded:
                tot += f * b.numel * 4 + f * b.numel * 2
            else:
                tot += 2 * f * b.numel * 4
        return int(tot)
