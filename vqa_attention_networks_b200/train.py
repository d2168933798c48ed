"""The Solver-equivalent training iteration around the drop-in modules, eager or captured in CUDA graphs.

``TrainStep`` is the body of the reference's train loop (solver.py:68-94 for the MFB / MFH nets, train_hfd.py:109-126 for
HieCoAtten): ``forward -> criterion -> zero_grad -> backward -> [gradient all-reduce] -> Adam``.  ``GraphedTrainStep``
captures that whole iteration -- forward, loss, backward, NCCL all-reduce, optimizer -- once per input slot and replays
it: a MHBCoAtt step is ~160 kernel launches that Python needs 4.5-6.9 ms to enqueue for 6.3-7.5 ms of GPU work, so eight
ranks on a 16-core host are bound by the host, not by the GPUs.  Under replay the host does one ``cudaGraphLaunch``.

What makes the iteration capturable (nothing in it may depend on host state that changes from step to step):
  * dropout: every fused-epilogue dropout site keeps the host seed it drew at capture time and salts it with a DEVICE
    step counter (``seed_dev``, include/vqa_b200.h) that the graph increments first thing;
  * Adam: the bias corrections are formed on the device from the same kind of counter (``vqa_b200_adam_step_dev``);
  * memory: activations, gradients and workspaces live in the graph's private pool (static addresses);
  * inputs: one graph per input SLOT (static ``img / questions / target`` tensors the caller copies into, or H2D-copies
    straight into) sharing one pool, so a slot can be refilled while another one is being consumed.

Timing inside a captured step: ``segment_tags`` names kernel launches that are kept OUT of the graphs -- capture ends
in front of such a launch and a new graph begins behind it; at replay the launch is issued through the C ABI with the
arguments recorded at capture time, bracketed by ordinary CUDA events on the launching stream (bench.py's live roofline
numbers are taken this way, inside the timed region).
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

import torch

from . import ops


class TrainStep:
    """One eager training iteration.  ``reducer``: optional ddp.GradientAllReducer (data-parallel ranks);
    ``optimizer``: torch.optim.Adam or optim.FusedAdam (per-bucket stepping is used when both support it)."""

    def __init__(self, model: torch.nn.Module, criterion: Callable, optimizer: torch.optim.Optimizer, reducer=None,
                 bucket_step: bool = True):
        self.model, self.criterion, self.optimizer, self.reducer = model, criterion, optimizer, reducer
        self.bucket_step = bucket_step and reducer is not None and hasattr(optimizer, "attach")
        # The reference's soft-answer loss -- nn.KLDivLoss() on the model's log_softmax output (solver.py:26-29) -- runs as
        # one fused kernel per direction when the model can hand its logits over (``_log_softmax`` of the drop-in classes);
        # any other criterion / model is called exactly as the solver calls it.  VQA_B200_LOSS=stock keeps the stock ops.
        self._fused_loss_model = None
        if (type(criterion) is torch.nn.KLDivLoss and criterion.reduction == "mean" and not criterion.log_target
                and hasattr(model, "_log_softmax") and os.environ.get("VQA_B200_LOSS", "fast") != "stock"):
            self._fused_loss_model = model

    def forward(self, img, questions):
        out = self.model(img, questions)
        return out[0] if isinstance(out, tuple) else out          # HieCoAtten returns (x, av, aq)

    def loss(self, img, questions, target) -> torch.Tensor:
        m = self._fused_loss_model
        if (m is None or not target.is_cuda or target.dim() != 2 or not target.is_floating_point()):
            return self.criterion(self.forward(img, questions), target)
        m.defer_log_softmax, m.deferred_log_softmax = True, False
        try:
            out = self.forward(img, questions)
        finally:
            m.defer_log_softmax = False
        if not m.deferred_log_softmax:
            return self.criterion(out, target)        # the model's output did not come from _log_softmax (MFB: logits)
        if out.dtype != torch.float32 or out.shape != target.shape:
            out = torch.nn.functional.log_softmax(out, dim=1)
            return self.criterion(out, target)
        return ops.KLDivLogSoftmaxFn.apply(out, target)

    def __call__(self, img, questions, target) -> torch.Tensor:
        if self.reducer is not None and hasattr(self.reducer, "begin_step"):
            self.reducer.begin_step()                 # sharded optimizer: all-gather of the updated bf16 weights
        loss = self.loss(img, questions, target)
        if self.reducer is not None:
            self.reducer.prepare()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if self.bucket_step:
            self.reducer.finish(self.optimizer)       # Adam per bucket, right behind that bucket's all-reduce
        else:
            if self.reducer is not None:
                self.reducer.finish()
            self.optimizer.step()
        return loss


class _EagerCall:
    """A kernel launch kept out of the graphs (see module docstring): C-ABI entry point + the arguments of capture time."""
    __slots__ = ("name", "tag", "args", "events")

    def __init__(self, name, tag, args):
        self.name, self.tag, self.args, self.events = name, tag, args, []


class _Capture:
    """State of one multi-segment capture; ops._call consults `ops._capture` for every launch."""

    def __init__(self, pool, stream, tags, before_cut=None):
        self.pool, self.stream, self.tags = pool, stream, set(tags or ())
        self.segments: List[object] = []
        self.graph = None
        # called before a segment ends: work forked to other streams inside the segment (the sharded optimizer's
        # all-gathers on the NCCL stream) must be joined first -- a capture cannot end with unjoined streams
        self.before_cut = before_cut

    def begin(self):
        self.graph = torch.cuda.CUDAGraph()
        self.graph.capture_begin(pool=self.pool, capture_error_mode="global")

    def end(self):
        self.graph.capture_end()
        self.segments.append(self.graph)
        self.graph = None

    def intercept(self, name, tag, args) -> bool:
        if tag not in self.tags:
            return False
        if self.before_cut is not None:
            self.before_cut()
        self.end()
        self.segments.append(_EagerCall(name, tag, args))
        self.begin()
        return True


class GraphedTrainStep:
    """``GraphedTrainStep(step, slots)``: capture ``step`` once per input slot; ``replay(i)`` runs slot i's iteration and
    returns its (static) loss tensor.

    slots: list of ``(img, questions, target)`` CUDA tensors -- they BECOME the static inputs: refill them in place
    (``slot[0].copy_(host_img, non_blocking=True)`` on a copy stream, ordered by events against ``replay``).
    The model must be in train mode; dropout sites and the optimizer are switched to their device-counter forms."""

    def __init__(self, step: TrainStep, slots: Sequence[Sequence[torch.Tensor]], warmup: int = 3,
                 segment_tags: Optional[Sequence[str]] = None):
        if not slots or not all(t.is_cuda for s in slots for t in s):
            raise RuntimeError("GraphedTrainStep needs CUDA input slots (there is no CPU path)")
        self.step, self.slots = step, [tuple(s) for s in slots]
        dev = self.slots[0][0].device
        model, opt = step.model, step.optimizer
        # device step counter shared by every fused dropout site of the model (low 32 bits salt the seeds)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        for m in model.modules():
            if hasattr(m, "seed_counter"):
                m.seed_counter = self.step_count
        if hasattr(opt, "enable_device_step"):
            opt.enable_device_step(dev)
        elif not all(g.get("capturable", False) for g in opt.param_groups):
            raise RuntimeError("GraphedTrainStep needs optim.FusedAdam or a torch optimizer built with capturable=True")
        self.stream = torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            for i in range(max(1, warmup)):          # eager: weight caches, allocator pools, cuDNN plans, NCCL channels
                self._iteration(self.slots[i % len(self.slots)])
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        self.pool = torch.cuda.graph_pool_handle()
        self.programs, self.losses, self.launches = [], [], 0
        for slot in self.slots:
            prog, loss, n = self._capture(slot, segment_tags)
            self.programs.append(prog)
            self.losses.append(loss)
            self.launches = n
        self.replays = 0
        self.timing = False

    def _iteration(self, slot):
        self.step_count.add_(1)                      # first node of every graph: new dropout masks for this step
        return self.step(*slot)

    def _capture(self, slot, tags):
        cap = _Capture(self.pool, self.stream, tags, getattr(self.step.reducer, "_join_gathers", None))
        torch.cuda.synchronize()
        n0 = ops.LaunchStats.count
        with torch.cuda.stream(self.stream):
            cap.begin()
            ops._capture = cap
            try:
                loss = self._iteration(slot)
            finally:
                ops._capture = None
                if cap.graph is not None:
                    cap.end()
        torch.cuda.synchronize()
        return cap.segments, loss.detach(), ops.LaunchStats.count - n0

    def replay(self, i: int = 0) -> torch.Tensor:
        """Run the captured iteration of slot i on the CURRENT stream; returns the slot's static loss tensor."""
        for seg in self.programs[i]:
            if isinstance(seg, _EagerCall):
                if self.timing:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ops.launch_direct(seg.name, seg.args)          # same operands as at capture time, current stream
                    e1.record()
                    seg.events.append((e0, e1))
                else:
                    ops.launch_direct(seg.name, seg.args)
            else:
                seg.replay()
        self.replays += 1
        return self.losses[i]

    def kernel_times(self):
        """tag -> (launches, total_ms) of the segment launches timed since the last reset (after a synchronize)."""
        out = {}
        for prog in self.programs:
            for seg in prog:
                if isinstance(seg, _EagerCall) and seg.events:
                    n, ms = out.get(seg.tag, (0, 0.0))
                    out[seg.tag] = (n + len(seg.events), ms + sum(a.elapsed_time(b) for a, b in seg.events))
        return out

    def reset_times(self, timing: bool):
        self.timing = timing
        for prog in self.programs:
            for seg in prog:
                if isinstance(seg, _EagerCall):
                    seg.events = []

    def sync_python_state(self):
        """Bring the Python-side bookkeeping in line with what the replays did on the device: parameter version counters
        (replays update the weights without Python seeing it) and the optimizer's per-parameter step counts.  Call before
        checkpointing, evaluating, or going back to eager training."""
        for p in self.step.model.parameters():
            torch.autograd.graph.increment_version(p)
        opt = self.step.optimizer
        if hasattr(opt, "sync_step_from_device"):
            opt.sync_step_from_device()
        for m in self.step.model.modules():
            c = getattr(m, "_wcache", None)
            if c is not None:
                c.clear()
