"""Inference / validation around the drop-in modules.

``GraphedForward``: CUDA-graph replay of the eval forward (BASELINE config 5: MFH inference, batch 1..4096, 100-region
bottom-up features, batch-sharded across GPUs with no communication).  At small batch the forward is launch-bound (~50
kernel launches plus Python dispatch for a few microseconds of math per launch); capturing it once and replaying it
removes the host from the loop.  The captured work is exactly the eager eval forward of the module (same kernels, same
C-ABI calls, dropout off, the fused log-softmax + argmax tail); results equal eager mode up to the order of the fp32
atomic accumulations (per-sample sum |z| of the L2 norm).

``evaluate``: the ``no_grad`` validation loop of the reference's ``Solver.val`` (solver.py:119-182): eval mode, forward,
criterion, ``softmax(logits).max(1)[1]`` predictions, accuracy against the hard answer or the top soft answer.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch


class GraphedForward:
    """``GraphedForward(model, img_example, questions_example)(img, questions) -> log-probs / logits``.

    The module is put in ``eval()``; inputs must keep the example's shapes and dtypes.  One instance per batch shape.
    After a call, ``pred`` holds the predicted answer ids when the module's tail produced them (MHBCoAtt / MHB)."""

    def __init__(self, model: torch.nn.Module, img_example: torch.Tensor, questions_example: torch.Tensor, warmup: int = 3):
        if not img_example.is_cuda:
            raise RuntimeError("GraphedForward needs CUDA tensors (there is no CPU path)")
        self.model = model.eval()
        self.img = img_example.clone()
        self.q = questions_example.clone()
        side = torch.cuda.Stream(device=img_example.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                  # populates the weight cache, cuDNN plans, allocator pools
                self.model(self.img, self.q)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            out = self.model(self.img, self.q)
            self.out = out[0] if isinstance(out, tuple) else out
            self.pred = getattr(self.model, "last_pred", None)
            if self.pred is None:
                self.pred = self.out.argmax(1)

    @torch.no_grad()
    def __call__(self, img: torch.Tensor, questions: torch.Tensor) -> torch.Tensor:
        self.img.copy_(img, non_blocking=True)
        self.q.copy_(questions, non_blocking=True)
        self.graph.replay()
        return self.out


@torch.no_grad()
def evaluate(model: torch.nn.Module, batches: Iterable, criterion: Optional[Callable] = None, max_batches: int = 0):
    """``Solver.val`` (solver.py:119-182) for the drop-in modules: ``batches`` yields ``(img, questions, answers)`` CUDA
    tensors (answers: int64 ids, or soft rows ``[N, A]`` whose arg-max is the label, solver.py:150-151).  Returns
    ``{"loss": mean batch loss or None, "acc": exact-match accuracy, "n": samples}``.  The module is left in eval mode;
    predictions come from the fused classifier tail when the module provides it."""
    model.eval()
    tot_loss, n_loss, correct, n = 0.0, 0, None, 0
    for j, (img, q, a) in enumerate(batches):
        if max_batches and j >= max_batches:
            break
        out = model(img, q)
        logits = out[0] if isinstance(out, tuple) else out
        if criterion is not None:
            tot_loss = tot_loss + criterion(logits, a).detach()
            n_loss += 1
        pred = getattr(model, "last_pred", None)
        if pred is None:
            pred = logits.argmax(1)                              # softmax is monotone: solver.py:148-149
        label = a.max(1)[1] if a.dim() == 2 else a.long()        # solver.py:150-151
        c = (pred == label).sum()
        correct = c if correct is None else correct + c
        n += int(img.shape[0])
    return {"loss": float(tot_loss / n_loss) if n_loss else None, "acc": float(correct) / n if n else 0.0, "n": n}
