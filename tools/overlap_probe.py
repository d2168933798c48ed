"""Probe (one B200): how much do the latency-bound recurrence kernels slow down when bandwidth-bound work runs beside
them on a second stream, and how much of that work gets done for free?  Decides whether overlapping the recurrence with
the feature cast (forward) or the optimizer update (backward) pays.  Prints one JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vqa_attention_networks_b200 import ops          # noqa: E402
from vqa_attention_networks_b200.optim import FusedAdam  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
Bt, S, E, H = 26, 256, 300, 1024
lstm = torch.nn.LSTM(E, H, batch_first=True).to(dev)
cache = ops.WeightCache()
x = torch.randn(Bt, S, E, device=dev, requires_grad=True)
cot = torch.randn(Bt, S, H, device=dev)
X = torch.randn(256, 196, 2048, device=dev).relu_()
params = [torch.nn.Parameter(torch.randn(5000, 2048, device=dev)) for _ in range(8)]
for p in params:
    p.grad = torch.randn_like(p)
opt = FusedAdam(params, lr=1e-3)
opt.step()
side = torch.cuda.Stream()


def lstm_fb():
    out = ops.LstmFn.apply(x, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0, cache)
    out.backward(cot)


def timed(fn, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for _ in range(3):
    lstm_fb()
    ops.pack_bf16(X)
    opt.step()
res = {"lstm_fwd_bwd_alone_ms": timed(lstm_fb), "pack_alone_ms": timed(lambda: ops.pack_bf16(X)),
       "adam_alone_ms": timed(lambda: opt.step())}


def both(load, n_load):
    def run():
        ev = torch.cuda.Event()
        ev.record()
        side.wait_event(ev)
        with torch.cuda.stream(side):
            for _ in range(n_load):
                load()
        lstm_fb()
        torch.cuda.current_stream().wait_stream(side)
    return run


for name, load, n in (("pack", lambda: ops.pack_bf16(X), 4), ("adam", lambda: opt.step(), 2), ("pack1", lambda: ops.pack_bf16(X), 1)):
    res["lstm_with_%s_x%d_ms" % (name, n)] = timed(both(load, n))
print(json.dumps(res))
