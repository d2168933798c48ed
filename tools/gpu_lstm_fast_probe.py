"""Times the persistent LSTM recurrence (ops.LstmFn) against the stock cuDNN nn.LSTM on the reference's feed
([T=26 rows, N=256 steps, E=300]) and prints their difference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vqa_attention_networks_b200 import ops

dev = "cuda:0"
torch.manual_seed(0)
lstm = torch.nn.LSTM(input_size=300, hidden_size=1024, num_layers=1, batch_first=True).to(dev)
for n, p in lstm.named_parameters():
    if "bias" not in n:
        torch.nn.init.xavier_uniform_(p)
x = torch.tanh(torch.randn(256, 26, 300, device=dev)).permute(1, 0, 2).requires_grad_(True)
cot = torch.randn(26, 256, 1024, device=dev)
cache = ops.WeightCache()


def stock():
    lstm.zero_grad(set_to_none=True)
    o, _ = lstm(x)
    (o * cot).sum().backward()
    return o


def fast():
    lstm.zero_grad(set_to_none=True)
    o = ops.LstmFn.apply(x, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0, cache)
    (o * cot).sum().backward()
    return o


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


o1 = stock().detach().clone()
g1 = {n: p.grad.clone() for n, p in lstm.named_parameters()}
o2 = fast().detach().clone()
g2 = {n: p.grad.clone() for n, p in lstm.named_parameters()}
print("rel |fast - stock| out: %.3e" % float((o1 - o2).norm() / o1.norm()))
for n in g1:
    print("  grad %-14s rel %.3e" % (n, float((g1[n] - g2[n]).norm() / g1[n].norm())))
print("stock cuDNN LSTM fwd+bwd: %.3f ms" % timeit(stock))
print("persistent LSTM fwd+bwd:  %.3f ms" % timeit(fast))
ops.LaunchStats.reset(timing=True)
for _ in range(10):
    fast()
torch.cuda.synchronize()
for k, (n, ms) in sorted(ops.LaunchStats.summary().items(), key=lambda kv: -kv[1][1]):
    print("  %-28s %3d launches  %.3f ms each" % (k, n, ms / n))
ops.LaunchStats.reset(timing=False)
with torch.no_grad():
    print("persistent LSTM fwd only (no_grad): %.3f ms" % timeit(lambda: ops.LstmFn.apply(
        x.detach(), lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0, cache)))
    print("stock LSTM fwd only (no_grad):      %.3f ms" % timeit(lambda: lstm(x.detach())))
