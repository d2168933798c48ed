"""Where does a step of the persistent LSTM recurrence spend its time?  Per-phase cycle counters of CTA 0 / thread 0
(vqa_b200_debug_set_lstm) and experiment switches (results are wrong under the switches; only the timing matters)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vqa_attention_networks_b200 import ops, _lib

dev = "cuda:0"
L = _lib.load()
torch.manual_seed(0)
lstm = torch.nn.LSTM(input_size=300, hidden_size=1024, num_layers=1, batch_first=True).to(dev)
x = torch.tanh(torch.randn(256, 26, 300, device=dev)).permute(1, 0, 2).requires_grad_(True)
cot = torch.randn(26, 256, 1024, device=dev)
cache = ops.WeightCache()
dbg = torch.zeros(16, dtype=torch.int64, device=dev)
S = 256


def run():
    lstm.zero_grad(set_to_none=True)
    o = ops.LstmFn.apply(x, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0, cache)
    (o * cot).sum().backward()


names = ["wait", "sweep(+mma)", "mma/partials", "barrier|reduce", "reduce+math", "stores+prefetch", "barrier2", "retries"]
for mode, label in [(0x0000, "bwd cluster 4"), (0x0200, "bwd cluster 2"), (0x0100, "bwd single CTA"), (0x0002, "no mma")]:
    L.vqa_b200_debug_set_lstm(ctypes.c_void_p(dbg.data_ptr()), mode)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    ops.LaunchStats.reset(timing=True)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    sm = ops.LaunchStats.summary()
    ops.LaunchStats.reset(timing=False)
    d = dbg.cpu().tolist()
    f = " ".join("%s %d" % (n, d[i] / S) for i, n in enumerate(names))
    b = " ".join("%s %d" % (n, d[8 + i] / S) for i, n in enumerate(names))
    f += "  (total retries %d)" % d[7]
    b += "  (total retries %d)" % d[15]
    print("mode %-30s fwd %.3f ms  bwd %.3f ms" % (label, sm["lstm_fwd"][1] / 5, sm["lstm_bwd"][1] / 5))
    print("    fwd cycles/step: " + f)
    print("    bwd cycles/step: " + b)
L.vqa_b200_debug_set_lstm(None, 0)
