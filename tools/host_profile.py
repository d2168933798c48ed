"""Host-side (Python) cost of enqueuing one MHBCoAtt train step: cProfile over a few steps with the GPU running ahead.
The step is ~6.4 ms of GPU work; if enqueuing it costs more than that, the GPU starves."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vqa_attention_networks_b200 import MHBCoAtt
from vqa_attention_networks_b200.optim import FusedAdam

dev = "cuda:0"
torch.manual_seed(0)
model = MHBCoAtt(bench.cfg_ns())
for n, p in model.named_parameters():
    if n.find("bias") == -1:
        torch.nn.init.xavier_uniform_(p)
model = model.to(dev).train()
opt = FusedAdam(model.parameters(), lr=7e-4).attach(model)
crit = torch.nn.KLDivLoss()
img, q, tgt = [t.to(dev) for t in bench.synth_batch(torch, 256, 1)]


def step():
    loss = crit(model(img, q), tgt)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.2f ms/step; wall incl. drain %.2f ms/step" % ((t1 - t0) * 50, (t2 - t0) * 50))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
