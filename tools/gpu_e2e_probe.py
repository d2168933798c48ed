"""Where does the e2e step lose ~1.7 ms vs the resident step?  Variants of the copy / dependency pattern."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as Bm
from vqa_attention_networks_b200 import MHBCoAtt

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MHBCoAtt(Bm.cfg_ns())
for n, p in model.named_parameters():
    if n.find("bias") == -1:
        torch.nn.init.xavier_uniform_(p)
model = model.to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=7e-4, fused=True)
crit = torch.nn.KLDivLoss()
host = [Bm.synth_batch(torch, 256, 1 + i, pin=True) for i in range(2)]
res = [tuple(t.to(dev) for t in hb) for hb in host]
slots = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
cs = torch.cuda.Stream()
ready = [torch.cuda.Event() for _ in range(2)]
done = [torch.cuda.Event() for _ in range(2)]


def step(b):
    loss = crit(model(b[0], b[1]), b[2])
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


def run(mode, K=20):
    for i in range(3):
        step(res[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        if mode in ("copy_nodep", "copy_dep"):
            with torch.cuda.stream(cs):
                if mode == "copy_dep" and i >= 1:
                    cs.wait_event(done[(i + 1) % 2])
                for d, s in zip(slots[(i + 1) % 2], host[(i + 1) % 2]):
                    d.copy_(s, non_blocking=True)
                ready[(i + 1) % 2].record(cs)
        if mode == "copy_dep":
            if i > 0:
                torch.cuda.current_stream().wait_event(ready[i % 2])
            step(slots[i % 2] if i > 0 else res[0])
            done[i % 2].record()
        else:
            step(res[i % 2])
    e1.record()
    torch.cuda.synchronize()
    print("%-12s %.3f ms/step" % (mode, e0.elapsed_time(e1) / K))


for m in ("resident", "copy_nodep", "copy_dep", "resident"):
    run(m)
