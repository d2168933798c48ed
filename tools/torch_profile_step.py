"""Which ATen ops (outside this repo's kernels) are left in one eager MHBCoAtt train step?  torch.profiler table."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vqa_attention_networks_b200.optim import FusedAdam  # noqa: E402
from vqa_attention_networks_b200.train import TrainStep  # noqa: E402

dev = torch.device("cuda:0")
wl = bench.WORKLOADS["c2"]
model = bench.build_model(torch, wl).to(dev).train()
opt = FusedAdam(model.parameters(), lr=7e-4).attach(model)
step = TrainStep(model, torch.nn.KLDivLoss(), opt)
img, q, tgt = (t.to(dev) for t in bench.synth_batch(torch, 256, 1, L=196, target="soft"))
for _ in range(3):
    step(img, q, tgt)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    for _ in range(2):
        step(img, q, tgt)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.device_time_total > 0 and not e.key.startswith(("vqa_b200", "void vqa", "ProfilerStep")):
        rows.append((e.device_time_total / 2, e.count / 2, e.key, str(e.input_shapes)[:90]))
rows.sort(reverse=True)
for t, n, k, sh in rows[:45]:
    print("%8.1f us  x%-4.1f %-50s %s" % (t, n, k[:50], sh))
