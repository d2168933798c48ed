"""Times vqa_b200_mfb_bwd on the grid MFB's c2 shape (256 x 196 rows, N = 5000, dropout 0.1) alone (CUDA events, 20 launches; `single`: three launches for an ncu capture).  Traffic per launch:
keep 502 MB + dI 502 MB + y, g 100 MB each = 1.2 GB."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vqa_attention_networks_b200 import ops

dev = "cuda:0"
groups, rpg, N = 256, 196, 5000
M, No = groups * rpg, N // 5
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
kw = dict(g=rnd(M, No).bfloat16(), Y=rnd(M, No).bfloat16(), inv=rnd(groups).abs() + 0.5, t=rnd(groups), Q=rnd(groups, N),
          keep=rnd(M, N).bfloat16(), rows_per_group=rpg, di_dtype=torch.bfloat16, p=0.1, seed=7,
          dbias=torch.zeros(N, device=dev))
bytes_ = 2 * M * N * 2 + 2 * M * No * 2
if len(sys.argv) > 1 and sys.argv[1] == "single":        # for ncu: a few launches, nothing else
    for _ in range(3):
        ops.mfb_bwd(**kw)
    torch.cuda.synchronize()
    sys.exit(0)
for _ in range(3):
    ops.mfb_bwd(**kw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    ops.mfb_bwd(**kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
res = {"ms": round(ms, 4), "GBps": round(bytes_ / ms / 1e6, 1), "bytes": bytes_}
print(res)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/mfb_bwd_probe.json", "w"), indent=1)
