"""BASELINE config 5: MHBCoAtt eval forward, 100 regions x 2048, batch sweep, eager vs CUDA-graph replay (one GPU;
batch-sharding over GPUs needs no communication, so N GPUs serve N x these numbers)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vqa_attention_networks_b200 import MHBCoAtt
from vqa_attention_networks_b200.inference import GraphedForward

DEV = "cuda:0"
cfg = types.SimpleNamespace(model_name="mhb_coAtt", q_vocab_size=15000, emb_dim=300, hidden_dim=1024, num_layers=1,
                            img_feature_channel=2048, img_feature_dim=100, a_vocab_size=3000, glove=False)
torch.manual_seed(0)
model = MHBCoAtt(cfg)
for n, p in model.named_parameters():
    if n.find("bias") == -1:
        torch.nn.init.xavier_uniform_(p)
model = model.to(DEV).eval()


def timeit(fn, iters):
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


print("batch  eager_ms  graph_ms  graph_samples_per_s  max|eager-graph|")
for B in (1, 4, 16, 64, 256, 1024, 4096):
    img = torch.relu(torch.randn(B, 100, 2048, device=DEV))
    q = torch.randint(0, 15000, (B, 26), device=DEV)
    with torch.no_grad():
        eager = model(img, q).clone()
        t_e = timeit(lambda: model(img, q), 20 if B <= 256 else 5)
    g = GraphedForward(model, img, q)
    out = g(img, q)
    t_g = timeit(lambda: g(img, q), 20 if B <= 256 else 5)
    print("%5d  %8.3f  %8.3f  %12.0f  %.2e" % (B, t_e, t_g, B / (t_g / 1e3), float((out - eager).abs().max())))
