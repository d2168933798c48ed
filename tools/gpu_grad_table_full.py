"""Per-parameter gradient errors at BASELINE dimensions (batch 6) vs the fp64 oracle (z-injected), plus the
error of an fp32 evaluation of the oracle itself (how ill-conditioned each gradient is)."""
import os, sys, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from vqa_attention_networks_b200 import MHBCoAtt

DEV = "cuda:0"
L = int(sys.argv[1]) if len(sys.argv) > 1 else 196
cfg = types.SimpleNamespace(model_name="mhb_coAtt", q_vocab_size=15000, emb_dim=300, hidden_dim=1024, num_layers=1,
                            img_feature_channel=2048, img_feature_dim=L, a_vocab_size=3000, glove=False)
torch.manual_seed(0)
model = MHBCoAtt(cfg)
for n, p in model.named_parameters():
    if n.find("bias") == -1:
        torch.nn.init.xavier_uniform_(p)
model = model.to(DEV).train()
model.dropout_l.p = 0.0
model.dropout_m.p = 0.0
X = O.synthetic_inputs(6, L, 2048, 26, 15000, seed=1234, device=DEV)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
cot = torch.randn(6, 3000, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
for mode in ("fp32", "bf16"):
    model.precision = mode
    model.zero_grad(set_to_none=True)
    model.capture = {}
    out = model(X["img"], X["questions"])
    (out * cot).sum().backward()
    inj = {}
    for key, y in model.capture.items():
        y = y.detach().double()
        z = torch.sign(y) * y * y
        inj["z" + key[1:]] = z.reshape(6, -1, z.shape[-1]) if key == "y1" else z
    res = {}
    for dt in (torch.float64, torch.float32):
        P = {k: v.detach().clone().to(dt).requires_grad_(True) for k, v in sd.items()}
        ref = O.mhbcoatt_forward(P, X["img"].to(dt), X["questions"], None, inj)
        (ref * cot.to(dt)).sum().backward()
        res[dt] = (ref.detach(), {k: v.grad for k, v in P.items()})
    print("== L=%d [%s] out rel-err %.3e (oracle fp32 vs fp64: %.3e)" % (L, mode, O.rel_err(out, res[torch.float64][0]),
                                                                     O.rel_err(res[torch.float32][0], res[torch.float64][0])))
    for k, p in model.named_parameters():
        r64, r32 = res[torch.float64][1][k], res[torch.float32][1][k]
        print("   %-26s |g|=%.3e  ours %.3e   oracle-fp32 %.3e" % (k, float(r64.norm()), O.rel_err(p.grad, r64), O.rel_err(r32, r64)))
