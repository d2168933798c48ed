"""Debug aid: MfbSpatialCoAttFn backward intermediates vs torch autograd (fp64) on a tiny case."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from vqa_attention_networks_b200 import ops
from vqa_attention_networks_b200._lib import K_MAJOR, MN_MAJOR

torch.manual_seed(0)
dev = "cuda:0"
N, L, D, H2, Ah, G = 3, 6, 16, 16, 512, 2
X = torch.relu(torch.randn(N, L, D, device=dev))
qa = torch.randn(N, H2, device=dev)
def mk(*s, sc=0.1): return (torch.randn(*s, device=dev) * sc)
Wq1, bq1 = mk(5000, H2), mk(5000)
Wimg, bimg = mk(5000, D, 1, 1), mk(5000)
Wc1, bc1 = mk(Ah, 1000, 1, 1), mk(Ah)
Wc2, bc2 = mk(G, Ah, 1, 1, sc=1.0), mk(G)
cot = torch.randn(N, G * D, device=dev)

# reference in fp64 with retained intermediates
P = dict(Wq1=Wq1, bq1=bq1, Wimg=Wimg, bimg=bimg, Wc1=Wc1, bc1=bc1, Wc2=Wc2, bc2=bc2)
P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
qa64 = qa.double().requires_grad_(True)
X64 = X.double()
Q1 = O.linear(qa64, P64["Wq1"], P64["bq1"])
I = O.conv1x1(X64, P64["Wimg"], P64["bimg"])
z = O.mfb_pool(I * Q1[:, None, :]); z.retain_grad()
y = O.signed_sqrt(z); y.retain_grad()
nrm = torch.sqrt((y * y).sum(dim=(1, 2), keepdim=True))
yhat = y / nrm; yhat.retain_grad()
h2 = torch.relu(O.conv1x1(yhat, P64["Wc1"], P64["bc1"]))
cl = O.conv1x1(h2, P64["Wc2"], P64["bc2"]); cl.retain_grad()
ca, att = O.softmax_pool(cl.permute(0, 2, 1), X64)
(ca * cot.double()).sum().backward()

for mode in ("fp32", "bf16"):
    cfg = ops.StageCfg(mode=mode)
    ps = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    qa_ = qa.clone().requires_grad_(True)
    out, att_ = ops.MfbSpatialCoAttFn.apply(X, qa_, ps["Wq1"], ps["bq1"], ps["Wimg"], ps["bimg"], ps["Wc1"], ps["bc1"], None, None, ps["Wc2"], ps["bc2"], cfg)
    (out * cot).sum().backward()
    print("==", mode, "fwd", O.rel_err(out, ca))
    for k in P:
        print("   d%-6s %.3e   |ref|=%.3e |got|=%.3e" % (k, O.rel_err(ps[k].grad, P64[k].grad), float(P64[k].grad.norm()), float(ps[k].grad.norm())))
    print("   dqa     %.3e" % O.rel_err(qa_.grad, qa64.grad))

# manual replay of the backward pieces in fp32 mode against the retained intermediates
cfg = ops.StageCfg(mode="fp32")
M = N * L
Xc = X.reshape(M, D).contiguous()
Q1k = ops._linear_fwd(qa.contiguous(), Wq1, bq1, cfg, torch.float32)
print("Q1", O.rel_err(Q1k, Q1))
yk, ssq, keep = ops.mfb_fused(ops.prep(Xc, K_MAJOR, 0, "fp32"), cfg.cache.get(Wimg, K_MAJOR, 1, "fp32"), bimg, Q1k, L, torch.float32, True, 0.0, 0)
print("y", O.rel_err(yk, y.reshape(M, 1000)), "ssq", O.rel_err(ssq, (y * y).sum(dim=(1, 2))))
inv = ops.inv_norm(ssq)
print("inv", O.rel_err(inv, 1 / nrm.reshape(-1)))
hid = ops._linear_fwd(yk, Wc1, bc1, cfg, torch.float32, relu=True, row_scale=inv, rows_per_group=L)
print("hid", O.rel_err(hid, h2.reshape(M, Ah)))
dlog = cl.grad.reshape(M, G).float().contiguous()
dpre_s, dWc2, dbc2, dbc1 = ops.attn_logits_bwd(hid, Wc2, dlog, torch.float32, out_scale=inv, rows_per_group=L)
g = ops._dgrad(dpre_s, Wc1, cfg, out_dtype=torch.float32)
g_ref = (yhat.grad / nrm).reshape(M, 1000)
print("g", O.rel_err(g, g_ref))
t = ops.group_dot(g, yk, N, L)
print("t", O.rel_err(t, (y * (yhat.grad / nrm)).sum(dim=(1, 2))))
dI, dQ1, dbimg = ops.mfb_bwd(g, yk, inv, t, Q1k, keep, L, torch.float32, 0.0, 0)
print("dy-check: z.grad norm", float(z.grad.norm()))
dI_ref = (z.grad.reshape(M, 1000).repeat_interleave(5, 1) * Q1.detach().repeat_interleave(L, 0))
print("dI", O.rel_err(dI, dI_ref), "dbimg", O.rel_err(dbimg, P64["bimg"].grad))
