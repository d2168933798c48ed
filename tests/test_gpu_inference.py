"""GPU tests of the eval path: the fused log-softmax + argmax classifier tail, the no_grad
validation loop (solver.py:119-182) and its CUDA-graph replay."""
import types

import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("M,N", [(1, 3000), (256, 3000), (37, 56), (5, 7)])
def test_classifier_tail_matches_torch(M, N):
    from vqa_attention_networks_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N)
    x = torch.randn(M, N, device=DEV, generator=g) * 4.0
    x[0, N // 2] = x[0].max() + 1.0
    if M > 1:
        x[1, 3] = x[1, N - 1] = x[1].max() + 2.0                # a tie: torch.argmax returns the lowest index
    logp, pred, plp = ops.log_softmax_argmax(x)
    ref = F.log_softmax(x.double(), dim=1)
    assert O.rel_err(logp, ref) < 1e-6
    assert torch.equal(pred, x.argmax(1))
    assert torch.allclose(plp.double(), ref.gather(1, pred[:, None])[:, 0], atol=1e-5)
    # a strided view of a wider buffer (ld > N), predictions only
    wide = torch.randn(M, N + 8, device=DEV, generator=g)
    _, pred2, _ = ops.log_softmax_argmax(wide[:, :N], want_logp=False)
    assert torch.equal(pred2, wide[:, :N].argmax(1))


def _model():
    from vqa_attention_networks_b200 import MHBCoAtt
    cfg = types.SimpleNamespace(model_name="mhb_coAtt", q_vocab_size=200, emb_dim=32, hidden_dim=128, num_layers=1,
                                img_feature_channel=256, img_feature_dim=49, a_vocab_size=56, glove=False)
    torch.manual_seed(0)
    m = MHBCoAtt(cfg)
    for n, p in m.named_parameters():
        if n.find("bias") == -1:
            torch.nn.init.xavier_uniform_(p)
    with torch.no_grad():
        m.linear_pred.weight.mul_(32.0)                         # sharpen the nearly flat Xavier logits
    return m.to(DEV)


def test_eval_forward_uses_the_fused_tail_and_train_forward_does_not():
    m = _model().eval()
    X = O.synthetic_inputs(8, 49, 256, 26, 200, seed=3, device=DEV)
    with torch.no_grad():
        out = m(X["img"], X["questions"])
    assert m.last_pred is not None and torch.equal(m.last_pred, out.argmax(1))
    assert torch.allclose(out.exp().sum(1), torch.ones(8, device=DEV), atol=1e-5)
    m.train()
    out2 = m(X["img"], X["questions"])
    assert m.last_pred is None and out2.requires_grad            # autograd path: the stock op


def test_evaluate_is_the_val_loop_of_the_solver():
    """solver.py:119-182 by hand (eval(), forward, KLDivLoss, softmax().max(1)[1], soft answers' arg-max as labels)
    against inference.evaluate, eager and through GraphedForward."""
    from vqa_attention_networks_b200.inference import GraphedForward, evaluate
    m = _model()
    crit = torch.nn.KLDivLoss()
    batches = []
    for i in range(3):
        X = O.synthetic_inputs(8, 49, 256, 26, 200, seed=40 + i, device=DEV)
        batches.append((X["img"], X["questions"], O.soft_answers(8, 56, seed=50 + i).to(DEV)))
    res = evaluate(m, batches, crit)
    assert not m.training and res["n"] == 24
    tot, correct = 0.0, 0
    with torch.no_grad():
        for img, q, a in batches:
            logits = m.forward(img, q)
            tot += float(crit(logits, a))
            pred = F.softmax(logits, dim=1).max(1)[1]
            correct += int((pred == a.max(1)[1]).sum())
    # two separate forwards: equal up to the atomics' summation order (see test_gpu_solver_loop.py)
    assert abs(res["loss"] - tot / 3) < 1e-5 * max(1.0, abs(tot)) and abs(res["acc"] - correct / 24) < 1e-9
    g = GraphedForward(m, batches[0][0], batches[0][1])
    for img, q, a in batches:
        out = g(img, q)
        with torch.no_grad():
            eager = m(img, q)
        assert O.rel_err(out, eager) < 1e-4          # fp32 atomics (split-K, sum |z|) order differs between runs
        assert torch.equal(g.pred, eager.argmax(1))
    hard = [(b[0], b[1], b[2].max(1)[1]) for b in batches]
    assert abs(evaluate(m, hard)["acc"] - res["acc"]) < 1e-9 and evaluate(m, hard)["loss"] is None


def test_custom_ops_pass_opcheck():
    """torch.library.opcheck on two of the dispatcher ops (schema / declared mutation / fake implementation agree with
    what the kernels really do), and the direct path gives the same bits as the dispatcher path."""
    from vqa_attention_networks_b200 import ops
    ops._register_custom_ops()
    x = torch.randn(64, 40, device=DEV)
    out = torch.empty(64, 40, device=DEV, dtype=torch.bfloat16)
    args = (x, out, 1, 64, 40, 0, 40, 1, 0, 40)
    torch.library.opcheck(torch.ops.vqa_b200.pack_bf16.default, args, test_utils=("test_schema", "test_faketensor"))
    torch.ops.vqa_b200.pack_bf16(*args)
    assert torch.equal(out, x.to(torch.bfloat16))
    out2 = torch.empty_like(out)
    ops.launch_direct("vqa_b200_pack_bf16", (x, out2, 1, 64, 40, 0, 40, 1, 0, 40))
    assert torch.equal(out2, out)
    ssq = torch.rand(33, device=DEV) + 0.1
    inv = torch.empty_like(ssq)
    torch.library.opcheck(torch.ops.vqa_b200.inv_norm.default, (ssq, inv, 33), test_utils=("test_schema", "test_faketensor"))
    torch.ops.vqa_b200.inv_norm(ssq, inv, 33)
    assert torch.allclose(inv, ssq.rsqrt(), rtol=1e-6)
