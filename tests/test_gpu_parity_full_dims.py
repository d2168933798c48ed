"""GPU parity at BASELINE.json dimensions (run on the B200 with ``-m gpu``): every model family of the path against
the fp64 oracle evaluated ON THE GPU as the checker (same seeded inputs and weights, the kernels' own dropout masks
injected), forward and every parameter gradient, fp32 and bf16 modes.

  config 1 / 4  MFB('mfb') and MFB('mfb-multilayer'): hidden 1024, D 2048, L 196, T 26 (mfb.py:61-140), default
                (degenerate-softmax) behaviour, eval-like and train mode;
  config 2      MHBCoAtt at batch 256 (mhb_coAtt.py:61-151): output error and top-1 agreement vs the oracle;
                train mode with masks at batch 6;
  config 3      HieCoAtten(block 196, word 26, img 2048, E 512, A 3000) (hieCoAtten.py:18-55), batch 6 and 1:
                the multi-tile batched products with padded pitches (L = 196 -> 200, T = 26 -> 32).

Every measured error is recorded (tests/_parity_util.record -> gpurun_out/parity_measured.json; the table in
DESIGN.md section 4 is that file).
"""
import types

import pytest
import torch

from _parity_util import (CENTRED_TOL, DEV, GRAD_TOL, HIE_GRAD_TOL, OUT_TOL, centred, check_grads, mhb_masks, record,
                          xavier_, z_from_capture)
from oracle import oracle as O

pytestmark = pytest.mark.gpu

L, D, T, H, V, A, E = 196, 2048, 26, 1024, 15000, 3000, 512

# see tests/test_gpu_parity.py: doubly-cancelling gradients of the question attention's first conv (sum_t dlogits == 0
# and a nearly t-constant ReLU mask leave a ~1e-4 residual of the terms).  fp32 mode needs no allowance (measured 3.8e-5:
# the kernels' maths is right); under bf16 rounding a single ReLU sign flip moves the residual by tens of percent
# (measured 0.24 / 0.30 at batch 6 in train mode), so the bf16 bound only says "same direction and scale".
ILL_CONDITIONED = {"fp32": {}, "bf16": {"ques_att_conv1.weight": 0.5, "ques_att_conv1.bias": 0.5}}


@pytest.fixture(autouse=True)
def _strict_fp32_stock_ops():
    """The stock cuDNN LSTM (MFB, and MHBCoAtt in fp32 mode: left as-is by north_star) may use TF32 by default; the
    fp32-mode contract (1e-4) is about this repo's kernels, so the stock ops around them are held to true fp32."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _cfg(name, Lr=L):
    return types.SimpleNamespace(model_name=name, q_vocab_size=V, emb_dim=300, hidden_dim=H, num_layers=1,
                                 img_feature_channel=D, img_feature_dim=Lr, a_vocab_size=A, glove=False)


def _sd64(model, grad=False):
    return {k: v.detach().double().clone().requires_grad_(grad and v.is_floating_point())
            for k, v in model.state_dict().items()}


def _fixed_seeds(monkeypatch, ops, seeds):
    it = iter(seeds)
    used = []

    def fake():
        s = next(it)
        used.append(s)
        return s

    monkeypatch.setattr(ops, "new_seed", fake)
    return used


# ------------------------------------------------------------------------------------------------------------
# configs 1 / 4: MFB and MFB-multilayer
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("train_masks", [False, True])
@pytest.mark.parametrize("name", ["mfb", "mfb-multilayer"])
def test_mfb_full_dims_vs_oracle(name, train_masks, monkeypatch):
    """mfb.py:61-140 at hidden 1024 / D 2048 / L 196 / T 26, batch 6, in the reference's default (degenerate softmax)
    behaviour: both glimpses are sum-pools, the first stage is dead (exactly-zero gradients), the MFB vector block
    (ques_proj2 x img_proj2, K = 2048 / 4096, dropout in train mode) carries everything."""
    from vqa_attention_networks_b200 import MFB, ops
    N = 6
    model = xavier_(MFB(_cfg(name))).to(DEV).train()
    if not train_masks:
        model.dropout_m.p = 0.0
    X = O.synthetic_inputs(N, L, D, T, V, seed=4321, device=DEV)
    cot = torch.randn(N, A, device=DEV, generator=torch.Generator(device=DEV).manual_seed(6))
    multi = name == "mfb-multilayer"
    tag = "c%s_%s_batch6_%s" % ("4" if multi else "1", name, "train" if train_masks else "eval")
    for mode in ("fp32", "bf16"):
        model.precision = mode
        model.zero_grad(set_to_none=True)
        model.capture = {}
        # LSTM-output dropout (mfb.py:70): bf16 mode applies it inside the recurrence kernels with a counter-hash mask that
        # can be injected into the oracle; fp32 mode runs the stock nn.LSTM + nn.Dropout (torch's RNG): switched off there
        model.dropout_l.p = 0.3 if (train_masks and mode == "bf16") else 0.0
        used = _fixed_seeds(monkeypatch, ops, [901, 902, 903, 904])
        out = model(X["img"], X["questions"])
        masks = {}
        if train_masks:
            # the spatial stage draws a seed too (its mask is dead code in degenerate mode); the last one is the vector block's
            masks["m2"] = ops.dropout_mask(N, 5000, 0.1, used[-1], DEV).double()
            if mode == "bf16":
                assert model.last_lstm_drop_seed == used[0]
                # mask rows are time-major: T steps over N rows -> the oracle's [N, T, H] output
                masks["l"] = ops.dropout_mask(T * N, H, 0.3, used[0], DEV).double().view(T, N, H).permute(1, 0, 2)
        with torch.no_grad():
            ref = O.mfb_forward(_sd64(model), X["img"].double(), X["questions"], multi, masks)
        raw, cen = O.rel_err(out, ref), O.rel_err(centred(out.double()), centred(ref))
        record(tag, mode + ":out_raw", raw)
        record(tag, mode + ":out_centred", cen)
        assert raw < OUT_TOL[mode], (mode, raw)
        assert cen < CENTRED_TOL[mode], (mode, cen)
        (out * cot).sum().backward()
        P64 = _sd64(model, grad=True)
        inj = {**masks, **z_from_capture(model.capture, N)}
        ref2 = O.mfb_forward(P64, X["img"].double(), X["questions"], multi, inj)
        (ref2 * cot.double()).sum().backward()
        worst, wname = check_grads(model, {k: v.grad for k, v in P64.items()}, GRAD_TOL[mode], tag=tag + ":" + mode)
        record(tag, mode + ":grad_worst", worst)
        # the dead first stage: exact zeros, as the reference produces (mfb.py:84,118)
        for dead in ("img_conv1d.weight", "ques_proj1.weight", "co_att_conv1.weight", "ques_att_conv1.weight"):
            g = dict(model.named_parameters())[dead].grad
            assert g is not None and float(g.abs().max()) == 0.0, dead


# ------------------------------------------------------------------------------------------------------------
# config 2: MHBCoAtt
# ------------------------------------------------------------------------------------------------------------
def test_mhbcoatt_full_dims_train_masks_vs_oracle(monkeypatch):
    """mhb_coAtt.py:61-151 in train mode at full dimensions, batch 6: the three fused-epilogue dropouts (grid MFB,
    two vector blocks) with the kernels' masks injected into the fp64 oracle; forward and all gradients."""
    from vqa_attention_networks_b200 import MHBCoAtt, ops
    N = 6
    model = xavier_(MHBCoAtt(_cfg("mhb_coAtt"))).to(DEV).train()
    X = O.synthetic_inputs(N, L, D, T, V, seed=99, device=DEV)
    cot = torch.randn(N, A, device=DEV, generator=torch.Generator(device=DEV).manual_seed(8))
    tag = "c2_mhbcoatt_batch6_train"
    for mode in ("fp32", "bf16"):
        model.precision = mode
        model.zero_grad(set_to_none=True)
        model.capture = {}
        # LSTM-output dropout (mhb_coAtt.py:75): in bf16 mode it runs inside the recurrence kernel (first seed drawn) and its
        # mask is injected like the others; fp32 mode keeps the stock nn.LSTM + nn.Dropout (torch's RNG): switched off
        model.dropout_l.p = 0.3 if mode == "bf16" else 0.0
        used = _fixed_seeds(monkeypatch, ops, [10, 11, 12] if mode == "bf16" else [11, 12, 13])
        out = model(X["img"], X["questions"])
        assert len(used) == 3                                 # fp32: three blocks; bf16: LSTM dropout + grid + one vector launch
        lmask = None
        if mode == "bf16":
            assert model.last_lstm_drop_seed == used[0]
            # the recurrence runs over the batch axis: N steps over T rows -> the oracle's [T, N, H]
            lmask = ops.dropout_mask(N * T, H, 0.3, used[0], DEV).double().view(N, T, H).permute(1, 0, 2)
            used = used[1:]
        masks = mhb_masks(ops, used, N, L)
        if lmask is not None:
            masks["l"] = lmask
        with torch.no_grad():
            ref = O.mhbcoatt_forward(_sd64(model), X["img"].double(), X["questions"], None, masks)
        raw, cen = O.rel_err(out, ref), O.rel_err(centred(out.double()), centred(ref))
        record(tag, mode + ":out_raw", raw)
        record(tag, mode + ":out_centred", cen)
        assert raw < OUT_TOL[mode], (mode, raw)
        assert cen < CENTRED_TOL[mode], (mode, cen)
        (out * cot).sum().backward()
        P64 = _sd64(model, grad=True)
        ref2 = O.mhbcoatt_forward(P64, X["img"].double(), X["questions"], None,
                                  {**masks, **z_from_capture(model.capture, N)})
        (ref2 * cot.double()).sum().backward()
        worst, _ = check_grads(model, {k: v.grad for k, v in P64.items()}, GRAD_TOL[mode], loose=ILL_CONDITIONED[mode],
                               tag=tag + ":" + mode)
        record(tag, mode + ":grad_worst", worst)
        # For information: the same gradients against the oracle evaluated at ITS OWN z -- no injection.
        # d(signed-sqrt) = 1/(2 sqrt|z|) makes this a log-divergent comparison (see tests/test_gpu_parity.py), so the
        # figures are recorded, and only fp32 mode's direction is asserted: under bf16 rounding of z the handful of smallest
        # |z| -- which carry most of the norm -- are garbled, and the un-injected img_conv1d gradient keeps a cosine of only
        # ~0.3 with the fp64 one (measured; the reference's own fp32 gradient moves by tens of percent under bf16 / TF32).
        P64b = _sd64(model, grad=True)
        ref3 = O.mhbcoatt_forward(P64b, X["img"].double(), X["questions"], None, masks)
        (ref3 * cot.double()).sum().backward()
        for name in ("img_conv1d.weight", "ques_proj1.weight", "img_proj2.weight", "co_att_conv1.weight",
                     "linear_pred.weight", "lstm.weight_hh_l0"):
            got = dict(model.named_parameters())[name].grad.double().reshape(-1)
            ref = P64b[name].grad.reshape(-1)
            cos = float(torch.dot(got, ref) / (got.norm() * ref.norm()))
            record(tag + ":" + mode, "grad_uninjected:" + name, O.rel_err(got, ref))
            record(tag + ":" + mode, "grad_uninjected_cosine:" + name, cos)
            if mode == "fp32":
                assert cos > 0.98, (name, cos)                        # measured: 0.9956 (img_conv1d) .. 1.0000


def test_config2_batch256_vs_fp64_oracle():
    """BASELINE config 2 at its real size (MHBCoAtt, batch 256, L = 196), eval forward, against the fp64 oracle run on
    the same GPU as the checker: output error (raw and centred) in both modes, and top-1 answer agreement of the bf16
    mode with the ORACLE.  Xavier weights give nearly flat logits (top-1 / top-2 margin ~ 0.006), so the
    classifier weight is sharpened x32 in the state dict both sides load; agreement is asserted on the samples whose
    fp64 margin exceeds twice the largest centred-logit error seen (a sample inside that band has no defined winner
    at the contract's precision) and the unfiltered figure is recorded next to it."""
    from vqa_attention_networks_b200 import MHBCoAtt
    N = 256
    model = xavier_(MHBCoAtt(_cfg("mhb_coAtt")))
    with torch.no_grad():
        model.linear_pred.weight.mul_(32.0)
    model = model.to(DEV).eval()
    X = O.synthetic_inputs(N, L, D, T, V, seed=1234, device=DEV)
    tag = "c2_mhbcoatt_batch256_eval"
    with torch.no_grad():
        ref = O.mhbcoatt_forward(_sd64(model), X["img"].double(), X["questions"])
        top2 = ref.topk(2, dim=1)
        margin = top2.values[:, 0] - top2.values[:, 1]
        ref_top1 = top2.indices[:, 0]
        record(tag, "oracle_margin_median", float(margin.median()))
        for mode in ("fp32", "bf16"):
            model.precision = mode
            out = model(X["img"], X["questions"]).double()
            raw, cen = O.rel_err(out, ref), O.rel_err(centred(out), centred(ref))
            err = (centred(out) - centred(ref)).abs().max()
            agree = (out.argmax(1) == ref_top1)
            sure = margin > 2 * err
            filtered = float(agree[sure].double().mean()) if bool(sure.any()) else 0.0
            record(tag, mode + ":out_raw", raw)
            record(tag, mode + ":out_centred", cen)
            record(tag, mode + ":max_abs_centred_err", float(err))
            record(tag, mode + ":top1_agreement_all", float(agree.double().mean()))
            record(tag, mode + ":top1_agreement_margin_filtered", filtered)
            record(tag, mode + ":margin_filter_kept_frac", float(sure.double().mean()))
            assert raw < OUT_TOL[mode], (mode, raw)
            assert cen < CENTRED_TOL[mode], (mode, cen)
            assert filtered >= 0.995, (mode, filtered)
            assert float(sure.double().mean()) >= 0.8, (mode, float(sure.double().mean()))
            assert float(agree.double().mean()) >= 0.98, (mode, float(agree.double().mean()))


# ------------------------------------------------------------------------------------------------------------
# config 3: HieCoAtten
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,init", [(6, "default"), (6, "xavier"), (1, "default")])
def test_hiecoatten_full_dims_vs_oracle(N, init, monkeypatch):
    """hieCoAtten.py:18-55 at (block 196, word 26, img 2048, E 512, A 3000): img_emb GEMM M = N*196, K = 2048; the
    three per-sample products over padded pitches (L 196 -> 200, T 26 -> 32) with several row tiles per sample and
    their backward (multi-tile accumulate wgrads); the five always-on dropouts with the kernels' masks injected.
    'default' is torch's own init as train_hfd.py:62-66 uses it (N(0,1) embeddings: the affinity tanh saturates),
    'xavier' the train_models.py recipe (small embeddings: the affinity is near-linear)."""
    from vqa_attention_networks_b200 import HieCoAtten, ops
    torch.manual_seed(3)
    model = HieCoAtten(block_num=L, word_num=T, img_size=D, vocab_size=V, embed_size=E, output_size=A)
    if init == "xavier":
        xavier_(model, seed=3)
    model = model.to(DEV).eval()                # eval() must NOT switch the functional dropouts off
    X = O.synthetic_inputs(N, L, D, T, V, seed=55, device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(9)
    cot = torch.randn(N, A, device=DEV, generator=gen)
    cot_av = torch.randn(N, L, device=DEV, generator=gen)
    cot_aq = torch.randn(N, T, device=DEV, generator=gen)
    tag = "c3_hiecoatten_batch%d_%s" % (N, init)
    shapes = [(N * L, E), (N * T, E), (N * T, L), (N * L, E), (N * T, E)]
    views = [(N, L, E), (N, T, E), (N, T, L), (N, L, E), (N, T, E)]
    for mode in ("fp32", "bf16"):
        model.precision = mode
        model.zero_grad(set_to_none=True)
        _fixed_seeds(monkeypatch, ops, [21, 22, 23, 24, 25])
        x, av, aq = model(X["img"], X["questions"])
        assert model.last_seeds == [21, 22, 23, 24, 25]
        masks = [ops.dropout_mask(r, c, 0.5, s, DEV).double().reshape(v)
                 for (r, c), v, s in zip(shapes, views, model.last_seeds)]
        P64 = _sd64(model, grad=True)
        rx, rav, raq = O.hiecoatten_forward(P64, X["img"].double(), X["questions"], masks)
        if N == 1:                              # torch.squeeze drops the batch axis (hieCoAtten.py:42-50)
            assert tuple(av.shape) == (L,) and tuple(aq.shape) == (T,) and tuple(x.shape) == (1, A)
        else:
            assert tuple(av.shape) == (N, L) and tuple(aq.shape) == (N, T) and tuple(x.shape) == (N, A)
        for nme, got, want in (("x", x, rx), ("av", av.reshape(N, L), rav), ("aq", aq.reshape(N, T), raq)):
            e = O.rel_err(got, want)
            record(tag, "%s:out_%s" % (mode, nme), e)
            assert e < OUT_TOL[mode], (mode, nme, e)
        cen = O.rel_err(centred(x.double()), centred(rx))
        record(tag, mode + ":out_x_centred", cen)
        assert cen < CENTRED_TOL[mode], (mode, cen)
        ((x * cot).sum() + (av.reshape(N, L) * cot_av).sum() + (aq.reshape(N, T) * cot_aq).sum()).backward()
        ((rx * cot.double()).sum() + (rav * cot_av.double()).sum() + (raq * cot_aq.double()).sum()).backward()
        assert model.fc_Wbq.weight.grad is None          # dead layer (hieCoAtten.py:30-31)
        refg = {k: v.grad for k, v in P64.items()}
        worst = 0.0
        for name, p in model.named_parameters():
            if name.startswith("fc_Wbq"):
                assert refg[name] is None
                continue
            assert p.grad is not None, name
            if float(refg[name].norm()) < 1e-9:          # fc_Whv / fc_Whq biases in front of a softmax
                assert float(p.grad.norm()) < 1e-4, name
                continue
            e = O.rel_err(p.grad, refg[name])
            record(tag + ":" + mode, "grad:" + name, e)
            worst = max(worst, e)
            assert e < HIE_GRAD_TOL[mode], (mode, name, e)
        record(tag, mode + ":grad_worst", worst)
