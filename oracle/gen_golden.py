"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.pt`` by running the REAL reference.

Run in the build container (where /root/reference is mounted):

    python oracle/gen_golden.py

For each case the unmodified reference module (imported by ``oracle/ref_loader.py`` under the
``view -> reshape`` shim) is loaded with deterministic parameters
(``oracle/fixtures.make_params``), run forward on seeded inputs, and back-propagated through a
seeded cotangent.  Outputs and gradients are stored; parameters and inputs are *not* (they are
rebuilt from the seeds in the tests and verified against the stored checksum).

Dropout: ``eval`` cases run with dropout disabled; ``masks`` cases run in train mode with
``F.dropout`` replaced by multiplication with injected pre-scaled masks (the only way to make
the reference's always-on dropouts in hieCoAtten.py:26-46 reproducible).

``MHB`` (mhb_coAtt.py:153-217) is broken as shipped (hard ``.cuda()`` at :176, undefined
``mhb_22`` at :214); its case executes an in-memory patched copy (two token edits) and is
labelled ``patched-oracle``.  No reference source is written to this repository.
"""
from __future__ import annotations

import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures, ref_loader  # noqa: E402


def _cfg(**kw):
    base = dict(model_name="mhb_coAtt", q_vocab_size=20, emb_dim=6, hidden_dim=8, num_layers=1,
                img_feature_channel=16, img_feature_dim=6, a_vocab_size=7, glove=False)
    base.update(kw)
    return base


def _shapes(model):
    return {k: list(v.shape) for k, v in model.state_dict().items()}


def _grads(model):
    out = {}
    for k, p in model.named_parameters():
        out[k] = None if p.grad is None else fixtures.subsample(p.grad)
    return out


def _finish(name, case, model, outputs, extra=None):
    rec = {"case": case, "outputs": {k: v.detach().clone() for k, v in outputs.items()},
           "grads": _grads(model), "torch": torch.__version__}
    if extra:
        rec.update(extra)
    path = fixtures.save_fixture(name, rec)
    print("wrote", path, os.path.getsize(path), "bytes")


def _prep(case, model):
    case["shapes"] = _shapes(model)
    P = fixtures.make_params(case["shapes"], case["param_seed"])
    model.load_state_dict(P)
    X = fixtures.make_inputs(case)
    case["checksum"] = fixtures.checksum({**P, **{k: v.double() for k, v in X.items()}})
    return P, X


def gen_coatt(name, model_kind, cfg, N, T, train_masks=False, glove=False):
    """MHBCoAtt / MFB cases."""
    ns = types.SimpleNamespace(**cfg)
    L, D, H, A = cfg["img_feature_dim"], cfg["img_feature_channel"], cfg["hidden_dim"], cfg["a_vocab_size"]
    if model_kind == "mhbcoatt":
        model = ref_loader.load("mhb_coAtt").MHBCoAtt(ns)
    else:
        model = ref_loader.load("mfb").MFB(ns)
    inputs = {"img": ["relu_randn", [N, L, D]], "questions": ["randint", [N, T], cfg["q_vocab_size"]],
              "cot": ["randn", [N, A]]}
    if glove:
        inputs["glove"] = ["randn", [N, T, cfg["emb_dim"]]]
    n_vec = 2 if model_kind == "mhbcoatt" else 1
    if train_masks:
        lshape = [T, N, H] if model_kind == "mhbcoatt" else [N, T, H]
        inputs["mask_l"] = ["mask", lshape, 0.3]
        inputs["mask_m1"] = ["mask", [N, L, 5000], 0.1]
        for i in range(n_vec):
            inputs["mask_m%d" % (i + 2)] = ["mask", [N, 5000], 0.1]
    case = {"model": model_kind, "cfg": cfg, "N": N, "T": T, "param_seed": 11, "input_seed": 23,
            "inputs": inputs, "train_masks": train_masks}
    P, X = _prep(case, model)
    args = [X["img"], X["questions"]]
    kwargs = {}
    if glove:
        kwargs["glove_matrix"] = X["glove"]
    if train_masks:
        model.train()
        # call order in forward: dropout_l, dropout_m (grid, reference layout [N,5000,L,1]), dropout_m per vector block
        masks = [X["mask_l"], X["mask_m1"].permute(0, 2, 1).unsqueeze(3)] + [X["mask_m%d" % (i + 2)] for i in range(n_vec)]
        ctx = ref_loader.injected_dropout(masks)
    else:
        model.eval()
        import contextlib
        ctx = contextlib.nullcontext()
    with ref_loader.view_shim(), ctx:
        out = model(*args, **kwargs)
        (out * X["cot"]).sum().backward()
    _finish(name, case, model, {"out": out})


def gen_hie(name, N, L=6, T=5, D=16, E=8, V=20, A=7, block_kw=None):
    hc = ref_loader.load("hieCoAtten")
    model = hc.HieCoAtten(block_num=L, word_num=T, img_size=D, vocab_size=V, embed_size=E, output_size=A)
    inputs = {"img": ["relu_randn", [N, L, D]], "questions": ["randint", [N, T], V],
              "cot": ["randn", [N, A]], "cot_av": ["randn", [N, L]], "cot_aq": ["randn", [N, T]],
              "mask0": ["mask", [N, L, E], 0.5], "mask1": ["mask", [N, T, E], 0.5],
              "mask2": ["mask", [N, T, L], 0.5], "mask3": ["mask", [N, L, E], 0.5],
              "mask4": ["mask", [N, T, E], 0.5]}
    case = {"model": "hiecoatten", "ctor": dict(block_num=L, word_num=T, img_size=D, vocab_size=V,
                                                 embed_size=E, output_size=A),
            "N": N, "T": T, "param_seed": 5, "input_seed": 29, "inputs": inputs}
    P, X = _prep(case, model)
    with ref_loader.injected_dropout([X["mask%d" % i] for i in range(5)]):
        x, av, aq = model(X["img"], X["questions"])
        loss = (x * X["cot"]).sum() + (av.reshape(N, L) * X["cot_av"]).sum() + (aq.reshape(N, T) * X["cot_aq"]).sum()
        loss.backward()
    _finish(name, case, model, {"x": x, "av": av, "aq": aq})


def gen_modules(name, kind, N, L=6, T=5, D=8):
    mods = ref_loader.load("modules")
    if kind == "attention_1":
        model = mods.Attention_1(D)
    elif kind == "attention_2":
        model = mods.Attention_2(D)
    elif kind == "attention_layer_1":
        model = mods.Attention_layer(D, 1)
    elif kind == "attention_layer_2":
        model = mods.Attention_layer(D, 2)
    elif kind == "nonlinear_layer":
        model = mods.Nonlinear_layer(D)
    else:
        raise ValueError(kind)
    inputs = {"f1": ["randn", [N, L, D]], "f2": ["randn", [N, T, D]],
              "cot_f": ["randn", [N, T, D]], "cot_att": ["randn", [N, T, L]], "cot_f1": ["randn", [N, L, D]]}
    case = {"model": kind, "D": D, "N": N, "param_seed": 3, "input_seed": 31, "inputs": inputs}
    P, X = _prep(case, model)
    f1 = X["f1"].clone().requires_grad_(True)
    f2 = X["f2"].clone().requires_grad_(True)
    if kind == "nonlinear_layer":
        o = model(f1)
        (o * X["cot_f1"]).sum().backward()
        outs = {"o": o}
    elif kind.startswith("attention_layer"):
        a, b, att = model(f1, f2)
        ((a * X["cot_f1"]).sum() + (b * X["cot_f"]).sum() + (att * X["cot_att"]).sum()).backward()
        outs = {"f1e": a, "f2e": b, "att": att}
    else:
        f_hat, att = model(f1, f2)
        ((f_hat * X["cot_f"]).sum() + (att * X["cot_att"]).sum()).backward()
        outs = {"f_hat": f_hat, "att": att}
    outs["d_f1"] = f1.grad
    if f2.grad is not None:
        outs["d_f2"] = f2.grad
    _finish(name, case, model, outs)


def gen_mhb(name, N=3, T=5):
    """patched-oracle: two token edits applied to an in-memory copy of mhb_coAtt.py."""
    root = ref_loader.reference_root()
    src = open(os.path.join(root, "mhb_coAtt.py")).read()
    src = src.replace("dtype=torch.float).cuda()", "dtype=torch.float)").replace("self.linear_out(mhb_22)", "self.linear_out(mhb_12)")
    mod = types.ModuleType("_vqa_ref_mhb_patched")
    exec(compile(src, "mhb_coAtt_patched", "exec"), mod.__dict__)
    cfg = _cfg(model_name="mhb", img_feature_dim=196)
    model = mod.MHB(types.SimpleNamespace(**cfg))
    L, D, A = 196, cfg["img_feature_channel"], cfg["a_vocab_size"]
    inputs = {"img": ["relu_randn", [N, L, D]], "questions": ["randint", [N, T], cfg["q_vocab_size"]],
              "cot": ["randn", [N, A]]}
    case = {"model": "mhb", "cfg": cfg, "N": N, "T": T, "param_seed": 13, "input_seed": 37, "inputs": inputs,
            "q_length": [T, 2, 3][:N], "note": "patched-oracle (mhb_coAtt.py:176 .cuda() removed, :214 mhb_22->mhb_12)"}
    P, X = _prep(case, model)
    model.eval()
    with ref_loader.view_shim():
        out = model(X["img"], X["questions"], torch.tensor(case["q_length"]))
        (out * X["cot"]).sum().backward()
    _finish(name, case, model, {"out": out})


def main():
    torch.manual_seed(0)
    assert ref_loader.available(), "reference not mounted"
    gen_coatt("mhbcoatt_eval", "mhbcoatt", _cfg(), N=3, T=5)
    gen_coatt("mhbcoatt_train_masks", "mhbcoatt", _cfg(), N=3, T=5, train_masks=True)
    gen_coatt("mhbcoatt_glove_eval", "mhbcoatt", _cfg(glove=True), N=2, T=4, glove=True)
    gen_coatt("mfb_eval", "mfb", _cfg(model_name="mfb"), N=3, T=5)
    gen_coatt("mfb_multilayer_eval", "mfb", _cfg(model_name="mfb-multilayer"), N=2, T=5)
    gen_coatt("mfb_train_masks", "mfb", _cfg(model_name="mfb"), N=2, T=4, train_masks=True)
    gen_hie("hiecoatten_n4", N=4)
    gen_hie("hiecoatten_n3", N=3)
    gen_hie("hiecoatten_n1", N=1)
    for kind in ("attention_1", "attention_2", "attention_layer_1", "attention_layer_2", "nonlinear_layer"):
        gen_modules(kind + "_n2", kind, N=2)
    gen_modules("attention_1_n1", "attention_1", N=1)
    gen_mhb("mhb_patched_eval")


if __name__ == "__main__":
    main()
