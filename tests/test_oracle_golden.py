"""Pin the oracle (oracle/oracle.py) against the golden vectors produced by the REAL reference
modules (oracle/gen_golden.py).  CPU only.  Tolerances: the reference ran in fp32, so the fp32
oracle must agree to 2e-5 (different summation order) and the fp64 oracle to 2e-5 as well."""
import pytest
import torch

from oracle import fixtures, oracle as O

FIX = fixtures.list_fixtures()
TOL = 2e-5
GTOL = 1e-3   # reference grads are fp32 through 1/sqrt|z| (signed-sqrt) amplification


def _rebuild(rec, dtype):
    case = rec["case"]
    P = fixtures.make_params(case["shapes"], case["param_seed"])
    X = fixtures.make_inputs(case)
    cs = fixtures.checksum({**P, **{k: v.double() for k, v in X.items()}})
    assert abs(cs - case["checksum"]) <= 1e-9 * max(1.0, abs(case["checksum"])), "seeded tensors differ from generation time"
    P = {k: v.to(dtype).requires_grad_(v.is_floating_point()) for k, v in P.items()}
    X = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in X.items()}
    return case, P, X


def _check_grads(rec, P, tol):
    for k, g in rec["grads"].items():
        if g is None:
            assert P[k].grad is None or float(P[k].grad.abs().max()) == 0.0, k
            continue
        got = P[k].grad if P[k].grad is not None else torch.zeros_like(P[k])
        if g["norm"] == 0.0:
            assert float(got.abs().max()) < 1e-7, k   # exactly-dead branches
            continue
        if g["norm"] < 1e-6:                          # shift-invariant biases: rounding noise only
            assert float(got.norm()) < 1e-5, k
            continue
        assert fixtures.compare_subsample(got, g) < tol, (k, fixtures.compare_subsample(got, g))


def test_fixture_inventory():
    need = {"mhbcoatt_eval", "mhbcoatt_train_masks", "mhbcoatt_glove_eval", "mfb_eval", "mfb_multilayer_eval",
            "mfb_train_masks", "hiecoatten_n4", "hiecoatten_n3", "hiecoatten_n1", "attention_1_n2", "attention_2_n2",
            "attention_layer_1_n2", "attention_layer_2_n2", "nonlinear_layer_n2", "attention_1_n1", "mhb_patched_eval"}
    assert need <= set(FIX)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", [f for f in FIX if f.startswith(("mhbcoatt", "mfb"))])
def test_coatt_models(name, dtype):
    rec = fixtures.load_fixture(name)
    case, P, X = _rebuild(rec, dtype)
    masks = None
    if case["train_masks"]:
        masks = {"l": X["mask_l"], "m1": X["mask_m1"], "m2": X["mask_m2"], "m3": X.get("mask_m3")}
    if case["model"] == "mhbcoatt":
        out = O.mhbcoatt_forward(P, X["img"], X["questions"], X.get("glove"), masks)
    else:
        out = O.mfb_forward(P, X["img"], X["questions"], case["cfg"]["model_name"] == "mfb-multilayer", masks)
    assert O.rel_err(out, rec["outputs"]["out"]) < TOL
    (out * X["cot"]).sum().backward()
    # two fp32 evaluations of d(signed-sqrt) = 1/(2 sqrt|z|) differ by up to ~4e-3 on the first-stage
    # parameters (measured: reference fp32 vs fp64 truth 1.7e-4, fp32 vs fp32 3.6e-3)
    _check_grads(rec, P, GTOL if dtype == torch.float64 else 1e-2)


@pytest.mark.parametrize("name", [f for f in FIX if f.startswith("hiecoatten")])
def test_hiecoatten(name):
    rec = fixtures.load_fixture(name)
    case, P, X = _rebuild(rec, torch.float64)
    N, L, T = case["N"], case["ctor"]["block_num"], case["T"]
    x, av, aq = O.hiecoatten_forward(P, X["img"], X["questions"], [X["mask%d" % i] for i in range(5)])
    assert O.rel_err(x, rec["outputs"]["x"]) < TOL
    assert O.rel_err(av, rec["outputs"]["av"].reshape(N, L)) < TOL
    assert O.rel_err(aq, rec["outputs"]["aq"].reshape(N, T)) < TOL
    ((x * X["cot"]).sum() + (av * X["cot_av"]).sum() + (aq * X["cot_aq"]).sum()).backward()
    _check_grads(rec, P, GTOL)
    assert rec["grads"]["fc_Wbq.weight"] is None      # dead layer (hieCoAtten.py:30-31)


@pytest.mark.parametrize("name", [f for f in FIX if f.startswith(("attention", "nonlinear"))])
def test_modules(name):
    rec = fixtures.load_fixture(name)
    case, P, X = _rebuild(rec, torch.float64)
    f1 = X["f1"].clone().requires_grad_(True)
    f2 = X["f2"].clone().requires_grad_(True)
    kind = case["model"]
    outs = rec["outputs"]
    if kind == "nonlinear_layer":
        o = O.nonlinear_layer(P, f1)
        assert O.rel_err(o, outs["o"]) < TOL
        (o * X["cot_f1"]).sum().backward()
    elif kind.startswith("attention_layer"):
        a, b, att = O.attention_layer(P, f1, f2, int(kind[-1]))
        for got, key in ((a, "f1e"), (b, "f2e"), (att, "att")):
            assert O.rel_err(got, outs[key].reshape(got.shape)) < TOL
        ((a * X["cot_f1"]).sum() + (b * X["cot_f"]).sum() + (att * X["cot_att"]).sum()).backward()
    else:
        f_hat, att = (O.attention_1 if kind == "attention_1" else O.attention_2)(P, f1, f2)
        assert O.rel_err(f_hat, outs["f_hat"]) < TOL
        assert O.rel_err(att, outs["att"].reshape(att.shape)) < TOL
        ((f_hat * X["cot_f"]).sum() + (att * X["cot_att"]).sum()).backward()
    assert O.rel_err(f1.grad, outs["d_f1"]) < GTOL
    if "d_f2" in outs and float(outs["d_f2"].abs().max()) > 1e-6:
        assert O.rel_err(f2.grad, outs["d_f2"]) < GTOL
    _check_grads(rec, P, GTOL)


def test_mhb_patched():
    rec = fixtures.load_fixture("mhb_patched_eval")
    case, P, X = _rebuild(rec, torch.float64)
    out = O.mhb_forward(P, X["img"], X["questions"], case["q_length"])
    assert O.rel_err(out, rec["outputs"]["out"]) < TOL
    (out * X["cot"]).sum().backward()
    _check_grads(rec, P, GTOL)


def test_mfb_dead_branch_pinned():
    """mfb.py's softmax over a singleton axis kills the gradient of the first stage."""
    rec = fixtures.load_fixture("mfb_eval")
    for k in ("img_conv1d.weight", "ques_proj1.weight", "co_att_conv1.weight", "ques_att_conv1.weight"):
        assert rec["grads"][k] is not None and rec["grads"][k]["norm"] == 0.0
