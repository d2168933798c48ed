"""Per-parameter gradient error table of the CUDA modules against the fp64 oracle (debug aid, run under gpurun)."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fixtures, oracle as O  # noqa: E402


def main():
    from vqa_attention_networks_b200 import MFB, MHBCoAtt
    names = sys.argv[1:] or ["mhbcoatt_eval", "mfb_multilayer_eval"]
    for name in names:
        rec = fixtures.load_fixture(name)
        case = rec["case"]
        P = fixtures.make_params(case["shapes"], case["param_seed"])
        X = fixtures.make_inputs(case)
        P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
        if case["model"] == "mhbcoatt":
            ref = O.mhbcoatt_forward(P64, X["img"].double(), X["questions"], X["glove"].double() if "glove" in X else None)
        else:
            ref = O.mfb_forward(P64, X["img"].double(), X["questions"], case["cfg"]["model_name"] == "mfb-multilayer")
        for mode in ("fp32", "bf16"):
            model = (MHBCoAtt if case["model"] == "mhbcoatt" else MFB)(types.SimpleNamespace(**case["cfg"]))
            model.load_state_dict(P)
            model.precision = mode
            model = model.cuda().train()
            model.dropout_l.p = 0.0
            model.dropout_m.p = 0.0
            args = [X["img"].cuda(), X["questions"].cuda()] + ([X["glove"].cuda()] if "glove" in X else [])
            model.capture = {}
            out = model(*args)
            (out * X["cot"].cuda()).sum().backward()
            inj = {}
            for key, y in model.capture.items():
                y = y.detach().double().cpu()
                z = torch.sign(y) * y * y
                inj["z" + key[1:]] = z.reshape(X["img"].shape[0], -1, z.shape[-1]) if key == "y1" else z
            P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
            if case["model"] == "mhbcoatt":
                ref2 = O.mhbcoatt_forward(P64, X["img"].double(), X["questions"], X["glove"].double() if "glove" in X else None, inj)
            else:
                ref2 = O.mfb_forward(P64, X["img"].double(), X["questions"], case["cfg"]["model_name"] == "mfb-multilayer", inj)
            (ref2 * X["cot"].double()).sum().backward()
            print("== %s [%s] out rel-err %.3e" % (name, mode, O.rel_err(out, ref)))
            for k, p in model.named_parameters():
                r = P64[k].grad
                g = p.grad
                rn = float(r.norm()) if r is not None else 0.0
                gn = float(g.norm()) if g is not None else 0.0
                e = O.rel_err(g, r) if (r is not None and g is not None and rn > 0) else float("nan")
                print("   %-28s ref|g|=%.3e got|g|=%.3e rel-err=%.3e" % (k, rn, gn, e))


if __name__ == "__main__":
    main()
