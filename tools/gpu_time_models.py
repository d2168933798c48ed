"""Forward+backward time of the drop-in modules at BASELINE sizes vs the oracle's torch-eager formulation on the
same GPU (an indication only: the oracle is a checker, not a tuned baseline)."""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import oracle as O
import vqa_attention_networks_b200 as V

DEV = "cuda:0"


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def xavier(model):
    torch.manual_seed(0)
    for n, p in model.named_parameters():
        if n.find("bias") == -1:
            torch.nn.init.xavier_uniform_(p)
    return model.to(DEV).train()


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    X = O.synthetic_inputs(N, 196, 2048, 26, 15000, seed=3, device=DEV)
    # ---- HieCoAtten (config 3)
    hie = xavier(V.HieCoAtten(block_num=196, word_num=26, img_size=2048, vocab_size=15000, embed_size=512, output_size=3000))

    def hie_step():
        hie.zero_grad(set_to_none=True)
        x, av, aq = hie(X["img"], X["questions"])
        x.sum().backward()

    P = {k: v.detach().clone().requires_grad_(True) for k, v in hie.state_dict().items()}

    def hie_ref():
        for v in P.values():
            v.grad = None
        masks = [torch.ones(1, device=DEV)] * 5
        x, av, aq = O.hiecoatten_forward(P, X["img"], X["questions"], None)
        x.sum().backward()

    print("HieCoAtten N=%d fwd+bwd: ours %.3f ms | torch-eager oracle (no dropout) %.3f ms" % (N, timeit(hie_step), timeit(hie_ref)))
    # ---- MFB / MFB-multilayer (configs 1 / 4)
    for name in ("mfb", "mfb-multilayer"):
        cfg = types.SimpleNamespace(model_name=name, q_vocab_size=15000, emb_dim=300, hidden_dim=1024, num_layers=1,
                                    img_feature_channel=2048, img_feature_dim=196, a_vocab_size=3000, glove=False)
        m = xavier(V.MFB(cfg))

        def step():
            m.zero_grad(set_to_none=True)
            m(X["img"], X["questions"]).sum().backward()
        print("MFB(%s) N=%d fwd+bwd: ours %.3f ms" % (name, N, timeit(step)))
    # ---- MHBCoAtt block vs oracle eager on GPU
    cfg = types.SimpleNamespace(model_name="mhb_coAtt", q_vocab_size=15000, emb_dim=300, hidden_dim=1024, num_layers=1,
                                img_feature_channel=2048, img_feature_dim=196, a_vocab_size=3000, glove=False)
    mh = xavier(V.MHBCoAtt(cfg))
    qf = mh.question_features(X["questions"]).detach()
    Pm = {k: v.detach().clone().requires_grad_(True) for k, v in mh.state_dict().items()}

    def blk():
        mh.zero_grad(set_to_none=True)
        mh.fused_block(X["img"], qf).sum().backward()

    def blk_ref():
        for v in Pm.values():
            v.grad = None
        f, _, _ = O.coatt_block(Pm, X["img"], qf, None, n_blocks=2)
        f.sum().backward()
    t_ours = timeit(blk)
    try:
        t_ref = timeit(blk_ref, iters=2)
    except RuntimeError as e:
        t_ref = float("nan")
        print("oracle eager failed:", str(e)[:80])
    print("MHBCoAtt fused block N=%d fwd+bwd: ours %.3f ms | torch-eager oracle (fp32/TF32 off, no dropout) %.3f ms" % (N, t_ours, t_ref))


if __name__ == "__main__":
    main()
