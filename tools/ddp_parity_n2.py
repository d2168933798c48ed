"""Multi-GPU gradient parity: run under torchrun with WORLD_SIZE ranks, one per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_parity_n2.py

Every rank holds the same MHBCoAtt (full BASELINE dimensions, bf16 mode, dropout off) and its own shard of the global
batch.  (1) The data-parallel step -- wgrad GEMMs writing into the reducer's bucket views, NCCL all-reduce(AVG) launched
from the gradient hooks, per-bucket fused Adam -- must leave in `p.grad` the AVERAGE over ranks of the per-shard
gradients: each rank recomputes every shard's single-GPU gradients locally (no reducer, no communication) and compares.
(2) After the per-bucket Adam step all ranks must hold identical parameters, equal to a single-process Adam step on the
averaged gradients.  Per-shard semantics as SURVEY 8e defines them (MHBCoAtt's LSTM runs over the batch axis of the
shard).  Prints one JSON line on rank 0; exit code 1 on a violated bound."""
import json
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from vqa_attention_networks_b200 import MHBCoAtt
    from vqa_attention_networks_b200.ddp import GradientAllReducer
    from vqa_attention_networks_b200.optim import FusedAdam
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = int(os.environ.get("PARITY_BATCH", "32"))
    cfg = types.SimpleNamespace(model_name="mhb_coAtt", q_vocab_size=15000, emb_dim=300, hidden_dim=1024, num_layers=1,
                                img_feature_channel=2048, img_feature_dim=196, a_vocab_size=3000, glove=False)

    def build():
        torch.manual_seed(0)
        m = MHBCoAtt(cfg)
        for n, p in m.named_parameters():
            if n.find("bias") == -1:
                torch.nn.init.xavier_uniform_(p)
        m = m.to(dev).train()
        m.dropout_l.p = 0.0
        m.dropout_m.p = 0.0
        return m

    def shard(r):
        g = torch.Generator().manual_seed(100 + r)
        img = torch.randn(B, 196, 2048, generator=g).relu_().to(dev)
        q = torch.randint(0, 15000, (B, 26), generator=g).to(dev)
        tgt = torch.rand(B, 3000, generator=g)
        tgt = (tgt / tgt.sum(1, keepdim=True)).to(dev)
        return img, q, tgt

    crit = torch.nn.KLDivLoss()
    # ---- reference: single-GPU gradients of every shard, averaged (computed locally on every rank)
    ref_model = build()
    avg = {n: torch.zeros_like(p) for n, p in ref_model.named_parameters()}
    for r in range(world):
        ref_model.zero_grad(set_to_none=True)
        img, q, tgt = shard(r)
        crit(ref_model(img, q), tgt).backward()
        for n, p in ref_model.named_parameters():
            if p.grad is not None:
                avg[n] += p.grad / world
    # ---- data-parallel step on this rank's shard
    model = build()
    opt = FusedAdam(model.parameters(), lr=7e-4).attach(model)
    defer = [p for n, p in model.named_parameters() if not n.startswith(("lstm.", "word_embedding."))]
    reducer = GradientAllReducer(model, defer_params=defer)
    img, q, tgt = shard(rank)
    loss = crit(model(img, q), tgt)
    reducer.prepare()
    loss.backward()
    reducer.finish()                      # gradients only: compare before the optimizer touches anything
    torch.cuda.synchronize()
    worst, worst_name = 0.0, ""
    in_place = 0
    for n, p in model.named_parameters():
        ref = avg[n]
        e = float((p.grad.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30))
        if float(ref.norm()) < 1e-9:
            continue
        if e > worst:
            worst, worst_name = e, n
        bi, pi = reducer._index[p]
        in_place += int(p.grad.data_ptr() == reducer.buckets[bi].views[pi].data_ptr())
    # ---- optimizer: per-bucket fused Adam on the averaged gradients == single-process Adam on `avg`
    opt.step()
    ref_opt = FusedAdam(ref_model.parameters(), lr=7e-4).attach(ref_model)
    for n, p in ref_model.named_parameters():
        p.grad = avg[n].clone()
    ref_opt.step()
    torch.cuda.synchronize()
    pw = 0.0
    for (n, p), (_, rp) in zip(model.named_parameters(), ref_model.named_parameters()):
        moved = float((rp.double() - dict(build().named_parameters())[n].double()).norm()) if False else 1.0
        pw = max(pw, float((p.double() - rp.double()).abs().max()))
    # all ranks hold the same parameters
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    spread = float((hi - lo).abs().max())
    t = torch.tensor([worst, pw, spread], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = float(t[0]) < 2e-2 and float(t[2]) == 0.0 and float(t[1]) <= 2.5 * 7e-4
    if rank == 0:
        print(json.dumps({"test": "ddp_gradient_parity", "world": world, "batch_per_rank": B,
                          "worst_grad_rel_err_vs_avg_of_shard_grads": float(t[0]), "worst_param": worst_name,
                          "max_abs_param_diff_vs_single_process_adam": float(t[1]),
                          "max_param_spread_across_ranks": float(t[2]), "grads_written_in_place": in_place,
                          "params": len(list(model.parameters())), "allreduce_bytes": reducer.bytes_per_step(),
                          "ok": ok}))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
