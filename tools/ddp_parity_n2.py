"""Multi-GPU gradient parity: run under torchrun with WORLD_SIZE ranks, one per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_parity_n2.py

Every rank holds the same MHBCoAtt (full BASELINE dimensions, bf16 mode, dropout off) and its own shard of the global
batch.  (1) The data-parallel step -- wgrad GEMMs writing into the reducer's bucket views, NCCL all-reduce(AVG) launched
from the gradient hooks, per-bucket fused Adam -- must leave in `p.grad` the AVERAGE over ranks of the per-shard
gradients: each rank recomputes every shard's single-GPU gradients locally (no reducer, no communication) and compares.
(2) After the per-bucket Adam step all ranks must hold identical parameters, equal to a single-process Adam step on the
averaged gradients; two more steps must see the same losses as that reference (the updated weights are really used).
PARITY_SHARD=1 runs the sharded optimizer (reduce-scatter -> Adam on the shard -> all-gather of the bf16 copies).  Per-shard semantics as SURVEY 8e defines them (MHBCoAtt's LSTM runs over the batch axis of the
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
    sharded = os.environ.get("PARITY_SHARD", "0") == "1"
    STEPS = 3
    # ---- reference: a single process doing the data-parallel maths by hand: per-shard gradients, averaged, one Adam step
    ref_model = build()
    ref_opt = FusedAdam(ref_model.parameters(), lr=7e-4).attach(ref_model)
    p_init = {n: p.detach().clone() for n, p in ref_model.named_parameters()}
    ref_losses, avg1, ref_after1 = [], None, None
    for it in range(STEPS):
        avg = {n: torch.zeros_like(p) for n, p in ref_model.named_parameters()}
        my_loss = None
        for r in range(world):
            ref_model.zero_grad(set_to_none=True)
            img, q, tgt = shard(r)
            loss = crit(ref_model(img, q), tgt)
            loss.backward()
            if r == rank:
                my_loss = float(loss.detach())
            for n, p in ref_model.named_parameters():
                if p.grad is not None:
                    avg[n] += p.grad / world
        ref_losses.append(my_loss)
        for n, p in ref_model.named_parameters():
            p.grad = avg[n].clone()
        ref_opt.step()
        if it == 0:
            avg1 = avg
            ref_after1 = {n: p.detach().clone() for n, p in ref_model.named_parameters()}
    # ---- the data-parallel run on this rank's shard
    model = build()
    opt = FusedAdam(model.parameters(), lr=7e-4).attach(model)
    defer = [p for n, p in model.named_parameters() if not n.startswith(("lstm.", "word_embedding."))]
    reducer = GradientAllReducer(model, defer_params=defer, shard_optimizer=opt if sharded else None)
    n_sharded = sum(len(b.params) for b in reducer.buckets if b.sharded)
    img, q, tgt = shard(rank)
    losses, worst, worst_name, in_place, pw = [], 0.0, "", 0, 0.0
    for it in range(STEPS):
        if hasattr(reducer, "begin_step"):
            reducer.begin_step()
        loss = crit(model(img, q), tgt)
        losses.append(float(loss.detach()))
        reducer.prepare()
        loss.backward()
        if it == 0:
            # gradients before the optimizer touches anything: finish() without an optimizer is only legal unsharded,
            # so in sharded mode the check reads this rank's slice of every bucket after finish(opt)
            pass
        reducer.finish(opt)
        torch.cuda.synchronize()
        if it == 0:
            for bkt in reducer.buckets:
                lo = reducer.rank * bkt.shard if bkt.sharded else 0
                hi = lo + bkt.shard if bkt.sharded else bkt.numel
                for pi, (p, o) in enumerate(zip(bkt.params, bkt.offsets)):
                    n = [k for k, v in model.named_parameters() if v is p][0]
                    a, b_ = max(lo, o), min(hi, o + p.numel())
                    if a >= b_:
                        continue
                    got = bkt.flat[a:b_].double()
                    ref = avg1[n].reshape(-1)[a - o:b_ - o].double()
                    if float(ref.norm()) < 1e-9:
                        continue
                    e = float((got - ref).norm() / ref.norm())
                    if e > worst:
                        worst, worst_name = e, n
                    in_place += int(p.grad is not None and p.grad.data_ptr() == bkt.views[pi].data_ptr())
            reducer.sync_master_weights()
            for n, p in model.named_parameters():
                pw = max(pw, float((p.double() - ref_after1[n].double()).abs().max()))
    reducer.sync_master_weights()
    torch.cuda.synchronize()
    # all ranks hold the same parameters
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    lo_, hi_ = flat.clone(), flat.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    spread = float((hi_ - lo_).abs().max())
    # ... and they are where the hand-made data-parallel reference is, up to Adam's amplification of rounding noise
    # (aggregated over all parameters: a per-parameter maximum of this ratio is itself noise -- Adam turns the sign of a
    # near-zero gradient, which the fp32 atomics of two runs disagree on, into a full +-lr step)
    d2 = m2 = 0.0
    for n, p in model.named_parameters():
        m2 += float((ref_model.get_parameter(n).double() - p_init[n].double()).norm()) ** 2
        d2 += float((p.double() - ref_model.get_parameter(n).double()).norm()) ** 2
    far = (d2 / m2) ** 0.5
    loss_err = max(abs(a - b_) / abs(b_) for a, b_ in zip(losses, ref_losses))
    t = torch.tensor([worst, pw, spread, far, loss_err], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = (float(t[0]) < 2e-2 and float(t[2]) == 0.0 and float(t[1]) <= 2.5 * 7e-4 and float(t[3]) < 0.25
          and float(t[4]) < 2e-3)
    if rank == 0:
        print(json.dumps({"test": "ddp_gradient_parity", "world": world, "batch_per_rank": B, "steps": STEPS,
                          "sharded_optimizer": sharded, "sharded_params": n_sharded,
                          "worst_grad_rel_err_vs_avg_of_shard_grads": float(t[0]), "worst_param": worst_name,
                          "max_abs_param_diff_vs_reference_after_step_1": float(t[1]),
                          "max_param_spread_across_ranks": float(t[2]),
                          "param_distance_over_distance_moved_after_%d_steps" % STEPS: float(t[3]),
                          "max_rel_loss_diff_vs_reference": float(t[4]), "losses": losses, "reference_losses": ref_losses,
                          "grads_written_in_place": in_place, "params": len(list(model.parameters())),
                          "allreduce_payload_bytes": reducer.bytes_per_step(),
                          "wire_bytes_per_rank": reducer.wire_bytes_per_step(), "ok": ok}))
    reducer.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
