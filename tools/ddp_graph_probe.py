"""Probe: does the captured data-parallel iteration (train.GraphedTrainStep + ddp.GradientAllReducer over NCCL) replay?
Small model, hard time limits, Python stacks dumped on a hang.  torchrun --nproc-per-node 2 tools/ddp_graph_probe.py"""
import faulthandler
import os
import sys
import time
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
faulthandler.dump_traceback_later(int(os.environ.get("PROBE_LIMIT", "50")), exit=True)


def main():
    from vqa_attention_networks_b200 import MHBCoAtt, train
    from vqa_attention_networks_b200.ddp import GradientAllReducer
    from vqa_attention_networks_b200.optim import FusedAdam
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    big = os.environ.get("PROBE_BIG", "0") == "1"
    H, D, L, A, V, B = (1024, 2048, 196, 3000, 15000, 64) if big else (128, 256, 49, 56, 200, 8)
    cfg = types.SimpleNamespace(model_name="mhb_coAtt", q_vocab_size=V, emb_dim=32 if not big else 300, hidden_dim=H,
                                num_layers=1, img_feature_channel=D, img_feature_dim=L, a_vocab_size=A, glove=False)
    torch.manual_seed(0)
    m = MHBCoAtt(cfg).to(dev).train()
    opt = FusedAdam(m.parameters(), lr=1e-3).attach(m)
    defer = [p for n, p in m.named_parameters() if not n.startswith(("lstm.", "word_embedding."))]
    red = GradientAllReducer(m, defer_params=defer if os.environ.get("PROBE_DEFER", "1") == "1" else None,
                             shard_optimizer=opt if os.environ.get("PROBE_SHARD", "0") == "1" else None)
    step = train.TrainStep(m, torch.nn.KLDivLoss(), opt, red)
    slots = []
    for i in range(2):
        g = torch.Generator().manual_seed(10 * rank + i)
        tgt = torch.rand(B, A, generator=g)
        slots.append((torch.randn(B, L, D, generator=g).relu_().to(dev), torch.randint(0, V, (B, 26), generator=g).to(dev),
                      (tgt / tgt.sum(1, keepdim=True)).to(dev)))
    t0 = time.time()
    print("rank %d: eager steps" % rank, flush=True)
    for i in range(2):
        step(*slots[i % 2])
    torch.cuda.synchronize()
    print("rank %d: eager ok %.1fs; capturing" % (rank, time.time() - t0), flush=True)
    tags = ["mfb_fused_spatial"] if os.environ.get("PROBE_SEGMENT", "0") == "1" else None
    g = train.GraphedTrainStep(step, slots, warmup=1, segment_tags=tags)
    print("rank %d: captured %.1fs; replaying" % (rank, time.time() - t0), flush=True)
    for i in range(6):
        loss = g.replay(i % 2)
        torch.cuda.synchronize()
        print("rank %d: replay %d loss %.6f" % (rank, i, float(loss)), flush=True)
    dist.barrier()
    print("rank %d: done %.1fs" % (rank, time.time() - t0), flush=True)
    del g
    import gc
    gc.collect()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
