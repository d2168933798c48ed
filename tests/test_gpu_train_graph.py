"""GPU tests of the captured training iteration (train.GraphedTrainStep): replaying the CUDA graphs must train exactly
like the eager loop (solver.py:68-94), draw new dropout masks at every replay, and keep Adam's bias corrections in step."""
import types

import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(seed=0, p_drop=0.0, p_lstm=0.0):
    from vqa_attention_networks_b200 import MHBCoAtt
    cfg = types.SimpleNamespace(model_name="mhb_coAtt", q_vocab_size=200, emb_dim=32, hidden_dim=128, num_layers=1,
                                img_feature_channel=256, img_feature_dim=49, a_vocab_size=56, glove=False)
    torch.manual_seed(seed)
    m = MHBCoAtt(cfg)
    for n, p in m.named_parameters():
        if n.find("bias") == -1:
            torch.nn.init.xavier_uniform_(p)
    m = m.to(DEV).train()
    m.dropout_l.p = p_lstm                    # applied inside the recurrence kernel (device-salted seed, like dropout_m)
    m.dropout_m.p = p_drop
    return m


def _batches(n, B=8):
    out = []
    for i in range(n):
        X = O.synthetic_inputs(B, 49, 256, 26, 200, seed=10 + i, device=DEV)
        out.append((X["img"], X["questions"], O.soft_answers(B, 56, seed=20 + i).to(DEV)))
    return out


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("segment", [False, True])
def test_graph_replay_trains_like_the_eager_loop(segment):
    """Same weights, same batches, dropout off: 6 replays (two slots, alternating) == 6 eager steps, parameters and Adam
    state included.  With `segment_tags` the tagged launches run between graph segments and are timed."""
    from vqa_attention_networks_b200.optim import FusedAdam
    from vqa_attention_networks_b200.train import GraphedTrainStep, TrainStep
    crit = torch.nn.KLDivLoss()
    data = _batches(2)
    ma, mb = _model(), _model()
    oa = FusedAdam(ma.parameters(), lr=2e-3).attach(ma)
    ob = FusedAdam(mb.parameters(), lr=2e-3).attach(mb)
    eager = TrainStep(ma, crit, oa)
    WARM = 2
    losses_a = [float(eager(*data[i % 2]).detach()) for i in range(WARM + 6)]
    slots = [tuple(t.clone() for t in d) for d in data]
    g = GraphedTrainStep(TrainStep(mb, crit, ob), slots, warmup=WARM,
                         segment_tags=["mfb_fused_spatial", "softmax_pool_fwd_regions"] if segment else None)
    assert len(g.programs[0]) == (5 if segment else 1)
    g.reset_times(segment)
    losses_b = [float(g.replay(i % 2)) for i in range(6)]
    torch.cuda.synchronize()
    for la, lb in zip(losses_a[WARM:], losses_b):
        assert abs(la - lb) <= 2e-3 * abs(la) + 1e-7, (losses_a, losses_b)
    # Parameters: the fp32 atomics of the split-K GEMMs make two runs differ in the last bits of every gradient, and
    # Adam turns a sign flip of a near-zero gradient into a full +-lr step, so bit equality is not on offer between
    # ANY two runs (per-parameter ratios are themselves noise: rarely used embedding rows, softmax-invariant biases).
    # Bound the distance between the two runs, over all parameters, by a fraction of the distance they travelled.
    p0 = dict(_model().named_parameters())
    d2 = m2 = 0.0
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        m2 += float((pa.double() - p0[n].double()).norm()) ** 2
        d2 += float((pb.double() - pa.double()).norm()) ** 2
    assert d2 ** 0.5 <= 0.25 * m2 ** 0.5, (d2 ** 0.5, m2 ** 0.5)
    g.sync_python_state()
    assert all(int(st["step"]) == WARM + 6 for st in ob.state.values())
    assert int(ob.step_count.item()) == WARM + 6
    if segment:
        kt = g.kernel_times()
        assert kt["mfb_fused_spatial"][0] == 6 and kt["mfb_fused_spatial"][1] > 0.0
        assert kt["softmax_pool_fwd_regions"][0] == 6
    assert g.launches > 50


def test_replays_draw_new_dropout_masks_and_are_reproducible():
    """lr = 0 freezes the weights, so with dropout on the loss of one slot changes from replay to replay only through the
    masks; the mask of a replay is the function of (host seed, device step count) that ops.dropout_mask reproduces."""
    from vqa_attention_networks_b200 import ops
    from vqa_attention_networks_b200.optim import FusedAdam
    from vqa_attention_networks_b200.train import GraphedTrainStep, TrainStep
    m = _model(p_drop=0.3, p_lstm=0.3)
    opt = FusedAdam(m.parameters(), lr=0.0).attach(m)
    data = _batches(1)
    g = GraphedTrainStep(TrainStep(m, torch.nn.KLDivLoss(), opt), data, warmup=1)
    losses = [float(g.replay(0)) for _ in range(4)]
    assert len({round(l, 9) for l in losses}) == 4, losses
    # salted mask: differs between counter values, equal for equal (seed, counter)
    c = torch.tensor([5], dtype=torch.int64, device=DEV)
    m5 = ops.dropout_mask(64, 500, 0.3, 1234, DEV, c)
    m5b = ops.dropout_mask(64, 500, 0.3, 1234, DEV, c.clone())
    c.add_(1)
    m6 = ops.dropout_mask(64, 500, 0.3, 1234, DEV, c)
    plain = ops.dropout_mask(64, 500, 0.3, 1234, DEV)
    assert torch.equal(m5, m5b) and not torch.equal(m5, m6) and not torch.equal(m5, plain)
    assert abs(float((m6 > 0).float().mean()) - 0.7) < 2e-2
    both = float(((m5 > 0) & (m6 > 0)).float().mean())
    assert abs(both - 0.49) < 2e-2                     # consecutive steps are independent masks


def test_device_step_adam_matches_torch_adam_under_replay():
    """The bias corrections of replay k must be those of step k (they are formed on the device from the counter)."""
    from vqa_attention_networks_b200.optim import FusedAdam
    g_ = torch.Generator().manual_seed(3)
    base = [torch.randn(300, 40, generator=g_), torch.randn(17, generator=g_)]
    pa = [torch.nn.Parameter(b.clone().to(DEV)) for b in base]
    pb = [torch.nn.Parameter(b.clone().to(DEV)) for b in base]
    oa = FusedAdam(pa, lr=1e-2)
    ob = torch.optim.Adam(pb, lr=1e-2)
    grads = [[torch.randn(b.shape, generator=g_).to(DEV) for b in base] for _ in range(5)]
    static = [torch.zeros_like(p) for p in pa]
    for p, s in zip(pa, static):
        p.grad = s
    oa.enable_device_step(DEV)
    # one eager step (allocates the state), then capture the update and replay it
    for s, gr in zip(static, grads[0]):
        s.copy_(gr)
    oa.step()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=side):
        oa.step()
    # capture itself does not execute: state is still at step 1
    assert int(oa.step_count.item()) == 1
    for k in range(1, 5):
        for s, gr in zip(static, grads[k]):
            s.copy_(gr)
        graph.replay()
    torch.cuda.synchronize()
    for k in range(5):
        for p, gr in zip(pb, grads[k]):
            p.grad = gr.clone()
        ob.step()
    for x, y in zip(pa, pb):
        assert _rel(x, y) < 1e-6, _rel(x, y)
    assert int(oa.step_count.item()) == 5


def test_feed_into_the_captured_step(tmp_path):
    """The documented end-to-end path: packed bf16 shard -> feed.ShardFeed -> the captured iteration's static input
    slots (the feed refills them in place, two batches ahead, ordered by events) -> graph replay.  Same losses as the
    eager loop fed the same batches from fp32 tensors (bf16 mode rounds the features to the shard's values anyway)."""
    from vqa_attention_networks_b200 import feed
    from vqa_attention_networks_b200.optim import FusedAdam
    from vqa_attention_networks_b200.train import GraphedTrainStep, TrainStep
    B, NB = 8, 4
    data = _batches(NB, B)
    path = str(tmp_path / "train.shard")
    with feed.ShardWriter(path, 49, 256, 26, 56, soft_answer=True) as w:
        for img, q, tgt in data:
            w.append(img.cpu(), q.cpu(), tgt.cpu())
    reader = feed.ShardReader(path)
    crit = torch.nn.KLDivLoss()
    ma, mb = _model(), _model()
    oa = FusedAdam(ma.parameters(), lr=2e-3).attach(ma)
    ob = FusedAdam(mb.parameters(), lr=2e-3).attach(mb)
    eager = TrainStep(ma, crit, oa)
    slots = [(torch.empty(B, 49, 256, dtype=torch.bfloat16, device=DEV), torch.empty(B, 26, dtype=torch.int64, device=DEV),
              torch.empty(B, 56, device=DEV)) for _ in range(3)]
    fd = feed.ShardFeed(reader, B, DEV, device_slots=slots, depth=2, ring_slots=2, bind_numa=False)   # streaming ring
    WARM = 3
    for i in range(WARM):                         # real data in every slot before anything is captured
        d, _ = fd.next()
        fd.done(d)
    torch.cuda.synchronize()
    # the feed runs ahead: the released slots 0 and 1 already hold batches 3 and 4.  The eager twin's warm-up sees exactly
    # what the graphed one's eager warm-up iterations see in the slots
    warm = [tuple(t.clone() for t in s_) for s_ in slots]
    g = GraphedTrainStep(TrainStep(mb, crit, ob), slots, warmup=WARM)
    for i in range(WARM):
        eager(*warm[i])
    la, lb = [], []
    for i in range(7):
        d, _ = fd.next()
        lb.append(float(g.replay(d)))
        fd.done(d)
        la.append(float(eager(*data[(WARM + i) % NB]).detach()))
    fd.close()
    for a, b in zip(la, lb):
        assert abs(a - b) <= 5e-3 * abs(a) + 1e-7, (la, lb)
