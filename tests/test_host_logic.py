"""CPU-side checks of host logic that needs no GPU: the weight-cache validity rules, the regime test of the
persistent LSTM kernels (a pure function of the C ABI) and the optimizer's refusal of CPU parameters."""
import pytest
import torch


def test_weight_cache_validity_rules():
    """Entries are stamped with the parameter's version counter; in training mode they must additionally have been
    produced or vouched for since the previous forward began (torch's fused Adam updates parameters without bumping
    their version counters)."""
    from vqa_attention_networks_b200.ops import WeightCache
    calls = []

    def derive(w):
        calls.append(1)
        return w.clone()

    c = WeightCache()
    w = torch.nn.Parameter(torch.zeros(4, 4))
    # eval mode: the version stamp alone decides
    c.begin_forward(False)
    a = c.get_fn(w, "t", derive)
    assert c.get_fn(w, "t", derive) is a and len(calls) == 1
    with torch.no_grad():
        w.add_(1.0)                                  # ordinary in-place update: version bump -> re-derived
    b = c.get_fn(w, "t", derive)
    assert b is not a and len(calls) == 2
    # training mode: a new forward distrusts everything from before it ...
    c.begin_forward(True)
    d = c.get_fn(w, "t", derive)
    assert d is not b and len(calls) == 3
    assert c.get_fn(w, "t", derive) is d             # ... but re-uses what this forward produced (backward pass)
    c.begin_forward(True)
    assert c.get_fn(w, "t", derive) is not d and len(calls) == 4
    # a silent update (no version bump, like torch.optim.Adam(fused=True)) is therefore harmless in training mode
    w.data.mul_(2.0)
    c.begin_forward(True)
    e = c.get_fn(w, "t", derive)
    assert torch.equal(e, w.detach()) and len(calls) == 5
    # an optimizer that wrote the kernel-form copy itself vouches for it: valid for exactly the next forward
    key = (w.data_ptr(), w.numel(), w.device.index, 0, 0, "bf16")
    from vqa_attention_networks_b200.ops import Operand
    c._d[key] = (w._version, Operand(torch.zeros(4, 4, dtype=torch.bfloat16), 0, 4, 4), c._epoch)
    assert c.bf16_entry(w) is not None
    c.refreshed(w)
    c.begin_forward(True)
    assert c._lookup(key, w._version) is not None
    c.begin_forward(True)
    assert c._lookup(key, w._version) is None


def test_module_mode_switch_clears_the_cache():
    import types
    from vqa_attention_networks_b200 import MFB
    cfg = types.SimpleNamespace(model_name="mfb", q_vocab_size=20, emb_dim=6, hidden_dim=8, num_layers=1,
                                img_feature_channel=16, img_feature_dim=6, a_vocab_size=7, glove=False)
    m = MFB(cfg)
    m._wcache._d["x"] = (0, None, 0)
    m.train()                       # no change of mode: kept
    assert "x" in m._wcache._d
    m.eval()
    assert not m._wcache._d
    m._wcache._d["x"] = (0, None, 0)
    m.train()
    assert not m._wcache._d


def test_lstm_regime():
    from vqa_attention_networks_b200 import ops
    assert ops.lstm_supported(26, 1024) and ops.lstm_supported(1, 128) and ops.lstm_supported(32, 512)
    assert not ops.lstm_supported(33, 1024)          # more rows per step than one tile
    assert not ops.lstm_supported(26, 1000)          # hidden size not one of 128/256/512/1024
    assert not ops.lstm_supported(0, 1024)


def test_fused_adam_refuses_cpu_parameters():
    from vqa_attention_networks_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedAdam([p]).step()
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)


def test_dropout_seed_contract(monkeypatch):
    """ops.new_seed: reproducible under torch.manual_seed, different per data-parallel rank, and it must not advance
    the user's global CPU RNG stream."""
    from vqa_attention_networks_b200 import ops
    monkeypatch.setenv("RANK", "0")
    torch.manual_seed(123)
    ops._seed_state["base"] = None
    a = [ops.new_seed() for _ in range(4)]
    before = torch.get_rng_state().clone()
    ops.new_seed()
    assert torch.equal(before, torch.get_rng_state())
    torch.manual_seed(123)
    ops._seed_state["base"] = None
    assert [ops.new_seed() for _ in range(4)] == a
    monkeypatch.setenv("RANK", "1")
    ops._seed_state["base"] = None
    b = [ops.new_seed() for _ in range(4)]
    assert b != a and len(set(a + b)) == 8
    assert all(0 <= s < 2 ** 31 for s in a + b)
    ops._seed_state["base"] = None


def test_concurrent_builds_are_serialised(tmp_path):
    """build._build_lock is an inter-process lock (N torchrun ranks racing nvcc into one output)."""
    import multiprocessing as mp
    import time
    from vqa_attention_networks_b200 import build as b

    def hold(q):
        with b._build_lock():
            q.put(time.time())
            time.sleep(0.5)
            q.put(time.time())

    ctx = mp.get_context("fork")
    q1, q2 = ctx.Queue(), ctx.Queue()
    p1 = ctx.Process(target=hold, args=(q1,))
    p1.start()
    t_in1 = q1.get(timeout=10)
    p2 = ctx.Process(target=hold, args=(q2,))
    p2.start()
    t_out1 = q1.get(timeout=10)
    t_in2 = q2.get(timeout=10)
    p1.join(10)
    p2.join(10)
    assert t_in1 <= t_out1 <= t_in2 + 1e-3


def test_shard_segments_partition_the_bucket():
    from vqa_attention_networks_b200.ddp import shard_segments
    offs, nums = [0, 12, 40, 44], [10, 28, 2, 20]          # 16-byte aligned starts, payloads shorter than the gaps
    total = 64
    for world in (1, 2, 4, 8):
        shard = total // world
        seen = []
        for r in range(world):
            for i, s, f, n in shard_segments(offs, nums, r * shard, (r + 1) * shard):
                assert offs[i] + s == f and 0 < n and s + n <= nums[i] and r * shard <= f and f + n <= (r + 1) * shard
                seen += list(range(f, f + n))
        assert sorted(seen) == sorted(e for o, n in zip(offs, nums) for e in range(o, o + n))    # every element once


def test_reducer_bucket_layout_for_fused_groups_and_shards():
    """Bucket layout rules of ddp.GradientAllReducer: the weights one wgrad GEMM produces together are adjacent and in
    order; sharded (bf16-only) weights never share a bucket with all-reduced ones; the modules' weight caches adopt the
    flat bf16 buffers, so a projection group is ONE contiguous operand."""
    import types
    from vqa_attention_networks_b200 import MHBCoAtt, ops
    from vqa_attention_networks_b200.ddp import GradientAllReducer
    from vqa_attention_networks_b200.optim import FusedAdam
    cfg = types.SimpleNamespace(model_name="mhb_coAtt", q_vocab_size=20, emb_dim=8, hidden_dim=128, num_layers=1,
                                img_feature_channel=16, img_feature_dim=6, a_vocab_size=8, glove=False)
    m = MHBCoAtt(cfg)
    opt = FusedAdam(m.parameters(), lr=1e-3)
    red = GradientAllReducer(m, bucket_mb=0.5, shard_optimizer=opt)
    try:
        names = {id(p): n for n, p in m.named_parameters()}
        sharded = {names[id(p)] for b in red.buckets if b.sharded for p in b.params}
        plain = {names[id(p)] for b in red.buckets if not b.sharded for p in b.params}
        assert sharded == {"ques_att_conv1.weight", "ques_proj1.weight", "ques_proj2.weight", "ques_proj3.weight",
                           "img_conv1d.weight", "co_att_conv1.weight", "img_proj2.weight", "img_proj3.weight",
                           "linear_pred.weight", "lstm.weight_hh_l0"}
        assert not (sharded & plain) and len(sharded | plain) == len(names)
        for group in m.fused_param_groups():
            b = red.buckets[red._index[group[0]][0]]
            idx = [[id(q) for q in b.params].index(id(p)) for p in group]
            assert idx == list(range(idx[0], idx[0] + len(group)))                       # adjacent, in order
            for a, c in zip(idx, idx[1:]):
                assert b.offsets[a] + b.params[a].numel() == b.offsets[c]                # no padding in between
            op = m._wcache.get_group(group, ops.K_MAJOR)
            assert op.t.data_ptr() == b.w16_views[idx[0]].data_ptr()
            assert op.t.shape[0] == sum(p.shape[0] for p in group)
        for b in red.buckets:
            if b.sharded:
                assert b.numel % 8 == 0 and b.shard == b.numel and sum(n for *_, n in b.segments) == b.payload
                for p, v in zip(b.params, b.w16_views):                                  # bf16 copies filled and adopted
                    assert torch.equal(v.view_as(p), p.detach().to(torch.bfloat16))
                    assert m._wcache.bf16_entry(p).data_ptr() == v.data_ptr()
        # pinned copies survive the train/eval cache flush and are never re-cast from the (possibly stale) fp32 master
        m.eval()
        m.train()
        w = m.img_conv1d.weight
        before = m._wcache.get(w, 0, 1, "bf16").t.clone()
        with torch.no_grad():
            w.add_(1.0)                                 # a version bump that would invalidate an ordinary entry
        m._wcache.begin_forward(True)
        assert torch.equal(m._wcache.get(w, 0, 1, "bf16").t, before)
    finally:
        red.close()
    assert m._wcache.bf16_entry(m.img_conv1d.weight) is None


def test_lstm_form_selection_and_stock_fallback():
    """ops.run_lstm: the two native regimes by rows per step, and the stock module (clean output, dropout left to the
    caller) for everything the kernels do not take -- here: CPU tensors."""
    from vqa_attention_networks_b200 import ops
    assert ops.lstm_steps_supported(64, 1024) and ops.lstm_steps_supported(33, 128)
    assert not ops.lstm_steps_supported(32, 1024)          # the persistent kernels' regime
    assert not ops.lstm_steps_supported(64, 100)           # hidden size must keep the bf16 TMA pitch rule
    lstm = torch.nn.LSTM(6, 8, batch_first=True)
    x = torch.randn(3, 5, 6)
    out, dropped = ops.run_lstm(lstm, x, ops.WeightCache(), "bf16", drop_p=0.3, seed=7)
    assert not dropped and torch.equal(out, lstm(x)[0])


def test_functions_see_the_callers_autograd_mode():
    """ctx.needs_input_grad is True for Parameters even under torch.no_grad(), and grad mode is always off inside
    Function.forward: the modules' forward scope records the caller's mode (ops.set_outer_grad) and ops._need_grad
    combines the two."""
    import types
    from vqa_attention_networks_b200 import MHBCoAtt, ops
    seen = []

    class Probe(torch.autograd.Function):
        @staticmethod
        def forward(ctx, w):
            seen.append((any(ctx.needs_input_grad), ops._need_grad(ctx), torch.is_grad_enabled()))
            return w * 2

        @staticmethod
        def backward(ctx, g):
            return g * 2

    w = torch.nn.Parameter(torch.ones(2))
    Probe.apply(w)                                         # no scope: assume autograd is on
    with torch.no_grad():
        Probe.apply(w)                                     # a Function on its own cannot tell
    cfg = types.SimpleNamespace(model_name="mhb_coAtt", q_vocab_size=10, emb_dim=4, hidden_dim=8, num_layers=1,
                                img_feature_channel=8, img_feature_dim=4, a_vocab_size=5, glove=False)
    m = MHBCoAtt(cfg)
    with torch.no_grad(), m._forward_scope():
        Probe.apply(w)
        with m._forward_scope():                           # re-entrant: the outermost scope decides
            Probe.apply(w)
    with m._forward_scope():
        Probe.apply(w)
    Probe.apply(w)                                         # scope left: back to "unknown"
    assert seen == [(True, True, False), (True, True, False), (True, False, False), (True, False, False),
                    (True, True, False), (True, True, False)]


def test_train_step_picks_the_fused_loss_only_for_the_solvers_criterion(monkeypatch):
    """train.TrainStep: nn.KLDivLoss() (reduction 'mean', solver.py:26-29) on a drop-in model that can hand its logits over
    -> fused loss; every other criterion / model / VQA_B200_LOSS=stock -> the criterion is called as the solver calls it."""
    import types
    from vqa_attention_networks_b200 import MHBCoAtt
    from vqa_attention_networks_b200.train import TrainStep
    cfg = types.SimpleNamespace(model_name="mhb_coAtt", q_vocab_size=10, emb_dim=4, hidden_dim=8, num_layers=1,
                                img_feature_channel=8, img_feature_dim=4, a_vocab_size=5, glove=False)
    m = MHBCoAtt(cfg)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    assert TrainStep(m, torch.nn.KLDivLoss(), opt)._fused_loss_model is m
    assert TrainStep(m, torch.nn.KLDivLoss(reduction="batchmean"), opt)._fused_loss_model is None
    assert TrainStep(m, torch.nn.KLDivLoss(log_target=True), opt)._fused_loss_model is None
    assert TrainStep(m, torch.nn.CrossEntropyLoss(), opt)._fused_loss_model is None
    assert TrainStep(torch.nn.Linear(2, 2), torch.nn.KLDivLoss(), opt)._fused_loss_model is None
    monkeypatch.setenv("VQA_B200_LOSS", "stock")
    assert TrainStep(m, torch.nn.KLDivLoss(), opt)._fused_loss_model is None
    # the deferral is only honoured with autograd on and CUDA logits: a CPU / no-grad forward returns log-probabilities
    m.defer_log_softmax = True
    logits = torch.randn(3, 5)
    assert torch.allclose(m._log_softmax(logits).exp().sum(1), torch.ones(3), atol=1e-6)
