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
