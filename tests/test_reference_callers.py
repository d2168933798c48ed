"""Drop-in proof against the reference's OWN callers, as far as it can go without a GPU: the real
``solver.py`` / ``networks.py`` / ``utils.py`` / ``cfg.py`` are imported from /root/reference (``tensorboardX``, ``spacy``
and ``easydict`` stubbed) and handed THIS package's classes.

Runs in the build container only (the reference does not travel to the GPU box: the tests skip there).  What executes
kernels -- train steps, the val loop, concurrent replicas -- is covered on the B200 by tests/test_gpu_solver_loop.py,
which follows the same call sites line by line.
"""
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "solver.py")), reason="reference not mounted")


class _Edict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class _Writer:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, *a, **k):
        self.scalars.append(a)

    def add_scalars(self, *a, **k):
        self.scalars.append(a)


@pytest.fixture()
def ref(monkeypatch):
    """The reference's modules importable by their own names, with the three absent third-party packages stubbed."""
    stubs = {"tensorboardX": types.ModuleType("tensorboardX"), "spacy": types.ModuleType("spacy"),
             "easydict": types.ModuleType("easydict")}
    stubs["tensorboardX"].SummaryWriter = _Writer
    stubs["spacy"].load = lambda *a, **k: None
    stubs["easydict"].EasyDict = _Edict
    for k, v in stubs.items():
        monkeypatch.setitem(sys.modules, k, v)
    monkeypatch.syspath_prepend(REF)
    names = ("solver", "networks", "modules", "utils", "cfg", "data_loader", "mhb_coAtt", "mfb")
    saved = {n: sys.modules.pop(n, None) for n in names}
    import importlib
    mods = types.SimpleNamespace()
    yield mods, importlib
    for n in names:
        sys.modules.pop(n, None)
        if saved[n] is not None:
            sys.modules[n] = saved[n]


def _cfg(ref_cfg, **kw):
    c = _Edict(ref_cfg)
    c.update(dict(q_vocab_size=40, a_vocab_size=12, hidden_dim=16, emb_dim=8, img_feature_channel=32, img_feature_dim=6,
                  glove=False, batch_size=2, num_workers=0, num_answer=12, mode="training", early_stopping=False))
    c.update(kw)
    return c


def _qa_data(n=4):
    rec = [{"image_id": i, "question": [1, 2, 3, 0, 0], "ques_length": 3, "answer": 1, "answers": {"1": 1.0}} for i in range(n)]
    return {"train": rec, "val": rec, "question_vocab": {"a": 1}, "answer_vocab": {"x": 0}}


@pytest.mark.parametrize("name,cls", [("mhb_coAtt", "MHBCoAtt"), ("mfb", "MFB"), ("mfb-multilayer", "MFB")])
def test_solver_takes_the_drop_in_modules(ref, name, cls, tmp_path, monkeypatch):
    """train_models.py:44-60 (factory, Xavier loop) + solver.py:16-43 (Adam over model.parameters(), DataParallel /
    .to(device), datasets) + solver.py:185-192 (save through utils.clean_state_dict) with this package's module, and the
    checkpoint it writes loads STRICTLY into the reference's own class (and back)."""
    mods, importlib = ref
    import vqa_attention_networks_b200 as V
    solver_mod = importlib.import_module("solver")
    ref_cfg = importlib.import_module("cfg").cfg
    cfg = _cfg(ref_cfg, model_name=name, out_dir=str(tmp_path / "models"), soft_answer=int(name == "mhb_coAtt"))
    model = getattr(V, cls)(cfg)
    for pname, param in model.named_parameters():                 # train_models.py:54-56, verbatim semantics
        if pname.find("bias") == -1:
            torch.nn.init.xavier_uniform_(param)
    s = solver_mod.Solver(model, cfg, _qa_data())
    assert isinstance(s.criterion, torch.nn.KLDivLoss if name == "mhb_coAtt" else torch.nn.CrossEntropyLoss)
    got = {id(p) for g in s.optimizer.param_groups for p in g["params"]}
    assert got == {id(p) for p in model.parameters()}
    assert len(s.data_loader["train"]) == 2
    s.adjust_learning_rate()
    assert abs(s.optimizer.param_groups[0]["lr"] - cfg.lr * cfg.decay_rate) < 1e-12
    s.save()                                                      # solver.py:185-192 -> utils.clean_state_dict
    ckpt = torch.load(os.path.join(cfg.out_dir, name + ".pth"))
    ref_cls = getattr(importlib.import_module("mhb_coAtt" if cls == "MHBCoAtt" else "mfb"), cls)
    theirs = ref_cls(cfg)
    theirs.load_state_dict(ckpt, strict=True)                     # our checkpoint -> the reference module
    model.load_state_dict(theirs.state_dict(), strict=True)       # ... and the reference's -> ours (testing mode, :58-60)
    # the DataParallel spelling of the same checkpoint (solver.py:34-36 wraps when several GPUs are visible)
    wrapped = torch.nn.DataParallel(model)
    cleaned = importlib.import_module("utils").clean_state_dict(wrapped.state_dict())
    assert list(cleaned) == list(ckpt)
    theirs.load_state_dict(cleaned, strict=True)


def test_attentionnet_is_built_from_the_drop_in_attention_layer(ref, monkeypatch):
    """networks.py:4 does `from modules import Attention_layer` and :35-42 instantiates it six times by name: with this
    package's modules.py in its place the reference's AttentionNet constructs, and its state dict has exactly the
    names and shapes of the one built from the reference's own modules.py."""
    mods, importlib = ref
    theirs = importlib.import_module("networks").AttentionNet(block_num=6, word_num=5, img_size=16, vocab_size=30,
                                                                embed_size=8, att_num=6, output_size=7)
    want = {k: tuple(v.shape) for k, v in theirs.state_dict().items()}
    for n in ("networks", "modules"):
        sys.modules.pop(n, None)
    import vqa_attention_networks_b200.modules as ours
    monkeypatch.setitem(sys.modules, "modules", ours)
    net = importlib.import_module("networks").AttentionNet(block_num=6, word_num=5, img_size=16, vocab_size=30,
                                                            embed_size=8, att_num=6, output_size=7)
    assert type(net._modules["att0"]).__module__ == "vqa_attention_networks_b200.modules"
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == want
    net.load_state_dict(theirs.state_dict(), strict=True)
    # and the product path still refuses to compute on the CPU instead of silently falling back
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(2, 6, 16), torch.zeros(2, 5, dtype=torch.long))
