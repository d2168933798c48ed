"""Drop-in replacement for the reference's ``mfb.py`` (class ``MFB``; ``cfg.model_name`` 'mfb' or
'mfb-multilayer').  Same constructor / forward signature / parameter names; the fusion and
co-attention stages run on the sm_100a kernels.

Bug-compatible by default: the reference takes ``softmax(dim=3)`` over a size-1
axis (mfb.py:84,118), so every attention weight is exactly 1, both glimpses are plain sum-pools and
the whole first stage (``img_conv1d``, ``ques_proj1``, ``co_att_conv*``, ``ques_att_conv*``) neither
influences the output nor receives gradient (exact zeros).  The kernels are therefore told
``degenerate=True``; the dead image projection is not executed.

Reference: /root/reference/mfb.py:6-140.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .mhb_coAtt import _FusionBase, _lstm_with_dropout, _scoped


class MFB(_FusionBase):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        multi = cfg.model_name == 'mfb-multilayer'
        self.word_embedding = nn.Embedding(cfg.q_vocab_size, cfg.emb_dim)
        self.lstm = nn.LSTM(input_size=cfg.emb_dim, hidden_size=cfg.hidden_dim, num_layers=cfg.num_layers,
                            batch_first=True)
        self.dropout_l = nn.Dropout(p=0.3)
        self.ques_att_conv1 = nn.Conv2d(cfg.hidden_dim, 1024, [1, 1])
        if multi:
            self.ques_att_multiconv = nn.Conv2d(1024, 512, [1, 1])
            self.ques_att_conv2 = nn.Conv2d(512, 2, [1, 1])
        else:
            self.ques_att_conv2 = nn.Conv2d(1024, 2, [1, 1])
        self.ques_proj1 = nn.Linear(2 * cfg.hidden_dim, 5000)
        self.img_conv1d = nn.Conv2d(cfg.img_feature_channel, 5000, [1, 1])
        self.dropout_m = nn.Dropout(p=0.1)
        self.co_att_conv1 = nn.Conv2d(1000, 1024, [1, 1])
        if multi:
            self.co_att_multiconv = nn.Conv2d(1024, 512, [1, 1])
            self.co_att_conv2 = nn.Conv2d(512, 2, [1, 1])
        else:
            self.co_att_conv2 = nn.Conv2d(1024, 2, [1, 1])
        self.ques_proj2 = nn.Linear(2 * cfg.hidden_dim, 5000)
        self.img_proj2 = nn.Linear(2 * cfg.img_feature_channel, 5000)
        self.linear_pred = nn.Linear(1000, cfg.a_vocab_size)
        # Opt-in, parity-unpinned: softmax over the region axis as in mhb_coAtt.py (not the reference's behaviour).
        self.corrected_softmax = False

    def dead_parameters(self):
        """Parameters whose gradient is EXACTLY zero on every step and every rank under the reference's singleton-axis
        softmax (mfb.py:84,118): the whole first stage.  Adam leaves such a parameter where it is
        (m = v = 0 => update 0), so an optimizer may skip it and a data-parallel reducer need not exchange it
        (`optim.FusedAdam.attach`, `ddp.GradientAllReducer`).  Empty with `corrected_softmax`."""
        if self.corrected_softmax:
            return []
        mods = [self.ques_att_conv1, self.ques_att_conv2, self.ques_proj1, self.img_conv1d, self.co_att_conv1,
                self.co_att_conv2]
        for name in ("ques_att_multiconv", "co_att_multiconv"):
            if hasattr(self, name):
                mods.append(getattr(self, name))
        return [p for m in mods for p in m.parameters()]

    def bf16_only_weights(self):
        """See MHBCoAtt.bf16_only_weights: the live projection weights, the classifier and the recurrent weight of the
        question encoder (ops.run_lstm reads W_hh through its bf16 copy at every batch size when the hidden size is one
        the persistent kernels take too)."""
        if self.precision != "bf16":
            return []
        ws = [self.ques_proj2.weight, self.img_proj2.weight]
        lp = self.linear_pred
        import os
        if not (lp.out_features % 8 or lp.in_features % 8 or os.environ.get("VQA_B200_CLASSIFIER", "fast") == "stock"):
            ws.append(lp.weight)
        if (self.lstm.num_layers == 1 and not self.lstm.bidirectional
                and os.environ.get("VQA_B200_LSTM", "fast") != "stock" and self.lstm.hidden_size in (128, 256, 512, 1024)):
            ws.append(self.lstm.weight_hh_l0)
        return ws

    @_scoped
    def question_features(self, questions):
        que_embedded = torch.tanh(self._embed(self.word_embedding, questions))       # mfb.py:68
        # proper batch_first here (mfb.py:69): T steps over N rows -- bf16 mode runs the per-step GEMM + cell form
        # (ops.LstmStepFn; N <= 32: the persistent kernels) on the module's own parameters; the dropout of mfb.py:70 is
        # applied by the same kernels
        return _lstm_with_dropout(self, self.lstm, self.dropout_l, que_embedded)      # [N, T, H]

    @_scoped
    def fused_block(self, img_features, ques_feature):
        multi = self.cfg.model_name == 'mfb-multilayer'
        deg = not self.corrected_softmax
        p = self.dropout_m.p
        qm = (self.ques_att_multiconv.weight, self.ques_att_multiconv.bias) if multi else (None, None)
        cm = (self.co_att_multiconv.weight, self.co_att_multiconv.bias) if multi else (None, None)
        qa, self.last_ques_att = ops.AttnPoolFn.apply(
            ques_feature, self.ques_att_conv1.weight, self.ques_att_conv1.bias, qm[0], qm[1],
            self.ques_att_conv2.weight, self.ques_att_conv2.bias, self._stage(degenerate=deg))
        ca, self.last_co_att = ops.MfbSpatialCoAttFn.apply(
            img_features, qa, self.ques_proj1.weight, self.ques_proj1.bias, self.img_conv1d.weight,
            self.img_conv1d.bias, self.co_att_conv1.weight, self.co_att_conv1.bias, cm[0], cm[1],
            self.co_att_conv2.weight, self.co_att_conv2.bias, self._stage(degenerate=deg, drop_p=p, key="y1"))
        return ops.MfbVectorFn.apply(qa, ca, self.ques_proj2.weight, self.ques_proj2.bias, self.img_proj2.weight,
                                     self.img_proj2.bias, self._stage(drop_p=p, key="y2"))

    @_scoped
    def forward(self, img_features, questions, is_training=True):
        ques_feature = self.question_features(questions)
        att_normed = self.fused_block(img_features, ques_feature)
        return self._classify(self.linear_pred, att_normed)             # mfb.py:140 returns the logits
