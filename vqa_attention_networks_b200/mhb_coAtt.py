"""Drop-in replacements for the reference's ``mhb_coAtt.py`` (classes ``MHBCoAtt`` and ``MHB``).

Same constructors, ``forward`` signatures, parameter names / shapes (so ``state_dict``s interchange
and ``train_models.py:54-56``'s Xavier loop and ``solver.py`` work unchanged); the bodies run the
fusion / co-attention stages on the sm_100a kernels (ops.py).  The word embedding, ``self.lstm`` and
``self.linear_pred`` stay the stock ``torch.nn`` modules as far as parameters and state dict go
(``north_star``: left as-is); in bf16 mode on CUDA their execution is widened in: the
LSTM's recurrence runs on the persistent kernels of csrc/lstm.cu and the classifier on the tcgen05
GEMM, with ``VQA_B200_LSTM=stock`` / ``VQA_B200_CLASSIFIER=stock`` selecting the stock calls.

Reference: /root/reference/mhb_coAtt.py:6-151 (MHBCoAtt), :153-217 (MHB).
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def default_precision() -> str:
    return os.environ.get("VQA_B200_PRECISION", "bf16")


def _scoped(fn):
    """Method decorator: run inside the module's forward scope (weight-cache bracketing, see _FusionBase)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        with self._forward_scope():
            return fn(self, *a, **kw)
    return wrapper


def _lstm_with_dropout(owner, lstm, dropout, x):
    """dropout(lstm(x)[0]) through ops.run_lstm: in training mode with p > 0 the native recurrence applies the dropout
    itself (one counter-hash seed per call, salted by the captured iteration's step counter like the MFB dropouts);
    otherwise -- and on the stock path -- the nn.Dropout module runs as in the reference."""
    p = float(dropout.p) if (dropout.training and not getattr(dropout, "inplace", False)) else 0.0
    if not 0.0 < p < 1.0:
        p = 0.0                                   # p = 1 (everything dropped) and in-place modules stay with nn.Dropout
    seed = ops.new_seed() if p > 0.0 else 0
    out, dropped = ops.run_lstm(lstm, x, owner._wcache, owner.precision, p, seed,
                                getattr(owner, "seed_counter", None) if p > 0.0 else None)
    owner.last_lstm_drop_seed = seed if dropped else None
    return out if dropped else dropout(out)


class _FusionBase(nn.Module):
    """Shared plumbing: precision mode, kernel-form weight cache, per-call dropout seeds."""

    def __init__(self):
        super().__init__()
        self.precision = default_precision()      # "bf16" | "fp32"
        self._wcache = ops.WeightCache()          # not a buffer: never enters the state dict
        self.capture = None                       # test hook: dict that receives the MFB blocks' y tensors
        self.last_pred = self.last_pred_logp = None   # set by _log_softmax() outside autograd
        self._scope_depth = 0
        # optional device step counter (int64 [1]) that salts every fused dropout seed: set by train.GraphedTrainStep so
        # that a captured step draws new masks at every replay (ops.StageCfg.seed_dev); None = a new host seed per call
        self.seed_counter = None

    @contextlib.contextmanager
    def _forward_scope(self):
        """Brackets one forward pass for the weight cache (ops.WeightCache.begin_forward); re-entrant, so that
        forward() -> question_features() / fused_block() counts once while direct calls of the parts still work."""
        outer = self._scope_depth == 0
        if outer:
            self._wcache.begin_forward(self.training and torch.is_grad_enabled())
            prev = ops.set_outer_grad(torch.is_grad_enabled())      # the Functions cannot see the caller's grad mode
        self._scope_depth += 1
        try:
            yield
        finally:
            self._scope_depth -= 1
            if outer:
                ops.set_outer_grad(prev)

    def _classify(self, linear: nn.Linear, feat):
        """The answer classifier (mhb_coAtt.py:147, mfb.py:137-140) with the SAME nn.Linear
        parameters on the tcgen05 GEMM (forward, dgrad, wgrad): stock PyTorch runs it as three fp32 SIMT GEMMs
        (0.24 ms per train step at N = 256).  Answer vocabularies / feature widths that are not multiples of 8 break
        TMA's 16-byte pitch rule for the gradient operands and keep the stock call, as does VQA_B200_CLASSIFIER=stock."""
        if (not feat.is_cuda or linear.out_features % 8 or linear.in_features % 8
                or os.environ.get("VQA_B200_CLASSIFIER", "fast") == "stock"):
            return linear(feat)
        return ops.LinearFn.apply(feat, linear.weight, linear.bias, ops.StageCfg(mode=self.precision, cache=self._wcache))

    def _embed(self, emb: nn.Embedding, idx):
        """`emb(idx)` with the scatter-add backward of ops.EmbeddingFn for plain CUDA embeddings (no padding_idx / max_norm
        / sparse gradients -- the reference's, mhb_coAtt.py:25-26); anything else is the stock module call."""
        if (emb.weight.is_cuda and emb.padding_idx is None and emb.max_norm is None and not emb.sparse
                and not emb.scale_grad_by_freq and emb.weight.dtype == torch.float32):
            return ops.EmbeddingFn.apply(idx, emb.weight)
        return emb(idx)

    def _log_softmax(self, logits):
        """F.log_softmax(logits, dim=1) (mhb_coAtt.py:149-151).  Outside autograd (the val loop / inference) the fused
        tail kernel also leaves the predicted answers in ``self.last_pred`` (solver.py:148-149's softmax + max) and their
        log-probabilities in ``self.last_pred_logp``; with autograd on it is the stock op."""
        if getattr(self, "defer_log_softmax", False) and torch.is_grad_enabled() and logits.is_cuda:
            # train.TrainStep fuses log-softmax with the solver's KLDivLoss (ops.KLDivLogSoftmaxFn): hand the logits over
            self.last_pred = self.last_pred_logp = None
            self.deferred_log_softmax = True          # tells TrainStep that what it got back are the logits
            return logits
        if logits.is_cuda and logits.dtype == torch.float32 and not (torch.is_grad_enabled() and logits.requires_grad):
            logp, self.last_pred, self.last_pred_logp = ops.log_softmax_argmax(logits.contiguous())
            return logp
        self.last_pred = self.last_pred_logp = None
        return F.log_softmax(logits, dim=1)

    def train(self, mode: bool = True):
        if mode != self.training:
            for cb in list(getattr(self, "_mode_switch_callbacks", ())):
                cb()                              # e.g. a sharded optimizer completing the fp32 weights on every rank
            self._wcache.clear()                  # eval-mode entries are trusted by version stamp only: start clean
        return super().train(mode)

    def _stage(self, degenerate=False, drop_p=0.0, key=""):
        p = drop_p if self.training else 0.0
        seed = ops.new_seed() if p > 0.0 else 0
        return ops.StageCfg(mode=self.precision, cache=self._wcache, degenerate=degenerate, drop_p=p, seed=seed,
                            capture=self.capture, key=key, seed_dev=self.seed_counter if p > 0.0 else None)


class MHBCoAtt(_FusionBase):
    """MFH co-attention network (reference mhb_coAtt.py:6-151): question attention -> MFB over the
    14x14 grid -> co-attention -> two independent MFB vector blocks -> classifier -> log-softmax."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.word_embedding = nn.Embedding(cfg.q_vocab_size, cfg.emb_dim)
        in_sz = cfg.emb_dim * 2 if cfg.glove else cfg.emb_dim                 # mhb_coAtt.py:27-36
        self.lstm = nn.LSTM(input_size=in_sz, hidden_size=cfg.hidden_dim, num_layers=cfg.num_layers, batch_first=True)
        self.dropout_l = nn.Dropout(p=0.3)
        self.ques_att_conv1 = nn.Conv2d(cfg.hidden_dim, 512, [1, 1])
        self.ques_att_conv2 = nn.Conv2d(512, 2, [1, 1])
        self.ques_proj1 = nn.Linear(2 * cfg.hidden_dim, 5000)
        self.img_conv1d = nn.Conv2d(cfg.img_feature_channel, 5000, [1, 1])
        self.dropout_m = nn.Dropout(p=0.1)
        self.co_att_conv1 = nn.Conv2d(1000, 512, [1, 1])
        self.co_att_conv2 = nn.Conv2d(512, 2, [1, 1])
        self.ques_proj2 = nn.Linear(2 * cfg.hidden_dim, 5000)
        self.ques_proj3 = nn.Linear(2 * cfg.hidden_dim, 5000)
        self.img_proj2 = nn.Linear(2 * cfg.img_feature_channel, 5000)
        self.img_proj3 = nn.Linear(2 * cfg.img_feature_channel, 5000)
        self.linear_pred = nn.Linear(2000, cfg.a_vocab_size)

    # -- the part north_star leaves as-is: embedding + tanh + LSTM (+ dropout), mhb_coAtt.py:69-75.
    # NB: the LSTM is batch_first but is fed [T, N, E] -> the recurrence runs over the batch axis
    # (SURVEY.md fact 5); reproduced verbatim.
    def question_features(self, questions, glove_matrix=None):
        que_embedded = torch.tanh(self._embed(self.word_embedding, questions))
        if self.cfg.glove:
            assert glove_matrix is not None, 'glove should not be NoneType.'
            que_embedded = torch.cat((que_embedded, glove_matrix), dim=2)
        with self._forward_scope():
            lstm_o = self._run_lstm(que_embedded.permute(1, 0, 2))
        return lstm_o.permute(1, 0, 2)                        # [N, T, H] view of the [T, N, H] output

    def _run_lstm(self, x):
        """`self.dropout_l(self.lstm(x)[0])` for the batch_first x = [T, N, E] the reference feeds (a recurrence of N steps
        over T rows; mhb_coAtt.py:72-74).  bf16 mode on CUDA runs it on the persistent recurrence kernel (ops.LstmFn,
        csrc/lstm.cu) with the SAME `self.lstm` parameters -- the stock path costs 2 launches per step, 8.4 of the
        13.4 ms train step at N = 256 -- and the kernel applies the dropout to its output (seed in
        ``self.last_lstm_drop_seed``; mask rows are time-major).  fp32 mode, unsupported shapes and VQA_B200_LSTM=stock
        keep the stock modules (north_star: left as-is)."""
        return _lstm_with_dropout(self, self.lstm, self.dropout_l, x)

    def fused_block(self, img_features, ques_feature):
        """The hot path (mhb_coAtt.py:77-145): [N,L,D] features + [N,T,H] question states -> [N, 2000]."""
        with self._forward_scope():
            return self._fused_block(img_features, ques_feature)

    def _fused_block(self, img_features, ques_feature):
        with ops.pack_scope():          # the question vector feeds three projections, the image vector two
            return self._fused_block_body(img_features, ques_feature)

    def bf16_only_weights(self):
        """Weights that the forward / backward kernels read ONLY through their cached bf16 copies (bf16 mode): a sharded
        data-parallel optimizer may then keep the fp32 master of each on one rank and all-gather just the bf16 copies
        (ddp.GradientAllReducer(shard_optimizer=...)).  Biases, the attention convs' 2-row weights (read as fp32 by the
        logits kernel), the embedding and W_ih (its kernel form is a padded copy derived from the fp32 tensor) are not."""
        if self.precision != "bf16" or os.environ.get("VQA_B200_FUSED_BLOCK", "1") == "0":
            return []
        ws = [self.ques_att_conv1.weight, self.ques_proj1.weight, self.ques_proj2.weight, self.ques_proj3.weight,
              self.img_conv1d.weight, self.co_att_conv1.weight, self.img_proj2.weight, self.img_proj3.weight]
        lp = self.linear_pred
        if not (lp.out_features % 8 or lp.in_features % 8 or os.environ.get("VQA_B200_CLASSIFIER", "fast") == "stock"):
            ws.append(lp.weight)
        if (self.lstm.num_layers == 1 and not self.lstm.bidirectional
                and os.environ.get("VQA_B200_LSTM", "fast") != "stock" and self.lstm.hidden_size in (128, 256, 512, 1024)):
            ws.append(self.lstm.weight_hh_l0)
        return ws

    def fused_param_groups(self):
        """Weights whose gradients one wgrad GEMM produces side by side (fused_block.MhbFusedBlockFn): a data-parallel
        reducer keeps each group adjacent, in this order, inside one bucket (ddp.GradientAllReducer)."""
        return [[self.ques_proj1.weight, self.ques_proj2.weight, self.ques_proj3.weight],
                [self.img_proj2.weight, self.img_proj3.weight]]

    def _fused_block_body(self, img_features, ques_feature):
        if self.precision == "bf16" and os.environ.get("VQA_B200_FUSED_BLOCK", "1") != "0":
            # one autograd node for the whole block: question / image projections batched across the three MFB blocks
            from .fused_block import BlockCfg, MhbFusedBlockFn
            p = self.dropout_m.p if self.training else 0.0
            cfg = BlockCfg(self._wcache, p, ops.new_seed() if p > 0.0 else 0, ops.new_seed() if p > 0.0 else 0,
                           self.seed_counter if p > 0.0 else None, self.capture)
            out, self.last_ques_att, self.last_co_att = MhbFusedBlockFn.apply(
                img_features, ques_feature, self.ques_att_conv1.weight, self.ques_att_conv1.bias,
                self.ques_att_conv2.weight, self.ques_att_conv2.bias, self.ques_proj1.weight, self.ques_proj1.bias,
                self.ques_proj2.weight, self.ques_proj2.bias, self.ques_proj3.weight, self.ques_proj3.bias,
                self.img_conv1d.weight, self.img_conv1d.bias, self.co_att_conv1.weight, self.co_att_conv1.bias,
                self.co_att_conv2.weight, self.co_att_conv2.bias, self.img_proj2.weight, self.img_proj2.bias,
                self.img_proj3.weight, self.img_proj3.bias, cfg)
            return out
        p = self.dropout_m.p
        qa, self.last_ques_att = ops.AttnPoolFn.apply(
            ques_feature, self.ques_att_conv1.weight, self.ques_att_conv1.bias, None, None,
            self.ques_att_conv2.weight, self.ques_att_conv2.bias, self._stage())
        ca, self.last_co_att = ops.MfbSpatialCoAttFn.apply(
            img_features, qa, self.ques_proj1.weight, self.ques_proj1.bias, self.img_conv1d.weight,
            self.img_conv1d.bias, self.co_att_conv1.weight, self.co_att_conv1.bias, None, None,
            self.co_att_conv2.weight, self.co_att_conv2.bias, self._stage(drop_p=p, key="y1"))
        o2 = ops.MfbVectorFn.apply(qa, ca, self.ques_proj2.weight, self.ques_proj2.bias, self.img_proj2.weight,
                                   self.img_proj2.bias, self._stage(drop_p=p, key="y2"))
        o3 = ops.MfbVectorFn.apply(qa, ca, self.ques_proj3.weight, self.ques_proj3.bias, self.img_proj3.weight,
                                   self.img_proj3.bias, self._stage(drop_p=p, key="y3"))
        return torch.cat([o2, o3], 1)

    def forward(self, img_features, questions, glove_matrix=None, is_training=True):
        with self._forward_scope():
            ques_feature = self.question_features(questions, glove_matrix)
            att_normed_23 = self.fused_block(img_features, ques_feature)
            logits = self._classify(self.linear_pred, att_normed_23)
        return self._log_softmax(logits)                      # implicit dim of mhb_coAtt.py:149 is 1 for 2-D


class MHB(_FusionBase):
    """MFH baseline without attention (reference mhb_coAtt.py:153-217): mean-pooled image feature, two *cascaded* MFB
    blocks (block 2 is multiplied by block 1's dropped-out product before pooling, :204-205 -- the true high-order
    coupling) -> classifier -> log-softmax.

    The reference class is broken as shipped (hard ``.cuda()`` at :176, undefined ``mhb_22`` at :214); this
    implementation follows the two-token patch the oracle / golden fixture use (``mhb_22`` -> ``mhb_12``).  The four
    projections run on the tcgen05 GEMM (forward, dgrad, wgrad), the 14x14 mean-pool on the pooling kernel, and the
    cascade coupling, both dropouts, the k-pools and the signed square roots inside the GEMM epilogues."""

    def __init__(self, cfg):
        super().__init__()
        self.model_name = cfg.model_name
        self.cfg = cfg
        self.mean_pool = nn.AvgPool2d((14, 14))            # kept for state/attribute parity; pooling runs on the kernel
        self.Embedding = nn.Embedding(cfg.q_vocab_size, cfg.emb_dim)
        self.LSTM = nn.LSTM(input_size=cfg.emb_dim, hidden_size=cfg.hidden_dim, num_layers=1, batch_first=False)
        self.linear_q_1 = nn.Linear(cfg.hidden_dim, 5000)
        self.linear_q_2 = nn.Linear(cfg.hidden_dim, 5000)
        self.linear_i_1 = nn.Linear(cfg.img_feature_channel, 5000)
        self.linear_i_2 = nn.Linear(cfg.img_feature_channel, 5000)
        self.lstm_dropout = nn.Dropout(0.3)
        self.mfb_dropout = nn.Dropout(0.1)
        self.linear_out = nn.Linear(2000, cfg.a_vocab_size)

    @_scoped
    def forward(self, img_feature, questions, q_length):
        batch_size, max_len = questions.size()
        N, Lr, D = img_feature.shape
        # mean over the 14x14 grid (:178-180): uniform softmax weights (all-zero logits) on the pooling kernel
        x3 = img_feature if img_feature.dtype in (torch.float32, torch.bfloat16) else img_feature.float()
        zeros = torch.zeros((N * Lr, 1), device=img_feature.device, dtype=torch.float32)
        i_mean, _ = ops.softmax_pool_fwd(x3.contiguous(), zeros, 1, False)        # [N, D]
        q_embedded = self.Embedding(questions).permute(1, 0, 2)                   # :181-182  [T, N, V]
        lstm_outs, _ = self.LSTM(q_embedded)                                      # :183
        idx = torch.as_tensor(q_length, device=lstm_outs.device).long() - 1
        lstm_out = lstm_outs[idx, torch.arange(batch_size, device=lstm_outs.device)]      # :185-186
        lstm_out = self.lstm_dropout(lstm_out)
        # the cascade (:189-212): both MFB blocks on the fused-epilogue kernels, block 2 multiplied by block 1's
        # dropped-out product inside the epilogue (fused_block.MhbCascadeFn)
        from .fused_block import MhbCascadeFn
        p = self.mfb_dropout.p if self.training else 0.0
        cfg = ops.StageCfg(mode=self.precision, cache=self._wcache, drop_p=p, seed=ops.new_seed() if p > 0 else 0,
                           capture=self.capture, seed_dev=self.seed_counter if p > 0 else None)
        mhb_12 = MhbCascadeFn.apply(lstm_out, i_mean, self.linear_q_1.weight, self.linear_q_1.bias,
                                    self.linear_q_2.weight, self.linear_q_2.bias, self.linear_i_1.weight,
                                    self.linear_i_1.bias, self.linear_i_2.weight, self.linear_i_2.bias, cfg,
                                    ops.new_seed() if p > 0 else 0)
        logits = self.linear_out(mhb_12)                                          # :213-214 (patched mhb_22 -> mhb_12)
        return self._log_softmax(logits)
