"""Drop-in replacement for the reference's ``hieCoAtten.py`` (class ``HieCoAtten``): parallel co-attention
with the affinity matrix ``C = tanh(Q^T W_b V)``.

Same constructor, ``forward(img_features, que_features) -> (x, av, aq)``, parameter names and shapes
(reference hieCoAtten.py:5-55).  Bug-compatible by default:
  * ``fc_Wbv`` is applied to BOTH modalities and ``fc_Wbq`` is dead (grad ``None``)      (hieCoAtten.py:30-31)
  * the five ``F.dropout`` calls are functional -> always on, p = 0.5, even in ``eval()``  (hieCoAtten.py:26-46)
  * ``cat((v, q), 0).view(N, -1)`` mixes samples                                          (hieCoAtten.py:52-53)
  * ``torch.squeeze`` drops the batch axis when N == 1                                     (hieCoAtten.py:42-50)
Every Linear runs on the tcgen05 GEMM with ReLU / tanh / dropout / residual-add fused into its epilogue; the three
per-sample products run on the same kernel through rank-3 TMA descriptors.  The embedding and the answer
classifier ``fc`` stay stock PyTorch.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import K_MAJOR, MN_MAJOR
from .mhb_coAtt import _FusionBase, _scoped

_RELU, _TANH = 1, 2


class HieCoAtten(_FusionBase):
    def __init__(self, block_num=196, word_num=22, img_size=1024, vocab_size=15881, embed_size=512, att_num=6,
                 output_size=3000):
        super().__init__()
        self.img_emb = nn.Linear(img_size, embed_size, bias=True)
        self.que_emb = nn.Embedding(vocab_size, embed_size)
        self.fc_Wbv = nn.Linear(embed_size, embed_size)
        self.fc_Wbq = nn.Linear(embed_size, embed_size)        # registered but unused, as in the reference
        self.fc_Wv = nn.Linear(embed_size, embed_size)
        self.fc_Wq = nn.Linear(embed_size, embed_size)
        self.fc_Whv = nn.Linear(embed_size, 1)
        self.fc_Whq = nn.Linear(embed_size, 1)
        self.fc = nn.Linear(2 * embed_size, output_size)
        self.dropout_p = 0.5            # F.dropout's default; always on (functional dropout ignores eval())
        self.last_seeds = []            # test hook: the five dropout seeds of the last forward, in call order

    @_scoped
    def forward(self, img_features, que_features):
        with ops.pack_scope():          # `img`, `que`, `C`, `img_`, `que_` are each consumed by several GEMMs
            return self._forward(img_features, que_features)

    def _forward(self, img_features, que_features):
        cfg = ops.StageCfg(mode=self.precision, cache=self._wcache, seed_dev=self.seed_counter)
        p = self.dropout_p
        seeds = [ops.new_seed() if p > 0 else 0 for _ in range(5)]
        self.last_seeds = seeds
        batch_size = img_features.size(0)
        lin = ops.LinearActFn.apply
        img = lin(img_features, self.img_emb.weight, self.img_emb.bias, cfg, _RELU, p, seeds[0],
                  "hie_img_emb")                                                                         # :25-26
        que = ops.ActFn.apply(self.que_emb(que_features), None, 0, p, seeds[1], self.seed_counter)       # :27-28
        Cv = lin(img, self.fc_Wbv.weight, self.fc_Wbv.bias, cfg, 0, 0.0, 0)                              # :30
        Cq = lin(que, self.fc_Wbv.weight, self.fc_Wbv.bias, cfg, 0, 0.0, 0)                              # :31 (Wbv!)
        C = ops.BmmActFn.apply(Cq, K_MAJOR, Cv, K_MAJOR, None, cfg, _TANH, p, seeds[2],
                               "hie_affinity")                                                          # :32-33 [N,T,L]
        img_ = lin(img, self.fc_Wv.weight, self.fc_Wv.bias, cfg, 0, 0.0, 0)                              # :35
        que_ = lin(que, self.fc_Wq.weight, self.fc_Wq.bias, cfg, 0, 0.0, 0)                              # :36
        # Hv[l,:] = tanh(img_[l,:] + sum_t C[t,l] que_[t,:])                                              # :38-39
        Hv = ops.BmmActFn.apply(C, MN_MAJOR, que_, MN_MAJOR, img_, cfg, _TANH, p, seeds[3])
        v, av = ops.LogitsPoolFn.apply(Hv, self.fc_Whv.weight, self.fc_Whv.bias, img)                    # :40-43
        # Hq[t,:] = tanh(que_[t,:] + sum_l C[t,l] img_[l,:])                                              # :45-46
        Hq = ops.BmmActFn.apply(C, K_MAJOR, img_, MN_MAJOR, que_, cfg, _TANH, p, seeds[4])
        q, aq = ops.LogitsPoolFn.apply(Hq, self.fc_Whq.weight, self.fc_Whq.bias, que)                    # :47-50
        v, q, av, aq = (torch.squeeze(t) for t in (v, q, av, aq))                                        # :42-43,49-50
        x = torch.cat((v, q), 0)                                                                         # :52
        x = x.view(batch_size, -1)                                                                       # :53
        x = self.fc(x)                                                                                   # :54
        return x, av, aq
