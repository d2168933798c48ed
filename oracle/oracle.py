"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's fusion / co-attention path.

This file is the *oracle*: a plain, functional torch-on-CPU restatement (fp32 or fp64,
differentiable through autograd) of the algorithms in klory/vqa-attention-networks'
``mfb.py``, ``mhb_coAtt.py``, ``hieCoAtten.py`` and ``modules.py``.  It is written from the
maths of those files (each function cites the file:line it follows) and is laid out the way
the CUDA path is (row-major ``[N*L, C]`` matrices, explicit pooling over the k factors) rather
than the way the reference is (NCHW 1x1 convs, permutes/views).

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the *real, unmodified reference modules* executed in the
build container by ``oracle/gen_golden.py`` and committed under ``tests/golden/``
(``tests/test_oracle_golden.py`` checks every fixture, forward and gradients).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product package never does.

All functions take a ``P`` mapping of parameter name -> tensor using the reference's
``state_dict`` names, so one state dict drives the reference, the oracle
and the CUDA modules.
"""
from __future__ import annotations

import math
from typing import Mapping, Optional

import torch

K_FACTOR = 5          # mhb_coAtt.py:43 "k * o = 5000, k = 5"
O_DIM = 1000
EPS_NORM = 1e-12      # F.normalize default eps (mhb_coAtt.py:107)


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def conv1x1(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """1x1 Conv2d applied to channel-last rows: x [..., Cin], w [Cout, Cin, 1, 1] -> [..., Cout].
    (mhb_coAtt.py:81,83,98,111,113 -- a 1x1 conv over an (L,1) grid is a per-position Linear.)"""
    w2 = w.reshape(w.shape[0], w.shape[1])
    y = x @ w2.t()
    return y if b is None else y + b


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def signed_sqrt(z: torch.Tensor) -> torch.Tensor:
    """sqrt(relu(z)) - sqrt(relu(-z))  (mhb_coAtt.py:106,131,143; mfb.py:104,133).
    Written through relu so autograd yields exactly 0 at z == 0, as the reference does."""
    return torch.sqrt(torch.relu(z)) - torch.sqrt(torch.relu(-z))


def l2_normalize_rows(y: torch.Tensor) -> torch.Tensor:
    """F.normalize(y) on a 2-D tensor: y / max(||y||_2, eps) per row."""
    n = torch.sqrt((y * y).sum(dim=1, keepdim=True))
    return y / torch.clamp(n, min=EPS_NORM)


def softmax_pool(logits: torch.Tensor, feats: torch.Tensor, degenerate: bool = False):
    """Per-glimpse softmax over the region/token axis followed by weighted pooling.

    logits [N, G, L], feats [N, L, D] -> pooled [N, G*D] (glimpse-major, mhb_coAtt.py:89,119),
    att [N, G, L].  ``degenerate=True`` reproduces mfb.py:84,118 where the softmax is taken over
    a size-1 axis, i.e. every weight is exactly 1 (and d(att)/d(logits) == 0)."""
    if degenerate:
        att = torch.softmax(logits.unsqueeze(-1), dim=-1).squeeze(-1)  # == 1, keeps the zero-grad edge
    else:
        att = torch.softmax(logits, dim=2)
    pooled = torch.einsum("ngl,nld->ngd", att, feats)
    return pooled.reshape(pooled.shape[0], -1), att


def mfb_pool(fused: torch.Tensor) -> torch.Tensor:
    """Sum-pool over the k=5 *adjacent* channels c = 5*o + j (view(...,1000,5).sum(-1),
    mhb_coAtt.py:102-103,129-130).  fused [..., 5000] -> [..., 1000]."""
    return fused.reshape(*fused.shape[:-1], O_DIM, K_FACTOR).sum(-1)


def _inject(z, z_forced):
    """Straight-through replacement of a forward value: returns a tensor whose VALUE is ``z_forced`` and whose
    gradient flows to ``z``.  Used by the parity tests to evaluate d(signed-sqrt) = 1/(2 sqrt|z|) -- singular
    at 0, so its L2 norm is log-divergently sensitive to rounding of z -- at exactly the z the kernels saw."""
    if z_forced is None:
        return z
    return z + (z_forced.to(z.dtype).reshape(z.shape) - z).detach()


def mfb_spatial(X, w_img, b_img, Q, mask=None, z_forced=None):
    """MFB block over the region grid (mhb_coAtt.py:94-108 / mfb.py:92-106).

    X [N, L, D] raw image features, w_img [5000, D, 1, 1], Q [N, 5000] projected question
    vector, mask [N, L, 5000] pre-scaled dropout mask (None in eval).
    Returns yhat [N, L, 1000]: signed-sqrt of the k-pooled Hadamard product, L2-normalised
    over all 1000*L elements of a sample (the reference's [N,1000,L] flattening order only
    permutes the elements inside the norm)."""
    I = conv1x1(X, w_img, b_img)                  # [N, L, 5000]
    F_ = I * Q[:, None, :]
    if mask is not None:
        F_ = F_ * mask
    z = _inject(mfb_pool(F_), z_forced)          # [N, L, 1000]
    y = signed_sqrt(z)
    n = torch.sqrt((y * y).sum(dim=(1, 2), keepdim=True))
    return y / torch.clamp(n, min=EPS_NORM)


def mfb_vector(q_proj, i_proj, mask=None, extra=None, z_forced=None):
    """MFB block on pooled vectors (mhb_coAtt.py:124-133,136-145; mfb.py:126-135).
    q_proj, i_proj [N, 5000] -> [N, 1000].  ``extra`` multiplies the product before dropout
    (the high-order coupling of MHB, mhb_coAtt.py:204-205).  Also returns the dropped-out
    product (MHB feeds it to the next block)."""
    f = q_proj * i_proj
    if extra is not None:
        f = f * extra
    fd = f if mask is None else f * mask
    y = signed_sqrt(_inject(mfb_pool(fd), z_forced))
    return l2_normalize_rows(y), fd


def lstm_batch_first(x, w_ih, w_hh, b_ih, b_hh):
    """Single-layer nn.LSTM(batch_first=True) on x [B, S, E] -> [B, S, H]; gate order i,f,g,o."""
    B, S, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = []
    for s in range(S):
        g = x[:, s] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, gg, o = g.split(H, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1)


def _lstm_params(P, prefix="lstm"):
    return (P[prefix + ".weight_ih_l0"], P[prefix + ".weight_hh_l0"],
            P[prefix + ".bias_ih_l0"], P[prefix + ".bias_hh_l0"])


# ----------------------------------------------------------------------------------------------
# MHBCoAtt  (mhb_coAtt.py:6-151)
# ----------------------------------------------------------------------------------------------
def mhbcoatt_question_features(P, questions, glove_matrix=None, mask_l=None):
    """mhb_coAtt.py:69-79.  Returns ques_feature as [N, T, H].
    NB: the LSTM is batch_first but is fed [T, N, E], so the recurrence runs
    over the *batch* axis; reproduced exactly."""
    emb = torch.tanh(P["word_embedding.weight"][questions])        # [N, T, E]
    if glove_matrix is not None:
        emb = torch.cat((emb, glove_matrix), dim=2)
    x = emb.permute(1, 0, 2)                                        # [T, N, E] fed as (batch=T, seq=N)
    lstm_o = lstm_batch_first(x, *_lstm_params(P))                  # [T, N, H]
    if mask_l is not None:
        lstm_o = lstm_o * mask_l
    return lstm_o.permute(1, 0, 2)                                  # [N, T, H]


def coatt_block(P, X, qfeat, masks=None, n_blocks=2, degenerate=False, multilayer=False):
    """The hot path shared by MFB and MHBCoAtt, from question attention to the MFB vector blocks.

    X [N, L, D] image features, qfeat [N, T, H] (dropped-out) LSTM outputs.
    Returns (att_normed [N, 1000*n_blocks], ques_att [N,2,T], co_att [N,2,L])."""
    masks = masks or {}
    # question attention (mhb_coAtt.py:81-91 / mfb.py:76-89)
    h = torch.relu(conv1x1(qfeat, P["ques_att_conv1.weight"], P["ques_att_conv1.bias"]))
    if multilayer:
        h = torch.relu(conv1x1(h, P["ques_att_multiconv.weight"], P["ques_att_multiconv.bias"]))
    ql = conv1x1(h, P["ques_att_conv2.weight"], P["ques_att_conv2.bias"])       # [N, T, 2]
    qa, q_att = softmax_pool(ql.permute(0, 2, 1), qfeat, degenerate)            # [N, 2H]
    # MFB #1 over the grid (mhb_coAtt.py:94-108)
    Q1 = linear(qa, P["ques_proj1.weight"], P["ques_proj1.bias"])
    yhat = mfb_spatial(X, P["img_conv1d.weight"], P["img_conv1d.bias"], Q1, masks.get("m1"), masks.get("z1"))
    # co-attention (mhb_coAtt.py:111-121)
    h2 = torch.relu(conv1x1(yhat, P["co_att_conv1.weight"], P["co_att_conv1.bias"]))
    if multilayer:
        h2 = torch.relu(conv1x1(h2, P["co_att_multiconv.weight"], P["co_att_multiconv.bias"]))
    cl = conv1x1(h2, P["co_att_conv2.weight"], P["co_att_conv2.bias"])          # [N, L, 2]
    ca, c_att = softmax_pool(cl.permute(0, 2, 1), X, degenerate)                # [N, 2D]
    # MFB vector blocks (mhb_coAtt.py:124-145)
    outs = []
    for bi in range(n_blocks):
        s = str(bi + 2)
        qp = linear(qa, P["ques_proj" + s + ".weight"], P["ques_proj" + s + ".bias"])
        ip = linear(ca, P["img_proj" + s + ".weight"], P["img_proj" + s + ".bias"])
        o, _ = mfb_vector(qp, ip, masks.get("m" + s), z_forced=masks.get("z" + s))
        outs.append(o)
    return torch.cat(outs, dim=1), q_att, c_att


def mhbcoatt_forward(P, img_features, questions, glove_matrix=None, masks=None):
    """MHBCoAtt.forward (mhb_coAtt.py:61-151) -> log-probabilities [N, A]."""
    masks = masks or {}
    qfeat = mhbcoatt_question_features(P, questions, glove_matrix, masks.get("l"))
    feat, _, _ = coatt_block(P, img_features, qfeat, masks, n_blocks=2)
    logits = linear(feat, P["linear_pred.weight"], P["linear_pred.bias"])
    return torch.log_softmax(logits, dim=1)         # implicit dim -> 1 for 2-D (mhb_coAtt.py:149)


# ----------------------------------------------------------------------------------------------
# MFB  (mfb.py:6-140)
# ----------------------------------------------------------------------------------------------
def mfb_forward(P, img_features, questions, multilayer=False, masks=None):
    """MFB.forward (mfb.py:61-140) -> *logits* [N, A] (mfb.py:140 returns logits, not probs).
    The attention softmaxes are degenerate (dim=3 over a size-1 axis, mfb.py:84,118)."""
    masks = masks or {}
    emb = torch.tanh(P["word_embedding.weight"][questions])
    lstm_o = lstm_batch_first(emb, *_lstm_params(P))                # [N, T, H], proper batch_first here
    if masks.get("l") is not None:
        lstm_o = lstm_o * masks["l"]
    feat, _, _ = coatt_block(P, img_features, lstm_o, masks, n_blocks=1, degenerate=True,
                             multilayer=multilayer)
    return linear(feat, P["linear_pred.weight"], P["linear_pred.bias"])


# ----------------------------------------------------------------------------------------------
# MHB (patched oracle: mhb_coAtt.py:153-217 with `mhb_22` -> `mhb_12` and no hard .cuda())
# ----------------------------------------------------------------------------------------------
def mhb_forward(P, img_feature, questions, q_length, masks=None):
    masks = masks or {}
    N = questions.shape[0]
    i_mean = img_feature.mean(dim=1)                                # AvgPool2d(14,14) over the grid (:178-180)
    emb = P["Embedding.weight"][questions]                          # [N, T, E]  (:181-183: seq-first LSTM over T)
    outs = lstm_batch_first(emb, *_lstm_params(P, "LSTM"))          # [N, T, H]
    lstm_out = outs[torch.arange(N), torch.as_tensor(q_length) - 1]           # (:185-186)
    if masks.get("l") is not None:
        lstm_out = lstm_out * masks["l"]
    q1 = linear(lstm_out, P["linear_q_1.weight"], P["linear_q_1.bias"])
    i1 = linear(i_mean, P["linear_i_1.weight"], P["linear_i_1.bias"])
    o1, f1d = mfb_vector(q1, i1, masks.get("m1"), z_forced=masks.get("z1"))
    q2 = linear(lstm_out, P["linear_q_2.weight"], P["linear_q_2.bias"])
    i2 = linear(i_mean, P["linear_i_2.weight"], P["linear_i_2.bias"])
    o2, _ = mfb_vector(q2, i2, masks.get("m2"), extra=f1d, z_forced=masks.get("z2"))          # (:204-205)
    logits = linear(torch.cat((o1, o2), 1), P["linear_out.weight"], P["linear_out.bias"])
    return torch.log_softmax(logits, dim=1)


# ----------------------------------------------------------------------------------------------
# HieCoAtten  (hieCoAtten.py:5-55)
# ----------------------------------------------------------------------------------------------
def hiecoatten_forward(P, img_features, que_tokens, masks=None):
    """Returns (x [N, A], av [N, L], aq [N, T]).  ``masks`` is the list of the five pre-scaled
    always-on dropout masks in call order (hieCoAtten.py:26,28,33,39,46); None = no dropout
    (not reachable in the reference, which always drops, but useful for unit checks).
    Quirks reproduced: fc_Wbv is applied to both modalities (:30-31, fc_Wbq unused), and the
    dim-0 cat + view mixes samples (:52-53)."""
    m = list(masks) if masks is not None else [None] * 5
    N = img_features.shape[0]
    img = torch.relu(linear(img_features, P["img_emb.weight"], P["img_emb.bias"]))      # [N, L, E]
    if m[0] is not None:
        img = img * m[0]
    que = P["que_emb.weight"][que_tokens]                                                # [N, T, E]
    if m[1] is not None:
        que = que * m[1]
    Cv = linear(img, P["fc_Wbv.weight"], P["fc_Wbv.bias"])
    Cq = linear(que, P["fc_Wbv.weight"], P["fc_Wbv.bias"])
    C = torch.tanh(torch.einsum("nte,nle->ntl", Cq, Cv))                                 # [N, T, L]
    if m[2] is not None:
        C = C * m[2]
    img_ = linear(img, P["fc_Wv.weight"], P["fc_Wv.bias"])
    que_ = linear(que, P["fc_Wq.weight"], P["fc_Wq.bias"])
    Hv = torch.tanh(img_ + torch.einsum("nte,ntl->nle", que_, C))                        # [N, L, E]
    if m[3] is not None:
        Hv = Hv * m[3]
    av = torch.softmax(linear(Hv, P["fc_Whv.weight"], P["fc_Whv.bias"]), dim=1)          # [N, L, 1]
    v = torch.einsum("nl,nle->ne", av[..., 0], img)
    Hq = torch.tanh(que_ + torch.einsum("nle,ntl->nte", img_, C))                        # [N, T, E]
    if m[4] is not None:
        Hq = Hq * m[4]
    aq = torch.softmax(linear(Hq, P["fc_Whq.weight"], P["fc_Whq.bias"]), dim=1)
    q = torch.einsum("nt,nte->ne", aq[..., 0], que)
    x = torch.cat((v, q), 0).reshape(N, -1)          # sample-mixing head (:52-53)
    x = linear(x, P["fc.weight"], P["fc.bias"])
    return x, av[..., 0], aq[..., 0]


# ----------------------------------------------------------------------------------------------
# modules.py
# ----------------------------------------------------------------------------------------------
def attention_1(P, f1, f2, prefix=""):
    """Attention_1.forward (modules.py:41-77): att = softmax_L(fc(f1[:,None] + f2[:,:,None])).
    Restated through its algebraic collapse: fc is linear, so the f2 and bias terms are constant
    along L and cancel in the softmax -> att[n,t,:] = softmax_l(w . f1[n,l]) for every t.
    The f2 term is kept in the logits (it cancels numerically) so autograd returns the same
    exactly-zero-in-exact-arithmetic gradients the reference produces."""
    w, b = P[prefix + "fc.weight"], P[prefix + "fc.bias"]
    s1 = (f1 @ w.t())[..., 0]                          # [N, L]
    s2 = (f2 @ w.t())[..., 0] + b                      # [N, T]
    att = torch.softmax(s1[:, None, :] + s2[:, :, None], dim=2)
    return att @ f1, att


def attention_2(P, f1, f2, prefix=""):
    """Attention_2.forward (modules.py:85-95): att = softmax_L(f2 (fc1 f1)^T); fc2 is unused."""
    g = f1 @ P[prefix + "fc1.weight"].t()
    att = torch.softmax(f2 @ g.transpose(1, 2), dim=2)
    return att @ f1, att


def attention_layer(P, f1, f2, att_type=1, prefix="att_layer."):
    """Attention_layer.forward (modules.py:26-33)."""
    a, b = torch.relu(f1), torch.relu(f2)
    f_hat, att = (attention_1 if att_type == 1 else attention_2)(P, a, b, prefix)
    return a, torch.relu(b + f_hat), att


def nonlinear_layer(P, x):
    """Nonlinear_layer.forward (modules.py:103-109): gated tanh."""
    return torch.tanh(linear(x, P["fc1.weight"], P["fc1.bias"])) * torch.sigmoid(linear(x, P["fc2.weight"], P["fc2.bias"]))


# ----------------------------------------------------------------------------------------------
# helpers shared by tests / golden generation / benches
# ----------------------------------------------------------------------------------------------
def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 in fp64 (the tolerance metric used by every parity test)."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    d = (a - b).norm().item()
    n = b.norm().item()
    return d / n if n > 0 else d


def synthetic_inputs(N, L, D, T, q_vocab, seed=1234, glove_dim=0, device="cpu"):
    """relu(N(0,1)) features, uniform token ids (generator-seeded)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.relu(torch.randn(N, L, D, generator=g))
    q = torch.randint(0, q_vocab, (N, T), generator=g)
    out = {"img": img.to(device), "questions": q.to(device)}
    if glove_dim:
        out["glove"] = torch.randn(N, T, glove_dim, generator=g).to(device)
    return out


def soft_answers(N, A, seed=7):
    """Soft answer rows as utils.py:250-265 produces: <=10 non-zeros, rows sum to 1."""
    g = torch.Generator().manual_seed(seed)
    a = torch.zeros(N, A)
    for n in range(N):
        k = int(torch.randint(1, 11, (1,), generator=g))
        idx = torch.randperm(A, generator=g)[:k]
        w = torch.rand(k, generator=g) + 0.1
        a[n, idx] = w / w.sum()
    return a
