"""The whole MFH co-attention block of ``MHBCoAtt`` (mhb_coAtt.py:77-145) as ONE autograd node (bf16 mode).

Stage for stage it launches the same kernels as ``ops.AttnPoolFn`` -> ``ops.MfbSpatialCoAttFn`` -> 2 x ``ops.MfbVectorFn``.
What owning the whole block buys is batching across the stages, which separate autograd nodes cannot do because their
gradients are due at different times:

  * ``ques_proj1 / 2 / 3`` read the same attended question vector: ONE GEMM over the row-concatenated weights
    (N = 15000) forward, ONE wgrad (``[15000, 2048]``) and ONE dgrad (K = 15000) backward, instead of 3 + 3 + 3 launches
    and two autograd additions;
  * ``img_proj2 / 3`` read the same attended image vector: ONE fused MFB launch (N = 10000, two L2-norm segments), ONE
    ``mfb_bwd``, ONE wgrad and ONE dgrad (K = 10000);
  * the concatenated bf16 weight copies live in ``ops.WeightCache`` groups whose per-parameter slices the fused Adam
    refreshes in place; the weight gradients are row slices of one buffer per group (inside a data-parallel reducer: the
    span of adjacent bucket views, ``ddp.GradientAllReducer(contiguous_groups=...)``);
  * every small zero-initialised accumulator of a pass (sum |z|, t = sum y g, bias / attention-conv gradients) is a slice
    of one workspace: one memset per pass instead of ~15 fills.

The vector blocks' dropout uses ONE seed for the [N, 10000] product (block 2 = columns 0..4999, block 3 = 5000..9999).
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import K_MAJOR, MN_MAJOR
from .ops import Operand


def _al4(n: int) -> int:
    return (n + 3) // 4 * 4


class _Workspace:
    """Zero-filled fp32 scratch carved into 16-byte aligned views."""

    def __init__(self, sizes, device):
        offs, off = [], 0
        for n in sizes:
            offs.append(off)
            off += _al4(n)
        self.flat = torch.zeros(off, device=device, dtype=torch.float32)
        self.views = [self.flat[o:o + n] for o, n in zip(offs, sizes)]


class BlockCfg:
    """Non-tensor settings of one forward of the block."""

    def __init__(self, cache: ops.WeightCache, drop_p: float, seed_spatial: int, seed_vector: int, seed_dev=None,
                 capture=None):
        self.cache, self.drop_p = cache, drop_p
        self.seed_spatial, self.seed_vector, self.seed_dev = seed_spatial, seed_vector, seed_dev
        self.capture = capture


class MhbFusedBlockFn(torch.autograd.Function):
    """(img_features [N,L,D], ques_feature [N,T,H], parameters) -> (att_normed_23 [N,2000], ques_att [N,2,T],
    co_att [N,2,L]); bf16 operands / stored activations, fp32 accumulation."""

    @staticmethod
    def forward(ctx, X, qf, Wqa1, bqa1, Wqa2, bqa2, Wq1, bq1, Wq2, bq2, Wq3, bq3, Wimg, bimg, Wc1, bc1, Wc2, bc2,
                Wi2, bi2, Wi3, bi3, cfg: BlockCfg):
        ops._cuda(X, qf, Wimg)
        if ctx.needs_input_grad[0]:
            raise RuntimeError("vqa_b200: gradients w.r.t. the image features are not part of the path")
        N, T, H = qf.shape
        _, Lr, D = X.shape
        M = N * Lr
        G = Wc2.shape[0]
        KO = Wq1.shape[0]                                     # k * o = 5000
        bf, f32 = torch.bfloat16, torch.float32
        sc = ops.StageCfg(mode="bf16", cache=cfg.cache)
        need_grad = ops._need_grad(ctx)
        ws = _Workspace([N, 2 * N], X.device)
        ssq1, ssq23 = ws.views
        # ---- question attention (mhb_coAtt.py:78-91)
        f2 = ops.pack_bf16(qf).view(N * T, H)
        hid_q = ops._linear_fwd(f2, Wqa1, bqa1, sc, bf, relu=True, tag="gemm_ques_att_conv1")
        logits_q = ops.attn_logits_fwd(hid_q, Wqa2, bqa2)
        qa, q_att = ops.softmax_pool_fwd(f2.view(N, T, H), logits_q, G, False)
        qa_b = ops.pack_bf16(qa)
        # ---- the three question projections in one GEMM (:94, :124, :136)
        wq = cfg.cache.get_group([Wq1, Wq2, Wq3], K_MAJOR)
        bq = torch.cat([bq1.detach(), bq2.detach(), bq3.detach()])
        Q123 = ops.gemm(Operand(qa_b, K_MAJOR, N, qa_b.shape[1]), K_MAJOR, wq, K_MAJOR, "bf16", out_dtype=f32, bias=bq,
                        tag="gemm_ques_proj123")
        Q1, Q23 = Q123[:, :KO], Q123[:, KO:]
        # ---- MFB over the grid (:97-108) + co-attention (:111-121)
        Xc = ops.pack_bf16(X).view(M, D)
        y1, _, keep1 = ops.mfb_fused(ops.prep(Xc, K_MAJOR, 0, "bf16"), cfg.cache.get(Wimg, K_MAJOR, 1, "bf16"), bimg, Q1,
                                     Lr, bf, bf if need_grad else None, cfg.drop_p, cfg.seed_spatial,
                                     tag="mfb_fused_spatial", seed_dev=cfg.seed_dev, ssq=ssq1)
        inv1 = ops.inv_norm(ssq1)
        hid_c = ops._linear_fwd(y1, Wc1, bc1, sc, bf, relu=True, row_scale=inv1, rows_per_group=Lr,
                                tag="gemm_co_att_conv1")
        logits_c = ops.attn_logits_fwd(hid_c, Wc2, bc2)
        ca, c_att = ops.softmax_pool_fwd(Xc.view(N, Lr, D), logits_c, G, False, tag="softmax_pool_fwd_regions")
        ca_b = ops.pack_bf16(ca)
        # ---- both MFB vector blocks in one launch (:124-145): N = 10000 columns, two L2-norm segments
        wi = cfg.cache.get_group([Wi2, Wi3], K_MAJOR)
        bi = torch.cat([bi2.detach(), bi3.detach()])
        y23, _, keep23 = ops.mfb_fused(Operand(ca_b, K_MAJOR, N, ca_b.shape[1]), wi, bi, Q23, 1, f32,
                                       bf if need_grad else None, cfg.drop_p, cfg.seed_vector, tag="mfb_fused_vector",
                                       seed_dev=cfg.seed_dev, seg_cols=KO, ssq=ssq23)
        inv23 = ops.inv_norm(ssq23)
        No = KO // 5
        out = ops.scale_rows(y23.view(2 * N, No), inv23, 1).view(N, 2 * No)      # == cat((block 2, block 3), 1), :145
        if cfg.capture is not None:
            cfg.capture["y1"], cfg.capture["y2"], cfg.capture["y3"] = y1, y23[:, :No], y23[:, No:]
        ctx.cfg, ctx.dims = cfg, (N, T, H, Lr, D, G, KO)
        ctx.save_for_backward(f2, hid_q, q_att, qa_b, Q123, Xc, y1, inv1, keep1, hid_c, c_att, ca_b, y23, inv23, keep23,
                              Wqa1, Wqa2, Wq1, Wq2, Wq3, Wimg, Wc1, Wc2, Wi2, Wi3)
        ctx.mark_non_differentiable(q_att, c_att)
        return out, q_att, c_att

    @staticmethod
    def backward(ctx, dout, _dq_att, _dc_att):
        (f2, hid_q, q_att, qa_b, Q123, Xc, y1, inv1, keep1, hid_c, c_att, ca_b, y23, inv23, keep23,
         Wqa1, Wqa2, Wq1, Wq2, Wq3, Wimg, Wc1, Wc2, Wi2, Wi3) = ctx.saved_tensors
        cfg: BlockCfg = ctx.cfg
        N, T, H, Lr, D, G, KO = ctx.dims
        No = KO // 5
        bf, f32 = torch.bfloat16, torch.float32
        dev = dout.device
        sc = ops.StageCfg(mode="bf16", cache=cfg.cache)
        Jc, Jq = hid_c.shape[1], hid_q.shape[1]
        ws = _Workspace([2 * N, 2 * KO, N, KO, G * Jc, G, Jc, G * Jq, G, Jq, 3 * KO], dev)
        t23, dbi23, t1, dbimg, dWc2, dbc2, dbc1, dWqa2, dbqa2, dbqa1, dbq123 = ws.views
        Q1, Q23 = Q123[:, :KO], Q123[:, KO:]
        # ---- the two vector blocks (:124-145 in reverse), one launch each step
        dout = dout.contiguous()
        g23, _ = ops.norm_bwd_prep(dout.view(2 * N, No), y23.view(2 * N, No), inv23, 1, t=t23)
        dI23, dQ23, _ = ops.mfb_bwd(g23.view(N, 2 * No), y23, inv23, t23, Q23, keep23, 1, bf, cfg.drop_p, cfg.seed_vector,
                                    cfg.seed_dev, seg_cols=KO, dbias=dbi23)
        wi_mn = cfg.cache.get_group([Wi2, Wi3], MN_MAJOR)
        dWi_base, (dWi2, dWi3) = ops.grad_buffer_group([Wi2, Wi3], ca_b.shape[1], dev)
        ops.gemm(Operand(dI23, MN_MAJOR, 2 * KO, N), MN_MAJOR, Operand(ca_b, MN_MAJOR, ca_b.shape[1], N), MN_MAJOR, "bf16",
                 out_dtype=f32, out=dWi_base, tag="wgrad_img_proj23")
        ops.grads_enqueued([Wi2, Wi3])               # 164 MB of gradients, final: their exchange can start now
        dca = ops.gemm(Operand(dI23, K_MAJOR, N, 2 * KO), K_MAJOR, wi_mn, MN_MAJOR, "bf16",
                       acc_into=torch.zeros((N, ca_b.shape[1]), device=dev, dtype=f32), tag="dgrad_img_proj23")
        # ---- co-attention and the MFB over the grid (:97-121 in reverse)
        dlogits_c, _ = ops.softmax_pool_bwd(Xc.view(N, Lr, D), c_att, dca, G, False, want_dx=False)
        dpre_s, _, _, _ = ops.attn_logits_bwd(hid_c, Wc2, dlogits_c, bf, out_scale=inv1, rows_per_group=Lr,
                                              zeroed=(dWc2.view(G, Jc), dbc2, dbc1))
        dWc1 = ops.wgrad(dpre_s, y1, "bf16", Wc1.shape, dest_for=Wc1, tag="wgrad_co_att_conv1")
        g1 = ops._dgrad(dpre_s, Wc1, sc, out_dtype=bf, dot_with=y1, dot_out=t1, rows_per_group=Lr,
                        tag="dgrad_co_att_conv1")
        dI1, dQ1, _ = ops.mfb_bwd(g1, y1, inv1, t1, Q1, keep1, Lr, bf, cfg.drop_p, cfg.seed_spatial, cfg.seed_dev,
                                  dbias=dbimg)
        # ---- the three question projections (:94, :124, :136 in reverse), merged
        dQ123 = torch.empty((N, 3 * KO), device=dev, dtype=bf)
        ops.pack_bf16(dQ1, out=dQ123[:, :KO])
        ops.pack_bf16(dQ23, out=dQ123[:, KO:])
        ops.colsum(dQ1, out=dbq123[:KO])
        ops.colsum(dQ23, out=dbq123[KO:])
        dWq_base, (dWq1, dWq2, dWq3) = ops.grad_buffer_group([Wq1, Wq2, Wq3], qa_b.shape[1], dev)
        ops.gemm(Operand(dQ123, MN_MAJOR, 3 * KO, N), MN_MAJOR, Operand(qa_b, MN_MAJOR, qa_b.shape[1], N), MN_MAJOR, "bf16",
                 out_dtype=f32, out=dWq_base, tag="wgrad_ques_proj123")
        ops.grads_enqueued([Wq1, Wq2, Wq3])          # 123 MB more, exchanged under the long img_conv1d wgrad below
        wq_mn = cfg.cache.get_group([Wq1, Wq2, Wq3], MN_MAJOR)
        dqa = ops.gemm(Operand(dQ123, K_MAJOR, N, 3 * KO), K_MAJOR, wq_mn, MN_MAJOR, "bf16",
                       acc_into=torch.zeros((N, qa_b.shape[1]), device=dev, dtype=f32), tag="dgrad_ques_proj123")
        # ---- question attention (:78-91 in reverse)
        need_qf = ctx.needs_input_grad[1]
        dlogits_q, dXq = ops.softmax_pool_bwd(f2.view(N, T, H), q_att, dqa, G, False, want_dx=need_qf)
        dh, _, _, _ = ops.attn_logits_bwd(hid_q, Wqa2, dlogits_q, bf, relu_mask=True,
                                          zeroed=(dWqa2.view(G, Jq), dbqa2, dbqa1))
        dWqa1 = ops.wgrad(dh, f2, "bf16", Wqa1.shape, dest_for=Wqa1, tag="wgrad_ques_att_conv1")
        if need_qf:
            ops._dgrad(dh, Wqa1, sc, acc_into=dXq.view(N * T, H), tag="dgrad_ques_att_conv1")
        # the longest GEMM of the backward pass comes last: everything the question side needs is already enqueued, and
        # the exchange of the projection gradients announced above runs beside it
        dWimg = ops.wgrad(dI1, Xc, "bf16", Wimg.shape, tag="gemm_wgrad_img_conv1d", dest_for=Wimg)
        return (None, dXq, dWqa1, dbqa1, dWqa2.view(Wqa2.shape), dbqa2,
                dWq1, dbq123[:KO], dWq2, dbq123[KO:2 * KO], dWq3, dbq123[2 * KO:],
                dWimg, dbimg, dWc1, dbc1, dWc2.view(Wc2.shape), dbc2,
                dWi2, dbi23[:KO], dWi3, dbi23[KO:], None)


class MhbCascadeFn(torch.autograd.Function):
    """The two cascaded MFB blocks of ``MHB`` (mhb_coAtt.py:189-214) on the fused-epilogue kernels:

        block 1:  (i1 + b) * mask1 * q1                      -> k-pool -> signed sqrt -> L2      (:193-203)
        block 2:  (i2 + b) * mask2 * q2 * [block 1's product] -> k-pool -> signed sqrt -> L2      (:204-212)

    The high-order coupling -- block 2 is multiplied by block 1's DROPPED-OUT product before the pooling -- happens in
    the GEMM epilogue (``extra`` / ``prod`` of vqa_b200_mfb_fused), and its two gradient paths in ``vqa_b200_mfb_bwd``
    (``dExtra`` of block 2 arrives at block 1 as ``dprod_in``).  (lstm_out [N,H], i_mean [N,D]) -> [N, 2000]."""

    @staticmethod
    def forward(ctx, qv, iv, Wq1, bq1, Wq2, bq2, Wi1, bi1, Wi2, bi2, cfg: ops.StageCfg, seed2: int):
        ops._cuda(qv, iv, Wq1, Wi1)
        mode = cfg.mode
        ad = ops._act_dtype(mode)
        need_grad = ops._need_grad(ctx)
        qv_c, iv_c = qv.contiguous(), iv.contiguous()
        if mode == "bf16":
            qv_c, iv_c = ops.pack_bf16(qv_c), ops.pack_bf16(iv_c)     # cast once: four GEMMs forward, four wgrads backward
        q1 = ops._linear_fwd(qv_c, Wq1, bq1, cfg, torch.float32)
        q2 = ops._linear_fwd(qv_c, Wq2, bq2, cfg, torch.float32)
        iop = ops.prep(iv_c, K_MAJOR, 0, mode)
        keep_dt = ad if need_grad else None
        y1, ssq1, keep1, F1 = ops.mfb_fused(iop, cfg.cache.get(Wi1, K_MAJOR, 1, mode), bi1, q1, 1, torch.float32, keep_dt,
                                            cfg.drop_p, cfg.seed, tag="mfb_fused_vector", seed_dev=cfg.seed_dev,
                                            want_prod=True)
        y2, ssq2, keep2 = ops.mfb_fused(iop, cfg.cache.get(Wi2, K_MAJOR, 1, mode), bi2, q2, 1, torch.float32, keep_dt,
                                        cfg.drop_p, seed2, tag="mfb_fused_vector", seed_dev=cfg.seed_dev, extra=F1)
        inv1, inv2 = ops.inv_norm(ssq1), ops.inv_norm(ssq2)
        out = torch.cat((ops.scale_rows(y1, inv1, 1), ops.scale_rows(y2, inv2, 1)), 1)        # :213
        if cfg.capture is not None:
            cfg.capture["y1"], cfg.capture["y2"] = y1, y2
        ctx.cfg, ctx.seed2 = cfg, seed2
        ctx.save_for_backward(qv_c, iv_c, q1, q2, F1, y1, y2, inv1, inv2, keep1, keep2, Wq1, Wq2, Wi1, Wi2)
        return out

    @staticmethod
    def backward(ctx, dout):
        qv_c, iv_c, q1, q2, F1, y1, y2, inv1, inv2, keep1, keep2, Wq1, Wq2, Wi1, Wi2 = ctx.saved_tensors
        cfg = ctx.cfg
        mode = cfg.mode
        ad = ops._act_dtype(mode)
        No = y1.shape[1]
        g2, t2 = ops.norm_bwd_prep(dout[:, No:], y2, inv2, 1)
        dI2, dQ2, dbi2, dF1 = ops.mfb_bwd(g2, y2, inv2, t2, q2, keep2, 1, ad, cfg.drop_p, ctx.seed2, cfg.seed_dev,
                                          extra=F1, want_dextra=True)
        g1, t1 = ops.norm_bwd_prep(dout[:, :No], y1, inv1, 1)
        dI1, dQ1, dbi1 = ops.mfb_bwd(g1, y1, inv1, t1, q1, keep1, 1, ad, cfg.drop_p, cfg.seed, cfg.seed_dev, dprod_in=dF1)
        iv_mn, qv_mn = ops._as_mn(iv_c), ops._as_mn(qv_c)
        dWi1 = ops.wgrad(dI1, iv_mn, mode, Wi1.shape, dest_for=Wi1)
        dWi2 = ops.wgrad(dI2, iv_mn, mode, Wi2.shape, dest_for=Wi2)
        div = None
        if ctx.needs_input_grad[1]:
            div = ops._dgrad(dI1, Wi1, cfg)
            div = div + ops._dgrad(dI2, Wi2, cfg)
        dQ1_w, dQ1_d = ops._both_layouts(dQ1, mode)
        dQ2_w, dQ2_d = ops._both_layouts(dQ2, mode)
        dWq1 = ops.wgrad(dQ1_w, qv_mn, mode, Wq1.shape, dest_for=Wq1)
        dWq2 = ops.wgrad(dQ2_w, qv_mn, mode, Wq2.shape, dest_for=Wq2)
        dqv = None
        if ctx.needs_input_grad[0]:
            dqv = ops._dgrad(dQ1_d, Wq1, cfg)
            dqv = dqv + ops._dgrad(dQ2_d, Wq2, cfg)
        return dqv, div, dWq1, ops.colsum(dQ1), dWq2, ops.colsum(dQ2), dWi1, dbi1, dWi2, dbi2, None, None
