"""Raw C-ABI kernel probe on a B200 (run under gpurun): each case runs in its own process with a
timeout so that a trap / hang in one kernel cannot take the others (or the box) down.

    python tools/gpu_probe.py            # run every case
    python tools/gpu_probe.py --case nt  # one case, in-process

torch is used only to allocate memory and as the fp32 checker.
"""
from __future__ import annotations

import argparse
import ctypes
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["nt", "nt_big", "nn", "tn", "mn_sweep", "splitk", "epi", "mfb", "mfb_bwd", "misc", "perf"]   # + "profile" (ncu target)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def rel(a, b):
    a = a.double().reshape(-1)
    b = b.double().reshape(-1)
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def run_case(case: str) -> int:
    import torch
    from vqa_attention_networks_b200 import _lib
    L = _lib.load()
    dev = torch.device("cuda:0")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cpu").manual_seed(0)
    bad = 0

    def randn(*shape):
        return torch.randn(*shape, generator=g).to(dev)

    def gemm(A, a_l, B, b_l, M, N, K, c_dtype=_lib.F32, bias=None, row_scale=None, rpg=1, relu=0, acc_into=None,
             k_split=0, dot_with=None, dot_out=None):
        if acc_into is not None:
            C = acc_into
        else:
            C = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if c_dtype == _lib.BF16 else torch.float32)
        rc = L.vqa_b200_gemm(_p(A), a_l, A.stride(0), _p(B), b_l, B.stride(0), _p(C), c_dtype, C.stride(0), M, N, K,
                             _p(bias), _p(row_scale), rpg, relu, 1 if acc_into is not None else 0, k_split,
                             _p(dot_with), dot_with.stride(0) if dot_with is not None else 0, _p(dot_out), st)
        _lib.check(rc, "gemm")
        return C

    def report(name, err, tol):
        nonlocal bad
        ok = err < tol
        bad += 0 if ok else 1
        print("  %-46s rel_err=%.3e  %s" % (name, err, "ok" if ok else "FAIL (tol %.1e)" % tol), flush=True)

    if case in ("nt", "nt_big"):
        shapes = [(128, 128, 64), (128, 128, 256), (256, 256, 512), (300, 200, 136), (15, 512, 8), (6656, 512, 1024),
                  (130, 1000, 1000)] if case == "nt" else [(50176, 512, 1000), (8192, 4096, 2048)]
        for (M, N, K) in shapes:
            A = randn(M, K).bfloat16()
            B = randn(N, K).bfloat16()
            ref = A.float() @ B.float().t()
            C = gemm(A, 0, B, 0, M, N, K)
            torch.cuda.synchronize()
            report("NT f32 M=%d N=%d K=%d" % (M, N, K), rel(C, ref), 1e-5)
            Cb = gemm(A, 0, B, 0, M, N, K, c_dtype=_lib.BF16)
            torch.cuda.synchronize()
            report("NT bf16-out M=%d N=%d K=%d" % (M, N, K), rel(Cb.float(), ref), 4e-3)
    elif case == "nn":
        for (M, N, K) in [(128, 128, 64), (256, 256, 128), (300, 1000, 512), (50176 // 8, 1000, 512)]:
            A = randn(M, K).bfloat16()
            Bm = randn(K, N).bfloat16()           # MN-major B: memory [K, N]
            ref = A.float() @ Bm.float()
            C = gemm(A, 0, Bm, 1, M, N, K)
            torch.cuda.synchronize()
            report("NN (B MN-major) M=%d N=%d K=%d" % (M, N, K), rel(C, ref), 1e-5)
    elif case == "tn":
        for (M, N, K) in [(128, 128, 64), (256, 256, 128), (512, 1000, 640), (5000, 2048, 1960), (136, 72, 200)]:
            Am = randn(K, M).bfloat16()           # MN-major A: memory [K, M]
            Bm = randn(K, N).bfloat16()
            ref = Am.float().t() @ Bm.float()
            C = gemm(Am, 1, Bm, 1, M, N, K)
            torch.cuda.synchronize()
            report("TN (A,B MN-major) M=%d N=%d K=%d" % (M, N, K), rel(C, ref), 1e-5)
            A = randn(M, K).bfloat16()
            ref2 = A.float() @ Bm.float()
            C2 = gemm(A, 0, Bm, 1, M, N, K)
            report("NN again M=%d N=%d K=%d" % (M, N, K), rel(C2, ref2), 1e-5)
    elif case == "mn_sweep":
        # which (LBO, SBO, k-advance) describes the TMA-written MN-major tile?  (default is first)
        M, N, K = 256, 256, 128
        Am = randn(K, M).bfloat16()
        Bm = randn(K, N).bfloat16()
        ref = Am.float().t() @ Bm.float()
        for (lbo, sbo, adv) in [(8192, 1024, 2048), (1024, 8192, 2048), (8192, 1024, 256), (1024, 8192, 256),
                                (128, 1024, 2048), (8192, 128, 2048), (16, 1024, 2048), (1024, 1024, 2048)]:
            L.vqa_b200_debug_set_mn_desc(lbo, sbo, adv)
            C = gemm(Am, 1, Bm, 1, M, N, K)
            torch.cuda.synchronize()
            report("TN lbo=%d sbo=%d kadv=%d" % (lbo, sbo, adv), rel(C, ref), 1e-5)
        L.vqa_b200_debug_set_mn_desc(0, 0, 0)
        bad = 0
    elif case == "splitk":
        for (M, N, K, ks) in [(512, 1000, 50176 // 4, 0), (256, 256, 4096, 7), (5000, 2048, 2048, 0)]:
            Am = randn(K, M).bfloat16()
            Bm = randn(K, N).bfloat16()
            ref = Am.float().t() @ Bm.float()
            C = torch.zeros(M, N, device=dev)
            gemm(Am, 1, Bm, 1, M, N, K, acc_into=C, k_split=ks)
            torch.cuda.synchronize()
            report("TN split-K accumulate M=%d N=%d K=%d ks=%d" % (M, N, K, ks), rel(C, ref), 1e-5)
            gemm(Am, 1, Bm, 1, M, N, K, acc_into=C, k_split=ks)
            torch.cuda.synchronize()
            report("   second accumulate (2x)", rel(C, 2 * ref), 1e-5)
    elif case == "epi":
        M, N, K, rpg = 392, 520, 264, 196
        A = randn(M, K).bfloat16()
        B = randn(N, K).bfloat16()
        bias = randn(N)
        rs = torch.rand(M // rpg, generator=g).to(dev) + 0.5
        dw = randn(M, N).bfloat16()
        ref = torch.relu((A.float() @ B.float().t()) * rs.repeat_interleave(rpg)[:, None] + bias)
        dot = torch.zeros(M // rpg, device=dev)
        C = gemm(A, 0, B, 0, M, N, K, bias=bias, row_scale=rs, rpg=rpg, relu=1, dot_with=dw, dot_out=dot)
        torch.cuda.synchronize()
        report("bias+row_scale+relu f32", rel(C, ref), 1e-5)
        report("fused group dot", rel(dot, (ref * dw.float()).reshape(M // rpg, -1).sum(1)), 1e-4)
        Cb = gemm(A, 0, B, 0, M, N, K, c_dtype=_lib.BF16, bias=bias, row_scale=rs, rpg=rpg, relu=1)
        torch.cuda.synchronize()
        report("bias+row_scale+relu bf16", rel(Cb.float(), ref), 4e-3)
    elif case == "mfb":
        for (Nb, Lr, D, p) in [(3, 6, 16, 0.0), (4, 196, 256, 0.0), (4, 196, 256, 0.1), (5, 100, 2048, 0.1), (7, 1, 64, 0.1)]:
            Nk = 5000
            M = Nb * Lr
            X = torch.relu(randn(M, D)).bfloat16()
            W = (randn(Nk, D) * 0.05).bfloat16()
            bias = randn(Nk) * 0.1
            Q = randn(Nb, Nk)
            for ydt in (_lib.F32, _lib.BF16):
                Y = torch.empty(M, Nk // 5, device=dev, dtype=torch.bfloat16 if ydt == _lib.BF16 else torch.float32)
                ssq = torch.zeros(Nb, device=dev)
                keep = torch.empty(M, Nk, device=dev, dtype=torch.bfloat16)
                seed = 1234
                rc = L.vqa_b200_mfb_fused(_p(X), X.stride(0), _p(W), W.stride(0), _p(bias), _p(Q), Q.stride(0), Lr,
                                          _p(Y), ydt, Y.stride(0), _p(ssq), _p(keep), 1, M, Nk, D, p, seed, st)
                _lib.check(rc, "mfb_fused")
                mask = torch.empty(M, Nk, device=dev)
                _lib.check(L.vqa_b200_dropout_mask(_p(mask), M, Nk, p, seed, st), "mask")
                torch.cuda.synchronize()
                I = (X.float() @ W.float().t() + bias) * mask
                z = (I * Q.repeat_interleave(Lr, 0)).reshape(M, Nk // 5, 5).sum(-1)
                yref = torch.sign(z) * torch.sqrt(z.abs())
                tol = 1e-4 if ydt == _lib.F32 else 4e-3
                report("mfb_fused N=%d L=%d D=%d p=%.1f y=%s" % (Nb, Lr, D, p, "bf16" if ydt else "f32"),
                       rel(Y.float(), yref), tol)
                report("   ssq", rel(ssq, z.abs().reshape(Nb, -1).sum(1)), 1e-4)
                report("   keep", rel(keep.float(), I), 4e-3)
                if p > 0:
                    kr = float((mask > 0).float().mean())
                    report("   keep-rate %.4f vs %.4f" % (kr, 1 - p), abs(kr - (1 - p)), 5e-3)
    elif case == "mfb_bwd":
        from vqa_attention_networks_b200 import ops
        for (Nb, Lr, p) in [(3, 6, 0.0), (4, 1, 0.0), (5, 196, 0.1)]:
            Nk, No = 5000, 1000
            M = Nb * Lr
            seed = 99
            mask = ops.dropout_mask(M, Nk, p, seed, dev)
            acc = randn(M, Nk)
            keep16 = (acc * mask).bfloat16()
            keep = keep16.float().requires_grad_(True)             # d/dkeep == d/d(acc*mask)
            Q = randn(Nb, Nk).requires_grad_(True)
            z = (keep * Q.repeat_interleave(Lr, 0)).reshape(M, No, 5).sum(-1)
            # (z == 0 happens when all five factors of a pool are dropped: sqrt'(0) = inf would turn the reference's
            #  0 * inf into NaN, while the defined gradient there -- and the kernel's -- is 0)
            y = torch.sign(z) * torch.sqrt(z.abs() + (z == 0).float())
            nrm = y.reshape(Nb, -1).norm(dim=1)
            yhat = y / nrm.repeat_interleave(Lr)[:, None]
            C = randn(M, No)
            (yhat * C).sum().backward()
            inv = (1.0 / nrm).detach()
            gg = (C * inv.repeat_interleave(Lr)[:, None]).contiguous()
            yd = y.detach().contiguous()
            t = (yd * gg).reshape(Nb, -1).sum(1).contiguous()
            dI, dQ, dbias = ops.mfb_bwd(gg, yd, inv.contiguous(), t, Q.detach(), keep16.float(), Lr, torch.float32, p, seed)
            torch.cuda.synchronize()
            dI_ref = keep.grad * mask              # d/dacc = d/dkeep * mask
            tag = "N=%d L=%d p=%.1f" % (Nb, Lr, p)
            report("mfb_bwd dI " + tag, rel(dI, dI_ref), 1e-4)
            report("mfb_bwd dQ " + tag, rel(dQ, Q.grad), 1e-4)
            report("mfb_bwd dbias " + tag, rel(dbias, dI_ref.sum(0)), 1e-4)
    elif case == "misc":
        # pack / split3
        x = randn(5, 7, 16)
        xp = x.permute(1, 0, 2)
        out = torch.empty(7, 5, 16, device=dev, dtype=torch.bfloat16)
        _lib.check(L.vqa_b200_pack_bf16(_p(xp), _p(out), 7, 5, 16, xp.stride(0), xp.stride(1), xp.stride(2), 0, 0, st), "pack")
        report("pack_bf16 strided", rel(out.float(), xp.bfloat16().float()), 1e-7)
        A = randn(33, 24)
        B = randn(17, 24)
        a3 = torch.empty(33, 72, device=dev, dtype=torch.bfloat16)
        b3 = torch.empty(17, 72, device=dev, dtype=torch.bfloat16)
        _lib.check(L.vqa_b200_split3_bf16(_p(A), 24, 0, _p(a3), 72, 0, 1, 33, 24, 0, 0, st), "split3")
        _lib.check(L.vqa_b200_split3_bf16(_p(B), 24, 0, _p(b3), 72, 0, 1, 17, 24, 1, 0, st), "split3")
        C = gemm(a3, 0, b3, 0, 33, 17, 72)
        torch.cuda.synchronize()
        report("split3 fp32-emulating GEMM", rel(C, A.double() @ B.double().t()), 3e-5)
        # logits + softmax pool fwd/bwd against autograd
        for (Nb, Lr, D, G, J, bf) in [(3, 6, 16, 2, 8, True), (4, 196, 2048, 2, 512, True), (4, 26, 1024, 2, 512, False),
                                      (2, 100, 512, 1, 512, True)]:
            X32 = torch.relu(randn(Nb, Lr, D))
            X = X32.bfloat16() if bf else X32
            H32 = torch.relu(randn(Nb * Lr, J))
            H = H32.bfloat16() if bf else H32
            W2 = (randn(G, J) * 0.2).requires_grad_(True)
            b2 = randn(G).requires_grad_(True)
            Hr = H.float().requires_grad_(True)
            Xr = X.float().requires_grad_(True)
            logits_ref = Hr @ W2.t() + b2
            att_ref = torch.softmax(logits_ref.reshape(Nb, Lr, G), dim=1)
            pooled_ref = torch.einsum("nlg,nld->ngd", att_ref, Xr).reshape(Nb, -1)
            cot = randn(Nb, G * D)
            (pooled_ref * cot).sum().backward()
            logits = torch.empty(Nb * Lr, G, device=dev)
            _lib.check(L.vqa_b200_attn_logits_fwd(_p(H), int(bf), H.stride(0), _p(W2.detach()), _p(b2.detach()), _p(logits),
                                                  Nb * Lr, J, G, st), "logits")
            att = torch.empty(Nb, G, Lr, device=dev)
            pooled = torch.empty(Nb, G * D, device=dev)
            _lib.check(L.vqa_b200_softmax_pool_fwd(_p(X), int(bf), _p(logits), _p(att), _p(pooled), Nb, Lr, D, G, 0, st), "pool")
            torch.cuda.synchronize()
            tag = "N=%d L=%d D=%d G=%d %s" % (Nb, Lr, D, G, "bf16" if bf else "f32")
            report("logits fwd " + tag, rel(logits, logits_ref), 1e-5)
            report("softmax_pool fwd " + tag, rel(pooled, pooled_ref), 1e-5)
            report("   att", rel(att, att_ref.permute(0, 2, 1)), 1e-5)
            dlog = torch.empty(Nb * Lr, G, device=dev)
            dX = torch.zeros(Nb, Lr, D, device=dev)
            _lib.check(L.vqa_b200_softmax_pool_bwd(_p(X), int(bf), _p(att), _p(cot), None, _p(dlog), _p(dX), Nb, Lr, D, G, 0, 0, st), "poolbwd")
            dH = torch.empty(Nb * Lr, J, device=dev, dtype=torch.bfloat16 if bf else torch.float32)
            dW2 = torch.zeros(G, J, device=dev)
            db2 = torch.zeros(G, device=dev)
            dbh = torch.zeros(J, device=dev)
            _lib.check(L.vqa_b200_attn_logits_bwd(_p(H), int(bf), H.stride(0), _p(W2.detach()), _p(dlog), _p(dH), int(bf),
                                                  dH.stride(0), None, 1, 1, _p(dW2), _p(db2), _p(dbh), Nb * Lr, J, G, st), "logbwd")
            torch.cuda.synchronize()
            report("softmax_pool bwd dX", rel(dX, Xr.grad), 1e-5)
            report("logits bwd dW2", rel(dW2, W2.grad), 1e-4)
            # db2 = sum of dlogits, and a softmax backward sums to zero over the sequence: the true value is pure
            # cancellation noise, so it is checked on the scale of what was summed, not of itself
            report("logits bwd db2 (abs / sum|dlogits|)",
                   float((db2 - b2.grad).abs().max() / dlog.abs().sum().clamp_min(1e-30)), 1e-5)
            dH_ref = Hr.grad * (H.float() > 0)
            report("logits bwd dH (relu-masked)", rel(dH.float(), dH_ref), 4e-3 if bf else 1e-5)
            report("logits bwd dbias_h", rel(dbh, dH_ref.sum(0)), 1e-4)
    elif case == "profile":
        # one launch of each heavy kernel at config-2 sizes (N=256) -- the target of `ncu --set full`
        from vqa_attention_networks_b200 import ops
        Nb, Lr, D, Nk = 256, 196, 2048, 5000
        M = Nb * Lr
        X = torch.relu(randn(M, D)).bfloat16()
        W = (randn(Nk, D) * 0.05).bfloat16()
        bias = randn(Nk) * 0.1
        Q = randn(Nb, Nk)
        for rep in range(2):
            Y, ssq, keep = ops.mfb_fused(ops.Operand(X, 0, M, D), ops.Operand(W, 0, Nk, D), bias, Q, Lr, torch.bfloat16,
                                         torch.bfloat16, 0.1, 7)
            inv = ops.inv_norm(ssq)
            gq = randn(M, Nk // 5).bfloat16()
            t = ops.group_dot(gq, Y, Nb, Lr)
            dI, dQ, db = ops.mfb_bwd(gq, Y, inv, t, Q, keep, Lr, torch.bfloat16, 0.1, 7)
            dW = ops.wgrad(dI, X, "bf16")
            Wc1 = (randn(512, 1000) * 0.05).bfloat16()
            hid = ops.gemm(Y, 0, Wc1, 0, "bf16", out_dtype=torch.bfloat16, bias=randn(512), row_scale=inv, rows_per_group=Lr, relu=True)
            logits = ops.attn_logits_fwd(hid, randn(2, 512), randn(2))
            pooled, att = ops.softmax_pool_fwd(X.view(Nb, Lr, D), logits, 2)
            dlog, _ = ops.softmax_pool_bwd(X.view(Nb, Lr, D), att, randn(Nb, 2 * D), 2)
            dh, dW2, db2, dbh = ops.attn_logits_bwd(hid, randn(2, 512), dlog, torch.bfloat16, out_scale=inv, rows_per_group=Lr)
            g2 = ops.gemm(dh, 0, Wc1, 1, "bf16", out_dtype=torch.bfloat16)
            dWc1 = ops.wgrad(dh, Y, "bf16")
            xp = ops.pack_bf16(randn(M // 4, D))
        torch.cuda.synchronize()
    elif case == "perf":
        def timeit(fn, iters=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        for (M, N, K) in [(8192, 8192, 8192), (50176, 5120, 2048), (50176, 512, 1024)]:
            A = randn(M, K).bfloat16()
            B = randn(N, K).bfloat16()
            C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            ms = timeit(lambda: _lib.check(L.vqa_b200_gemm(_p(A), 0, K, _p(B), 0, K, _p(C), 1, N, M, N, K, None, None, 1, 0, 0, 0,
                                                           None, 0, None, st)))
            print("  NT bf16 %dx%dx%d: %.3f ms  %.1f TFLOP/s" % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
            ms2 = timeit(lambda: torch.matmul(A, B.t()))
            print("     (cuBLAS same shape: %.3f ms  %.1f TFLOP/s)" % (ms2, 2.0 * M * N * K / ms2 / 1e9), flush=True)
        Nb, Lr, D, Nk = 256, 196, 2048, 5000
        M = Nb * Lr
        X = torch.relu(randn(M, D)).bfloat16()
        W = (randn(Nk, D) * 0.05).bfloat16()
        bias = randn(Nk) * 0.1
        Q = randn(Nb, Nk)
        Y = torch.empty(M, Nk // 5, device=dev, dtype=torch.bfloat16)
        ssq = torch.zeros(Nb, device=dev)
        keep = torch.empty(M, Nk, device=dev, dtype=torch.bfloat16)
        for (kp, p) in [(None, 0.0), (keep, 0.0), (keep, 0.1)]:
            ms = timeit(lambda: _lib.check(L.vqa_b200_mfb_fused(_p(X), D, _p(W), D, _p(bias), _p(Q), Nk, Lr, _p(Y), 1, Nk // 5,
                                                                _p(ssq), _p(kp), 1, M, Nk, D, p, 7, st)))
            print("  mfb_fused N=256 keep=%s p=%.1f: %.3f ms  %.1f TFLOP/s" % (kp is not None, p, ms, 2.0 * M * Nk * D / ms / 1e9),
                  flush=True)
        # wgrad-shaped TN split-K
        dI = randn(M, Nk).bfloat16()
        dW = torch.zeros(Nk, D, device=dev)
        ms = timeit(lambda: _lib.check(L.vqa_b200_gemm(_p(dI), 1, Nk, _p(X), 1, D, _p(dW), 0, D, Nk, D, M, None, None, 1, 0, 1, 0,
                                                       None, 0, None, st)))
        print("  TN wgrad 5000x2048x50176: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * Nk * D / ms / 1e9), flush=True)
        # pooling bandwidth
        logits = randn(M, 2)
        att = torch.empty(Nb, 2, Lr, device=dev)
        pooled = torch.empty(Nb, 2 * D, device=dev)
        Xb = X.reshape(Nb, Lr, D)
        ms = timeit(lambda: _lib.check(L.vqa_b200_softmax_pool_fwd(_p(Xb), 1, _p(logits), _p(att), _p(pooled), Nb, Lr, D, 2, 0, st)))
        byt = M * D * 2 + M * 2 * 4 + Nb * 2 * D * 4
        print("  softmax_pool fwd bf16 N=256: %.3f ms  %.0f GB/s" % (ms, byt / ms / 1e6), flush=True)
        dlog = torch.empty(M, 2, device=dev)
        cot = randn(Nb, 2 * D)
        ms = timeit(lambda: _lib.check(L.vqa_b200_softmax_pool_bwd(_p(Xb), 1, _p(att), _p(cot), None, _p(dlog), None, Nb, Lr, D, 2, 0, 0, st)))
        print("  softmax_pool bwd bf16 N=256: %.3f ms  %.0f GB/s" % (ms, byt / ms / 1e6), flush=True)
    else:
        raise SystemExit("unknown case " + case)
    torch.cuda.synchronize()
    print("case %s: %s" % (case, "OK" if bad == 0 else "%d FAILED" % bad), flush=True)
    return 1 if bad else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--timeout", type=int, default=240)
    a = ap.parse_args()
    if a.case:
        sys.exit(run_case(a.case))
    summary = {}
    for c in a.cases.split(","):
        t0 = time.time()
        print("=== case %s" % c, flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c], timeout=a.timeout)
            summary[c] = "rc=%d" % r.returncode
        except subprocess.TimeoutExpired:
            summary[c] = "TIMEOUT"
        print("=== case %s -> %s (%.1fs)" % (c, summary[c], time.time() - t0), flush=True)
    print("SUMMARY", summary, flush=True)


if __name__ == "__main__":
    main()
