"""Host-side operators: thin wrappers that allocate outputs and call the C-ABI kernels, plus the
``torch.autograd.Function``s that stitch them into the stages of the reference's forward pass.

Precision modes (``north_star``):
  * ``"bf16"`` -- bf16 operands / stored activations, fp32 accumulation (tcgen05 kind::f16);
  * ``"fp32"`` -- fp32 activations; every GEMM runs as ONE bf16 tcgen05 GEMM over a 3x longer
    contraction built from the bf16 hi/lo split of both operands (hi*hi + hi*lo + lo*hi), which
    reproduces an fp32 GEMM to ~1e-5 relative.

CUDA only: CPU tensors raise.  Nothing here falls back to PyTorch eager math for the hot path.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional

import torch

from . import _lib
from ._lib import BF16, F32, K_MAJOR, MN_MAJOR

_KO_FACTOR = 5


# --------------------------------------------------------------------------------------------
# plumbing
# --------------------------------------------------------------------------------------------
def _p(t: Optional[torch.Tensor]):
    """A pointer argument of a kernel entry point: handed on as the TENSOR (the custom op / the direct call turns it into
    its data pointer at launch time, and a recorded launch keeps its operands alive)."""
    return t


def _raw(a):
    return ctypes.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor) else a


def _st():
    """cudaStream_t of the calling thread's current stream on its current device.  Goes through the two raw C entry
    points: torch.cuda.current_stream() costs ~15 us of Python per call, 1.3 ms per 80-launch train step (cProfile)."""
    try:
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    except AttributeError:          # private entry points moved: fall back to the public (slow) API
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError("vqa_b200: unsupported dtype %s" % t.dtype)


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vqa_b200 operators run on CUDA tensors only (there is no CPU fallback)")


def _act_dtype(mode: str):
    return torch.bfloat16 if mode == "bf16" else torch.float32


def _w2d(w: torch.Tensor) -> torch.Tensor:
    """Conv2d 1x1 weights [out, in, 1, 1] are used as [out, in] matrices (a view, never a copy)."""
    return w.reshape(w.shape[0], -1) if w.dim() != 2 else w


_seed_state = {"base": None, "gen": None}


def new_seed() -> int:
    """31-bit dropout seed for one fused-epilogue dropout site.

    Determinism contract: the stream of seeds is a function of ``torch.initial_seed()`` (so ``torch.manual_seed(s)``
    makes a run reproducible) and of the data-parallel rank (``RANK``), so that replicas draw DIFFERENT masks for
    their shards, as independent per-replica RNGs would.  It comes from a private generator: drawing a seed does not
    advance the user's global CPU RNG stream."""
    base = torch.initial_seed()
    st = _seed_state
    if st["base"] != base:
        rank = int(os.environ.get("RANK", "0") or 0)
        g = torch.Generator()
        g.manual_seed((base * 1000003 + 7919 * (rank + 1)) % (2 ** 63 - 1))
        st["base"], st["gen"] = base, g
    return int(torch.randint(0, 2 ** 31 - 1, (1,), generator=st["gen"]).item())


# --------------------------------------------------------------------------------------------
# launch accounting: every C-ABI call below launches exactly one kernel of this library
# --------------------------------------------------------------------------------------------
class LaunchStats:
    """Counts kernel launches and, when ``timing`` is on, brackets launches with CUDA events on the launching stream
    (bench.py reads these for the live roofline numbers).  ``only`` restricts the bracketing to a set of tags: two
    event records per launch cost host time, and a 160-launch step that is timed end to end should not pay for 320
    of them."""
    count = 0
    timing = False
    only = None          # None = every launch, else a set of tags
    events = {}          # tag -> [(start_event, end_event), ...]

    @classmethod
    def reset(cls, timing=False, only=None):
        cls.count, cls.timing, cls.events = 0, timing, {}
        cls.only = set(only) if only is not None else None

    @classmethod
    def wants(cls, tag) -> bool:
        return cls.timing and (cls.only is None or tag in cls.only)

    @classmethod
    def summary(cls):
        """tag -> (launches, total_ms); call after torch.cuda.synchronize()."""
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in cls.events.items()}


# set by train.GraphedTrainStep while it captures an iteration: launches whose tag it wants to time at replay are kept out
# of the graphs (recorded, not issued) -- see train._Capture.intercept
_capture = None

# --------------------------------------------------------------------------------------------
# the torch custom-op layer: every kernel entry point of include/vqa_b200.h is ONE dispatcher op
# `torch.ops.vqa_b200.<name>` -- Tensor / int / float arguments in the C order, written operands declared as mutated, the
# stream taken from the caller's current stream.  The ops are thin on purpose: outputs and workspaces are allocated by the
# callers below (the C ABI allocates nothing), autograd lives in the stage-level torch.autograd.Functions, and a fake
# (meta) implementation that does nothing makes the ops traceable.  VQA_B200_DISPATCH=direct bypasses the dispatcher
# (ctypes straight from the wrappers: ~10-20 us less host time per launch for un-captured eager loops).
# --------------------------------------------------------------------------------------------
_ops_registered = False
_DIRECT = os.environ.get("VQA_B200_DISPATCH", "torch") == "direct"


def launch_direct(fn_name: str, args):
    """The kernel launch itself: tensors -> device pointers, stream = the calling thread's current stream."""
    fn = getattr(_lib.load(), fn_name)
    _lib.check(fn(*[_raw(a) for a in args], _st()), fn_name)


def _schema_of(fn_name: str) -> tuple:
    res, argtypes = _lib._PROTOTYPES[fn_name]
    mutated = _lib.MUTATED_ARGS[fn_name]
    parts, names, letters = [], [], iter("abcdefghijklmnop")
    for i, t in enumerate(argtypes[:-1]):                       # the trailing void* is the stream
        if t is ctypes.c_void_p:
            parts.append(("Tensor(%s!)? a%d" % (next(letters), i)) if i in mutated else ("Tensor? a%d" % i))
        elif t in (ctypes.c_float, ctypes.c_double):
            parts.append("float a%d" % i)
        else:
            parts.append("int a%d" % i)
        names.append("a%d" % i)
    return "(" + ", ".join(parts) + ") -> ()", [names[i] for i in mutated]


def _register_custom_ops():
    global _ops_registered
    if _ops_registered:
        return
    for fn_name in _lib.MUTATED_ARGS:
        schema, mutated = _schema_of(fn_name)
        short = fn_name[len("vqa_b200_"):]

        def impl(*args, _n=fn_name):
            launch_direct(_n, args)

        op = torch.library.custom_op("vqa_b200::" + short, impl, mutates_args=mutated, schema=schema, device_types="cuda")
        op.register_fake(lambda *a: None)
    _ops_registered = True


def _launch(fn_name: str, args):
    if _DIRECT:
        launch_direct(fn_name, args)
        return
    _register_custom_ops()
    getattr(torch.ops.vqa_b200, fn_name[len("vqa_b200_"):])(*args)


def _call(fn_name: str, tag: Optional[str], *args):
    """One kernel launch through the custom-op layer.  `args` are the entry point's arguments WITHOUT the trailing stream
    (every call site below still passes `_st()` last, which is dropped here: the op supplies the stream itself)."""
    args = args[:-1]
    LaunchStats.count += 1
    cap = _capture
    if cap is not None:
        if cap.intercept(fn_name, tag or fn_name, args):
            return
        _launch(fn_name, args)
        return
    if LaunchStats.wants(tag or fn_name):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _launch(fn_name, args)
        e1.record()
        LaunchStats.events.setdefault(tag or fn_name, []).append((e0, e1))
    else:
        _launch(fn_name, args)


# --------------------------------------------------------------------------------------------
# kernel wrappers
# --------------------------------------------------------------------------------------------
def pack_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 (any <=3-D strided view) -> contiguous bf16 of the same shape.  `out`: optional 2-D bf16 destination with
    stride(1) == 1 and any row pitch (a column slice of a wider matrix)."""
    _cuda(x)
    if out is not None:
        assert x.dim() == 2 and x.dtype == torch.float32 and out.dtype == torch.bfloat16 and out.shape == x.shape
        assert out.stride(1) == 1
        _call("vqa_b200_pack_bf16", None, _p(x), _p(out), 1, x.shape[0], x.shape[1], 0, x.stride(0), x.stride(1),
              0, out.stride(0), _st())
        return out
    if x.dtype == torch.bfloat16:
        return x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    if x.dim() > 3:
        x = x.contiguous().reshape(-1, x.shape[-1])
    shape = tuple(x.shape)
    dims = (1,) * (3 - x.dim()) + shape
    strides = (0,) * (3 - x.dim()) + tuple(x.stride())
    out = torch.empty(shape, device=x.device, dtype=torch.bfloat16)
    if out.numel() == 0:
        return out
    _call("vqa_b200_pack_bf16", None, _p(x), _p(out), dims[0], dims[1], dims[2], strides[0], strides[1], strides[2],
          0, 0, _st())
    return out


def split3(x: torch.Tensor, role: int, concat_rows: bool) -> torch.Tensor:
    """bf16 hi/lo split of an fp32 [R, C] matrix, tripled along the contraction axis."""
    _cuda(x)
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    R, C = x.shape
    out = torch.empty((3 * R, C) if concat_rows else (R, 3 * C), device=x.device, dtype=torch.bfloat16)
    _call("vqa_b200_split3_bf16", None, _p(x), x.stride(0), 0, _p(out), out.stride(0), 0, 1, R, C, role,
          int(concat_rows), _st())
    return out


# Autograd mode of the caller.  Inside autograd.Function.forward grad mode is always off, and ctx.needs_input_grad is True
# for every Parameter even under torch.no_grad() -- so a Function alone cannot tell an inference forward from a training
# one, and would save its backward state (the bf16 `keep` copy, the recurrence's gates and cell states) and pick its
# training launch forms during evaluation.  The modules record the caller's mode for the duration of a forward
# (_FusionBase._forward_scope); Functions used on their own see None = "assume autograd is on".
_outer_grad = threading.local()


def set_outer_grad(flag):
    prev = getattr(_outer_grad, "v", None)
    _outer_grad.v = flag
    return prev


def _need_grad(ctx) -> bool:
    v = getattr(_outer_grad, "v", None)
    return any(ctx.needs_input_grad) and (v is None or bool(v))


class Operand:
    """A GEMM operand already in kernel form (bf16, layout, contraction multiplier)."""
    __slots__ = ("t", "layout", "rows", "k")

    def __init__(self, t, layout, rows, k):
        self.t, self.layout, self.rows, self.k = t, layout, rows, k


def prep(x, layout: int, role: int, mode: str) -> Operand:
    """Bring a 2-D matrix into kernel form.  K-major: x is [rows, K]; MN-major: x is [K, rows]."""
    if isinstance(x, Operand):
        return x
    _cuda(x)
    assert x.dim() == 2
    rows, k = (x.shape[0], x.shape[1]) if layout == K_MAJOR else (x.shape[1], x.shape[0])
    if mode == "fp32":
        xf = x if x.dtype == torch.float32 else x.float()
        if xf.stride(1) != 1:
            xf = xf.contiguous()
        t = split3(xf, role, concat_rows=(layout == MN_MAJOR))
        return Operand(t, layout, rows, 3 * k)
    if x.dtype == torch.bfloat16 and x.stride(1) == 1:
        return Operand(x, layout, rows, k)
    memo = pack_scope.memo()              # inside a pack_scope the same tensor object is cast once
    key = ("prep", id(x), x._version)
    if memo is not None and key in memo:
        return Operand(memo[key][1], layout, rows, k)
    t = pack_bf16(x)
    if memo is not None:
        memo[key] = (x, t)                # holds x: its id cannot be re-used while the entry lives
    return Operand(t, layout, rows, k)


class WeightCache:
    """Kernel-form copies of parameters (bf16 casts / hi-lo splits / padded or transposed forms).  Lives outside the
    state dict (re-layouts must never be persisted).

    Validity.  An entry is keyed on the parameter's storage and stamped with its autograd version counter, which
    every ordinary in-place update (load_state_dict, non-fused optimizers, manual edits) bumps.  That is NOT enough
    in training: ``torch.optim.Adam(fused=True)`` updates the parameters without touching their version counters
    (measured: no re-cast happened after its step, i.e. the GEMMs would keep multiplying by the initial weights).  So
    the owning module brackets every forward with ``begin_forward(training)``; in training mode an entry is only
    used if it was produced -- or vouched for by ``refreshed()`` (optim.FusedAdam writes the new bf16 weights itself)
    -- since the previous forward began.  In eval mode the version stamp alone decides and ``module.train(mode)``
    transitions clear the cache."""

    def __init__(self):
        self._d = {}
        self._groups = {}
        self._pinned = {}           # key -> None | pending collective: entries kept up to date by someone else (adopt())
        self._epoch = 0
        self._strict = False

    def begin_forward(self, training: bool):
        self._strict = bool(training)
        if training:
            self._epoch += 1

    def _lookup(self, key, ver):
        hit = self._d.get(key)
        if hit is not None and key in self._pinned:
            pend = self._pinned[key]
            if pend is not None:
                pend.wait()          # the collective that delivers the copy: the current stream waits for it (once)
                self._pinned[key] = None
            return hit[1]
        if hit is not None and hit[0] == ver and (not self._strict or hit[2] >= self._epoch):
            return hit[1]
        return None

    # ---- externally maintained copies (ddp.GradientAllReducer with a sharded optimizer): the bf16 copy of `w` is a view of
    # a flat buffer that the optimizer shards write and an all-gather completes; it is NEVER re-derived from the fp32
    # parameter (which is stale on the ranks that do not own its shard) while it is pinned.
    @staticmethod
    def _bf16_key(w):
        return (w.data_ptr(), w.numel(), w.device.index, 0, 0, "bf16")

    def adopt(self, w: torch.Tensor, view: torch.Tensor):
        key = self._bf16_key(w)
        v2 = view.view(_w2d(w).shape)
        self._d[key] = (w._version, Operand(v2, K_MAJOR, v2.shape[0], v2.shape[1]), self._epoch)
        self._pinned[key] = None
        for gkey in [g for g in self._groups if (w.data_ptr(), w.numel()) in g]:
            del self._groups[gkey]       # rebuilt over the adopted views at the next get_group()

    def set_pending(self, w: torch.Tensor, pending):
        key = self._bf16_key(w)
        if key in self._pinned:
            self._pinned[key] = pending

    def unpin_all(self):
        self._pinned.clear()

    def get(self, w: torch.Tensor, layout: int, role: int, mode: str) -> Operand:
        key = (w.data_ptr(), w.numel(), w.device.index, layout if mode == "fp32" else 0, role if mode == "fp32" else 0,
               mode)
        ver = w._version
        op = self._lookup(key, ver)
        if op is not None:
            if mode != "fp32":      # bf16 copy is layout-agnostic: same memory serves K-major and MN-major
                rows, k = (op.t.shape[0], op.t.shape[1]) if layout == K_MAJOR else (op.t.shape[1], op.t.shape[0])
                return Operand(op.t, layout, rows, k)
            return op
        op = prep(_w2d(w.detach()), layout, role, mode)
        self._d[key] = (ver, op, self._epoch)
        return op

    def get_group(self, ws, layout: int) -> Operand:
        """bf16 copies of several [rows_i, K] weights as ONE row-concatenated operand (ques_proj1/2/3 share their input,
        as do img_proj2/3: one GEMM over [sum rows, K] instead of one launch per layer).  The per-parameter bf16 entries
        become row-slice views of the group's buffer, so validity, `bf16_entry()` and the optimizer's in-place refresh
        work per parameter exactly as for stand-alone copies."""
        gkey = tuple((w.data_ptr(), w.numel()) for w in ws)
        grp = self._groups.get(gkey)
        if grp is None:
            k = _w2d(ws[0]).shape[1]
            rows = [_w2d(w).shape[0] for w in ws]
            keys = [self._bf16_key(w) for w in ws]
            base = None
            if all(kk in self._pinned for kk in keys):
                # adopted copies: the group is the span of their (adjacent) views inside the owner's flat buffer
                vs = [self._d[kk][1].t for kk in keys]
                if all(vs[i].data_ptr() + vs[i].numel() * 2 == vs[i + 1].data_ptr() for i in range(len(vs) - 1)):
                    base = vs[0].new_empty(0).set_(vs[0].untyped_storage(), vs[0].storage_offset(), (sum(rows), k), (k, 1))
            if base is None:
                if any(kk in self._pinned for kk in keys):
                    raise RuntimeError("WeightCache.get_group: adopted weight copies of a group must be adjacent")
                base = torch.empty((sum(rows), k), device=ws[0].device, dtype=torch.bfloat16)
            views, r0 = [], 0
            for r in rows:
                views.append(base[r0:r0 + r])
                r0 += r
            grp = self._groups[gkey] = (base, views)
        base, views = grp
        for w, v in zip(ws, views):
            key = self._bf16_key(w)
            if key in self._pinned:
                self._lookup(key, w._version)             # waits for a pending all-gather, never re-casts
                continue
            hit = self._d.get(key)
            ok = (hit is not None and hit[1].t.data_ptr() == v.data_ptr() and hit[0] == w._version
                  and (not self._strict or hit[2] >= self._epoch))
            if not ok:
                v.copy_(_w2d(w.detach()))                 # fp32 -> bf16, round to nearest even (as pack_bf16)
                self._d[key] = (w._version, Operand(v, K_MAJOR, v.shape[0], v.shape[1]), self._epoch)
        rows, k = (base.shape[0], base.shape[1]) if layout == K_MAJOR else (base.shape[1], base.shape[0])
        return Operand(base, layout, rows, k)

    def bf16_entry(self, w: torch.Tensor):
        """The cached plain bf16 copy of parameter w (same shape as _w2d(w)), or None.  Used by optim.FusedAdam, which
        writes the updated weights straight into this tensor and then calls refreshed()."""
        hit = self._d.get((w.data_ptr(), w.numel(), w.device.index, 0, 0, "bf16"))
        return hit[1].t if hit is not None else None

    def refreshed(self, w: torch.Tensor):
        """The bf16 entry of w now holds the values of w's CURRENT version (written by the fused optimizer kernel): it
        is valid for the next forward."""
        key = (w.data_ptr(), w.numel(), w.device.index, 0, 0, "bf16")
        hit = self._d.get(key)
        if hit is not None:
            self._d[key] = (w._version, hit[1], self._epoch + 1)

    def get_fn(self, w: torch.Tensor, tag: str, fn):
        """Any other kernel-form derivative of a parameter (padded / transposed bf16 copies), same validity rules."""
        key = (w.data_ptr(), w.numel(), w.device.index, tag)
        ver = w._version
        v = self._lookup(key, ver)
        if v is not None:
            return v
        v = fn(w.detach())
        self._d[key] = (ver, v, self._epoch)
        return v

    def clear(self):
        """Drop every derived copy -- except the adopted ones, which their owner keeps current."""
        keep = {k: v for k, v in self._d.items() if k in self._pinned}
        self._d.clear()
        self._d.update(keep)
        if not self._pinned:
            self._groups.clear()


def gemm(A, a_layout, B, b_layout, mode, out_dtype=torch.float32, bias=None, row_scale=None, rows_per_group=1,
         relu=False, acc_into=None, k_split=0, dot_with=None, dot_out=None, tag=None, out=None) -> torch.Tensor:
    """C[m, n] = epi(sum_k A(m,k) B(n,k)) on the tcgen05 kernel.  A is the role-0 and B the role-1 operand.
    `out`: optional caller-owned [M, N] result buffer (store mode); `acc_into`: fp32 buffer accumulated into."""
    a = prep(A, a_layout, 0, mode)
    b = prep(B, b_layout, 1, mode)
    if a.k != b.k:
        raise ValueError("gemm: contraction mismatch %d vs %d" % (a.k, b.k))
    M, N, K = a.rows, b.rows, a.k
    dev = a.t.device
    if acc_into is not None:
        C = acc_into
        assert C.dtype == torch.float32 and C.shape == (M, N) and C.stride(1) == 1
    elif out is not None:
        C = out
        assert C.shape == (M, N) and C.stride(1) == 1 and C.dtype == out_dtype
    else:
        C = torch.empty((M, N), device=dev, dtype=out_dtype)
    if M == 0 or N == 0:
        return C
    _call("vqa_b200_gemm", tag, _p(a.t), a.layout, a.t.stride(0), _p(b.t), b.layout, b.t.stride(0), _p(C), _dt(C), C.stride(0),
                         M, N, K, _p(bias), _p(row_scale), rows_per_group, int(relu),
                         1 if acc_into is not None else 0, k_split,
                         _p(dot_with), dot_with.stride(0) if dot_with is not None else 0, _p(dot_out), _st())
    return C


# Gradient destinations (data-parallel training).  A reducer (ddp.GradientAllReducer) gives a weight a destination -- an
# fp32 view of the parameter's shape inside one of ITS flat buckets -- by setting two attributes on the Parameter object:
#   p._vqa_grad_dest       the view;  wgrad() then writes the weight gradient straight into the bucket (autograd adopts
#                          the returned tensor as .grad without a copy) instead of into a fresh tensor that the reducer's
#                          hook has to copy over (396 MB of device copies per step for MHBCoAtt)
#   p._vqa_grad_dest_used  True once the destination has been handed out since the reducer's prepare(): a parameter that
#                          is used twice in the graph (hieCoAtten's fc_Wbv) gets it once; the second gradient is a fresh
#                          tensor that autograd ADDS to the first
# The state lives with the parameter and is owned by the reducer that registered it (removed by its close()): nothing
# here is process-global, two reducers over two models do not see each other.
def grad_dest_of(p):
    return getattr(p, "_vqa_grad_dest", None) if p is not None else None


def grad_dest_taken(p) -> bool:
    return bool(getattr(p, "_vqa_grad_dest_used", False))


def _grad_buffer(dest_for, n_out, k_in, dev):
    dst = grad_dest_of(dest_for)
    if (dst is None or grad_dest_taken(dest_for) or dst.numel() != n_out * k_in or dst.device != dev
            or dst.dtype != torch.float32 or not dst.is_contiguous()):
        return None
    dest_for._vqa_grad_dest_used = True
    return dst.view(n_out, k_in)


# Callables `f(params)` told that the weight gradients of `params` have just been ENQUEUED in their final place (a
# data-parallel reducer may start exchanging the bucket right away instead of waiting for autograd to hand the
# gradients over at the end of the node's backward): fused_block.MhbFusedBlockFn calls grads_enqueued().
grad_ready_hooks = []


def grads_enqueued(params):
    for f in list(grad_ready_hooks):
        f(params)


def grad_buffer_group(params, k_in, dev):
    """One [sum rows, k_in] fp32 buffer for the weight gradients of several layers computed by ONE wgrad GEMM, plus the
    per-parameter row-slice views.  Inside a data-parallel reducer the buffer is the span of the parameters' adjacent
    bucket views (ddp.GradientAllReducer lays `contiguous_groups` out that way); otherwise a fresh tensor."""
    rows = [p.shape[0] for p in params]
    dsts = [grad_dest_of(p) for p in params]
    base = None
    if all(d is not None and not grad_dest_taken(p) and d.is_contiguous() and d.dtype == torch.float32
           and d.device == dev and d.numel() == r * k_in for d, p, r in zip(dsts, params, rows)):
        adjacent = all(dsts[i].data_ptr() + dsts[i].numel() * 4 == dsts[i + 1].data_ptr() for i in range(len(dsts) - 1))
        if adjacent and dsts[0].data_ptr() % 16 == 0:
            base = dsts[0].new_empty(0).set_(dsts[0].untyped_storage(), dsts[0].storage_offset(), (sum(rows), k_in),
                                             (k_in, 1))
            for p in params:
                p._vqa_grad_dest_used = True
    if base is None:
        base = torch.empty((sum(rows), k_in), device=dev, dtype=torch.float32)
    views, r0 = [], 0
    for p, r in zip(params, rows):
        views.append(base[r0:r0 + r].view(p.shape))
        r0 += r
    return base, views


def wgrad(dY, Xin, mode, out_shape=None, tag=None, dest_for=None) -> torch.Tensor:
    """dW[n_out, k_in] = sum_m dY[m, n_out] * Xin[m, k_in]: both operands MN-major, split-K, fp32 atomics.
    dest_for: the parameter this is the gradient of (see grad_dest_of)."""
    n_out = dY.rows if isinstance(dY, Operand) else dY.shape[1]
    k_in = Xin.rows if isinstance(Xin, Operand) else Xin.shape[1]
    dev = dY.t.device if isinstance(dY, Operand) else dY.device
    k_rows = dY.k if isinstance(dY, Operand) else dY.shape[0]
    dst = _grad_buffer(dest_for, n_out, k_in, dev)
    if k_rows <= 2048 and n_out * k_in >= (1 << 20):
        # short contraction, weight-sized output (vector MFB blocks at K = batch): plain stores, no zero-fill/atomics
        dW = gemm(dY, MN_MAJOR, Xin, MN_MAJOR, mode, out_dtype=torch.float32, tag=tag or "gemm_wgrad", out=dst)
        return dW if out_shape is None else dW.view(out_shape)
    dW = dst.zero_() if dst is not None else torch.zeros((n_out, k_in), device=dev, dtype=torch.float32)
    gemm(dY, MN_MAJOR, Xin, MN_MAJOR, mode, acc_into=dW, tag=tag or "gemm_wgrad")
    return dW if out_shape is None else dW.view(out_shape)


def mfb_fused(X: Operand, W: Operand, bias, Q, rows_per_group, y_dtype, keep, p: float, seed: int, tag=None,
              seed_dev=None, seg_cols=0, ssq=None, extra=None, want_prod=False):
    """keep: None (inference) or the dtype of the saved (acc + bias) * mask copy used by the backward pass.
    seed_dev: optional device step counter salting the seed (include/vqa_b200.h, "dropout")."""
    M, N, K = X.rows, W.rows, X.k
    dev = X.t.device
    groups = (M + rows_per_group - 1) // rows_per_group
    Y = torch.empty((M, N // _KO_FACTOR), device=dev, dtype=y_dtype)
    nseg = N // seg_cols if seg_cols else 1
    if ssq is None:                    # caller-provided: a zeroed slice of a pooled workspace
        ssq = torch.zeros(groups * nseg, device=dev, dtype=torch.float32)
    kp = torch.empty((M, N), device=dev, dtype=keep) if keep is not None else None
    prod = torch.empty((M, N), device=dev, dtype=torch.float32) if want_prod else None
    if extra is not None:
        assert extra.dtype == torch.float32 and extra.stride(0) == Q.stride(0) and extra.stride(1) == 1
    _call("vqa_b200_mfb_fused", tag, _p(X.t), X.t.stride(0), _p(W.t), W.t.stride(0), _p(bias), _p(Q), Q.stride(0),
                              rows_per_group, _p(Y), _dt(Y), Y.stride(0), _p(ssq), _p(kp), _dt(kp) if kp is not None else BF16, M, N, K, int(seg_cols),
                              _p(extra), _p(prod), float(p), int(seed) & 0xFFFFFFFF, _p(seed_dev), _st())
    if want_prod:
        return Y, ssq, kp, prod
    return Y, ssq, kp


def dropout_mask(M, N, p, seed, device, seed_dev=None) -> torch.Tensor:
    """The pre-scaled mask mfb_fused applies (test hook: lets the oracle run with the identical mask)."""
    mask = torch.empty((M, N), device=device, dtype=torch.float32)
    _call("vqa_b200_dropout_mask", None, _p(mask), M, N, float(p), int(seed) & 0xFFFFFFFF, _p(seed_dev), _st())
    return mask


def inv_norm(ssq):
    inv = torch.empty_like(ssq)
    _call("vqa_b200_inv_norm", None, _p(ssq), _p(inv), ssq.numel(), _st())
    return inv


def scale_rows(Y, inv, rows_per_group):
    M, No = Y.shape
    out = torch.empty((M, No), device=Y.device, dtype=torch.float32)
    _call("vqa_b200_scale_rows", None, _p(Y), _dt(Y), Y.stride(0), _p(inv), rows_per_group, _p(out), out.stride(0), M, No,
                                     _st())
    return out


def attn_logits_fwd(H, W2, b2):
    M, J = H.shape
    G = W2.shape[0]
    W2 = _w2d(W2).contiguous()
    logits = torch.empty((M, G), device=H.device, dtype=torch.float32)
    _call("vqa_b200_attn_logits_fwd", None, _p(H), _dt(H), H.stride(0), _p(W2), _p(b2), _p(logits), M, J, G, _st())
    return logits


def attn_logits_bwd(H, W2, dlogits, out_dtype, out_scale=None, rows_per_group=1, relu_mask=True, zeroed=None):
    """zeroed: optional (dW2 [G, J], db2 [G], dbh [J]) views of an already zero-filled workspace."""
    M, J = H.shape
    G = W2.shape[0]
    W2 = _w2d(W2).contiguous()
    dH = torch.empty((M, J), device=H.device, dtype=out_dtype)
    if zeroed is not None:
        dW2, db2, dbh = zeroed
    else:
        dW2 = torch.zeros((G, J), device=H.device, dtype=torch.float32)
        db2 = torch.zeros(G, device=H.device, dtype=torch.float32)
        dbh = torch.zeros(J, device=H.device, dtype=torch.float32)
    _call("vqa_b200_attn_logits_bwd", None, _p(H), _dt(H), H.stride(0), _p(W2), _p(dlogits), _p(dH), _dt(dH),
                                          dH.stride(0), _p(out_scale), rows_per_group, int(relu_mask), _p(dW2),
                                          _p(db2), _p(dbh), M, J, G, _st())
    return dH, dW2, db2, dbh


def softmax_pool_fwd(X3, logits, G, degenerate=False, tag=None):
    N, Lr, D = X3.shape
    att = torch.empty((N, G, Lr), device=X3.device, dtype=torch.float32)
    pooled = torch.empty((N, G * D), device=X3.device, dtype=torch.float32)
    _call("vqa_b200_softmax_pool_fwd", tag, _p(X3), _dt(X3), _p(logits), _p(att), _p(pooled), N, Lr, D, G,
                                           int(degenerate), _st())
    return pooled, att


def softmax_pool_bwd(X3, att, dpooled, G, degenerate=False, want_dx=False, datt_extra=None):
    N, Lr, D = X3.shape
    # dlogits and the N completion tickets of the kernel back to back: the library zeroes both with one memset
    buf = torch.empty(N * Lr * G + N, device=X3.device, dtype=torch.float32)
    dlogits = buf[:N * Lr * G].view(N * Lr, G)
    done = buf[N * Lr * G:]
    dX = torch.empty((N, Lr, D), device=X3.device, dtype=torch.float32) if want_dx else None
    dpooled = dpooled.contiguous()
    _call("vqa_b200_softmax_pool_bwd", None, _p(X3), _dt(X3), _p(att), _p(dpooled), _p(datt_extra), _p(dlogits), _p(done),
                                           _p(dX), N, Lr, D, G, int(degenerate), 0, _st())
    return dlogits, dX


def mfb_bwd(g, Y, inv, t, Q, keep, rows_per_group, di_dtype, p, seed, seed_dev=None, seg_cols=0, dbias=None,
            extra=None, dprod_in=None, want_dextra=False):
    M, No = Y.shape
    N = No * _KO_FACTOR
    groups = (M + rows_per_group - 1) // rows_per_group
    dI = torch.empty((M, N), device=Y.device, dtype=di_dtype)
    dQ = torch.empty((groups, N), device=Y.device, dtype=torch.float32)
    if dbias is None:
        dbias = torch.zeros(N, device=Y.device, dtype=torch.float32)
    dextra = torch.empty((groups, N), device=Y.device, dtype=torch.float32) if want_dextra else None
    if dprod_in is not None:
        dprod_in = dprod_in.contiguous()
    _call("vqa_b200_mfb_bwd", None, _p(g), _dt(g), g.stride(0), _p(Y), _dt(Y), Y.stride(0), _p(inv), _p(t), _p(Q),
                                  Q.stride(0), _p(keep), _dt(keep), _p(dI), _dt(dI), _p(dQ), _p(dbias), rows_per_group, M, N,
                                  int(seg_cols), _p(extra), _p(dprod_in), _p(dextra), float(p), int(seed) & 0xFFFFFFFF,
                                  _p(seed_dev), _st())
    if want_dextra:
        return dI, dQ, dbias, dextra
    return dI, dQ, dbias


def norm_bwd_prep(d, Y, inv, rows_per_group, t=None):
    M, No = Y.shape
    d = d.contiguous()
    g = torch.empty((M, No), device=Y.device, dtype=torch.float32)
    if t is None:                      # else: a zeroed slice of a pooled workspace
        t = torch.zeros(inv.numel(), device=Y.device, dtype=torch.float32)
    _call("vqa_b200_norm_bwd_prep", None, _p(d), d.stride(0), _p(Y), _dt(Y), Y.stride(0), _p(inv), _p(g), g.stride(0),
                                        _p(t), rows_per_group, M, No, _st())
    return g, t


def group_dot(A, B, groups, rows_per_group):
    M, No = A.shape
    t = torch.zeros(groups, device=A.device, dtype=torch.float32)
    _call("vqa_b200_group_dot", None, _p(A), _dt(A), A.stride(0), _p(B), _dt(B), B.stride(0), _p(t), rows_per_group, M,
                                    No, _st())
    return t


def colsum(X, out=None):
    M, J = X.shape
    if out is None:                    # else: a zeroed slice of a pooled workspace
        out = torch.zeros(J, device=X.device, dtype=torch.float32)
    _call("vqa_b200_colsum", None, _p(X), _dt(X), X.stride(0), _p(out), M, J, _st())
    return out


def log_softmax_argmax(logits: torch.Tensor, want_logp: bool = True):
    """(log_softmax(logits, 1), argmax(logits, 1) int64, log-prob of the argmax) for fp32 logits [M, N]: the classifier
    tail of the eval path (mhb_coAtt.py:149-151, solver.py:148-149) in one kernel."""
    _cuda(logits)
    assert logits.dim() == 2 and logits.dtype == torch.float32 and logits.stride(1) == 1
    M, N = logits.shape
    logp = torch.empty((M, N), device=logits.device, dtype=torch.float32) if want_logp else None
    pred = torch.empty(M, device=logits.device, dtype=torch.int64)
    plp = torch.empty(M, device=logits.device, dtype=torch.float32)
    _call("vqa_b200_logsoftmax_argmax", None, _p(logits), logits.stride(0), _p(logp), N, _p(pred), _p(plp), M, N, _st())
    return logp, pred, plp


def relu_bwd(D, H, out_dtype, scale=None, rows_per_group=1):
    M, J = H.shape
    out = torch.empty((M, J), device=H.device, dtype=out_dtype)
    dbias = torch.zeros(J, device=H.device, dtype=torch.float32)
    _call("vqa_b200_relu_bwd", None, _p(D), _dt(D), D.stride(0), _p(H), _dt(H), H.stride(0), _p(out), _dt(out),
                                   out.stride(0), _p(scale), rows_per_group, _p(dbias), M, J, _st())
    return out, dbias


# --------------------------------------------------------------------------------------------
# stage-level autograd functions
# --------------------------------------------------------------------------------------------
class StageCfg:
    """Non-tensor settings threaded through the autograd functions."""

    def __init__(self, mode="bf16", cache: Optional[WeightCache] = None, degenerate=False, drop_p=0.0, seed=0,
                 capture: Optional[dict] = None, key: str = "", seed_dev: Optional[torch.Tensor] = None):
        if mode not in ("bf16", "fp32"):
            raise ValueError("precision mode must be 'bf16' or 'fp32'")
        self.mode, self.cache, self.degenerate = mode, cache or WeightCache(), degenerate
        self.drop_p, self.seed = drop_p, seed
        # device step counter that salts every dropout seed of this stage (CUDA-graph replay: the host seed is frozen
        # into the captured launches, the counter is incremented by the graph itself); None = host seeds only
        self.seed_dev = seed_dev
        # test hook: when a dict is given, the signed-sqrt outputs y of the MFB blocks are stored under `key`
        # (tests inject z = sign(y) y^2 into the oracle so that d(signed-sqrt) is evaluated at identical points)
        self.capture, self.key = capture, key


def _both_layouts(x: torch.Tensor, mode: str):
    """bf16 mode: cast a [rows, cols] gradient ONCE and hand it out both as the MN-major wgrad operand and as the K-major
    dgrad operand (same memory); fp32 mode: the raw tensor (the hi/lo splits differ per role)."""
    if mode != "bf16":
        return x, x
    t = x if (x.dtype == torch.bfloat16 and x.stride(1) == 1) else pack_bf16(x)
    return Operand(t, MN_MAJOR, t.shape[1], t.shape[0]), Operand(t, K_MAJOR, t.shape[0], t.shape[1])


def _as_mn(x: torch.Tensor):
    """A saved [rows, cols] activation as the MN-major wgrad operand: the bf16 cast the forward already made is used
    as it is, anything else goes through prep()."""
    if x.dtype == torch.bfloat16 and x.stride(1) == 1:
        return Operand(x, MN_MAJOR, x.shape[1], x.shape[0])
    return x


def _few_tiles(M: int, N: int, K: int) -> bool:
    """Small-M weight-streaming GEMM (the [N=256 samples, 2048..5000] vector projections): so few output tiles that a
    whole-tile decomposition leaves most SMs idle while each busy one walks a long contraction alone."""
    return ((M + 127) // 128) * ((N + 127) // 128) <= 74 and K >= 1024


def _linear_fwd(x, W, b, cfg: StageCfg, out_dtype, relu=False, row_scale=None, rows_per_group=1, tag=None):
    """x [M, K] @ W[N, K]^T + b with the epilogue fused (nn.Linear / 1x1 nn.Conv2d forward)."""
    wop = cfg.cache.get(W, K_MAJOR, 1, cfg.mode)
    xop = prep(x, K_MAJOR, 0, cfg.mode)
    if out_dtype == torch.float32 and not relu and row_scale is None and _few_tiles(xop.rows, wop.rows, xop.k):
        # split the contraction over the idle SMs: C starts as the broadcast bias and the K-slices accumulate into it
        # (.clone(), not .contiguous(): for a single row the expanded bias IS contiguous and would alias the parameter)
        C = (b.detach().to(torch.float32).expand(xop.rows, wop.rows).clone(memory_format=torch.contiguous_format)
             if b is not None else torch.zeros((xop.rows, wop.rows), device=xop.t.device, dtype=torch.float32))
        return gemm(xop, K_MAJOR, wop, K_MAJOR, cfg.mode, acc_into=C, tag=tag or "gemm_fwd")
    return gemm(xop, K_MAJOR, wop, K_MAJOR, cfg.mode, out_dtype=out_dtype, bias=b, relu=relu, row_scale=row_scale,
                rows_per_group=rows_per_group, tag=tag or "gemm_fwd")


def _dgrad(dY, W, cfg: StageCfg, out_dtype=torch.float32, acc_into=None, **kw):
    """dX[M, K] = dY[M, N] @ W[N, K]: W is consumed as the MN-major B operand (no transposed copy)."""
    wop = cfg.cache.get(W, MN_MAJOR, 1, cfg.mode)
    kw.setdefault("tag", "gemm_dgrad")
    dop = prep(dY, K_MAJOR, 0, cfg.mode)
    if (acc_into is None and out_dtype == torch.float32 and kw.get("dot_with") is None
            and _few_tiles(dop.rows, wop.rows, dop.k)):
        acc_into = torch.zeros((dop.rows, wop.rows), device=dop.t.device, dtype=torch.float32)
    return gemm(dop, K_MAJOR, wop, MN_MAJOR, cfg.mode, out_dtype=out_dtype, acc_into=acc_into, **kw)


class LinearFn(torch.autograd.Function):
    """nn.Linear on the tcgen05 GEMM: forward NT, dgrad NN, wgrad TN (hieCoAtten.py:25,30-36; mhb_coAtt.py:94)."""

    @staticmethod
    def forward(ctx, x, W, b, cfg: StageCfg, relu=False):
        _cuda(x, W)
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        xin = x2 if cfg.mode == "fp32" else pack_bf16(x2)
        y = _linear_fwd(xin, W, b, cfg, torch.float32, relu=relu)
        ctx.cfg, ctx.relu, ctx.shp = cfg, relu, shp
        ctx.save_for_backward(xin, W, y if relu else None)
        ctx.has_bias = b is not None
        return y.view(*shp[:-1], W.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xin, W, y = ctx.saved_tensors
        cfg = ctx.cfg
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        db = None
        if ctx.relu:
            dy2, db = relu_bwd(dy2, y, _act_dtype(cfg.mode))
        elif ctx.has_bias:
            db = colsum(dy2)
        dyo = dy2 if cfg.mode == "fp32" else pack_bf16(dy2)
        dW = wgrad(dyo, xin, cfg.mode, W.shape, dest_for=W) if ctx.needs_input_grad[1] else None
        dx = _dgrad(dyo, W, cfg).view(ctx.shp) if ctx.needs_input_grad[0] else None
        return dx, dW, (db if ctx.has_bias else None), None, None


class AttnPoolFn(torch.autograd.Function):
    """1x1 conv -> ReLU -> [1x1 conv -> ReLU] -> 1x1 conv(->G) -> softmax over the sequence axis ->
    G-glimpse weighted pooling of the SAME features (question attention, mhb_coAtt.py:78-91, mfb.py:73-89)."""

    @staticmethod
    def forward(ctx, feat, W1, b1, Wm, bm, W2, b2, cfg: StageCfg):
        _cuda(feat, W1, W2)
        N, T, H = feat.shape
        ad = _act_dtype(cfg.mode)
        f2 = (pack_bf16(feat) if cfg.mode == "bf16" else feat.float().contiguous()).view(N * T, H)
        G = W2.shape[0]
        if cfg.degenerate:
            # mfb.py:84 -- softmax over a size-1 axis: weights are exactly 1, the logits are dead code
            logits = torch.zeros((N * T, G), device=feat.device, dtype=torch.float32)
            pooled, att = softmax_pool_fwd(f2.view(N, T, H), logits, G, True)
            ctx.cfg, ctx.dims = cfg, (N, T, H, G)
            ctx.save_for_backward(f2, None, None, att, W1, Wm, W2)
            ctx.mark_non_differentiable(att)
            return pooled, att
        hid = _linear_fwd(f2, W1, b1, cfg, ad, relu=True)
        hid2 = _linear_fwd(hid, Wm, bm, cfg, ad, relu=True) if Wm is not None else None
        last = hid2 if hid2 is not None else hid
        logits = attn_logits_fwd(last, W2, b2)
        pooled, att = softmax_pool_fwd(f2.view(N, T, H), logits, G, cfg.degenerate)
        ctx.cfg, ctx.dims = cfg, (N, T, H, G)
        ctx.save_for_backward(f2, hid, hid2, att, W1, Wm, W2)
        ctx.mark_non_differentiable(att)
        return pooled, att

    @staticmethod
    def backward(ctx, dpooled, _datt):
        f2, hid, hid2, att, W1, Wm, W2 = ctx.saved_tensors
        cfg = ctx.cfg
        N, T, H, G = ctx.dims
        ad = _act_dtype(cfg.mode)
        need_x = ctx.needs_input_grad[0]
        dlogits, dX = softmax_pool_bwd(f2.view(N, T, H), att, dpooled, G, cfg.degenerate, want_dx=need_x)
        if cfg.degenerate:
            # softmax over a singleton axis (mfb.py:84): no gradient reaches the logits -> exact zeros
            z = torch.zeros_like
            return (dX, z(W1), z(W1[:, 0, 0, 0] if W1.dim() == 4 else W1[:, 0]),
                    z(Wm) if Wm is not None else None,
                    z(Wm[:, 0, 0, 0] if Wm.dim() == 4 else Wm[:, 0]) if Wm is not None else None,
                    z(W2), z(W2[:, 0, 0, 0] if W2.dim() == 4 else W2[:, 0]), None)
        last = hid2 if hid2 is not None else hid
        dh, dW2, db2, dblast = attn_logits_bwd(last, W2, dlogits, ad, relu_mask=True)
        dWm = dbm = None
        if hid2 is not None:
            dWm = wgrad(dh, hid, cfg.mode, Wm.shape, dest_for=Wm)
            dbm = dblast
            dhid = _dgrad(dh, Wm, cfg, out_dtype=ad)
            dpre, db1 = relu_bwd(dhid, hid, ad)
        else:
            dpre, db1 = dh, dblast
        dW1 = wgrad(dpre, f2, cfg.mode, W1.shape, dest_for=W1)
        if need_x:
            _dgrad(dpre, W1, cfg, acc_into=dX.view(N * T, H))
        return dX, dW1, db1, dWm, dbm, dW2.view(W2.shape), db2, None


class MfbSpatialCoAttFn(torch.autograd.Function):
    """ques_proj1 -> [img_conv1d GEMM + Hadamard + dropout + k-pool + signed sqrt] -> L2 norm folded into
    co_att_conv1 -> ReLU -> [multiconv] -> co_att_conv2 -> softmax over regions -> G-glimpse pooling of the raw
    image features (mhb_coAtt.py:94-121, mfb.py:92-123)."""

    @staticmethod
    def forward(ctx, X, qa, Wq1, bq1, Wimg, bimg, Wc1, bc1, Wcm, bcm, Wc2, bc2, cfg: StageCfg):
        _cuda(X, qa, Wimg)
        if ctx.needs_input_grad[0]:
            raise RuntimeError("vqa_b200: gradients w.r.t. the image features are not part of the path")
        N, Lr, D = X.shape
        M = N * Lr
        mode = cfg.mode
        ad = _act_dtype(mode)
        Xc = (pack_bf16(X) if mode == "bf16" else X.float().contiguous()).view(M, D)
        qa_c = qa.contiguous()
        G = Wc2.shape[0]
        if cfg.degenerate:
            # mfb.py:118 -- all-ones attention: the pooled feature is a plain sum over the regions and the
            # image projection / MFB / co-attention convs are dead code: not executed.
            logits = torch.zeros((M, G), device=X.device, dtype=torch.float32)
            ca, att = softmax_pool_fwd(Xc.view(N, Lr, D), logits, G, True, tag="softmax_pool_fwd_regions")
            ctx.cfg, ctx.dims = cfg, (N, Lr, D, G)
            ctx.save_for_backward(Xc, qa_c, None, None, None, None, None, None, att, Wq1, Wimg, Wc1, Wcm, Wc2)
            ctx.mark_non_differentiable(att)
            return ca, att
        qa_k = prep(qa_c, K_MAJOR, 0, mode) if mode == "bf16" else None     # cast once: forward GEMM and backward wgrad
        Q1 = _linear_fwd(qa_k if qa_k is not None else qa_c, Wq1, bq1, cfg, torch.float32)
        if qa_k is not None:
            qa_c = qa_k.t                 # the saved tensor (only its values / shape are used in backward)
        need_grad = _need_grad(ctx)
        xop = prep(Xc, K_MAJOR, 0, mode)
        wop = cfg.cache.get(Wimg, K_MAJOR, 1, mode)
        y, ssq, keep = mfb_fused(xop, wop, bimg, Q1, Lr, ad, ad if need_grad else None, cfg.drop_p, cfg.seed,
                                 tag="mfb_fused_spatial", seed_dev=cfg.seed_dev)
        if cfg.capture is not None:
            cfg.capture[cfg.key] = y
        inv = inv_norm(ssq)
        hid = _linear_fwd(y, Wc1, bc1, cfg, ad, relu=True, row_scale=inv, rows_per_group=Lr, tag="gemm_co_att_conv1")
        hid2 = _linear_fwd(hid, Wcm, bcm, cfg, ad, relu=True) if Wcm is not None else None
        last = hid2 if hid2 is not None else hid
        logits = attn_logits_fwd(last, Wc2, bc2)
        ca, att = softmax_pool_fwd(Xc.view(N, Lr, D), logits, G, cfg.degenerate, tag="softmax_pool_fwd_regions")
        ctx.cfg, ctx.dims = cfg, (N, Lr, D, G)
        ctx.save_for_backward(Xc, qa_c, Q1, y, inv, keep, hid, hid2, att, Wq1, Wimg, Wc1, Wcm, Wc2)
        ctx.mark_non_differentiable(att)
        return ca, att

    @staticmethod
    def backward(ctx, dca, _datt):
        Xc, qa_c, Q1, y, inv, keep, hid, hid2, att, Wq1, Wimg, Wc1, Wcm, Wc2 = ctx.saved_tensors
        cfg = ctx.cfg
        mode = cfg.mode
        ad = _act_dtype(mode)
        N, Lr, D, G = ctx.dims
        if cfg.degenerate:
            # mfb.py:118 -- softmax over a singleton axis: the whole first stage is dead;
            # the reference produces exactly-zero (not None) gradients here.
            z = torch.zeros_like
            return (None, torch.zeros(qa_c.shape, device=qa_c.device, dtype=torch.float32), z(Wq1), z(Wq1[:, 0]), z(Wimg),
                    z(Wimg[:, 0, 0, 0]), z(Wc1), z(Wc1[:, 0, 0, 0]),
                    z(Wcm) if Wcm is not None else None, z(Wcm[:, 0, 0, 0]) if Wcm is not None else None,
                    z(Wc2), z(Wc2[:, 0, 0, 0]), None)
        dlogits, _ = softmax_pool_bwd(Xc.view(N, Lr, D), att, dca, G, False, want_dx=False)
        dWcm = dbcm = None
        if hid2 is None:
            dpre_s, dWc2, dbc2, dbc1 = attn_logits_bwd(hid, Wc2, dlogits, ad, out_scale=inv, rows_per_group=Lr)
        else:
            dh2, dWc2, dbc2, dbcm = attn_logits_bwd(hid2, Wc2, dlogits, ad)
            dWcm = wgrad(dh2, hid, mode, Wcm.shape, dest_for=Wcm)
            dhid = _dgrad(dh2, Wcm, cfg, out_dtype=ad)
            dpre_s, dbc1 = relu_bwd(dhid, hid, ad, scale=inv, rows_per_group=Lr)
        # co_att_conv1: dW = (dpre * inv)^T y ;  g = (dpre * inv) W  (= d/dy_hat * inv)
        dWc1 = wgrad(dpre_s, y, mode, Wc1.shape, dest_for=Wc1)
        if mode == "bf16":
            t = torch.zeros(N, device=y.device, dtype=torch.float32)
            g = _dgrad(dpre_s, Wc1, cfg, out_dtype=ad, dot_with=y, dot_out=t, rows_per_group=Lr)
        else:
            g = _dgrad(dpre_s, Wc1, cfg, out_dtype=ad)
            t = group_dot(g, y, N, Lr)
        dI, dQ1, dbimg = mfb_bwd(g, y, inv, t, Q1, keep, Lr, ad, cfg.drop_p, cfg.seed, cfg.seed_dev)
        dWimg = wgrad(dI, Xc, mode, Wimg.shape, tag="gemm_wgrad_img_conv1d", dest_for=Wimg)
        dQ1_w, dQ1_d = _both_layouts(dQ1, mode)
        dWq1 = wgrad(dQ1_w, _as_mn(qa_c), mode, Wq1.shape, dest_for=Wq1)
        dbq1 = colsum(dQ1)
        dqa = _dgrad(dQ1_d, Wq1, cfg) if ctx.needs_input_grad[1] else None
        return (None, dqa, dWq1, dbq1, dWimg, dbimg, dWc1, dbc1, dWcm, dbcm, dWc2.view(Wc2.shape), dbc2, None)


class MfbVectorFn(torch.autograd.Function):
    """MFB block on pooled vectors: ques_proj ⊙ img_proj -> dropout -> k-pool -> signed sqrt -> L2 normalise
    (mhb_coAtt.py:124-133,136-145; mfb.py:126-135).  The image projection GEMM carries the fused epilogue."""

    @staticmethod
    def forward(ctx, qa, ca, Wq, bq, Wi, bi, cfg: StageCfg):
        _cuda(qa, ca, Wq, Wi)
        mode = cfg.mode
        qa_c, ca_c = qa.contiguous(), ca.contiguous()
        qa_k = prep(qa_c, K_MAJOR, 0, mode) if mode == "bf16" else None     # cast once: forward GEMMs and backward wgrads
        Qb = _linear_fwd(qa_k if qa_k is not None else qa_c, Wq, bq, cfg, torch.float32)
        need_grad = _need_grad(ctx)
        xop = prep(ca_c, K_MAJOR, 0, mode)
        if mode == "bf16":
            qa_c, ca_c = qa_k.t, xop.t
        wop = cfg.cache.get(Wi, K_MAJOR, 1, mode)
        y, ssq, keep = mfb_fused(xop, wop, bi, Qb, 1, torch.float32, _act_dtype(mode) if need_grad else None, cfg.drop_p,
                                 cfg.seed, tag="mfb_fused_vector", seed_dev=cfg.seed_dev)
        if cfg.capture is not None:
            cfg.capture[cfg.key] = y
        inv = inv_norm(ssq)
        out = scale_rows(y, inv, 1)
        ctx.cfg = cfg
        ctx.save_for_backward(qa_c, ca_c, Qb, y, inv, keep, Wq, Wi)
        return out

    @staticmethod
    def backward(ctx, dout):
        qa_c, ca_c, Qb, y, inv, keep, Wq, Wi = ctx.saved_tensors
        cfg = ctx.cfg
        mode = cfg.mode
        ad = _act_dtype(mode)
        g, t = norm_bwd_prep(dout, y, inv, 1)
        dI, dQ, dbi = mfb_bwd(g, y, inv, t, Qb, keep, 1, ad, cfg.drop_p, cfg.seed, cfg.seed_dev)
        dWi = wgrad(dI, _as_mn(ca_c), mode, Wi.shape, dest_for=Wi)
        dca = _dgrad(dI, Wi, cfg) if ctx.needs_input_grad[1] else None
        dQ_w, dQ_d = _both_layouts(dQ, mode)
        dWq = wgrad(dQ_w, _as_mn(qa_c), mode, Wq.shape, dest_for=Wq)
        dbq = colsum(dQ)
        dqa = _dgrad(dQ_d, Wq, cfg) if ctx.needs_input_grad[0] else None
        return dqa, dca, dWq, dbq, dWi, dbi, None


# --------------------------------------------------------------------------------------------
# batched / extended-epilogue GEMM and the blocks of hieCoAtten.py and modules.py
# --------------------------------------------------------------------------------------------
def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def alloc_padded(shape, dtype, device) -> torch.Tensor:
    """[..., R, C] tensor whose row pitch is a multiple of 8 elements (TMA: strides must be 16-byte multiples even
    when C is 196 regions or 26 tokens).  The returned tensor is a view with stride(-1) == 1."""
    *lead, R, C = shape
    Cp = _pad8(C)
    buf = torch.zeros((*lead, R, Cp), device=device, dtype=dtype) if Cp != C else \
        torch.empty((*lead, R, C), device=device, dtype=dtype)
    return buf[..., :C]


class pack_scope:
    """Within the scope, bf16 packs of the same tensor object are made once (HieCoAtten feeds `img` to three Linear
    layers and `C`, `img_`, `que_` to two products each).  Entries hold a reference to their source tensor, so an
    address can never be re-used while its memo entry is alive; the memo dies with the scope."""
    _tls = threading.local()

    def __enter__(self):
        self._prev = getattr(pack_scope._tls, "memo", None)
        pack_scope._tls.memo = {}
        return self

    def __exit__(self, *exc):
        pack_scope._tls.memo = self._prev
        return False

    @staticmethod
    def memo():
        return getattr(pack_scope._tls, "memo", None)


def _prep3(x: torch.Tensor, layout: int, role: int, mode: str):
    """3-D operand [B, R, C] (fp32 or bf16, stride(-1) == 1) -> (bf16 tensor, ld, bstride, rows, k)."""
    _cuda(x)
    assert x.dim() == 3
    Bn, R, C = x.shape
    rows, k = (R, C) if layout == K_MAJOR else (C, R)
    if mode == "fp32":
        xf = x if x.dtype == torch.float32 else x.float()
        if xf.stride(2) != 1:
            xf = xf.contiguous()
        concat_rows = layout == MN_MAJOR
        out = alloc_padded((Bn, 3 * R, C) if concat_rows else (Bn, R, 3 * C), torch.bfloat16, x.device)
        _call("vqa_b200_split3_bf16", None, _p(xf), xf.stride(1), xf.stride(0), _p(out), out.stride(1), out.stride(0), Bn,
              R, C, role, int(concat_rows), _st())
        return out, out.stride(1), out.stride(0), rows, 3 * k
    ok = (x.dtype == torch.bfloat16 and x.stride(2) == 1 and x.stride(1) % 8 == 0 and x.stride(0) % 8 == 0 and
          x.data_ptr() % 16 == 0)
    if ok:
        return x, x.stride(1), x.stride(0), rows, k
    memo = pack_scope.memo()
    key = (id(x), x._version)
    if memo is not None and key in memo:
        out = memo[key][1]
        return out, out.stride(1), out.stride(0), rows, k
    xf = x if x.dtype == torch.float32 else x.float()
    out = alloc_padded((Bn, R, C), torch.bfloat16, x.device)
    _call("vqa_b200_pack_bf16", None, _p(xf), _p(out), Bn, R, C, xf.stride(0), xf.stride(1), xf.stride(2), out.stride(0),
          out.stride(1), _st())
    if memo is not None:
        memo[key] = (x, out)
    return out, out.stride(1), out.stride(0), rows, k


def gemm_ex(A, a_layout, B, b_layout, mode, bias=None, act=0, add=None, drop_p=0.0, seed=0, out=None, accumulate=False,
            tag=None, seed_dev=None) -> torch.Tensor:
    """C[b] = dropout(act(A_b B_b^T + bias + add[b])) for 3-D operands (batch first); fp32 output [B, M, N] whose row
    pitch is padded to a multiple of 8 so that it can be re-used as a TMA operand."""
    a, lda, abs_, M, K = _prep3(A, a_layout, 0, mode)
    b, ldb, bbs, N, K2 = _prep3(B, b_layout, 1, mode)
    if K != K2 or A.shape[0] != B.shape[0]:
        raise ValueError("gemm_ex: shape mismatch")
    Bn = A.shape[0]
    C = out if out is not None else alloc_padded((Bn, M, N), torch.float32, A.device)
    if add is not None:
        assert add.dtype == torch.float32 and add.shape == C.shape and add.stride() == C.stride(), \
            "addend must share C's layout"
    _call("vqa_b200_gemm_batched", tag or "gemm_batched", _p(a), a_layout, lda, abs_, _p(b), b_layout, ldb, bbs, _p(C),
          _dt(C), C.stride(1), C.stride(0), Bn, M, N, K, _p(bias), act, _p(add), F32, float(drop_p),
          int(seed) & 0xFFFFFFFF, _p(seed_dev), int(accumulate), _st())
    return C


def act_fwd(x, add=None, bias=None, act=0, drop_p=0.0, seed=0, seed_dev=None):
    x = x.contiguous()
    cols = x.shape[-1]
    rows = x.numel() // cols
    out = torch.empty_like(x)
    _call("vqa_b200_act_fwd", None, _p(x), _p(add.contiguous() if add is not None else None), _p(bias), _p(out), rows,
          cols, act, float(drop_p), int(seed) & 0xFFFFFFFF, _p(seed_dev), _st())
    return out


def act_bwd(dout, h, act, drop_p=0.0, seed=0, want_dbias=False, seed_dev=None, out_dtype=torch.float32):
    """dout, h: [..., J] views with stride(-1) == 1 and a common row pitch structure (2-D after flattening).
    out_dtype bf16: the result only feeds GEMMs (bf16 mode) -- written once, in operand form, instead of fp32 + a cast."""
    J = h.shape[-1]
    M = h.numel() // J
    ld_h = h.stride(-2) if h.dim() > 1 else J
    if dout.stride(-1) != 1 or (dout.dim() > 2 and not dout.is_contiguous()):
        dout = dout.contiguous()
    ld_d = dout.stride(-2) if dout.dim() > 1 else J
    out = alloc_padded((M, J), out_dtype, h.device) if (ld_h != J or J % 8) else torch.empty((M, J), device=h.device,
                                                                                             dtype=out_dtype)
    dbias = torch.zeros(J, device=h.device, dtype=torch.float32) if want_dbias else None
    _call("vqa_b200_act_bwd", None, _p(dout), _dt(dout), ld_d, _p(h), _dt(h), ld_h, _p(out), _dt(out), out.stride(0),
          _p(dbias), M, J, act, float(drop_p), int(seed) & 0xFFFFFFFF, _p(seed_dev), _st())
    return out, dbias


class LinearActFn(torch.autograd.Function):
    """dropout(act(x W^T + b)) with everything after the GEMM in its epilogue (hieCoAtten.py:25-26,30-31,35-36;
    modules.py:89,104-105).  x: [..., K] -> [..., N] fp32."""

    @staticmethod
    def forward(ctx, x, W, b, cfg: StageCfg, act=0, drop_p=0.0, seed=0, tag=None):
        _cuda(x, W)
        shp = x.shape
        x2 = x.reshape(1, -1, shp[-1])
        if x2.stride(2) != 1:
            x2 = x2.contiguous()
        if cfg.mode == "fp32":
            xin = x2
        else:
            memo = pack_scope.memo()
            key = ("lin", id(x), x._version)
            if memo is not None and key in memo:
                xin = memo[key][1]
            else:
                xin = _prep3(x2, K_MAJOR, 0, "bf16")[0]
                if memo is not None:
                    memo[key] = (x, xin)
        wop = cfg.cache.get(W, K_MAJOR, 1, cfg.mode)
        y = gemm_ex(xin, K_MAJOR, wop.t.unsqueeze(0) if cfg.mode == "bf16" else _w2d(W.detach()).unsqueeze(0), K_MAJOR,
                    cfg.mode, bias=b, act=act, drop_p=drop_p, seed=seed, tag=tag or "gemm_fwd", seed_dev=cfg.seed_dev)
        ctx.cfg, ctx.shp, ctx.act, ctx.drop = cfg, shp, act, (drop_p, seed)
        ctx.has_bias = b is not None
        ctx.save_for_backward(xin, W, y)
        return y[0].view(*shp[:-1], W.shape[0]) if y.is_contiguous() else y[0].reshape(*shp[:-1], W.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xin, W, y = ctx.saved_tensors
        cfg = ctx.cfg
        N = W.shape[0]
        dy2 = dy.reshape(-1, N)
        bf = cfg.mode == "bf16"
        if ctx.act == 0 and ctx.drop[0] == 0.0:
            # a plain Linear (fc_Wbv / fc_Wv / fc_Wq): nothing to undo, the incoming gradient IS the pre-activation one
            if dy2.stride(-1) != 1:
                dy2 = dy2.contiguous()
            db = colsum(dy2) if ctx.has_bias else None
            dpre = _prep3(dy2.unsqueeze(0), K_MAJOR, 0, "bf16")[0][0] if bf else dy2
        else:
            # bf16 mode: dpre only feeds the two GEMMs below -> written once, as their bf16 operand
            dpre, db = act_bwd(dy2, y[0], ctx.act, ctx.drop[0], ctx.drop[1], want_dbias=ctx.has_bias,
                               seed_dev=cfg.seed_dev, out_dtype=torch.bfloat16 if bf else torch.float32)
        dW = None
        if ctx.needs_input_grad[1]:
            dWb = alloc_padded((1, N, xin.shape[2]), torch.float32, dy.device)
            dWb.zero_()
            gemm_ex(dpre.unsqueeze(0), MN_MAJOR, xin, MN_MAJOR, cfg.mode, out=dWb, accumulate=True, tag="gemm_wgrad")
            dW = dWb[0].reshape(W.shape)
        dx = None
        if ctx.needs_input_grad[0]:
            # bf16 mode: the cached bf16 weight copy serves as the MN-major operand (no re-cast per backward)
            w2 = cfg.cache.get(W, MN_MAJOR, 1, "bf16").t.unsqueeze(0) if bf else _w2d(W.detach()).unsqueeze(0)
            dx = gemm_ex(dpre.unsqueeze(0), K_MAJOR, w2, MN_MAJOR, cfg.mode, tag="gemm_dgrad")[0]
            dx = dx.reshape(ctx.shp)
        return dx, dW, db, None, None, None, None, None


class BmmActFn(torch.autograd.Function):
    """C[b] = dropout(act(A_b B_b^T + add[b])) for per-sample products (hieCoAtten.py:32-33,38-39,45-46;
    modules.py:91,94).  A / B are [B, R, C] tensors consumed K-major ([rows, K]) or MN-major ([K, rows])."""

    @staticmethod
    def forward(ctx, A, a_layout, B, b_layout, add, cfg: StageCfg, act=0, drop_p=0.0, seed=0, tag=None):
        addp = None
        if add is not None:
            Cp = _pad8(add.shape[-1])
            want = (add.shape[1] * Cp, Cp, 1)
            if add.dtype == torch.float32 and tuple(add.stride()) == want:
                addp = add                                   # already in the output's (padded-pitch) layout
            else:
                addp = alloc_padded(tuple(add.shape), torch.float32, add.device)
                addp.copy_(add)
        if cfg.mode == "bf16":
            # operands packed once, here; the backward GEMMs re-use the very same bf16 tensors
            A = _prep3(A, a_layout, 0, "bf16")[0]
            B = _prep3(B, b_layout, 1, "bf16")[0]
        C = gemm_ex(A, a_layout, B, b_layout, cfg.mode, act=act, add=addp, drop_p=drop_p, seed=seed, tag=tag or "gemm_bmm",
                    seed_dev=cfg.seed_dev)
        ctx.cfg, ctx.lay, ctx.act, ctx.drop = cfg, (a_layout, b_layout), act, (drop_p, seed)
        ctx.has_add = add is not None
        ctx.save_for_backward(A, B, C)
        return C

    @staticmethod
    def backward(ctx, dC):
        A, B, C = ctx.saved_tensors
        cfg = ctx.cfg
        la, lb = ctx.lay
        Bn, M, N = C.shape
        if ctx.act != 0 or ctx.drop[0] > 0:
            dpre = _act_bwd_strided(dC, C, ctx.act, ctx.drop, cfg.seed_dev)
        else:
            dpre = dC
        dadd = dpre if ctx.has_add else None                 # fp32: it is a gradient handed on to autograd
        if cfg.mode == "bf16" and (ctx.needs_input_grad[0] and ctx.needs_input_grad[2]):
            dpre = _prep3(dpre, K_MAJOR, 0, "bf16")[0]         # one cast for both products below
        dA = dB = None
        if ctx.needs_input_grad[0]:
            # dA_b(m,k) = sum_n dpre(m,n) B_b(n,k)
            if la == K_MAJOR:
                dA = gemm_ex(dpre, K_MAJOR, B, MN_MAJOR if lb == K_MAJOR else K_MAJOR, cfg.mode, tag="gemm_bmm_bwd")
            else:   # A stored [B, K, M]: dA^T(k,m) = sum_n B_b(n,k) dpre(m,n)
                dA = gemm_ex(B, MN_MAJOR if lb == K_MAJOR else K_MAJOR, dpre, K_MAJOR, cfg.mode, tag="gemm_bmm_bwd")
        if ctx.needs_input_grad[2]:
            # dB_b(n,k) = sum_m dpre(m,n) A_b(m,k)
            if lb == K_MAJOR:
                dB = gemm_ex(dpre, MN_MAJOR, A, MN_MAJOR if la == K_MAJOR else K_MAJOR, cfg.mode, tag="gemm_bmm_bwd")
            else:   # B stored [B, K, N]: dB^T(k,n) = sum_m A_b(m,k) dpre(m,n)
                dB = gemm_ex(A, MN_MAJOR if la == K_MAJOR else K_MAJOR, dpre, MN_MAJOR, cfg.mode, tag="gemm_bmm_bwd")
        return dA, None, dB, None, dadd, None, None, None, None, None


def _act_bwd_strided(dC, C, act, drop, seed_dev=None):
    """act_bwd for a padded-pitch [B, M, N] output: rows are (b, m) with pitch C.stride(1) (batch stride == M * pitch)."""
    Bn, M, N = C.shape
    assert C.stride(0) == M * C.stride(1)
    d = dC if (dC.stride() == C.stride()) else None
    if d is None:
        d = alloc_padded((Bn, M, N), torch.float32, C.device)
        d.copy_(dC)
    out = alloc_padded((Bn, M, N), torch.float32, C.device)
    _call("vqa_b200_act_bwd", None, _p(d), F32, d.stride(1), _p(C), F32, C.stride(1), _p(out), F32, out.stride(1), None,
          Bn * M, N, act, float(drop[0]), int(drop[1]) & 0xFFFFFFFF, _p(seed_dev), _st())
    return out


class LogitsPoolFn(torch.autograd.Function):
    """att = softmax_L(H w + b);  pooled = att^T X   with H and X different tensors (hieCoAtten.py:40-43,47-50;
    modules.py:59-65 after the algebraic collapse).  Returns (pooled [N, D], att [N, L])."""

    @staticmethod
    def forward(ctx, H, W2, b2, X):
        _cuda(H, X)
        N, Lr, J = H.shape
        Hc = H.contiguous().float()
        Xc = X.contiguous().float()
        logits = attn_logits_fwd(Hc.view(N * Lr, J), W2, b2)
        pooled, att = softmax_pool_fwd(Xc, logits, 1, False)
        ctx.save_for_backward(Hc, Xc, att, W2)
        return pooled, att.view(N, Lr)

    @staticmethod
    def backward(ctx, dpooled, datt):
        Hc, Xc, att, W2 = ctx.saved_tensors
        N, Lr, J = Hc.shape
        dext = datt.contiguous().view(N, 1, Lr).float() if datt is not None else None
        dlogits, dX = softmax_pool_bwd(Xc, att, dpooled.float(), 1, False, want_dx=ctx.needs_input_grad[3],
                                       datt_extra=dext)
        dH, dW2, db2, _ = attn_logits_bwd(Hc.view(N * Lr, J), W2, dlogits, torch.float32, relu_mask=False)
        return dH.view(N, Lr, J), dW2.view(W2.shape), db2, dX


class ActFn(torch.autograd.Function):
    """dropout(act(x + add)) elementwise (F.relu / F.dropout sites of hieCoAtten.py:28, modules.py:27-31)."""

    @staticmethod
    def forward(ctx, x, add, act=0, drop_p=0.0, seed=0, seed_dev=None):
        _cuda(x)
        out = act_fwd(x.float(), add.float() if add is not None else None, None, act, drop_p, seed, seed_dev)
        ctx.act, ctx.drop, ctx.has_add, ctx.seed_dev = act, (drop_p, seed), add is not None, seed_dev
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        J = out.shape[-1]
        d, _ = act_bwd(dout.contiguous().reshape(-1, J), out.reshape(-1, J), ctx.act, ctx.drop[0], ctx.drop[1],
                       seed_dev=ctx.seed_dev)
        d = d.view(out.shape)
        return d, (d if ctx.has_add else None), None, None, None, None


class RowSoftmaxFn(torch.autograd.Function):
    """softmax over the last axis (modules.py:90)."""

    @staticmethod
    def forward(ctx, x):
        _cuda(x)
        xc = x.contiguous().float()
        y = torch.empty_like(xc)
        cols = xc.shape[-1]
        _call("vqa_b200_row_softmax_fwd", None, _p(xc), _p(y), xc.numel() // cols, cols, _st())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dyc = dy.contiguous().float()
        dx = torch.empty_like(y)
        cols = y.shape[-1]
        _call("vqa_b200_row_softmax_bwd", None, _p(y), _p(dyc), _p(dx), y.numel() // cols, cols, _st())
        return dx


class GateFn(torch.autograd.Function):
    """tanh(a) * sigmoid(b) (modules.py:105-108)."""

    @staticmethod
    def forward(ctx, a, b):
        _cuda(a, b)
        ac, bc = a.contiguous().float(), b.contiguous().float()
        o = torch.empty_like(ac)
        _call("vqa_b200_gate_fwd", None, _p(ac), _p(bc), _p(o), ac.numel(), _st())
        ctx.save_for_backward(ac, bc)
        return o

    @staticmethod
    def backward(ctx, do):
        ac, bc = ctx.saved_tensors
        doc = do.contiguous().float()
        da, db = torch.empty_like(ac), torch.empty_like(bc)
        _call("vqa_b200_gate_bwd", None, _p(ac), _p(bc), _p(doc), _p(da), _p(db), ac.numel(), _st())
        return da, db


class KLDivLogSoftmaxFn(torch.autograd.Function):
    """``nn.KLDivLoss()(F.log_softmax(logits, 1), target)`` -- the solver's training loss on MHBCoAtt's output
    (solver.py:26-29,77-92; mhb_coAtt.py:149-151; reduction 'mean' over all M * N elements) -- as one kernel per direction
    instead of ATen's ten launches (the loss step around the block).  logits, target: fp32 [M, N];
    returns the 0-dim loss.  Gradient for the logits only (the soft answers are data)."""

    @staticmethod
    def forward(ctx, logits, target):
        _cuda(logits, target)
        if logits.dim() != 2 or logits.shape != target.shape or logits.dtype != torch.float32:
            raise ValueError("KLDivLogSoftmaxFn: fp32 [M, N] logits and a target of the same shape")
        x = logits if logits.stride(1) == 1 else logits.contiguous()
        t = target if (target.dtype == torch.float32 and target.stride(1) == 1) else target.float().contiguous()
        M, N = x.shape
        buf = torch.zeros(1 + 2 * M, device=x.device, dtype=torch.float32)      # loss | lse[M] | tsum[M]
        loss, lse, tsum = buf[:1], buf[1:1 + M], buf[1 + M:]
        _call("vqa_b200_kldiv_logsoftmax_fwd", "kldiv_logsoftmax_fwd", _p(x), x.stride(0), _p(t), t.stride(0), _p(loss),
              _p(lse), _p(tsum), M, N, _st())
        ctx.save_for_backward(x, t, lse, tsum)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        x, t, lse, tsum = ctx.saved_tensors
        M, N = x.shape
        d = torch.empty((M, N), device=x.device, dtype=torch.float32)
        gc = g.reshape(1).float().contiguous()
        _call("vqa_b200_kldiv_logsoftmax_bwd", "kldiv_logsoftmax_bwd", _p(x), x.stride(0), _p(t), t.stride(0), _p(lse),
              _p(tsum), _p(gc), _p(d), d.stride(0), M, N, _st())
        return d, None


class EmbeddingFn(torch.autograd.Function):
    """``nn.Embedding`` lookup (mhb_coAtt.py:69, mfb.py:68) whose backward is ONE scatter-add into a zeroed weight-shaped
    buffer (inside a data-parallel reducer: straight into the parameter's bucket view).  ATen's dense embedding backward
    sorts the indices first (8 radix-sort launches + 3 more, 0.1 ms per step for 6656 tokens); with 15 000 rows of 300
    floats the atomics of a plain scatter-add do not contend.  Same values up to the summation order of repeated tokens."""

    @staticmethod
    def forward(ctx, idx, W):
        _cuda(W)
        ctx.save_for_backward(idx)
        ctx.W = W
        return torch.nn.functional.embedding(idx, W)

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        W = ctx.W
        dst = _grad_buffer(W, W.shape[0], W.shape[1], dy.device)
        dW = dst.zero_() if dst is not None else torch.zeros(W.shape, device=dy.device, dtype=torch.float32)
        dW.index_add_(0, idx.reshape(-1), dy.reshape(-1, W.shape[1]).float())
        return None, dW


# --------------------------------------------------------------------------------------------
# question-encoder recurrence (mhb_coAtt.py:72-74): persistent LSTM kernels + tcgen05 GEMMs around them
# --------------------------------------------------------------------------------------------
def lstm_supported(Bt: int, H: int) -> bool:
    """True when (rows per step, hidden size) is inside the persistent kernels' regime (Bt <= 32, H in 128..1024)."""
    return bool(_lib.load().vqa_b200_lstm_supported(int(Bt), int(H)))


def _padded_bf16_2d(x2: torch.Tensor) -> torch.Tensor:
    """[R, C] fp32 (row-strided) -> bf16 [R, C] view whose row pitch is a multiple of 8 elements (TMA stride rule;
    C = 300-d embeddings), pad columns zero."""
    return _prep3(x2.unsqueeze(0), K_MAJOR, 0, "bf16")[0][0]


def _sentinel_bf16(shape, device) -> torch.Tensor:
    """bf16 buffer filled with the bit pattern 0xFFFF, the 'not written yet' mark of the recurrence's exchange buffers."""
    raw = torch.full(shape, -1, device=device, dtype=torch.int16)
    return raw.view(torch.bfloat16)


class LstmFn(torch.autograd.Function):
    """Single-layer nn.LSTM(batch_first=True), zero initial state, on x [Bt, S, E] -> [Bt, S, H] (all hidden states).

    x-projection, dW_ih, dW_hh, dx: tcgen05 GEMMs; the S-step recurrence: one persistent kernel per direction
    (csrc/lstm.cu: cooperative launch forward, clusters of 4 CTAs backward).  bf16 operands, fp32 accumulation / cell
    state / gate math.  Reference call site: mhb_coAtt.py:72-74."""

    @staticmethod
    def forward(ctx, x, W_ih, W_hh, b_ih, b_hh, cache: WeightCache, drop_p=0.0, seed=0, seed_dev=None):
        _cuda(x, W_ih, W_hh)
        Bt, S, E = x.shape
        H = W_hh.shape[1]
        dev = x.device
        xs = x.permute(1, 0, 2).reshape(S * Bt, E)                 # time-major rows (t, b); a view for the reference's feed
        if xs.dtype != torch.float32:
            xs = xs.float()
        xb = _padded_bf16_2d(xs)
        wih = cache.get_fn(W_ih, "lstm_ih", _padded_bf16_2d)       # bf16 [4H, E], padded pitch
        whh = cache.get(W_hh, K_MAJOR, 1, "bf16").t                # bf16 [4H, H]
        bias = b_ih.detach() + b_hh.detach() if b_ih is not None else None
        gates = gemm(Operand(xb, K_MAJOR, S * Bt, E), K_MAJOR, Operand(wih, K_MAJOR, 4 * H, E), K_MAJOR, "bf16",
                     out_dtype=torch.float32, bias=bias, tag="lstm_xproj")
        need_grad = _need_grad(ctx)
        out = torch.empty((S, Bt, H), device=dev, dtype=torch.float32)
        hb = _sentinel_bf16((S + 1, Bt, H), dev)                    # exchange buffer: 0xFFFF = "not written yet"
        hb[0].zero_()                                               # h_{-1} = 0
        c_all = torch.empty((S, Bt, H), device=dev, dtype=torch.float32) if need_grad else None
        # drop_p > 0: the dropout that follows the LSTM in the reference (mhb_coAtt.py:74) is applied to `out` by the kernel
        _call("vqa_b200_lstm_fwd", "lstm_fwd", _p(gates), _p(whh), _p(out), _p(hb), _p(c_all), S, Bt, H, float(drop_p),
              int(seed) & 0xFFFFFFFF, _p(seed_dev), _st())
        ctx.cache, ctx.dims, ctx.has_bias = cache, (Bt, S, E, H), b_ih is not None
        ctx.drop = (float(drop_p), int(seed) & 0xFFFFFFFF, seed_dev)
        if need_grad:
            ctx.save_for_backward(xb, gates, c_all, hb, W_ih, W_hh)
        return out.permute(1, 0, 2)

    @staticmethod
    def backward(ctx, dout):
        xb, gates, c_all, hb, W_ih, W_hh = ctx.saved_tensors
        Bt, S, E, H = ctx.dims
        dev = dout.device
        # dout arrives in the caller's [Bt, S, H] order (the reference's feed): read through its strides, no transposing
        # copy; the recurrent weight is read in the parameter's own [4H, H] layout from the cached bf16 copy (which
        # optim.FusedAdam refreshes in place), no per-step transposed copy either
        d = dout if (dout.dtype == torch.float32 and dout.stride(2) == 1) else dout.float().contiguous()
        whh = ctx.cache.get(W_hh, K_MAJOR, 1, "bf16").t
        dg = _sentinel_bf16((S * Bt, 4 * H), dev)
        _call("vqa_b200_lstm_bwd", "lstm_bwd", _p(gates), _p(c_all), _p(d), d.stride(1), d.stride(0), _p(whh), 1, _p(dg),
              S, Bt, H, ctx.drop[0], ctx.drop[1], _p(ctx.drop[2]), _st())
        dgo = Operand(dg, MN_MAJOR, 4 * H, S * Bt)
        dW_hh = dW_ih = db = dx = None
        if ctx.needs_input_grad[2]:
            dW_hh = wgrad(dgo, Operand(hb[:S].view(S * Bt, H), MN_MAJOR, H, S * Bt), "bf16", tag="lstm_wgrad", dest_for=W_hh)
        if ctx.needs_input_grad[1]:
            dW_ih = wgrad(dgo, Operand(xb, MN_MAJOR, E, S * Bt), "bf16", tag="lstm_wgrad", dest_for=W_ih)
        if ctx.has_bias and (ctx.needs_input_grad[3] or ctx.needs_input_grad[4]):
            db = colsum(dg)
        if ctx.needs_input_grad[0]:
            wih = ctx.cache.get_fn(W_ih, "lstm_ih", _padded_bf16_2d)
            dx = gemm(Operand(dg, K_MAJOR, S * Bt, 4 * H), K_MAJOR, Operand(wih, MN_MAJOR, E, 4 * H), MN_MAJOR, "bf16",
                      out_dtype=torch.float32, tag="lstm_dgrad")
            dx = dx.view(S, Bt, E).permute(1, 0, 2)
        return dx, dW_ih, dW_hh, db, (db.clone() if db is not None else None), None, None, None, None


def lstm_steps_supported(Bt: int, H: int) -> bool:
    """True when the wide-batch form of the recurrence applies (ops.LstmStepFn): more rows per step than the persistent
    kernels hold (Bt > 32) and a hidden size the bf16 TMA operands accept."""
    return Bt > 32 and H % 8 == 0


class LstmStepFn(torch.autograd.Function):
    """Single-layer nn.LSTM(batch_first=True), zero initial state, x [Bt, S, E] -> [Bt, S, H], for MANY rows per step
    (mfb.py:68-70: S = T = 26 steps over Bt = N = 64..512 rows; the persistent kernels of LstmFn cover Bt <= 32).

    Per step: ONE tcgen05 GEMM h_{t-1} W_hh^T accumulated onto the x-projection (4.3 GFLOP at Bt = 512) and ONE
    elementwise cell pass (vqa_b200_lstm_cell_fwd), which writes h_t straight into the caller's [Bt, S, H] result and as
    the bf16 operand of the next step; backward mirrors it (cell pass -> dg_t, GEMM dg_t W_hh -> recurrent dh).  The
    x-projection, dW_ih, dW_hh and dx are single large GEMMs over all steps, as in LstmFn.  bf16 operands, fp32
    accumulation / cell state / gate math.  Inside a captured iteration (train.GraphedTrainStep) the 2 S launches per
    direction are graph nodes."""

    @staticmethod
    def forward(ctx, x, W_ih, W_hh, b_ih, b_hh, cache: WeightCache, drop_p=0.0, seed=0, seed_dev=None):
        _cuda(x, W_ih, W_hh)
        Bt, S, E = x.shape
        H = W_hh.shape[1]
        dev = x.device
        drop = (float(drop_p), int(seed) & 0xFFFFFFFF, seed_dev)
        xs = x.permute(1, 0, 2).reshape(S * Bt, E)                 # time-major rows (t, b)
        if xs.dtype != torch.float32:
            xs = xs.float()
        xb = _padded_bf16_2d(xs)
        wih = cache.get_fn(W_ih, "lstm_ih", _padded_bf16_2d)       # bf16 [4H, E], padded pitch
        whh = cache.get(W_hh, K_MAJOR, 1, "bf16").t                # bf16 [4H, H]
        bias = b_ih.detach() + b_hh.detach() if b_ih is not None else None
        gates = gemm(Operand(xb, K_MAJOR, S * Bt, E), K_MAJOR, Operand(wih, K_MAJOR, 4 * H, E), K_MAJOR, "bf16",
                     out_dtype=torch.float32, bias=bias, tag="lstm_xproj").view(S, Bt, 4 * H)
        need_grad = _need_grad(ctx)
        out = torch.empty((Bt, S, H), device=dev, dtype=torch.float32)
        hb = torch.empty((S + 1, Bt, H), device=dev, dtype=torch.bfloat16)
        hb[0].zero_()                                               # h_{-1} = 0
        c_all = torch.empty((S if need_grad else 2, Bt, H), device=dev, dtype=torch.float32)
        wop = Operand(whh, K_MAJOR, 4 * H, H)
        for t in range(S):
            if t > 0:                                               # h_{-1} = 0: nothing to add at t = 0
                # inference: k_split = 1, one contribution per element, so the states do not depend on a summation order
                # (the val loop is repeatable); training lets the launcher split K to fill the machine at small Bt
                gemm(Operand(hb[t], K_MAJOR, Bt, H), K_MAJOR, wop, K_MAJOR, "bf16", acc_into=gates[t],
                     k_split=0 if need_grad else 1, tag="lstm_step_fwd")
            ci = t if need_grad else t & 1
            cprev = c_all[ci - 1 if need_grad else ci ^ 1] if t > 0 else None
            _call("vqa_b200_lstm_cell_fwd", "lstm_cell_fwd", _p(gates[t]), _p(cprev), _p(c_all[ci]), _p(out[:, t]),
                  S * H, _p(hb[t + 1]), Bt, H, 1 if need_grad else 0, t * Bt, drop[0], drop[1], _p(drop[2]), _st())
        ctx.cache, ctx.dims, ctx.has_bias, ctx.drop = cache, (Bt, S, E, H), b_ih is not None, drop
        if need_grad:
            ctx.save_for_backward(xb, gates, c_all, hb, W_ih, W_hh)
        return out

    @staticmethod
    def backward(ctx, dout):
        xb, gates, c_all, hb, W_ih, W_hh = ctx.saved_tensors
        Bt, S, E, H = ctx.dims
        dev = dout.device
        d = dout if (dout.dtype == torch.float32 and dout.stride(2) == 1 and dout.stride(0) % 4 == 0
                     and dout.stride(1) % 4 == 0 and dout.data_ptr() % 16 == 0) else dout.float().contiguous()
        whh = ctx.cache.get(W_hh, K_MAJOR, 1, "bf16").t
        wop = Operand(whh, MN_MAJOR, H, 4 * H)                     # W_hh itself as the [N = H, K = 4H] operand
        dg = torch.empty((S, Bt, 4 * H), device=dev, dtype=torch.bfloat16)
        state = torch.zeros((2, Bt, H), device=dev, dtype=torch.float32)          # recurrent dh, dc
        for t in range(S - 1, -1, -1):
            _call("vqa_b200_lstm_cell_bwd", "lstm_cell_bwd", _p(gates[t]), _p(c_all[t - 1] if t > 0 else None),
                  _p(c_all[t]), _p(d[:, t]), d.stride(0), _p(state[0]), _p(state[1]), _p(dg[t]), Bt, H, t * Bt,
                  ctx.drop[0], ctx.drop[1], _p(ctx.drop[2]), _st())
            if t > 0:
                gemm(Operand(dg[t], K_MAJOR, Bt, 4 * H), K_MAJOR, wop, MN_MAJOR, "bf16", acc_into=state[0],
                     tag="lstm_step_bwd")
        dg = dg.view(S * Bt, 4 * H)
        dgo = Operand(dg, MN_MAJOR, 4 * H, S * Bt)
        dW_hh = dW_ih = db = dx = None
        if ctx.needs_input_grad[2]:
            dW_hh = wgrad(dgo, Operand(hb[:S].view(S * Bt, H), MN_MAJOR, H, S * Bt), "bf16", tag="lstm_wgrad", dest_for=W_hh)
        if ctx.needs_input_grad[1]:
            dW_ih = wgrad(dgo, Operand(xb, MN_MAJOR, E, S * Bt), "bf16", tag="lstm_wgrad", dest_for=W_ih)
        if ctx.has_bias and (ctx.needs_input_grad[3] or ctx.needs_input_grad[4]):
            db = colsum(dg)
        if ctx.needs_input_grad[0]:
            wih = ctx.cache.get_fn(W_ih, "lstm_ih", _padded_bf16_2d)
            dx = gemm(Operand(dg, K_MAJOR, S * Bt, 4 * H), K_MAJOR, Operand(wih, MN_MAJOR, E, 4 * H), MN_MAJOR, "bf16",
                      out_dtype=torch.float32, tag="lstm_dgrad")
            dx = dx.view(S, Bt, E).permute(1, 0, 2)
        return dx, dW_ih, dW_hh, db, (db.clone() if db is not None else None), None, None, None, None


def run_lstm(lstm, x, cache: WeightCache, precision: str, drop_p: float = 0.0, seed: int = 0, seed_dev=None):
    """`lstm(x)[0]` for a batch_first nn.LSTM on x [Bt, S, E] (mhb_coAtt.py:72-74, mfb.py:68-70) with the module's own
    parameters.  bf16 mode on CUDA, one layer, one direction: the persistent recurrence kernels (Bt <= 32: MHBCoAtt's
    [T, N, E] feed) or the per-step GEMM + cell form (more rows: MFB).  fp32 mode, other shapes and VQA_B200_LSTM=stock
    keep the stock module (north_star: left as-is).

    Returns (output, dropped): with drop_p > 0 the native forms apply the dropout that follows the LSTM in the reference
    (`self.dropout_l`, time-major mask rows t * Bt + b) inside their kernels and report dropped = True; the stock path
    returns the clean output and dropped = False -- the caller applies its nn.Dropout then."""
    import os
    fast = (x.is_cuda and precision == "bf16" and lstm.num_layers == 1 and not lstm.bidirectional and lstm.batch_first
            and getattr(lstm, "proj_size", 0) == 0 and os.environ.get("VQA_B200_LSTM", "fast") != "stock")
    if fast:
        Bt, H = x.shape[0], lstm.hidden_size
        args = (x, lstm.weight_ih_l0, lstm.weight_hh_l0, getattr(lstm, "bias_ih_l0", None),
                getattr(lstm, "bias_hh_l0", None), cache, float(drop_p), seed, seed_dev if drop_p > 0.0 else None)
        if lstm_supported(Bt, H):
            return LstmFn.apply(*args), drop_p > 0.0
        if lstm_steps_supported(Bt, H):
            return LstmStepFn.apply(*args), drop_p > 0.0
    return lstm(x)[0], False
