"""Data feed for the drop-in modules: packed bf16 feature shards, a pinned host ring, and an asynchronous H2D pipeline
(replaces the per-item path of the reference's ``data_loader.py:27-57``).

The reference stores one fp32 ``[2048, 14, 14]`` ``.npy`` per image (extract_image_features.py:78-84) and, per ITEM,
loads it, transposes it to ``[196, 2048]`` and converts it to a float tensor (data_loader.py:29-32); soft answers are
expanded to a dense ``[num_answer]`` row on the host (data_loader.py:39-43).  At the > 4e4 samples/s per GPU the kernels
sustain that is > 65 GB/s of fp32 features per GPU -- more than a PCIe 5 x16 link moves (~55 GB/s measured) -- before any
host work.  This module keeps the SAME records in the layout the kernels consume:

  shard file  = header | features bf16 ``[N, L, D]`` (region-major rows: the transpose is done once, at shard-writing
                time) | questions int32 ``[N, T]`` | question lengths int32 ``[N]`` | answers: hard int32 ``[N]`` or soft
                sparse (index int32, weight fp32) ``[N, 10]`` (utils.py:250-265 yields <= 10 non-zeros per row)
  ShardReader = read-only memory map of that file, batches are contiguous row ranges (no per-item work)
  ShardFeed   = staging thread (shard -> pinned ring slot, first-touched on the GPU's NUMA node) + copy stream
                (pinned slot -> device slot, two batches ahead of the consumer) + CUDA events both ways; the dense soft
                answer rows and the int64 token ids the modules expect are rebuilt ON THE DEVICE from the packed forms.

bf16 features are what ``precision = "bf16"`` computes with anyway (the modules' first kernel would round the fp32 input to
exactly these values), so results are bit-identical to the fp32 feed in that mode; half the bytes cross PCIe.
CUDA only on the consuming side; there is no CPU fallback for the device half.
"""
from __future__ import annotations

import os
import struct
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

MAGIC = b"VQAB200S"
VERSION = 1
HEADER_BYTES = 4096
SOFT_NNZ = 10                      # utils.py:250-265: at most 10 distinct answers per question
_HDR = struct.Struct("<8sIQIIIIIQQQQQ")     # magic, version, N, L, D, T, A, target_kind, 5 section offsets
TARGET_NONE, TARGET_HARD, TARGET_SOFT = 0, 1, 2


def _align(n: int, a: int = 4096) -> int:
    return (n + a - 1) // a * a


def to_bf16_bits(x: torch.Tensor) -> np.ndarray:
    """fp32 / bf16 tensor -> uint16 numpy array holding the bf16 bit patterns (round-to-nearest-even, as the device pack)."""
    return x.detach().to(torch.bfloat16).contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


class ShardWriter:
    """Streaming writer.  ``append`` takes what ``VqaDataset.__getitem__`` yields, batched: features ``[n, L, D]`` (fp32 or
    bf16; a ``[n, D, 14, 14]`` extractor output is transposed here, data_loader.py:30-31), questions ``[n, T]``,
    question lengths ``[n]`` (optional) and answers (hard ``[n]`` ids or dense soft rows ``[n, A]``)."""

    def __init__(self, path: str, L: int, D: int, T: int, num_answer: int, soft_answer: bool):
        self.path, self.L, self.D, self.T, self.A = path, L, D, T, num_answer
        self.kind = TARGET_SOFT if soft_answer else TARGET_HARD
        self.f = open(path, "wb")
        self.f.write(b"\0" * HEADER_BYTES)
        self.n = 0
        self._q, self._ql, self._a_idx, self._a_w = [], [], [], []

    def append(self, features: torch.Tensor, questions: torch.Tensor, answers: torch.Tensor, ques_length=None):
        if features.dim() == 4:                                   # [n, D, h, w] -> [n, h*w, D]
            features = features.permute(0, 2, 3, 1).reshape(features.shape[0], -1, features.shape[1])
        n = features.shape[0]
        assert tuple(features.shape[1:]) == (self.L, self.D) and tuple(questions.shape) == (n, self.T)
        self.f.write(to_bf16_bits(features).tobytes())
        self._q.append(questions.to(torch.int32).cpu().numpy())
        ql = ques_length if ques_length is not None else torch.full((n,), self.T)
        self._ql.append(torch.as_tensor(ql).to(torch.int32).cpu().numpy())
        if self.kind == TARGET_HARD:
            self._a_idx.append(answers.reshape(n).to(torch.int32).cpu().numpy())
        else:
            a = answers.float().cpu()
            assert tuple(a.shape) == (n, self.A)
            if int((a != 0).sum(1).max()) > SOFT_NNZ:
                raise ValueError("soft answer rows with more than %d non-zeros are not representable" % SOFT_NNZ)
            w, idx = a.topk(SOFT_NNZ, dim=1)                     # zeros pad the tail (weight 0 adds nothing)
            self._a_idx.append(idx.to(torch.int32).numpy())
            self._a_w.append(w.numpy().astype(np.float32))
        self.n += n

    def close(self):
        offs = [HEADER_BYTES]
        pos = HEADER_BYTES + self.n * self.L * self.D * 2

        def section(arrs, dtype, shape_tail):
            nonlocal pos
            pos = _align(pos)
            self.f.seek(pos)
            off = pos
            data = (np.concatenate(arrs) if arrs else np.zeros((0,) + shape_tail, dtype)).astype(dtype, copy=False)
            self.f.write(data.tobytes())
            pos += data.nbytes
            return off

        offs.append(section(self._q, np.int32, (self.T,)))
        offs.append(section(self._ql, np.int32, ()))
        offs.append(section(self._a_idx, np.int32, (SOFT_NNZ,) if self.kind == TARGET_SOFT else ()))
        offs.append(section(self._a_w, np.float32, (SOFT_NNZ,)) if self.kind == TARGET_SOFT else 0)
        self.f.seek(0)
        self.f.write(_HDR.pack(MAGIC, VERSION, self.n, self.L, self.D, self.T, self.A, self.kind, *offs))
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class ShardReader:
    """Read-only memory map of one shard.  ``rows(i0, n)`` returns numpy VIEWS (no copy) of n consecutive records."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as f:
            hdr = f.read(_HDR.size)
        magic, ver, N, L, D, T, A, kind, o_feat, o_q, o_ql, o_ai, o_aw = _HDR.unpack(hdr)
        if magic != MAGIC or ver != VERSION:
            raise ValueError("%s is not a vqa_b200 feature shard (magic %r, version %d)" % (path, magic, ver))
        self.N, self.L, self.D, self.T, self.A, self.kind = N, L, D, T, A, kind
        mm = np.memmap(path, dtype=np.uint8, mode="r")
        self._mm = mm
        self.features = mm[o_feat:o_feat + N * L * D * 2].view(np.uint16).reshape(N, L, D)
        self.questions = mm[o_q:o_q + N * T * 4].view(np.int32).reshape(N, T)
        self.ques_length = mm[o_ql:o_ql + N * 4].view(np.int32)
        if kind == TARGET_SOFT:
            self.ans_idx = mm[o_ai:o_ai + N * SOFT_NNZ * 4].view(np.int32).reshape(N, SOFT_NNZ)
            self.ans_w = mm[o_aw:o_aw + N * SOFT_NNZ * 4].view(np.float32).reshape(N, SOFT_NNZ)
        else:
            self.ans_idx = mm[o_ai:o_ai + N * 4].view(np.int32)
            self.ans_w = None

    def __len__(self):
        return self.N

    def rows(self, i0: int, n: int):
        s = slice(i0, i0 + n)
        return (self.features[s], self.questions[s], self.ques_length[s], self.ans_idx[s],
                self.ans_w[s] if self.ans_w is not None else None)

    def bytes_per_record(self) -> int:
        return self.L * self.D * 2 + self.T * 4 + 4 + (SOFT_NNZ * 8 if self.kind == TARGET_SOFT else 4)


# --------------------------------------------------------------------------------------------------------------
# NUMA placement of the pinned ring
# --------------------------------------------------------------------------------------------------------------
def _cpu_list(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if part:
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def parse_nvidia_smi_topo(text: str, gpu: int) -> Tuple[Optional[int], Optional[set]]:
    """(numa node, cpu set) of GPU<gpu> from the table `nvidia-smi topo -m` prints: the header names the columns ("CPU
    Affinity", "NUMA Affinity"), the GPU's row holds one cell more (its own name first).  (None, None) if absent."""
    import re
    text = re.sub(r"\x1b\[[0-9;]*m", "", text)                        # the header is underlined with ANSI codes
    split = lambda line: [c.strip() for c in re.split(r"\t+", line.strip("\n")) if c.strip()]
    header = row = None
    for line in text.splitlines():
        cells = split(line)
        if not cells:
            continue
        if header is None and "CPU Affinity" in cells:
            header = cells
        elif cells[0] == "GPU%d" % gpu:
            row = cells
    if header is None or row is None:
        return None, None
    try:
        cpus = _cpu_list(row[header.index("CPU Affinity") + 1])
        node = None
        if "NUMA Affinity" in header:
            cell = row[header.index("NUMA Affinity") + 1]
            node = int(cell.split(",")[0].split("-")[0]) if cell[:1].isdigit() else None
        return node, (cpus or None)
    except (IndexError, ValueError):
        return None, None


def gpu_cpu_affinity(dev_index: int) -> Tuple[Optional[int], Optional[set]]:
    """(numa node, cpu set) the GPU is attached to: sysfs first (bare metal), then NVML's affinity mask (VMs often report
    numa_node = -1 in sysfs but a usable NVML mask), then the table of `nvidia-smi topo -m`; a host with a single NUMA
    node answers (0, all CPUs).  (None, None) when nothing is known."""
    try:
        pr = torch.cuda.get_device_properties(dev_index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node >= 0:
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            return node, cpus
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id,
                                                                      pr.pci_device_id)).encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        node = None
        try:
            node = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        if cpus and len(cpus) < (os.cpu_count() or 0):
            return node, cpus
    except Exception:
        pass
    try:
        import subprocess
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = int(vis.split(",")[dev_index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else dev_index
        node, cpus = parse_nvidia_smi_topo(out, phys)
        if cpus and (node is not None or len(cpus) < (os.cpu_count() or 0)):
            return node, cpus
    except Exception:
        pass
    try:
        if open("/sys/devices/system/node/online").read().strip() == "0":      # one node: nothing to choose
            return 0, set(range(os.cpu_count() or 1))
    except Exception:
        pass
    return None, None


def bind_to_gpu_numa_node(dev_index: int) -> Optional[int]:
    """Run this process -- and therefore first-touch its pinned buffers -- on the CPUs next to the GPU.  Pinned memory on
    the remote socket feeds the GPU at less than half the PCIe rate.  Returns the node (or -1 when only a CPU mask is
    known), None when the host exposes no topology."""
    node, cpus = gpu_cpu_affinity(dev_index)
    if not cpus:
        return None
    allowed = cpus & os.sched_getaffinity(0)
    if not allowed:
        return None
    os.sched_setaffinity(0, allowed)
    return node if node is not None else -1


# --------------------------------------------------------------------------------------------------------------
# pinned ring + H2D pipeline
# --------------------------------------------------------------------------------------------------------------
class _PinnedSlot:
    def __init__(self, B, L, D, T, soft):
        self.feat = torch.empty((B, L, D), dtype=torch.int16).pin_memory()          # bf16 bit patterns
        self.q = torch.empty((B, T), dtype=torch.int32).pin_memory()
        self.ql = torch.empty((B,), dtype=torch.int32).pin_memory()
        self.ai = torch.empty((B, SOFT_NNZ) if soft else (B,), dtype=torch.int32).pin_memory()
        self.aw = torch.empty((B, SOFT_NNZ), dtype=torch.float32).pin_memory() if soft else None
        self.feat.zero_()                                                            # first touch (NUMA placement)
        self.gen = -1                      # running number of the batch the slot holds (-1: nothing staged yet)
        self.issued_gen = -1               # running number of the last batch whose H2D copy was enqueued from this slot
        self.h2d_done: Optional[torch.cuda.Event] = None      # recorded behind that copy

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.feat, self.q, self.ql, self.ai, self.aw) if t is not None)


class ShardFeed:
    """Batches of one shard on the device, ``depth`` batches ahead of the consumer.

        feed = ShardFeed(reader, batch, device)                  # or device_slots=[(img, questions, target), ...]
        for _ in range(steps):
            slot, (img, questions, target, ques_length) = feed.next()      # current stream now waits for the copy
            loss = step(img, questions, target)
            feed.done(slot)                                      # the slot may be refilled once this step has run

    img is bf16 ``[B, L, D]``; questions int64 ``[B, T]``; target = int64 ``[B]`` (hard) or fp32 dense ``[B, A]`` (soft,
    data_loader.py:39-43, rebuilt on the device); ques_length int64 ``[B]``.  Batches are consecutive rows, the last
    partial batch of an epoch is dropped, epochs repeat.  ``device_slots`` lets a CUDA-graphed step own the destination
    tensors (its static inputs).  When the pinned ring is at least one epoch long every batch is staged exactly once
    (the ring is then a pinned cache of the shard); otherwise the staging thread keeps refilling it."""

    def __init__(self, reader: ShardReader, batch: int, device, device_slots: Optional[Sequence[Sequence[torch.Tensor]]] = None,
                 depth: int = 2, ring_slots: int = 4, bind_numa: bool = True):
        self.r, self.B = reader, int(batch)
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("ShardFeed feeds CUDA devices only (there is no CPU path)")
        self.per_epoch = len(reader) // self.B
        if self.per_epoch < 1:
            raise ValueError("shard holds %d records, fewer than one batch of %d" % (len(reader), self.B))
        self.numa_node = bind_to_gpu_numa_node(self.dev.index or 0) if bind_numa else None
        soft = reader.kind == TARGET_SOFT
        self.soft = soft
        self.cached = ring_slots >= self.per_epoch
        self.ring = [_PinnedSlot(self.B, reader.L, reader.D, reader.T, soft)
                     for _ in range(self.per_epoch if self.cached else max(2, ring_slots))]
        self.depth = max(1, depth)
        n_dev = self.depth + 1
        if device_slots is not None:
            n_dev = len(device_slots)
            if n_dev < self.depth + 1:
                raise ValueError("need at least depth + 1 = %d device slots" % (self.depth + 1))
            self.dslots = [tuple(s) for s in device_slots]
        else:
            tgt_shape, tgt_dtype = ((self.B, reader.A), torch.float32) if soft else ((self.B,), torch.int64)
            self.dslots = [(torch.empty((self.B, reader.L, reader.D), dtype=torch.bfloat16, device=self.dev),
                            torch.empty((self.B, reader.T), dtype=torch.int64, device=self.dev),
                            torch.empty(tgt_shape, dtype=tgt_dtype, device=self.dev)) for _ in range(n_dev)]
        self.qlen = [torch.empty((self.B,), dtype=torch.int64, device=self.dev) for _ in range(n_dev)]
        # packed forms land in these staging tensors and are widened on the copy stream
        self._q32 = [torch.empty((self.B, reader.T), dtype=torch.int32, device=self.dev) for _ in range(n_dev)]
        self._ql32 = [torch.empty((self.B,), dtype=torch.int32, device=self.dev) for _ in range(n_dev)]
        self._ai = [torch.empty((self.B, SOFT_NNZ) if soft else (self.B,), dtype=torch.int32, device=self.dev)
                    for _ in range(n_dev)]
        self._aw = [torch.empty((self.B, SOFT_NNZ), dtype=torch.float32, device=self.dev) if soft else None
                    for _ in range(n_dev)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.ready = [torch.cuda.Event() for _ in range(n_dev)]
        self.consumed: List[Optional[torch.cuda.Event]] = [None] * n_dev
        self._issued = 0                   # batches whose H2D has been enqueued
        self._taken = 0                    # batches handed to the consumer
        self._stop = False
        self._cv = threading.Condition()   # guards the slots' gen / issued_gen hand-over between the two threads
        self.staged_bytes = 0
        self.staging_seconds = 0.0
        self._thread = threading.Thread(target=self._stage_loop, name="vqa_b200_feed", daemon=True)
        self._thread.start()

    # ---- staging thread: shard (page cache) -> pinned slot
    def _stage_loop(self):
        import time
        k = 0
        R = len(self.ring)
        while not self._stop:
            if self.cached and k >= self.per_epoch:
                return
            slot = self.ring[k % R]
            if k >= R:
                # streaming ring: the slot still holds batch k - R.  Wait until the consumer has enqueued that batch's
                # H2D copy, then until the GPU has finished it, before overwriting the pinned memory
                with self._cv:
                    while slot.issued_gen != k - R and not self._stop:
                        self._cv.wait(0.05)
                if self._stop:
                    return
                slot.h2d_done.synchronize()
            b = k % self.per_epoch
            feat, q, ql, ai, aw = self.r.rows(b * self.B, self.B)
            t0 = time.perf_counter()
            np.copyto(slot.feat.numpy().view(np.uint16), feat)
            np.copyto(slot.q.numpy(), q)
            np.copyto(slot.ql.numpy(), ql)
            np.copyto(slot.ai.numpy(), ai)
            if aw is not None:
                np.copyto(slot.aw.numpy(), aw)
            self.staging_seconds += time.perf_counter() - t0
            self.staged_bytes += slot.nbytes()
            with self._cv:
                slot.gen = k
                self._cv.notify_all()
            k += 1

    def _issue(self):
        """Enqueue the H2D copy of the next batch on the copy stream (host side: waits for the staging thread only)."""
        i = self._issued
        d = i % len(self.dslots)
        slot = self.ring[i % len(self.ring)]
        with self._cv:                     # cached ring: staged once, valid for every epoch; streaming ring: exactly batch i
            while (slot.gen < 0) if self.cached else (slot.gen != i):
                self._cv.wait(0.05)
        img, q, tgt = self.dslots[d]
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[d] is not None:
                self.copy_stream.wait_event(self.consumed[d])       # never overwrite inputs a queued step still reads
            if img.dtype == torch.bfloat16:
                img.view(torch.int16).copy_(slot.feat, non_blocking=True)
            else:                                                    # fp32 static input: widen on the device
                stage = getattr(self, "_feat16", None)
                if stage is None:
                    stage = self._feat16 = [torch.empty((self.B, self.r.L, self.r.D), dtype=torch.bfloat16, device=self.dev)
                                            for _ in self.dslots]
                stage[d].view(torch.int16).copy_(slot.feat, non_blocking=True)
                img.copy_(stage[d])
            self._q32[d].copy_(slot.q, non_blocking=True)
            self._ql32[d].copy_(slot.ql, non_blocking=True)
            self._ai[d].copy_(slot.ai, non_blocking=True)
            q.copy_(self._q32[d])                                    # int32 -> int64 token ids (nn.Embedding's index type)
            self.qlen[d].copy_(self._ql32[d])
            if self.soft:
                self._aw[d].copy_(slot.aw, non_blocking=True)
                tgt.zero_()
                tgt.scatter_add_(1, self._ai[d].long(), self._aw[d])   # dense soft-answer rows (data_loader.py:39-43)
            else:
                tgt.copy_(self._ai[d])
            self.ready[d].record(self.copy_stream)
            if not self.cached:
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                with self._cv:             # the staging thread may refill the pinned slot once this event has passed
                    slot.h2d_done, slot.issued_gen = ev, i
                    self._cv.notify_all()
        self._issued += 1

    def next(self):
        while self._issued < self._taken + self.depth:
            self._issue()
        d = self._taken % len(self.dslots)
        torch.cuda.current_stream(self.dev).wait_event(self.ready[d])
        self._taken += 1
        if self._issued < self._taken + self.depth:                 # keep the copy engine `depth` batches ahead
            self._issue()
        return d, (self.dslots[d][0], self.dslots[d][1], self.dslots[d][2], self.qlen[d])

    def done(self, d: int):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.consumed[d] = ev

    def h2d_bytes_per_batch(self) -> int:
        return self.ring[0].nbytes()

    def staging_gbs(self) -> float:
        return self.staged_bytes / self.staging_seconds / 1e9 if self.staging_seconds > 0 else 0.0

    def close(self):
        self._stop = True
        with self._cv:
            self._cv.notify_all()
        self._thread.join(timeout=2.0)
