"""Packed feature shards and the H2D feed (vqa_attention_networks_b200/feed.py; SURVEY.md 8f rank 4, the bytes-level
counterpart of the reference's data_loader.py:27-57).  Writer / reader round trips run on the CPU; the device half of
the pipeline is a GPU test."""
import os

import numpy as np
import pytest
import torch

from vqa_attention_networks_b200 import feed


def _records(n, L=6, D=16, T=5, A=30, seed=0):
    g = torch.Generator().manual_seed(seed)
    img = torch.relu(torch.randn(n, L, D, generator=g))
    q = torch.randint(0, 100, (n, T), generator=g)
    ql = torch.randint(1, T + 1, (n,), generator=g)
    soft = torch.zeros(n, A)
    for i in range(n):
        k = int(torch.randint(1, 11, (1,), generator=g))
        idx = torch.randperm(A, generator=g)[:k]
        w = torch.rand(k, generator=g) + 0.1
        soft[i, idx] = w / w.sum()
    hard = torch.randint(0, A, (n,), generator=g)
    return img, q, ql, soft, hard


@pytest.mark.parametrize("soft", [True, False])
def test_shard_round_trip(tmp_path, soft):
    img, q, ql, sa, ha = _records(23)
    path = str(tmp_path / "s.bin")
    with feed.ShardWriter(path, 6, 16, 5, 30, soft_answer=soft) as w:
        w.append(img[:10], q[:10], (sa if soft else ha)[:10], ql[:10])
        w.append(img[10:], q[10:], (sa if soft else ha)[10:], ql[10:])
    r = feed.ShardReader(path)
    assert (len(r), r.L, r.D, r.T, r.A) == (23, 6, 16, 5, 30)
    f, qq, qll, ai, aw = r.rows(0, 23)
    # features are the bf16 rounding (nearest even) of the fp32 input: exactly what the device pack would produce
    want = img.to(torch.bfloat16)
    got = torch.from_numpy(np.ascontiguousarray(f).view(np.int16)).view(torch.bfloat16)
    assert torch.equal(got, want)
    assert np.array_equal(qq, q.numpy()) and np.array_equal(qll, ql.numpy())
    if soft:
        dense = torch.zeros(23, 30).scatter_add_(1, torch.from_numpy(np.ascontiguousarray(ai)).long(),
                                                torch.from_numpy(np.ascontiguousarray(aw)))
        assert torch.allclose(dense, sa, atol=0, rtol=0)
    else:
        assert np.array_equal(ai, ha.numpy()) and aw is None
    # a row range is a view of the map (no copy), and sections are page aligned
    f2 = r.rows(5, 4)[0]
    assert f2.base is not None and f2.shape == (4, 6, 16)
    assert r.bytes_per_record() == 6 * 16 * 2 + 5 * 4 + 4 + (80 if soft else 4)


def test_extractor_layout_is_transposed_once(tmp_path):
    """extract_image_features.py:78-84 stores [2048, 14, 14] per image and data_loader.py:30-31 transposes per item;
    the writer accepts the extractor layout and stores region-major rows."""
    g = torch.Generator().manual_seed(1)
    chw = torch.randn(3, 16, 2, 3, generator=g)                  # [n, D, h, w]
    path = str(tmp_path / "t.bin")
    with feed.ShardWriter(path, 6, 16, 5, 30, soft_answer=False) as w:
        w.append(chw, torch.zeros(3, 5, dtype=torch.long), torch.zeros(3, dtype=torch.long))
    r = feed.ShardReader(path)
    ref = np.transpose(chw.numpy(), (0, 2, 3, 1)).reshape(3, 6, 16)          # the reference's per-item transform
    got = torch.from_numpy(np.ascontiguousarray(r.features).view(np.int16)).view(torch.bfloat16).float().numpy()
    assert np.allclose(got, ref, rtol=2 ** -8, atol=0)


def test_rejects_foreign_files_and_dense_soft_rows(tmp_path):
    p = str(tmp_path / "x.bin")
    open(p, "wb").write(b"\0" * 8192)
    with pytest.raises(ValueError, match="not a vqa_b200 feature shard"):
        feed.ShardReader(p)
    w = feed.ShardWriter(str(tmp_path / "y.bin"), 6, 16, 5, 30, soft_answer=True)
    with pytest.raises(ValueError, match="non-zeros"):
        w.append(torch.zeros(1, 6, 16), torch.zeros(1, 5, dtype=torch.long), torch.full((1, 30), 1 / 30.0))
    w.close()
    with pytest.raises(RuntimeError, match="CUDA"):
        feed.ShardFeed(_tiny_reader(tmp_path), 2, "cpu")


def _tiny_reader(tmp_path):
    img, q, ql, sa, ha = _records(8)
    path = str(tmp_path / "tiny.bin")
    with feed.ShardWriter(path, 6, 16, 5, 30, soft_answer=True) as w:
        w.append(img, q, sa, ql)
    return feed.ShardReader(path)


@pytest.mark.gpu
@pytest.mark.parametrize("ring_slots,own_slots", [(8, False), (2, False), (8, True)])
def test_feed_delivers_the_shard_in_order(tmp_path, ring_slots, own_slots):
    """Two and a half epochs through the pipeline: every batch arrives intact and in order, as bf16 features / int64 ids /
    dense soft rows, both when the pinned ring caches the epoch and when it has to be refilled, and with caller-owned
    (static) device slots -- an fp32 one included."""
    dev = "cuda:0"
    img, q, ql, sa, ha = _records(26, seed=3)
    path = str(tmp_path / "f.bin")
    with feed.ShardWriter(path, 6, 16, 5, 30, soft_answer=True) as w:
        w.append(img, q, sa, ql)
    r = feed.ShardReader(path)
    B = 4
    slots = None
    if own_slots:
        slots = [(torch.empty(B, 6, 16, dtype=(torch.float32 if i == 0 else torch.bfloat16), device=dev),
                  torch.empty(B, 5, dtype=torch.int64, device=dev), torch.empty(B, 30, device=dev)) for i in range(3)]
    f = feed.ShardFeed(r, B, dev, device_slots=slots, ring_slots=ring_slots, bind_numa=False)
    assert f.per_epoch == 6 and f.cached == (ring_slots >= 6)
    want_img = img.to(torch.bfloat16).float()
    for step in range(15):
        b = step % 6
        d, (x, qq, tgt, qlen) = f.next()
        torch.cuda.current_stream().synchronize()
        rows = slice(b * B, (b + 1) * B)
        assert torch.equal(x.float().cpu(), want_img[rows]), step
        assert torch.equal(qq.cpu(), q[rows]) and qq.dtype == torch.int64
        assert torch.equal(qlen.cpu(), ql[rows])
        assert torch.equal(tgt.cpu(), sa[rows])
        f.done(d)
    # stress: no host synchronisation between steps, so the consumer runs far ahead of the GPU and of the staging thread
    # (a pinned slot must never be refilled before its H2D copy is done, nor handed out before its batch is staged)
    sums = []
    for step in range(15, 15 + 60):
        d, (x, qq, tgt, qlen) = f.next()
        sums.append(torch.stack([x.float().sum(), qq.sum().float(), tgt[:, :7].sum()]))
        f.done(d)
    got = torch.stack(sums).cpu()
    for k, step in enumerate(range(15, 15 + 60)):
        rows = slice((step % 6) * B, (step % 6 + 1) * B)
        want = torch.stack([want_img[rows].sum(), q[rows].sum().float(), sa[rows][:, :7].sum()])
        assert torch.allclose(got[k], want, rtol=1e-5, atol=1e-4), (step, got[k], want)
    f.close()
    assert f.h2d_bytes_per_batch() == B * r.bytes_per_record()


def test_nvidia_smi_topology_table_is_parsed():
    """feed.parse_nvidia_smi_topo: the fallback for hosts whose sysfs / NVML report no GPU locality (the
    driver's box answered numa_node = null).  Tables as `nvidia-smi topo -m` prints them (tab-separated, ANSI-underlined
    header; the one-GPU table is the one a B200 box of this pool printed)."""
    from vqa_attention_networks_b200 import feed
    one = "\t\x1b[4mGPU0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\x1b[0m\nGPU0\t X \t0-15\t0\t\tN/A\n\nLegend:\n\n  X    = Self\n"
    assert feed.parse_nvidia_smi_topo(one, 0) == (0, set(range(16)))
    assert feed.parse_nvidia_smi_topo(one, 1) == (None, None)
    two = ("\t\x1b[4mGPU0\tGPU1\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\x1b[0m\n"
           "GPU0\t X \tNV18\tPXB\t0-55,112-167\t0\t\tN/A\n"
           "GPU1\tNV18\t X \tSYS\t56-111,168-223\t1\t\tN/A\n"
           "NIC0\tPXB\tSYS\t X \t\t\t\t\n")
    node, cpus = feed.parse_nvidia_smi_topo(two, 1)
    assert node == 1 and min(cpus) == 56 and max(cpus) == 223 and len(cpus) == 112 and 112 not in cpus
    assert feed.parse_nvidia_smi_topo(two, 0)[0] == 0
    assert feed.parse_nvidia_smi_topo("no table here", 0) == (None, None)
