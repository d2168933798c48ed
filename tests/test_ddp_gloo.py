"""Host-side logic of the data-parallel gradient exchange, world_size 2 on CPU (gloo): bucketed hooks must
produce exactly the average of the per-rank gradients, handle a parameter that never gets a gradient (hieCoAtten's
dead fc_Wbq) and exactly-zero gradients (MFB's dead first stage), and leave the optimizer reading bucket views."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(16, 32)
        self.dead = torch.nn.Linear(16, 16)      # registered, never used -> grad None
        self.b = torch.nn.Linear(32, 24)
        self.zero = torch.nn.Linear(24, 24)      # used, but multiplied by 0 -> exactly-zero grads
        self.c = torch.nn.Linear(24, 5)

    def forward(self, x):
        h = torch.relu(self.b(torch.relu(self.a(x))))
        return self.c(h + 0.0 * self.zero(h))


class _BucketSGD:
    """Stand-in for optim.FusedAdam's ``step(only=params)`` protocol (the real one is CUDA-only)."""

    def __init__(self, params):
        self.params, self.calls, self.seen = list(params), 0, []

    @torch.no_grad()
    def step(self, closure=None, only=None):
        self.calls += 1
        for p in (self.params if only is None else only):
            if p.grad is not None:
                p -= 0.1 * p.grad
                self.seen.append(p)


def _worker(rank, world, port, q, defer):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vqa_attention_networks_b200.ddp import GradientAllReducer
    torch.manual_seed(0)
    net = Net()
    dp = [net.c.weight, net.c.bias, net.zero.weight, net.b.weight] if defer else None
    red = GradientAllReducer(net, bucket_mb=0.002, defer_params=dp)      # tiny buckets -> several of them
    assert len(red.buckets) >= 3
    ok = True
    for step in range(2):
        g = torch.Generator().manual_seed(100 + rank + 10 * step)
        x = torch.randn(8, 16, generator=g)
        y = torch.randn(8, 5, generator=g)
        # expected: average over ranks of the local gradients
        ref = Net()
        ref.load_state_dict(net.state_dict())
        ((ref(x) - y) ** 2).mean().backward()
        expect = {}
        for n, p in ref.named_parameters():
            gl = p.grad if p.grad is not None else torch.zeros_like(p)
            parts = [torch.zeros_like(gl) for _ in range(world)]
            dist.all_gather(parts, gl)
            expect[n] = sum(parts) / world
        red.prepare()
        ((net(x) - y) ** 2).mean().backward()
        if step == 0:
            red.finish()
        else:
            # per-bucket update protocol: every bucket is stepped exactly once, right behind its all-reduce, and
            # the reduced gradients it saw are the averages (checked below on .grad, which the step leaves alone)
            opt = _BucketSGD(net.parameters())
            before = {n: p.detach().clone() for n, p in net.named_parameters()}
            red.finish(opt)
            ok &= opt.calls == len(red.buckets)
            ok &= sorted(id(p) for p in opt.seen) == sorted(id(p) for p in net.parameters())
            for n, p in net.named_parameters():
                ok &= torch.allclose(p.detach(), before[n] - 0.1 * expect[n], atol=1e-7)
        for n, p in net.named_parameters():
            ok &= p.grad is not None and torch.allclose(p.grad, expect[n], atol=1e-7)
            bi, pi = red._index[p]
            ok &= p.grad.data_ptr() == red.buckets[bi].views[pi].data_ptr()
        ok &= float(net.dead.weight.grad.abs().max()) == 0.0 and float(net.zero.weight.grad.abs().max()) == 0.0
        if step == 0:
            with torch.no_grad():
                for p in net.parameters():
                    p -= 0.1 * p.grad
    # replicas stay bit-identical
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    parts = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    ok &= torch.equal(parts[0], parts[1])
    q.put((rank, bool(ok), red.bytes_per_step()))
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("defer", [False, True])
def test_bucketed_allreduce_world2_gloo(defer):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, defer)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == sum(p.numel() for p in Net().parameters()) * 4


def _dry_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vqa_attention_networks_b200.ddp import GradientAllReducer
    torch.manual_seed(0)
    net = Net()
    red = GradientAllReducer(net, bucket_mb=0.002)
    g = torch.Generator().manual_seed(100 + rank)
    x, y = torch.randn(8, 16, generator=g), torch.randn(8, 5, generator=g)
    ref = Net()
    ref.load_state_dict(net.state_dict())
    ((ref(x) - y) ** 2).mean().backward()
    red.dry_run = True                    # bench.py's exposed-time leg: same bucket traffic, no collective
    red.prepare()
    ((net(x) - y) ** 2).mean().backward()
    red.finish()
    ok = True
    for (n, p), (_, pr) in zip(net.named_parameters(), ref.named_parameters()):
        local = pr.grad if pr.grad is not None else torch.zeros_like(pr)
        ok &= p.grad is not None and torch.allclose(p.grad, local, atol=1e-7)       # the LOCAL gradient, in the bucket
        bi, pi = red._index[p]
        ok &= p.grad.data_ptr() == red.buckets[bi].views[pi].data_ptr()
    red.dry_run = False                   # and the exchange works again afterwards
    red.prepare()
    ((net(x) - y) ** 2).mean().backward()
    red.finish()
    for (n, p), (_, pr) in zip(net.named_parameters(), ref.named_parameters()):
        local = pr.grad if pr.grad is not None else torch.zeros_like(pr)
        parts = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        ok &= torch.allclose(p.grad, sum(parts) / world, atol=1e-7)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_dry_run_skips_the_collectives_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dry_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
