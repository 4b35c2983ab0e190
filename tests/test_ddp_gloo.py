"""CPU, world_size 2 over gloo: the data-parallel plumbing of the Trainer --
bucketed gradient averaging (== the reference's gradient_accumulate_every semantics,
histoGAN/histoGAN.py:924,977), rank-consistent NaN handling and weight broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from histogan_b200.trainer import _allreduce_mean_grads
    torch.manual_seed(0)
    shapes = [(7, 3), (5,), (64, 32, 3, 3), (1,), (1000, 13)]
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    g = torch.Generator().manual_seed(100 + rank)
    for p in params:
        p.grad = torch.randn(p.shape, generator=g)
    params.append(torch.nn.Parameter(torch.zeros(3)))           # a parameter without grad
    expected = []
    for s in shapes:
        acc = torch.zeros(s)
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            # replay rank r's stream up to this tensor
            for s2 in shapes:
                t = torch.randn(s2, generator=gr)
                if s2 is s:
                    acc += t
                    break
        expected.append(acc / world)
    _allreduce_mean_grads(params, bucket_bytes=4096)               # several buckets
    ok = all(torch.allclose(p.grad, e, atol=1e-6) for p, e in zip(params, expected))
    ok = ok and params[-1].grad is None
    # experimental overlapped exchange: every gradient is all-reduced by a hook as soon as autograd
    # has accumulated it; finish() turns the sums into means -> same result as above
    from histogan_b200.trainer import _GradOverlap
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))
    for q_ in net.parameters():
        dist.broadcast(q_.data, src=0)
    xs = [torch.randn(4, 6, generator=torch.Generator().manual_seed(7 + r)) for r in range(world)]
    want = [torch.zeros_like(q_) for q_ in net.parameters()]
    for r in range(world):
        net.zero_grad()
        net(xs[r]).pow(2).sum().backward()
        for acc, q_ in zip(want, net.parameters()):
            acc += q_.grad / world
    net.zero_grad()
    with _GradOverlap(net.parameters(), force=True) as ov:
        net(xs[rank]).pow(2).sum().backward()
    ov.finish()
    ok = ok and ov.enabled and all(torch.allclose(q_.grad, wnt, atol=1e-6)
                                   for q_, wnt in zip(net.parameters(), want))
    # flat gradient arena: slots alias one buffer, finalize() gathers gradients that were produced
    # elsewhere, ONE all-reduce averages everything (Trainer._exchange)
    from histogan_b200.trainer import GradArena
    from histogan_b200 import ops
    aps = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    aps[2].data = aps[2].data.contiguous(memory_format=torch.channels_last)        # a conv weight as stored
    arena = GradArena(aps)
    arena.begin()
    ga = torch.Generator().manual_seed(100 + rank)
    for i, p in enumerate(aps):
        gsrc = torch.randn(p.shape, generator=ga)
        if i == 2:                          # "the kernel wrote into the slot": adopt the slot as .grad
            slot = ops.grad_slot(p)
            assert slot is not None and ops.grad_slot(p) is None        # handed out once per backward
            slot.copy_(gsrc)
            p.grad = slot
        elif i != 3:
            p.grad = gsrc                   # produced by a torch op: gathered by finalize()
    arena.finalize()                        # parameter 3 had no gradient: zeros
    ok = ok and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(aps, arena.slots))
    arena.all_reduce_mean()
    for i, (p, e) in enumerate(zip(aps, expected)):
        want_i = torch.zeros_like(e) if i == 3 else e
        ok = ok and torch.allclose(p.grad, want_i, atol=1e-6)
    # pipelined exchange + update (Trainer._exchange_and_step): the arena reduced in pieces cut at slot
    # boundaries, each piece's parameters updated as soon as its piece has arrived == one all-reduce
    # followed by one optimiser step
    from histogan_b200.optim import DiffGrad
    pieces = arena.chunks(3)
    ok = ok and 1 < len(pieces) <= 3 and pieces[0][0] == 0 and pieces[-1][1] == arena.flat.numel()
    ok = ok and all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
    ok = ok and [id(q_) for _, _, ps in pieces for q_ in ps] == [id(q_) for q_ in aps]
    bps = [torch.nn.Parameter(q_.detach().clone()) for q_ in aps]
    ref_opt, pip_opt = DiffGrad(bps, lr=1e-2, betas=(0.5, 0.9)), DiffGrad(aps, lr=1e-2, betas=(0.5, 0.9))
    for it in range(2):
        gb = torch.Generator().manual_seed(500 + 10 * it + rank)
        for q_, v in zip(aps, arena.slots):
            v.copy_(torch.randn(q_.shape, generator=gb))
            q_.grad = v
        for q_, b_ in zip(aps, bps):
            b_.grad = q_.grad.detach().clone()
        _allreduce_mean_grads(bps)
        ref_opt.step()
        waits = [arena.all_reduce_mean_async(a, b) for a, b, _ in pieces]
        for wait, (_, _, ps) in zip(waits, pieces):
            wait()
            pip_opt.step(only=ps)
        ok = ok and all(torch.allclose(q_, b_, atol=1e-7) for q_, b_ in zip(aps, bps))
        ok = ok and all(pip_opt.state[q_]['step'] == it + 1 for q_ in aps)
    # NaN flag agreement (Trainer.train): MAX-reduce of a per-rank flag
    f = torch.tensor([1.0 if rank == 1 else 0.0])
    dist.all_reduce(f, op=dist.ReduceOp.MAX)
    ok = ok and bool(f.item() == 1.0)
    # replicas start from rank 0's weights (Trainer.init_GAN)
    w = torch.full((4,), float(rank))
    dist.broadcast(w, src=0)
    ok = ok and bool((w == 0).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res


def test_single_process_is_noop():
    from histogan_b200.trainer import _allreduce_mean_grads
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    _allreduce_mean_grads([p])
    assert torch.equal(p.grad, torch.ones(3))
