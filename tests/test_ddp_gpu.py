"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): after a few graph-replayed Trainer.train steps on
DIFFERENT data the replicas hold bit-identical weights (gradient arenas, D-side all-reduce under G1, G-side
exchange pipelined with DiffGrad, rank-consistent NaN flag), and a NaN on ONE rank makes BOTH raise."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp, q):
    import torch.distributed as dist
    from histogan_b200.trainer import NanException, SyntheticLoader, Trainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    torch.manual_seed(100 + rank)                     # different initial weights: init_GAN broadcasts rank 0's
    tr = Trainer("t", f"{tmp}/r{rank}/results", f"{tmp}/r{rank}/models", image_size=32, network_capacity=16,
                 batch_size=4, hist_insz=150, hist_resizing="interpolation", save_every=1000, cuda_graphs=True,
                 fast_rng=True)
    tr.loader = SyntheticLoader(4, 32, seed=10 + rank)
    tr.loader_evaluate = SyntheticLoader(4, 32, seed=1, eval_batch=4)
    tr.init_GAN()
    tr.steps = 2501
    for _ in range(6):                                # 2501..2506: plain steps and one gradient-penalty step
        tr.train(alpha=2)
    torch.cuda.synchronize()
    ok = True
    for p in list(tr.GAN.D.parameters()) + list(tr.GAN.G.parameters()) + list(tr.GAN.S.parameters()):
        both = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(both, p.detach().contiguous())
        ok = ok and torch.equal(both[0], both[1]) and bool(torch.isfinite(p).all())
    arenas = {k: all(s is not None for s in a.slots) for k, a in tr._arenas.items()}
    d_loss = tr.d_loss
    # NaN on rank 1 only -> the flag travels with the read-outs: both ranks reload and raise, one call later
    tr.save(2)
    if rank == 1:
        with torch.no_grad():
            next(iter(tr.GAN.D.parameters())).fill_(float("nan"))
    raised = 0
    try:
        tr.train(alpha=2)
        tr.train(alpha=2)
    except NanException:
        raised = 1
    q.put((rank, ok, arenas, d_loss == d_loss, raised, tr.steps))
    dist.destroy_process_group()


def test_two_rank_training_keeps_replicas_identical(tmp_path, cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    for rank, ok, arenas, finite, raised, steps in res:
        assert ok, (rank, "replicas diverged")
        assert arenas == {"d": True, "g": True}, arenas
        assert finite and raised == 1 and steps == 2000, (rank, finite, raised, steps)
