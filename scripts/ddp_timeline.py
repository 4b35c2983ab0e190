"""GPU (1 or N ranks under torchrun): where one plain graph-replayed train step spends its device time.
CUDA events around the phases of Trainer._train_graphed; prints the mean timeline of rank 0.
usage: [torchrun ...] python scripts/ddp_timeline.py <gpus>"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from histogan_b200.trainer import Trainer

gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dv = bench.Dist(gpus)
out_dir = os.path.join(bench.ROOT, "gpurun_out", f"tl_rank{dv.rank}")
torch.manual_seed(1234 + dv.rank)
tr = Trainer("tl", out_dir + "/results", out_dir + "/models", image_size=bench.S, network_capacity=bench.CAPACITY,
             batch_size=bench.B_PER_GPU, hist_insz=150, hist_resizing="interpolation", save_every=10 ** 9,
             fast_rng=True, cuda_graphs=True)
tr.loader = bench.DeviceLoader(dv.rank, dv.dev)
tr.loader_evaluate = bench.DeviceLoader(dv.rank, dv.dev, eval_only=True)
tr.steps = 2529
for _ in range(8):                         # capture every variant, warm up
    tr.train(alpha=bench.ALPHA)
marks = []


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


def wrap(obj, attr, name):
    fn = getattr(obj, attr)

    def w(*a, **k):
        mark(name(*a, **k) + ":begin") if callable(name) else mark(name + ":begin")
        r = fn(*a, **k)
        mark(name(*a, **k) + ":end") if callable(name) else mark(name + ":end")
        return r
    setattr(obj, attr, w)


wrap(tr, "_replay", lambda key, *a, **k: "replay " + str(key[0]))
wrap(tr.GAN.D_opt, "step", "D_opt.step")
if hasattr(tr, "_exchange_and_step"):
    wrap(tr, "_exchange_and_step", lambda kind, *a, **k: "exchange+step " + kind)
wrap(tr.GAN.G_opt, "step", "G_opt.step")
orig_async = tr._exchange_async


def async_wrapped(kind, params):
    mark("allreduce " + kind + " enqueued")
    w = orig_async(kind, params)
    if w is None:
        return None

    def wait():
        w()
        mark("allreduce " + kind + " waited")
    return wait


tr._exchange_async = async_wrapped
acc = collections.OrderedDict()
n = 0
for step in range(12):
    tr.steps = 2529 + 4 * step + (0 if True else 0)      # plain steps only (2529, 2533, ...: not multiples of 4)
    marks.clear()
    dv.barrier()
    mark("step begin")
    tr.train(alpha=bench.ALPHA)
    mark("step end")
    torch.cuda.synchronize()
    t0 = marks[0][1]
    for name, e in marks:
        acc.setdefault(name, []).append(t0.elapsed_time(e))
    n += 1
if dv.rank == 0:
    print(f"# {gpus} GPU(s), split_g={tr.split_g_phase} env HG_SPLIT_G={os.environ.get('HG_SPLIT_G')} chunks={getattr(tr, 'exchange_chunks', None)}"
          f" arenas={ {k: (v is not None and all(s is not None for s in v.slots)) for k, v in tr._arenas.items()} }")
    prev = 0.0
    for name, ts in acc.items():
        m = sorted(ts)[len(ts) // 2]
        print(f"{m:9.3f} ms  (+{m - prev:7.3f})  {name}")
        prev = m
dv.close()
