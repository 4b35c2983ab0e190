"""HistoGAN container + Trainer with the reference's public API
(``histoGAN/histoGAN.py:634-1139``) on the sm_100a kernels.

``Trainer.train(alpha)`` performs the same optimisation step as the reference
(D phase: hinge loss + R1-style gradient penalty every 4th step; G phase:
adversarial + Hellinger histogram loss + path-length regulariser every 32nd step;
EMA / NaN guard / checkpointing on the same schedule) but

* the histogram block and its loss are the fused CUDA kernels (hist.py),
* every convolution runs on the tcgen05 kernels (gan.py / ops.py),
* the generator pass of the D phase runs without building an autograd graph (the
  reference builds one and throws it away, :904-910),
* under ``torch.distributed`` (one process per GPU) gradients are averaged with a
  bucketed NCCL all-reduce before each optimiser step, which is exactly the
  reference's ``gradient_accumulate_every = world_size`` semantics (:924,977):
  the Hellinger loss keeps its per-micro-batch global sqrt.
"""
from __future__ import annotations

import json
import os
from math import floor, log2
from pathlib import Path
from random import random
from shutil import rmtree

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from .gan import (Discriminator, Generator, HistVectorizer, StyleVectorizer, EPS)
from .hist import RGBuvHistBlock, hellinger_loss
from .optim import DiffGrad

SCALE = 1 / np.sqrt(2.0)       # histoGAN/histoGAN.py:54


class NanException(Exception):
    pass


class EMA:
    def __init__(self, beta):
        self.beta = beta

    def update_average(self, old, new):
        if old is None:
            return new
        return old * self.beta + (1 - self.beta) * new


def default(value, d):
    return d if value is None else value


def cast_list(el):
    return el if isinstance(el, list) else [el]


def is_empty(t):
    if isinstance(t, torch.Tensor):
        return t.nelement() == 0
    return t is None


def raise_if_nan(t):
    if torch.isnan(t):
        raise NanException


def set_requires_grad(model, flag):
    for p in model.parameters():
        p.requires_grad = flag


# latent / noise helpers: random numbers are drawn on the CPU generator and copied, in the
# same order as the reference (histoGAN/histoGAN.py:166-189), so seeded runs see the same
# stream; `fast=True` draws on the device instead (no host round trip).

def noise(n, latent_dim, fast=False):
    if fast:
        return torch.randn(n, latent_dim, device='cuda')
    return torch.randn(n, latent_dim).cuda()


def noise_list(n, layers, latent_dim, fast=False):
    return [(noise(n, latent_dim, fast), layers)]


def mixed_list(n, layers, latent_dim, fast=False):
    tt = int(torch.rand(()).numpy() * layers)
    return noise_list(n, tt, latent_dim, fast) + noise_list(n, layers - tt, latent_dim, fast)


def image_noise(n, im_size, fast=False):
    if fast:
        return torch.rand(n, im_size, im_size, 1, device='cuda')
    return torch.FloatTensor(n, im_size, im_size, 1).uniform_(0.0, 1.0).cuda()


def latent_to_w(style_vectorizer, latent_descr):
    return [(style_vectorizer(z), num_layers) for z, num_layers in latent_descr]


def styles_def_to_tensor(styles_def):
    return torch.cat([t[:, None, :].expand(-1, n, -1) for t, n in styles_def], dim=1)


def evaluate_in_chunks(max_batch_size, model, *args):
    chunks = list(zip(*[a.split(max_batch_size, dim=0) for a in args]))
    outs = [model(*c) for c in chunks]
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)


def gradient_penalty(images, output, weight=10):
    """R1-style penalty on d D(x)/d x (histoGAN/histoGAN.py:156-163); needs the
    second-order gradients the conv ops provide."""
    from . import ops
    with ops.input_grads_only():             # d out / d images: no weight / bias gradients on the way
        (gradients,) = torch.autograd.grad(outputs=output, inputs=images,
                                           grad_outputs=torch.ones_like(output),
                                           create_graph=True, retain_graph=True, only_inputs=True)
    gradients = gradients.reshape(images.shape[0], -1)
    return weight * ((gradients.norm(2, dim=1) - 1) ** 2).mean()


class HistoGAN(nn.Module):
    """S, H, G, D + EMA twins SE, HE, GE and the two DiffGrad optimisers
    (histoGAN/histoGAN.py:634-715).  state_dict keys are the reference's."""

    def __init__(self, image_size, latent_dim=512, style_depth=8, network_capacity=16,
                 transparent=False, fp16=False, steps=1, lr=1e-4, fq_layers=[], fq_dict_size=256,
                 attn_layers=[], aug=False, hist=64):
        super().__init__()
        if fp16:
            raise NotImplementedError("Apex AMP (fp16=True) is not part of the sm_100a path; the "
                                      "convolutions already run TF32 operands / fp32 accumulate")
        self.lr = lr
        self.aug = aug
        self.steps = steps
        self.ema_updater = EMA(0.995)
        self.S = StyleVectorizer(latent_dim, style_depth)
        self.H = HistVectorizer(hist, latent_dim, int(style_depth))
        self.G = Generator(image_size, latent_dim, network_capacity, transparent=transparent)
        self.D = Discriminator(image_size, network_capacity, fq_layers=fq_layers,
                               fq_dict_size=fq_dict_size, attn_layers=attn_layers,
                               transparent=transparent)
        self.SE = StyleVectorizer(latent_dim, style_depth)
        self.HE = HistVectorizer(hist, latent_dim, int(style_depth))
        self.GE = Generator(image_size, latent_dim, network_capacity, transparent=transparent)
        if self.aug:
            raise NotImplementedError("DiffAugment (aug_prob > 0) is out of scope (off by default "
                                      "in the reference, histoGAN.py:254)")
        self.D_aug = None
        for m in (self.SE, self.HE, self.GE):
            set_requires_grad(m, False)
        g_params = list(self.G.parameters()) + list(self.S.parameters()) + list(self.H.parameters())
        self.G_opt = DiffGrad(g_params, lr=self.lr, betas=(0.5, 0.9))
        self.D_opt = DiffGrad(self.D.parameters(), lr=self.lr, betas=(0.5, 0.9))
        self._init_weights()
        self.reset_parameter_averaging()
        self.cuda()

    def _init_weights(self):
        for m in self.modules():
            if type(m) in {nn.Conv2d, nn.Linear}:
                nn.init.kaiming_normal_(m.weight, a=0, mode='fan_in', nonlinearity='leaky_relu')
        for block in self.G.blocks:
            for lin in (block.to_noise1, block.to_noise2):
                nn.init.zeros_(lin.weight)
                nn.init.zeros_(lin.bias)

    @torch.no_grad()
    def EMA(self):
        """SE/HE/GE <- beta * (SE/HE/GE) + (1 - beta) * (S/H/G)  (histoGAN.py:698-707) as ONE fused
        multi-tensor kernel (hg_ema_update): a single pass over the 100 M generator-side weights."""
        import ctypes as C
        from . import _lib
        beta = self.ema_updater.beta
        ma_p, cur_p = [], []
        for ma, cur in ((self.SE, self.S), (self.HE, self.H), (self.GE, self.G)):
            for a, c in zip(ma.parameters(), cur.parameters()):
                dense = a.is_contiguous() or (a.dim() == 4 and a.is_contiguous(memory_format=torch.channels_last))
                same = a.shape == c.shape and (a.stride() == c.stride()) and a.is_cuda and \
                    a.dtype == c.dtype == torch.float32 and dense
                if same:
                    ma_p.append(a); cur_p.append(c)
                else:               # layouts differ (e.g. a state_dict loaded into another format)
                    a.mul_(beta).add_(c, alpha=1 - beta)
        if not ma_p:
            return
        n = len(ma_p)
        arr = C.c_void_p * n
        dev = ma_p[0].device
        with torch.cuda.device(dev):
            rc = _lib.load().hg_ema_update(n, arr(*[t.data_ptr() for t in ma_p]),
                                           arr(*[t.data_ptr() for t in cur_p]),
                                           (C.c_int64 * n)(*[t.numel() for t in ma_p]), float(beta),
                                           _lib.current_stream_ptr(dev))
        _lib.check(rc, "hg_ema_update")
        torch.autograd.graph.increment_version(ma_p)

    def reset_parameter_averaging(self):
        self.SE.load_state_dict(self.S.state_dict())
        self.HE.load_state_dict(self.H.state_dict())
        self.GE.load_state_dict(self.G.state_dict())

    def forward(self, x):
        return x


class _PendingScalars:
    """the scalar read-outs of one graph-replayed step on their way to the host: ONE stack kernel, ONE
    device-to-host copy into pinned memory and an event -- the host does not wait.  `get()` blocks until the
    copy has landed.  (Four `.item()` calls at the end of every step kept the GPU idle during the whole
    host prologue of the next step: 0.8-1.0 ms of a 25 ms step, scripts/ddp_timeline.py.)"""

    _pool = []

    def __init__(self, named):
        self.names = list(named)
        dev = next(iter(named.values())).device
        stacked = torch.stack([t.detach().reshape(()).float() for t in named.values()])
        # slot 0: "the discriminator or the generator loss is NaN" (histoGAN.py:1003), max over the ranks
        flag = torch.isnan(stacked[:2]).any().float().reshape(1)
        if _ddp_active():
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        stacked = torch.cat((flag, stacked))
        self.buf = self._pool.pop() if self._pool else torch.empty(16, dtype=torch.float32).pin_memory()
        self.n = stacked.numel()
        self.buf[:self.n].copy_(stacked, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()
        self.values = None

    def get(self):
        """{'nan': bool, name: float, ...}"""
        if self.values is None:
            self.event.synchronize()
            v = self.buf[:self.n].tolist()
            self.values = dict(zip(self.names, v[1:]))
            self.values['nan'] = v[0] > 0
            self._pool.append(self.buf)
            self.buf = None
        return self.values


def _lazy_scalar(name):
    """attribute `name` of the Trainer: a plain float, except that after a graph-replayed step the value is
    still in flight (Trainer._pending) and is fetched on first access"""
    slot = '_lazy_' + name

    def get(self):
        pend = self.__dict__.get('_pending')
        if pend is not None and name in pend.names:
            self._adopt(pend)
        return self.__dict__.get(slot, 0)

    def set(self, value):
        self.__dict__[slot] = value

    return property(get, set)


class GradArena:
    """ONE contiguous float32 buffer holding the gradients of a parameter group (D, or G+S+H), so that
    the data-parallel exchange is a single all-reduce instead of ~100-200 small ones.

    Every parameter owns a slot (a view with the parameter's own memory layout).  The weight-gradient
    kernels write conv gradients straight into their slots (ops.grad_slot), autograd adopts those
    tensors as `.grad`; whatever arrived elsewhere (small tensors produced by torch ops) is copied
    into its slot by finalize(), which also re-points `.grad` at the slot -- the optimiser then reads
    the arena.  All of it is capturable in a CUDA graph (fixed addresses)."""

    ALIGN = 128          # floats: 512-byte slots keep every view 16-byte aligned for the vector kernels

    def __init__(self, params):
        from . import ops
        self.params = [p for p in params]
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.slots = []
        for p, o in zip(self.params, offs):
            flat = self.flat[o:o + p.numel()]
            if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
                co, ci, kh, kw = p.shape
                v = flat.view(co, kh, kw, ci).permute(0, 3, 1, 2)
            else:
                v = flat.view(p.shape) if p.is_contiguous() else None
            self.slots.append(v)
            if v is not None:
                setattr(p, ops.GRAD_SLOT, v)

    def begin(self):
        from . import ops
        ops.new_backward()

    def finalize(self):
        """after backward: every gradient lives in (and `.grad` points at) its slot"""
        src, dst = [], []
        for p, v in zip(self.params, self.slots):
            if v is None:
                continue
            g = p.grad
            if g is None:
                v.zero_()               # parameter unused in this phase: contributes zeros
            elif g.data_ptr() != v.data_ptr():
                src.append(g); dst.append(v)
            p.grad = v
        if src:
            torch._foreach_copy_(dst, src)

    def chunks(self, k):
        """the arena cut into <= k contiguous pieces of about equal size at slot boundaries:
        [(first float, one past the last float, [parameters])]"""
        bounds, total = [], 0
        for p in self.params:
            bounds.append(total)
            total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        bounds.append(total)
        out, first, want = [], 0, total / max(1, k)
        for i in range(len(self.params)):
            last = i == len(self.params) - 1
            if last or (bounds[i + 1] >= want * (len(out) + 1) and len(out) < k - 1):
                out.append((bounds[first], bounds[i + 1], self.params[first:i + 1]))
                first = i + 1
        return out

    def all_reduce_mean_async(self, a, b):
        """start averaging flat[a:b] over the ranks; returns the wait() of the collective"""
        piece = self.flat[a:b]
        if self.flat.is_cuda and dist.get_backend() == 'nccl':
            return dist.all_reduce(piece, op=dist.ReduceOp.AVG, async_op=True).wait
        w = dist.all_reduce(piece, op=dist.ReduceOp.SUM, async_op=True)
        world = dist.get_world_size()

        def wait():
            w.wait()
            piece.div_(world)
        return wait

    def all_reduce_mean(self):
        world = dist.get_world_size()
        if self.flat.is_cuda and dist.get_backend() == 'nccl':
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)       # one collective, mean inside NCCL
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world)
        leftovers = [p for p, v in zip(self.params, self.slots) if v is None and p.grad is not None]
        if leftovers:
            _allreduce_mean_grads(leftovers)


def _ddp_active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _allreduce_mean_grads(params, bucket_bytes=128 << 20):
    """bucketed NCCL all-reduce (mean) of .grad over the default process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    if grads[0].is_cuda and dist.get_backend() == 'nccl' and hasattr(dist, '_coalescing_manager'):
        # NCCL: one grouped launch that reduces every gradient tensor IN PLACE (no flatten /
        # copy-back passes over the 364 + 399 MB of gradients), then one multi-tensor divide
        try:
            with dist._coalescing_manager(device=grads[0].device, async_ops=True) as cm:
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.SUM)
            cm.wait()
            torch._foreach_div_(grads, world)
            return
        except (RuntimeError, TypeError, AttributeError):      # older torch: fall through
            pass
    bucket, size, works = [], 0, []

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, bucket))
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    for work, flat, tensors in works:
        work.wait()
        flat.div_(world)
        pieces = flat.split([t.numel() for t in tensors])
        torch._foreach_copy_(tensors, [p.view_as(t) for p, t in zip(pieces, tensors)])


class _GradOverlap:
    """EXPERIMENTAL, off unless HG_OVERLAP_ALLREDUCE=1 (not yet measured on >1 GPU): all-reduce each
    gradient as soon as autograd has accumulated it, on the process group's own stream, so that the
    exchange overlaps the rest of the backward pass instead of following it.  Used inside the
    captured phases (NCCL collectives are capturable), which puts the collectives INTO the CUDA
    graph; ``finish()`` waits for them and turns the sums into means.

        with _GradOverlap(params) as ov:
            loss.backward()
        ov.finish()
    """

    def __init__(self, params, force=None):
        on = os.environ.get('HG_OVERLAP_ALLREDUCE', '0') != '0' if force is None else force
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.enabled = bool(on) and self.world > 1
        self.params = [p for p in params if p.requires_grad]
        self.works, self.handles = [], []

    def _hook(self, p):
        self.works.append((dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=True), p))

    def __enter__(self):
        if self.enabled:
            self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        return self

    def __exit__(self, *exc):
        for h in self.handles:
            h.remove()
        self.handles = []
        return False

    def finish(self):
        if not self.enabled:
            return
        for w, _ in self.works:
            w.wait()
        grads = [p.grad for _, p in self.works]
        if grads:
            torch._foreach_div_(grads, self.world)
        self.works = []


class Trainer:
    """Same constructor / attributes / methods as the reference Trainer
    (histoGAN/histoGAN.py:718-1139)."""

    # the step's scalar read-outs (histoGAN.py:930,982-983 assign floats): same values, fetched lazily
    d_loss = _lazy_scalar('d_loss')
    g_loss = _lazy_scalar('g_loss')
    h_loss = _lazy_scalar('h_loss')
    last_gp_loss = _lazy_scalar('last_gp_loss')

    def _adopt(self, pend):
        """the values of a finished step become the attributes"""
        vals = pend.get()
        if self.__dict__.get('_pending') is pend:
            self._pending = None
        for k in ('d_loss', 'g_loss', 'h_loss', 'last_gp_loss'):
            if k in vals:
                self.__dict__['_lazy_' + k] = vals[k]
        return vals

    def __init__(self, name, results_dir, models_dir, image_size, network_capacity,
                 transparent=False, batch_size=4, mixed_prob=0.9, gradient_accumulate_every=1,
                 lr=2e-4, num_workers=None, save_every=1000, trunc_psi=0.6, fp16=False,
                 fq_layers=[], fq_dict_size=256, attn_layers=[], hist_method='inverse-quadratic',
                 hist_resizing='sampling', hist_sigma=0.02, hist_bin=64, hist_insz=150,
                 aug_prob=0.0, dataset_aug_prob=0.0, aug_types=None, *args, **kwargs):
        if aug_types is None:
            aug_types = ['translation', 'cutout']
        self.fast_rng = bool(kwargs.pop('fast_rng', False))
        # cuda_graphs=True: replay each phase (forward + backward) of a step as one CUDA graph
        # (see _train_graphed): D with / without gradient penalty, G with / without the
        # path-length regulariser
        self.cuda_graphs = bool(kwargs.pop('cuda_graphs', False))
        self._graphs = {}
        self._param_lists = None
        self._pending = None
        self._static = None
        self._arenas = {}
        # flat gradient arenas (one all-reduce per phase): on under torch.distributed, or forced
        self.use_grad_arena = kwargs.pop('grad_arena', None)
        # 'deferred' (default on the CUDA-graph path): train() returns without waiting for the GPU; a NaN loss
        # of step N raises NanException from the train() call of step N+1 (from the same call on
        # checkpoint / evaluation / path-length steps).  'immediate': the reference's timing, one host
        # sync per step.
        self.nan_check = kwargs.pop('nan_check', os.environ.get('HG_NAN_CHECK', 'deferred'))
        self._pending = None
        # DDP: pieces of the G-side gradient arena whose all-reduce is pipelined with the optimiser
        self.exchange_chunks = int(kwargs.pop('exchange_chunks', os.environ.get('HG_EXCHANGE_CHUNKS', 4)))
        # split the captured G phase into a D-independent part and the rest (overlaps the D-side
        # all-reduce); None = only under torch.distributed
        self.split_g_phase = kwargs.pop('split_g_phase', None)
        self.graph_replayed_launches = 0      # library kernels launched through graph replays
        self.GAN_params = [args, kwargs]
        self.GAN = None
        self.hist_method = hist_method
        self.hist_resizing = hist_resizing
        self.hist_sigma = hist_sigma
        self.hist_bin = hist_bin
        self.hist_insz = hist_insz
        self.histBlock = RGBuvHistBlock(insz=self.hist_insz, h=self.hist_bin,
                                        method=self.hist_method, resizing=self.hist_resizing,
                                        sigma=self.hist_sigma)
        self.name = name
        self.results_dir = Path(results_dir)
        self.models_dir = Path(models_dir)
        self.config_path = self.models_dir / name / '.config.json'
        assert log2(image_size).is_integer(), 'image size must be a power of 2 (64, 128, 256, 512, 1024)'
        self.image_size = image_size
        self.network_capacity = network_capacity
        self.transparent = transparent
        self.fq_layers = cast_list(fq_layers)
        self.fq_dict_size = fq_dict_size
        self.attn_layers = cast_list(attn_layers)
        self.aug_prob = aug_prob
        self.aug_types = aug_types
        self.dataset_aug_prob = dataset_aug_prob
        self.lr = lr
        self.batch_size = batch_size
        self.num_workers = num_workers
        self.mixed_prob = mixed_prob
        self.save_every = save_every
        self.steps = 0
        self.av = None
        self.trunc_psi = trunc_psi
        self.pl_mean = 0
        self.gradient_accumulate_every = gradient_accumulate_every
        assert not fp16, 'Apex mixed precision is not available on the sm_100a path'
        self.fp16 = fp16
        self.d_loss = 0
        self.g_loss = 0
        self.last_gp_loss = 0
        self.last_cr_loss = 0
        self.q_loss = 0
        self.pl_length_ma = EMA(0.99)
        self.init_folders()
        self.loader = None
        self.loader_evaluate = None

    # ------------------------------------------------------------ plumbing --
    def init_GAN(self):
        args, kwargs = self.GAN_params
        # captured graphs hold the OLD parameter / gradient tensors: a new GAN (first call, or
        # load() -> load_config() after a NaN) must be captured afresh
        self._graphs = {}
        self._param_lists = None
        self._pending = None
        self._static = None
        self._arenas = {}
        self.GAN = HistoGAN(lr=self.lr, image_size=self.image_size,
                            network_capacity=self.network_capacity, transparent=self.transparent,
                            fq_layers=self.fq_layers, fq_dict_size=self.fq_dict_size,
                            attn_layers=self.attn_layers, fp16=self.fp16, hist=self.hist_bin,
                            aug=self.aug_prob > 0, *args, **kwargs)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            for p in self.GAN.parameters():        # replicas start from rank 0's weights
                dist.broadcast(p.data, src=0)

    def config(self):
        return {'image_size': self.image_size, 'network_capacity': self.network_capacity,
                'transparent': self.transparent, 'fq_layers': self.fq_layers,
                'fq_dict_size': self.fq_dict_size, 'attn_layers': self.attn_layers}

    def write_config(self):
        self.config_path.write_text(json.dumps(self.config()))

    def load_config(self):
        config = self.config() if not self.config_path.exists() else json.loads(
            self.config_path.read_text())
        self.image_size = config['image_size']
        self.network_capacity = config['network_capacity']
        self.transparent = config['transparent']
        self.fq_layers = config['fq_layers']
        self.fq_dict_size = config['fq_dict_size']
        self.attn_layers = config.pop('attn_layers', [])
        del self.GAN
        self.init_GAN()

    def set_data_src(self, folder):
        from .data import make_loaders
        self.loader, self.loader_evaluate = make_loaders(self, folder)

    def _is_main(self):
        return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0

    def _arena(self, kind):
        """the flat gradient arena of the D ('d') or G+S+H ('g') parameters, or None when off"""
        on = self.use_grad_arena
        if on is None:                      # HG_GRAD_ARENA=1 / 0 forces it on / off; default: under DDP
            env = os.environ.get('HG_GRAD_ARENA')
            on = _ddp_active() if env is None else env != '0'
        if not on:
            return None
        if kind not in self._arenas:
            GAN = self.GAN
            params = list(GAN.D.parameters()) if kind == 'd' else \
                [p for grp in GAN.G_opt.param_groups for p in grp['params']]
            self._arenas[kind] = GradArena(params)
        return self._arenas[kind]

    def _exchange_async(self, kind, params):
        """start averaging one parameter group's gradients over the ranks; returns a callable that
        makes the current stream wait for the result (None on one GPU)."""
        if not _ddp_active():
            return None
        arena = self._arenas.get(kind)
        if arena is None or not (arena.flat.is_cuda and dist.get_backend() == 'nccl') or \
                any(v is None for v in arena.slots):
            self._exchange(kind, params)
            return None
        w = dist.all_reduce(arena.flat, op=dist.ReduceOp.AVG, async_op=True)   # on NCCL's own stream
        return w.wait

    def _exchange_and_step(self, kind, params, opt):
        """gradient exchange + optimiser step of one parameter group.  Under DDP with a flat arena the
        two are pipelined: the arena is reduced in `exchange_chunks` pieces on NCCL's stream and the
        fused DiffGrad kernel of piece i runs while piece i+1 is still on the wire (the update is
        element-wise, so the result is the one of exchange-then-step, bit for bit)."""
        arena = self._arenas.get(kind)
        k = self.exchange_chunks
        if not _ddp_active() or arena is None or k <= 1 or any(v is None for v in arena.slots):
            self._exchange(kind, params)
            opt.step()
            return
        pieces = arena.chunks(k)
        waits = [arena.all_reduce_mean_async(a, b) for a, b, _ in pieces]
        for wait, (_, _, ps) in zip(waits, pieces):
            wait()
            opt.step(only=ps)

    def _exchange(self, kind, params):
        """average the gradients of one parameter group over the ranks (no-op on one GPU)"""
        arena = self._arenas.get(kind)
        if not _ddp_active():
            return
        if arena is not None:
            arena.all_reduce_mean()
        else:
            _allreduce_mean_grads(params)

    # --------------------------------------------------------------- train --
    def _sample_latents(self, get_latents_fn, batch_size, num_layers, latent_dim, image_size):
        style = get_latents_fn(batch_size, num_layers - 2, latent_dim, self.fast_rng)
        return style, image_noise(batch_size, image_size, self.fast_rng)

    def _generate(self, style, hist_batch, inoise):
        GAN = self.GAN
        h_w = GAN.H(hist_batch).unsqueeze(1)
        h_w = torch.cat((h_w, h_w), dim=1)                         # last two blocks (:900-902)
        w_styles = styles_def_to_tensor(latent_to_w(GAN.S, style))
        return GAN.G(w_styles, h_w, inoise), w_styles, h_w

    def train(self, alpha=2):
        assert self.loader is not None, ('You must first initialize the data source with '
                                         '`. set_data_src(<folder of images>)`')
        torch.autograd.set_detect_anomaly(False)
        if self.GAN is None:
            self.init_GAN()
        GAN = self.GAN
        if not GAN.training:                 # the walk over ~600 modules costs 0.2 ms of GPU-idle host time
            GAN.train()
        batch_size = self.batch_size
        image_size, latent_dim, num_layers = GAN.G.image_size, GAN.G.latent_dim, GAN.G.num_layers
        accum = self.gradient_accumulate_every
        apply_gradient_penalty = self.steps % 4 == 0
        apply_path_penalty = self.steps % 32 == 0
        avg_pl_length = self.pl_mean
        if self.cuda_graphs and accum == 1:
            previous, pending = self._train_graphed(alpha, apply_gradient_penalty, apply_path_penalty)
            return self._finish_graphed_step(previous, pending, apply_path_penalty)

        total_disc_loss = torch.tensor(0.0).cuda()
        total_gen_loss = torch.tensor(0.0).cuda()
        total_hist_loss = torch.tensor(0.0).cuda()
        # ---------------------------------------------------- discriminator --
        GAN.D_opt.zero_grad()
        arena_d = self._arena('d')
        if arena_d is not None:
            arena_d.begin()
        for _ in range(accum):
            get_latents_fn = mixed_list if random() < self.mixed_prob else noise_list
            style, inoise = self._sample_latents(get_latents_fn, batch_size, num_layers,
                                                 latent_dim, image_size)
            batch = next(self.loader)
            image_batch = batch['images'].cuda(non_blocking=True)
            image_batch.requires_grad_()
            hist_batch = batch['histograms'].cuda(non_blocking=True)
            with torch.no_grad():                                   # graph is never used (:910)
                generated_images, _, _ = self._generate(style, hist_batch, inoise)
            fake_output, fake_q_loss = GAN.D(generated_images)
            real_output, real_q_loss = GAN.D(image_batch)
            divergence = (F.relu(1 + real_output) + F.relu(1 - fake_output)).mean()
            quantize_loss = (fake_q_loss + real_q_loss).mean()
            self.q_loss = float(quantize_loss.detach().item())
            disc_loss = divergence + quantize_loss
            if apply_gradient_penalty:
                gp = gradient_penalty(image_batch, real_output)
                self.last_gp_loss = gp.clone().detach().item()
                disc_loss = disc_loss + gp
            disc_loss = disc_loss / accum
            disc_loss.register_hook(raise_if_nan)
            disc_loss.backward()
            total_disc_loss += divergence.detach().item() / accum
        self.d_loss = float(total_disc_loss)
        if arena_d is not None:
            arena_d.finalize()
        self._exchange_and_step('d', list(GAN.D.parameters()), GAN.D_opt)

        # -------------------------------------------------------- generator --
        GAN.G_opt.zero_grad()
        g_params = [p for grp in GAN.G_opt.param_groups for p in grp['params']]
        arena_g = self._arena('g')
        if arena_g is not None:
            arena_g.begin()
        # the reference lets this phase's backward fill D's parameter gradients too, only to
        # zero them at the next D_opt.zero_grad() (:886): skip that dead wgrad work
        set_requires_grad(GAN.D, False)
        for _ in range(accum):
            style, inoise = self._sample_latents(get_latents_fn, batch_size, num_layers,
                                                 latent_dim, image_size)
            batch = next(self.loader)
            hist_batch = batch['histograms'].cuda(non_blocking=True)
            hist_batch.requires_grad_()
            generated_images, w_styles, h_w = self._generate(style, hist_batch, inoise)
            fake_output, _ = GAN.D(generated_images)
            generated_histograms = self.histBlock(F.relu(generated_images))      # :955
            histogram_loss = hellinger_loss(hist_batch, generated_histograms, alpha)  # :957-960
            loss = fake_output.mean()
            gen_loss = loss + histogram_loss
            if apply_path_penalty:
                std = 0.1 / (w_styles.std(dim=0, keepdim=True) + EPS)
                pert = torch.randn(w_styles.shape, device='cuda') if self.fast_rng else \
                    torch.randn(w_styles.shape).cuda()
                w_styles_2 = w_styles + pert / (std + EPS)
                pl_images = GAN.G(w_styles_2, h_w, inoise)
                pl_lengths = ((pl_images - generated_images) ** 2).mean(dim=(1, 2, 3))
                avg_pl_length = np.mean(pl_lengths.detach().cpu().numpy())
                if not is_empty(self.pl_mean):
                    pl_loss = ((pl_lengths - self.pl_mean) ** 2).mean()
                    if not torch.isnan(pl_loss):
                        gen_loss = gen_loss + pl_loss
            gen_loss = gen_loss / accum
            gen_loss.register_hook(raise_if_nan)
            gen_loss.backward()
            total_gen_loss += loss.detach().item() / accum
            total_hist_loss += histogram_loss.detach().item() / accum
        set_requires_grad(GAN.D, True)
        self.g_loss = float(total_gen_loss)
        self.h_loss = float(total_hist_loss)
        if arena_g is not None:
            arena_g.finalize()
        self._exchange_and_step('g', g_params, GAN.G_opt)

        return self._finish_step(total_disc_loss, total_gen_loss, total_hist_loss,
                                 apply_path_penalty, avg_pl_length)

    def _finish_graphed_step(self, previous, pending, apply_path_penalty):
        """bookkeeping of a graph-replayed step.  The host waits for THIS step's read-outs only when it needs
        them now (path-length mean, checkpoint, evaluation, nan_check='immediate'); otherwise it checks the
        PREVIOUS step's NaN flag -- by now long on the host -- and returns while the GPU still works."""
        save_now = self.steps % self.save_every == 0
        eval_now = self.steps % 1000 == 0 or (self.steps % 100 == 0 and self.steps < 2500)
        nan, nan_step = False, self.steps
        if previous is not None:
            pv = previous.get()
            if 'last_gp_loss' in pv:                  # persists over the non-penalty steps (:922)
                self.__dict__['_lazy_last_gp_loss'] = pv['last_gp_loss']
            if pv['nan']:
                nan, nan_step = True, self.steps - 1
        if apply_path_penalty or save_now or eval_now or self.nan_check != 'deferred' or nan:
            vals = self._adopt(pending)
            nan = nan or vals['nan']
            avg_pl = vals.get('avg_pl', self.pl_mean)
        else:
            avg_pl = self.pl_mean
        return self._finish_step(float('nan') if nan else 0.0, 0.0, 0.0, apply_path_penalty, avg_pl,
                                 nan_reduced=True, nan_step=nan_step)

    def _finish_step(self, total_disc_loss, total_gen_loss, total_hist_loss, apply_path_penalty,
                     avg_pl_length, nan_reduced=False, nan_step=None):
        GAN = self.GAN
        # ------------------------------------------------------ bookkeeping --
        if apply_path_penalty and not np.isnan(avg_pl_length):
            self.pl_mean = self.pl_length_ma.update_average(self.pl_mean, avg_pl_length)
        if self.steps % 10 == 0 and self.steps > 20000:
            GAN.EMA()
        if self.steps <= 25000 and self.steps % 1000 == 2:
            GAN.reset_parameter_averaging()

        checkpoint_num = floor(self.steps / self.save_every)
        if isinstance(total_gen_loss, float):        # graph path: the losses are already on the host
            nan_flag = total_gen_loss != total_gen_loss or total_disc_loss != total_disc_loss
            if _ddp_active() and not nan_reduced:
                nan_flag = torch.tensor(float(nan_flag), device='cuda')
        else:
            nan_flag = torch.isnan(total_gen_loss) | torch.isnan(total_disc_loss)
        if _ddp_active() and not nan_reduced:
            f = nan_flag.float()
            dist.all_reduce(f, op=dist.ReduceOp.MAX)               # all ranks retry together
            nan_flag = f > 0
        if bool(nan_flag):
            if nan_step is not None:                  # deferred detection: the step that produced the NaN
                checkpoint_num = floor(nan_step / self.save_every)
            print(f'NaN detected for generator or discriminator. Loading from checkpoint '
                  f'#{checkpoint_num}')
            self.load(checkpoint_num)
            raise NanException
        if self.steps % self.save_every == 0 and self._is_main():
            self.save(checkpoint_num)
        if (self.steps % 1000 == 0 or (self.steps % 100 == 0 and self.steps < 2500)) and self._is_main():
            self.evaluate(floor(self.steps / 1000))
        self.steps += 1
        self.av = None

    # ------------------------------------------------------- CUDA-graph path --
    def _mixed_styles(self, z1, z2, mask):
        """w_styles of mixed_list / noise_list (histoGAN.py:170-176,215-217) with the split
        point as a device-side 0/1 mask over the layers, so the graph topology is fixed."""
        S = self.GAN.S
        w1, w2 = S(z1)[:, None, :], S(z2)[:, None, :]
        m = mask[None, :, None]
        return w1 * m + w2 * (1 - m)

    def _device_draws(self, phase, with_pl=False):
        """latents / image noise of one captured phase, drawn on the device inside the graph
        (a replay advances the graph's Philox offset).  Tests pin them via _static['fixed']."""
        GAN, st = self.GAN, self._static
        fixed = st.get('fixed')
        if fixed is not None:
            return fixed[phase]
        B, S_, L = self.batch_size, GAN.G.image_size, GAN.G.num_layers - 2
        out = {'z1': torch.randn(B, GAN.G.latent_dim, device='cuda'),
               'z2': torch.randn(B, GAN.G.latent_dim, device='cuda'),
               'inoise': torch.rand(B, S_, S_, 1, device='cuda')}
        if with_pl:
            out['pl_noise'] = torch.randn(B, L, GAN.G.latent_dim, device='cuda')
        return out

    def _phase_d(self, apply_gp):
        GAN, st = self.GAN, self._static
        # the backward allocates this graph's own gradient tensors (kept alive by _graphed and
        # re-attached to the parameters after every replay): no zero-fill / accumulate kernels
        GAN.D_opt.zero_grad(set_to_none=True)
        arena = self._arena('d')
        if arena is not None:
            arena.begin()
        dr = self._device_draws('d')
        z1, z2, inoise = dr['z1'], dr['z2'], dr['inoise']
        with torch.no_grad():
            h_w = GAN.H(st['hists']).unsqueeze(1)
            fake = GAN.G(self._mixed_styles(z1, z2, st['mask']), torch.cat((h_w, h_w), dim=1), inoise)
        # a fresh leaf over the static buffer: its .grad is a graph-local temporary
        images = st['images'].detach().requires_grad_(True) if apply_gp else st['images']
        if apply_gp:
            fake_out, _ = GAN.D(fake)
            real_out, _ = GAN.D(images)
        else:
            # one discriminator pass over [fake; real] (the same per-sample arithmetic, histoGAN.py
            # :904-911): half the launches, and D's small-resolution layers run 2x fuller tiles.
            # Steps with the gradient penalty keep two passes: its d D(real)/d real backward would
            # otherwise drag the fake half along.
            both, _ = GAN.D(torch.cat((fake, images), dim=0))
            fake_out, real_out = both[:fake.shape[0]], both[fake.shape[0]:]
        divergence = (F.relu(1 + real_out) + F.relu(1 - fake_out)).mean()
        loss, gp = divergence, None
        if apply_gp:
            gp = gradient_penalty(images, real_out)
            loss = loss + gp
        with _GradOverlap(GAN.D.parameters()) as ov:
            loss.backward()
        ov.finish()
        if arena is not None:
            arena.finalize()
        return divergence.detach(), (gp.detach() if gp is not None else None)

    def _phase_g(self, alpha, apply_pl=False):
        """G phase (histoGAN.py:934-989).  apply_pl: + the path-length regulariser (:965-975)
        with `pl_mean` read from a device scalar, so that PL steps are capturable too; the
        reference's host-side `if not isnan(pl_loss)` becomes a select on the device.
        = _phase_g1 (everything that does not depend on D) followed by _phase_g2."""
        self._phase_g1(apply_pl)
        return self._phase_g2(alpha, apply_pl)

    def _phase_g1(self, apply_pl=False):
        """generator side of the G phase: latents -> S / H -> G forward (twice on path-length steps).
        Independent of the discriminator's weights, so under DDP it runs WHILE the D-side gradient
        all-reduce of the same step is in flight (see _train_graphed)."""
        GAN, st = self.GAN, self._static
        GAN.G_opt.zero_grad(set_to_none=True)
        arena = self._arena('g')
        if arena is not None:
            arena.begin()
        dr = self._device_draws('g', apply_pl)
        z1, z2, inoise = dr['z1'], dr['z2'], dr['inoise']
        hists = st['hists']      # the reference's hist_batch.requires_grad_() (:940) is never used
        h_w = GAN.H(hists).unsqueeze(1)
        h_w = torch.cat((h_w, h_w), dim=1)
        w_styles = self._mixed_styles(z1, z2, st['mask'])
        fake = GAN.G(w_styles, h_w, inoise)
        pl_images = None
        if apply_pl:
            std = 0.1 / (w_styles.std(dim=0, keepdim=True) + EPS)
            pl_images = GAN.G(w_styles + dr['pl_noise'] / (std + EPS), h_w, inoise)
        self._g1 = (fake, pl_images, hists)
        return ()

    def _phase_g2(self, alpha, apply_pl=False):
        """the rest of the G phase: D(fake), histogram loss, path-length loss, backward"""
        GAN, st = self.GAN, self._static
        fake, pl_images, hists = self._g1
        self._g1 = None
        arena = self._arena('g')
        set_requires_grad(GAN.D, False)          # D's parameter gradients are dead work here
        avg_pl = None
        try:
            fake_out, _ = GAN.D(fake)
            hist_loss = hellinger_loss(hists, self.histBlock(F.relu(fake)), alpha)
            loss = fake_out.mean()
            gen_loss = loss + hist_loss
            if apply_pl:
                pl_lengths = ((pl_images - fake) ** 2).mean(dim=(1, 2, 3))
                avg_pl = pl_lengths.detach().mean()
                pl_loss = ((pl_lengths - st['pl_mean']) ** 2).mean()
                gen_loss = gen_loss + torch.where(torch.isnan(pl_loss), torch.zeros_like(pl_loss), pl_loss)
            with _GradOverlap([p for grp in GAN.G_opt.param_groups for p in grp['params']]) as ov:
                gen_loss.backward()
            ov.finish()
            if arena is not None:
                arena.finalize()
        finally:
            set_requires_grad(GAN.D, True)
        return loss.detach(), hist_loss.detach(), avg_pl

    def _capture(self, keys, fns, params_list):
        """capture the functions `fns` (run in this order; later ones may consume tensors earlier ones
        left on `self`) as one CUDA graph each: an eager warm-up of the whole sequence on a side
        stream, then the captures, all sharing one memory pool."""
        from . import ops, _lib
        lib = _lib.load()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):           # eager warm-up on a side stream
            for fn in fns:
                fn()
        torch.cuda.current_stream().wait_stream(side)
        # the warm-up filled the packed-weight cache (stable addresses, refilled in place by the
        # optimiser step): the captured graphs read those tensors and pack nothing themselves
        for key, fn, params in zip(keys, fns, params_list):
            g = torch.cuda.CUDAGraph()
            n0 = lib.hg_launch_count()
            with torch.cuda.graph(g, pool=self._graphs.get('pool')):
                outs = fn()
            self._graphs.setdefault('pool', g.pool())
            # library kernels recorded in this graph (each replay launches them again)
            grads = [(p, p.grad) for p in params if p.grad is not None]
            packed = [p for p in self.GAN.parameters() if getattr(p, ops._PackCache.ATTR, None)]
            self._graphs[key] = (g, outs, int(lib.hg_launch_count() - n0), grads, packed)

    def _replay(self, key):
        from . import ops
        entry = self._graphs[key]
        ops._packs.refresh_stale(entry[4])      # no-op unless someone other than DiffGrad moved weights
        entry[0].replay()
        for p, gr in entry[3]:
            p.grad = gr
        self.graph_replayed_launches += entry[2]
        # The graphs share one memory pool: a graph captured EARLIER may use, as scratch, the block
        # in which a later-captured graph keeps an output.  Gradients are consumed (all-reduce +
        # optimiser) before the next replay; the scalar outputs are read at the end of the step,
        # so hand out copies that live outside the pool.
        return tuple(o.clone() if o is not None else None for o in entry[1])

    def _graphed(self, key, fn, params):
        """capture `fn` (one phase: zero_grad + forward + backward) once, then replay.

        Each graph writes the gradients of `params` into tensors of its own memory pool; they
        are kept alive (stable addresses, refreshed by every replay) and attached to the
        parameters after the replay, so the optimiser always reads what the graph just wrote
        -- whichever variant (with / without gradient penalty) ran."""
        if key not in self._graphs:
            self._capture([key], [fn], [params])
        return self._replay(key)

    def _to_static(self, key, src):
        """batch tensor -> the fixed-address graph input `key`.  A HOST batch goes through a copy stream and
        two device staging buffers: train() returns before the GPU has finished the step, so the copy of
        step N+1's batch is issued while step N still computes -- on the main stream it would queue behind
        it (25 MB of images per step: 0.5-1 ms of PCIe time on the critical path)."""
        st = self._static
        dst = st[key]
        if src.is_cuda:
            dst.copy_(src, non_blocking=True)
            return
        ring = st.get('ring_' + key)
        if ring is None:
            ring = st['ring_' + key] = {'buf': [torch.empty_like(dst) for _ in range(2)], 'free': [None, None], 'i': 0}
        if 'copy_stream' not in st:
            st['copy_stream'] = torch.cuda.Stream()
        cs, i = st['copy_stream'], ring['i']
        ring['i'] ^= 1
        with torch.cuda.stream(cs):
            if ring['free'][i] is not None:
                cs.wait_event(ring['free'][i])           # the step that read this staging buffer has copied it out
            ring['buf'][i].copy_(src, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(cs)
        main = torch.cuda.current_stream()
        main.wait_event(ready)
        dst.copy_(ring['buf'][i], non_blocking=True)
        ring['free'][i] = torch.cuda.Event()
        ring['free'][i].record(main)

    def _train_graphed(self, alpha, apply_gp, apply_pl=False):
        GAN = self.GAN
        B, S_, L = self.batch_size, GAN.G.image_size, GAN.G.num_layers - 2
        if self._static is None:
            self._static = {
                'images': torch.zeros(B, 3, S_, S_, device='cuda'),
                'hists': torch.zeros(B, 3, self.hist_bin, self.hist_bin, device='cuda'),
                'mask': torch.ones(L, device='cuda'),
                'pl_mean': torch.zeros((), device='cuda'),
                # one pinned staging buffer per phase: the G-phase mask must not overwrite the
                # D-phase one while its H2D copy may still be queued
                'mask_host': [torch.ones(L).pin_memory(), torch.ones(L).pin_memory()],
                'mask_event': [None, None],
            }
        st = self._static
        # mixed or not is drawn ONCE per step and reused by the G phase (histoGAN.py:891,936);
        # each phase draws its own split point, as mixed_list does (:174-176)
        get_mixed = random() < self.mixed_prob

        def stage(batch, phase):
            tt = int(torch.rand(()).numpy() * L) if get_mixed else L
            ev = st['mask_event'][phase]
            if ev is not None:
                ev.synchronize()             # last step's copy out of this pinned buffer is done
            st['mask_host'][phase].copy_((torch.arange(L) < tt).float())
            st['mask'].copy_(st['mask_host'][phase], non_blocking=True)
            st['mask_event'][phase] = torch.cuda.Event()
            st['mask_event'][phase].record()
            self._to_static('hists', batch['histograms'])
            if phase == 0:
                self._to_static('images', batch['images'])

        stage(next(self.loader), 0)
        if self._param_lists is None:
            self._param_lists = (list(GAN.D.parameters()),
                                 [p for grp in GAN.G_opt.param_groups for p in grp['params']])
        d_params, g_params = self._param_lists
        overlapped = _GradOverlap([]).enabled       # the collectives then live inside the graphs
        divergence, gp = self._graphed(('D', apply_gp), lambda: self._phase_d(apply_gp), d_params)
        split_g = self.split_g_phase
        if split_g is None:                 # HG_SPLIT_G=1 / 0 forces it; default: under DDP
            env = os.environ.get('HG_SPLIT_G')
            split_g = _ddp_active() if env is None else env != '0'
        if split_g:
            # the generator side of the G phase (G1) does not depend on D: it runs while the D-side
            # gradient all-reduce is in flight; D's update and the rest of the phase (G2) follow
            k1, k2 = ('G1', bool(apply_pl)), ('G2', float(alpha), bool(apply_pl))
            batch_g = next(self.loader)
            work = None if overlapped else self._exchange_async('d', d_params)
            stage(batch_g, 1)
            if apply_pl:
                st['pl_mean'].fill_(float(self.pl_mean))
            if k2 not in self._graphs:
                self._capture([k1, k2], [lambda: self._phase_g1(apply_pl),
                                         lambda: self._phase_g2(alpha, apply_pl)], [[], g_params])
            self._replay(k1)
            if work is not None:
                work()
            GAN.D_opt.step()
            g_loss, h_loss, avg_pl = self._replay(k2)
        else:
            if overlapped:
                GAN.D_opt.step()
            else:
                self._exchange_and_step('d', d_params, GAN.D_opt)
            stage(next(self.loader), 1)
            if apply_pl:
                st['pl_mean'].fill_(float(self.pl_mean))
            g_loss, h_loss, avg_pl = self._graphed(('G', float(alpha), bool(apply_pl)),
                                                   lambda: self._phase_g(alpha, apply_pl), g_params)
        if overlapped:
            GAN.G_opt.step()
        else:
            self._exchange_and_step('g', g_params, GAN.G_opt)
        # one stacked device-to-host copy of the read-outs, after everything has been queued
        self.q_loss = 0.0
        named = {'d_loss': divergence, 'g_loss': g_loss, 'h_loss': h_loss}
        if gp is not None:
            named['last_gp_loss'] = gp
        if avg_pl is not None:
            named['avg_pl'] = avg_pl
        previous, self._pending = self._pending, _PendingScalars(named)
        return previous, self._pending

    # ------------------------------------------------------------ evaluate --
    @torch.no_grad()
    def evaluate(self, num=0, hist_batch=None, num_image_tiles=4, latents=None, n=None,
                 save_noise_latent=False, load_noise_file=None, load_latent_file=None):
        self.GAN.eval()
        if hist_batch is None:
            hist_batch = next(self.loader_evaluate)['histograms'].cuda()
        ext = 'jpg' if not self.transparent else 'png'
        num_rows = num_image_tiles
        if latents is None and n is None:
            G = self.GAN.G
            n = torch.tensor(np.load(load_noise_file)).cuda() if load_noise_file is not None else \
                image_noise(num_rows ** 2, G.image_size)
            latents = np.load(load_latent_file) if load_latent_file is not None else \
                noise_list(num_rows ** 2, G.num_layers - 2, G.latent_dim)
        generated_images = self.generate_truncated(self.GAN.SE, self.GAN.HE, self.GAN.GE,
                                                   hist_batch, latents, n, trunc_psi=self.trunc_psi)
        if num is not None:
            import torchvision
            torchvision.utils.save_image(generated_images.contiguous(),
                                         str(self.results_dir / self.name / f'{str(num)}-ema.{ext}'),
                                         nrow=num_rows)
        if save_noise_latent:
            Path(f'temp/{self.name}').mkdir(parents=True, exist_ok=True)
            np.save(f'temp/{self.name}/{str(num)}-noise.npy', n.clone().cpu().numpy())
            np.save(f'temp/{self.name}/{str(num)}-latents.npy', latents)
        return generated_images

    @torch.no_grad()
    def generate_truncated(self, S, H, G, hist_batch, style, noi, trunc_psi=0.75):
        latent_dim = G.latent_dim
        if self.av is None:
            z = noise(2000, latent_dim)
            samples = evaluate_in_chunks(self.batch_size, S, z).cpu().numpy()
            self.av = np.expand_dims(np.mean(samples, axis=0), axis=0)
        av_torch = torch.from_numpy(self.av).cuda()
        w_space = [(trunc_psi * (S(t) - av_torch) + av_torch, nl) for t, nl in style]
        h_w = H(hist_batch).unsqueeze(1)
        h_w = torch.cat((h_w, h_w), dim=1)
        for _ in range(int(np.log2(np.sqrt(w_space[0][0].shape[0])))):
            h_w = torch.cat((h_w, h_w), dim=0)
        w_styles = styles_def_to_tensor(w_space)
        generated_images = evaluate_in_chunks(self.batch_size, G, w_styles, h_w, noi)
        return generated_images.clamp_(0.0, 1.0)

    def print_log(self):
        h = f' | H: {self.h_loss:.2f}' if hasattr(self, 'h_loss') else ''
        print(f'\nG: {self.g_loss:.2f}{h} | D: {self.d_loss:.2f} | GP: {self.last_gp_loss:.2f}'
              f' | PL: {self.pl_mean:.2f} | CR: {self.last_cr_loss:.2f} | Q: {self.q_loss:.2f}')

    # --------------------------------------------------------- checkpoints --
    def model_name(self, num):
        return str(self.models_dir / self.name / f'model_{num}.pt')

    def init_folders(self):
        (self.results_dir / self.name).mkdir(parents=True, exist_ok=True)
        (self.models_dir / self.name).mkdir(parents=True, exist_ok=True)

    def clear(self):
        rmtree(f'./models/{self.name}', True)
        rmtree(f'./results/{self.name}', True)
        rmtree(str(self.config_path), True)
        self.init_folders()

    def save(self, num):
        torch.save(self.GAN.state_dict(), self.model_name(num))
        self.write_config()

    def load(self, num=-1):
        self.load_config()
        name = num
        if num == -1:
            saved = sorted(int(p.stem.split('_')[1])
                           for p in Path(self.models_dir / self.name).glob('model_*.pt'))
            if len(saved) == 0:
                return
            name = saved[-1]
            print(f'continuing from previous epoch - {name}')
        self.steps = name * self.save_every
        self.GAN.load_state_dict(torch.load(self.model_name(name),
                                            map_location=f'cuda:{torch.cuda.current_device()}'))


class SyntheticLoader:
    """Infinite iterator of {'images', 'histograms'} batches resident on the device
    (benchmarks / smoke tests; replaces Trainer.set_data_src)."""

    def __init__(self, batch_size, image_size, hist_bin=64, seed=0, device='cuda', eval_batch=None):
        g = torch.Generator().manual_seed(seed)
        b = eval_batch or batch_size
        self.images = torch.rand(b, 3, image_size, image_size, generator=g).to(device)
        t = torch.rand(b, 3, hist_bin, hist_bin, generator=g)
        self.hists = (t / t.sum(dim=(1, 2, 3), keepdim=True)).to(device)
        self.eval_only = eval_batch is not None

    def __iter__(self):
        return self

    def __next__(self):
        if self.eval_only:
            return {'histograms': self.hists}
        return {'images': self.images, 'histograms': self.hists}
