"""DiffGrad optimiser (Dubey et al., "diffGrad: An Optimization Method for
Convolutional Neural Networks", 2019) as used by the reference through
``torch_optimizer.DiffGrad`` (histoGAN/histoGAN.py:28,670-671).

PARITY UNPINNED: torch-optimizer is not part of the reference tree (README.md:44
lists it unpinned) and is not installed here, so this is a restatement of the
published update as torch-optimizer 0.3.0 implements it:

    m_t = b1 m + (1-b1) g ;  v_t = b2 v + (1-b2) g^2
    xi  = sigmoid(|g_{t-1} - g_t|)                       ("friction" coefficient)
    p  -= lr * sqrt(1-b2^t)/(1-b1^t) * (m_t * xi) / (sqrt(v_t) + eps)

On CUDA the whole update is ONE fused multi-tensor kernel (hg_diffgrad_step, csrc/optim.cu:
a single pass over p, g, m, v, g_prev); on other devices (tests) the same arithmetic runs
through torch._foreach ops.
"""
from __future__ import annotations

import math

import torch
from torch.optim import Optimizer


def _dense(t):
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def _same_layout(a, b):
    return a.shape == b.shape and (a.stride() == b.stride() or (a.is_contiguous() and b.is_contiguous()))


class DiffGrad(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr <= 0.0:
            raise ValueError(f'Invalid learning rate: {lr}')
        if eps < 0.0:
            raise ValueError(f'Invalid epsilon value: {eps}')
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f'Invalid betas: {betas}')
        if weight_decay < 0.0:
            raise ValueError(f'Invalid weight_decay value: {weight_decay}')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, only=None):
        """`only`: update just these parameters (the trainer pipelines the update of one piece of the
        gradient arena with the all-reduce of the next); each parameter keeps its own step count."""
        only = None if only is None else {id(p) for p in only}
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            by_step = {}
            for p in group['params']:
                if p.grad is None or (only is not None and id(p) not in only):
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('DiffGrad does not support sparse gradients')
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['previous_grad'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['step'] += 1
                by_step.setdefault(st['step'], []).append(p)
            for step, ps in by_step.items():
                step_size = group['lr'] * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
                # the fused kernel walks raw memory: parameter, gradient and state must share one
                # dense layout (contiguous, or channels_last for the conv weights)
                if all(p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32
                       and _dense(p) and _same_layout(p, p.grad) for p in ps):
                    self._fused_step(ps, beta1, beta2, group['eps'], step_size, group['weight_decay'])
                    continue
                grads = [p.grad for p in ps]
                m = [self.state[p]['exp_avg'] for p in ps]
                v = [self.state[p]['exp_avg_sq'] for p in ps]
                prev = [self.state[p]['previous_grad'] for p in ps]
                if group['weight_decay'] != 0:
                    grads = torch._foreach_add(grads, ps, alpha=group['weight_decay'])
                torch._foreach_mul_(m, beta1)
                torch._foreach_add_(m, grads, alpha=1 - beta1)
                torch._foreach_mul_(v, beta2)
                torch._foreach_addcmul_(v, grads, grads, value=1 - beta2)
                denom = torch._foreach_sqrt(v)
                torch._foreach_add_(denom, group['eps'])
                # friction coefficient from the change of the gradient
                dfc = torch._foreach_sub(prev, grads)
                torch._foreach_abs_(dfc)
                torch._foreach_sigmoid_(dfc)
                torch._foreach_copy_(prev, grads)
                torch._foreach_mul_(dfc, m)
                torch._foreach_addcdiv_(ps, dfc, denom, value=-step_size)
        return loss

    def _fused_step(self, ps, beta1, beta2, eps, step_size, weight_decay):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        n = len(ps)
        arr = C.c_void_p * n

        def table(ts):
            return arr(*[t.data_ptr() for t in ts])

        st = [self.state[p] for p in ps]
        # the kernel indexes p, g and the three state tensors with ONE linear offset
        assert all(_same_layout(p, s[k]) for p, s in zip(ps, st)
                   for k in ('exp_avg', 'exp_avg_sq', 'previous_grad')), "optimizer state layout"
        numel = (C.c_int64 * n)(*[p.numel() for p in ps])
        dev = ps[0].device
        # conv weights whose tensor-core forward operand is a plain TF32-rounded copy: the kernel
        # writes that copy too (ops._PackCache keeps its address stable)
        from . import ops
        targets = [ops._packs.fused_forward_target(p) for p in ps]
        packed = arr(*[t.data_ptr() if t is not None else None for t in targets])
        with torch.cuda.device(dev):
            rc = lib.hg_diffgrad_step(n, table(ps), table([p.grad for p in ps]),
                                      table([s['exp_avg'] for s in st]),
                                      table([s['exp_avg_sq'] for s in st]),
                                      table([s['previous_grad'] for s in st]), packed, numel,
                                      beta1, beta2, eps, step_size, weight_decay,
                                      _lib.current_stream_ptr(dev))
        _lib.check(rc, "hg_diffgrad_step")
        # the kernel wrote through raw pointers: tell autograd the parameters changed in place
        # (saved-tensor checks, and the packed-weight caches of ops.py key on ``_version``)
        torch.autograd.graph.increment_version(ps)
        for p, t in zip(ps, targets):        # the remaining packed forms (dgrad transposes), in place
            ops._packs.refresh(p, done=(0,) if t is not None else ())
