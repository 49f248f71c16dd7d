"""FusedAdam: torch.optim.Adam's constructor and parameter-group semantics on one multi-tensor kernel (csrc/adam.cu).

The reference builds `self.optim_cls(params, **self.optim_params)` over six named parameter groups and adds the camera
extrinsics as a seventh (pc_nerf/trainer.py:229-300); with `optimizer_type: adam` (configs/bup20/best.yaml:114) that class is
torch.optim.Adam.  This one is a drop-in for it on CUDA fp32 parameters: same defaults (lr 1e-3, betas (0.9, 0.999), eps 1e-8,
weight_decay 0), same update rule (amsgrad off), per-group lr / weight_decay, `state_dict()` with `exp_avg` / `exp_avg_sq` /
`step`.  The step counter lives in device memory and one launch covers every tensor, so `step()` can be captured into the same
CUDA graph as the forward / backward (graph.GraphedStep) -- the reference's per-tensor Adam costs ~25 launches per step.
"""
import ctypes

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("csrc/adam.cu implements Adam without amsgrad (the reference never enables it)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step_dev = {}
        self.inv_scale = None      # optional device float: gradients are multiplied by it (fused unscale)

    def _state(self, p):
        st = self.state[p]
        if not st:
            st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['step'] = torch.zeros((), dtype=torch.int32, device=p.device)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        # one launch per distinct (betas, eps) -- a single one for the reference's groups, which differ in lr / weight_decay only
        buckets = {}
        for group in self.param_groups:
            key = (tuple(group['betas']), float(group['eps']))
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: CUDA fp32 contiguous parameters only (no CPU fallback)")
                buckets.setdefault((key, p.device), []).append((p, float(group['lr']), float(group['weight_decay'])))
        for ((betas, eps), dev), items in buckets.items():
            skey = (dev, betas, eps)      # one device-side step counter per launch bucket
            step = self._step_dev.get(skey)
            if step is None:
                step = self._step_dev[skey] = torch.zeros(1, dtype=torch.int32, device=dev)
            for c0 in range(0, len(items), 48):
                chunk = items[c0:c0 + 48]
                n = len(chunk)
                sts = [self._state(p) for p, _, _ in chunk]
                g = [p.grad.contiguous() for p, _, _ in chunk]
                vp = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
                numel = (ctypes.c_int64 * n)(*[p.numel() for p, _, _ in chunk])
                lr = (ctypes.c_float * n)(*[l for _, l, _ in chunk])
                wd = (ctypes.c_float * n)(*[w for _, _, w in chunk])
                if c0 == 0:
                    tick = step
                else:      # later chunks of the same bucket reuse the already advanced count: give them a scratch counter one behind
                    tick = (step - 1).contiguous()
                _lib.call("pag_adam_step", vp([p for p, _, _ in chunk]), vp(g), vp([s['exp_avg'] for s in sts]),
                          vp([s['exp_avg_sq'] for s in sts]), numel, lr, wd, n, float(betas[0]), float(betas[1]), float(eps),
                          _lib.ptr(tick), _lib.ptr(self.inv_scale))
                for s in sts:
                    s['step'] = step[0]
        return loss
