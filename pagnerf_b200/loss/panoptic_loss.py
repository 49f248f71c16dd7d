"""Fused photometric + panoptic NLL loss: w_rgb * L1(rgb) + w_sem * NLL(log(sem + eps)) + w_inst * NLL(log(inst + eps)) with mean
reduction, the combination the reference's step forms from torch ops (pc_nerf/trainer.py:442-480; `torch.log(x + 1e-27)` :459) --
one launch forward, one backward (csrc/loss.cu) instead of ~20 small torch kernels between the trace's forward and backward."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .._lib import call, ptr

_SCRATCH = {}


class _PanopticLossFn(Function):
    @staticmethod
    def forward(ctx, rgb, sem, inst, t_rgb, t_sem, t_inst, w_rgb, w_sem, w_inst, eps):
        ref = next(t for t in (rgb, sem, inst) if t is not None)
        if not ref.is_cuda:
            raise RuntimeError("pagnerf_b200 losses run on CUDA tensors only (no CPU fallback)")
        f = lambda t: t.detach().to(torch.float32).contiguous() if t is not None else None
        r, s, i = f(rgb), f(sem), f(inst)
        N, dev = ref.shape[0], ref.device
        tr = t_rgb.to(torch.float32).contiguous() if r is not None else None
        ts = t_sem.to(torch.int64).contiguous() if s is not None else None
        ti = t_inst.to(torch.int64).contiguous() if i is not None else None
        key = (str(dev), torch.cuda.current_stream().cuda_stream)
        sc = _SCRATCH.get(key)
        if sc is None:      # partial sums + a self-resetting ticket, per (device, stream)
            sc = _SCRATCH[key] = (torch.empty(1024, dtype=torch.float32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
        loss = torch.empty((), dtype=torch.float32, device=dev)
        Cs = s.shape[1] if s is not None else 0
        Ci = i.shape[1] if i is not None else 0
        call("pag_panoptic_loss_fwd", ptr(r), ptr(s), ptr(i), ptr(tr), ptr(ts), ptr(ti), N, Cs, Ci, float(w_rgb), float(w_sem), float(w_inst),
             float(eps), ptr(sc[0]), ptr(sc[1]), ptr(loss))
        ctx.save_for_backward(r, s, i, tr, ts, ti)
        ctx.cfg = (N, Cs, Ci, float(w_rgb), float(w_sem), float(w_inst), float(eps))
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        r, s, i, tr, ts, ti = ctx.saved_tensors
        N, Cs, Ci, w_rgb, w_sem, w_inst, eps = ctx.cfg
        gl = g.detach().to(torch.float32).reshape(1).contiguous()
        need = ctx.needs_input_grad
        g_rgb = torch.empty_like(r) if (r is not None and need[0]) else None
        g_sem = torch.empty_like(s) if (s is not None and need[1]) else None
        g_inst = torch.empty_like(i) if (i is not None and need[2]) else None
        call("pag_panoptic_loss_bwd", ptr(r), ptr(s), ptr(i), ptr(tr), ptr(ts), ptr(ti), N, Cs, Ci, w_rgb, w_sem, w_inst, eps, ptr(gl),
             ptr(g_rgb), ptr(g_sem), ptr(g_inst))
        return g_rgb, g_sem, g_inst, None, None, None, None, None, None, None


def panoptic_loss(rgb, sem, inst, t_rgb, t_sem, t_inst, w_rgb=1.0, w_sem=1.0, w_inst=1.0, eps=1e-27):
    """Scalar loss; any of (rgb, sem, inst) may be None."""
    return _PanopticLossFn.apply(rgb, sem, inst, t_rgb, t_sem, t_inst, w_rgb, w_sem, w_inst, eps)
