"""Oracle: the reference's autocast (fp16) training numerics restated on the CPU.  TEST INFRASTRUCTURE ONLY.

The reference trains under `torch.cuda.amp.autocast()` + GradScaler (pc_nerf/trainer.py:429, :582).  For the hot path
that means (torch autocast op lists + the reference's own call sites):
  * `PermutoGrid.interpolate` is `custom_fwd(cast_inputs=torch.half)` (grids/permuto_grid.py:65): the sample coordinates
    are rounded to fp16, the encoding itself runs in fp32 on `.type(torch.float)` coordinates (:71) -> fp32 features;
  * every `nn.Linear` of the four BasicDecoders casts input, weight and bias to fp16, accumulates in fp32 (cuBLAS) and
    rounds the result to fp16; ReLU / sigmoid keep fp16; `torch.cat` promotes to the widest input; `softmax` runs in fp32;
  * the backward mirrors it: the gradient w.r.t. a Linear's fp16 output is fp16, dX and dW are fp32-accumulated GEMMs
    rounded to fp16, the weight-cast's backward widens dW to fp32; GradScaler multiplies the loss by 2^16 first.
This module provides that Linear as an explicit autograd function so that tests can MEASURE how far the reference's own
autocast step is from exact fp32 arithmetic, and hold the tensor-core kernels to the same yardstick.
"""
import contextlib

import torch

_STATE = {"enabled": False}


def enabled():
    return _STATE["enabled"]


@contextlib.contextmanager
def emulate_fp16(on=True):
    prev = _STATE["enabled"]
    _STATE["enabled"] = bool(on)
    try:
        yield
    finally:
        _STATE["enabled"] = prev


class _Linear16(torch.autograd.Function):
    """y16 = round16(x16 @ W16^T + b16) with fp32 accumulation; backward likewise (what cuBLAS hgemm with fp32 compute does)."""

    @staticmethod
    def forward(ctx, x16, w16, b16):
        ctx.save_for_backward(x16, w16)
        y = x16.float() @ w16.float().t()
        if b16 is not None:
            y = y + b16.float()
        return y.half()

    @staticmethod
    def backward(ctx, g16):
        x16, w16 = ctx.saved_tensors
        g = g16.float()
        gx = (g @ w16.float()).half()
        gw = (g.t() @ x16.float()).half()
        gb = g.sum(0).half()
        return gx, gw, gb


def linear(x, weight, bias):
    """nn.Linear under CUDA autocast(fp16): casts are autograd ops (their backward widens / narrows the gradient)."""
    return _Linear16.apply(x.half(), weight.half(), bias.half() if bias is not None else None)
