"""wisp.tracers.BaseTracer / PackedRFTracer stand-ins (argument plumbing only).

`BaseTracer.forward` fills `trace()`'s optional arguments from the call kwargs, else from
same-named tracer attributes (the reference relies on that for raymarch_type / num_steps, which the
trainer mutates at pc_nerf/trainer.py:364-366).
"""
import inspect

import torch.nn as nn


class BaseTracer(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()

    def get_supported_channels(self):
        raise NotImplementedError

    def get_required_nef_channels(self):
        raise NotImplementedError

    def trace(self, nef, channels, extra_channels, *args, **kwargs):
        raise NotImplementedError

    def forward(self, nef, channels=None, **kwargs):
        nef_channels = nef.get_supported_channels()
        missing = self.get_required_nef_channels() - nef_channels
        if missing:
            raise Exception(f"The neural field class {type(nef)} does not output the required channels {missing}.")
        if channels is None:
            requested = self.get_supported_channels()
        elif isinstance(channels, str):
            requested = {channels}
        else:
            requested = set(channels)
        extra = requested - self.get_supported_channels()
        unsupported = extra - nef_channels
        if unsupported:
            raise Exception(f"Channels {unsupported} are not supported in the tracer {type(self)} or neural field {type(nef)}.")

        argspec = inspect.getfullargspec(self.trace)
        ndef = len(argspec.defaults) if argspec.defaults else 0
        required = argspec.args[:len(argspec.args) - ndef][4:]  # skip self, nef, channels, extra_channels
        optional = argspec.args[len(argspec.args) - ndef:]
        input_args = {}
        for a in required:
            if a not in kwargs:
                raise Exception(f"Argument {a} not found as input to in {type(self)}.trace()")
            input_args[a] = kwargs[a]
        for a in optional:
            if a in kwargs:
                input_args[a] = kwargs[a]
            else:
                default = getattr(self, a, None)
                if default is not None:
                    input_args[a] = default
        return self.trace(nef, requested, extra, **input_args)


class PackedRFTracer(BaseTracer):
    def __init__(self, raymarch_type='voxel', num_steps=64, step_size=1.0, bg_color='white', **kwargs):
        super().__init__(**kwargs)
        self.raymarch_type = raymarch_type
        self.num_steps = num_steps
        self.step_size = step_size
        self.bg_color = bg_color
