"""wisp BasicDecoder / PositionalEmbedder / small helpers -- parameter CONTAINERS.

These modules own the decoder weights under the reference's parameter names
(`decoder_density.layers.0.weight`, `decoder_density.lout.bias`, ... -- the trainer groups
optimiser parameters by those names, pc_nerf/trainer.py:240-258; pc_nerf/panoptic_nef.py:123 pokes
`lout.bias`).  Their `forward` is NOT a product compute path: the hot path reads the weights
directly in csrc/decoder.cu.  Calling `forward` raises on CUDA tensors to make an accidental
library (cuBLAS) fallback loud; on CPU tensors it is allowed for host-side unit tests.
"""
import torch
import torch.nn as nn


class BasicDecoder(nn.Module):
    def __init__(self, input_dim, output_dim, activation=torch.relu, bias=True, layer=nn.Linear,
                 num_layers=1, hidden_dim=128, skip=None):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.activation, self.bias, self.layer = activation, bias, layer
        self.num_layers, self.hidden_dim, self.skip = num_layers, hidden_dim, list(skip or [])
        layers = []
        for i in range(num_layers):
            if i == 0:
                layers.append(layer(input_dim, hidden_dim, bias=bias))
            elif i in self.skip:
                layers.append(layer(hidden_dim + input_dim, hidden_dim, bias=bias))
            else:
                layers.append(layer(hidden_dim, hidden_dim, bias=bias))
        self.layers = nn.ModuleList(layers)
        self.lout = layer(hidden_dim, output_dim, bias=bias)

    def forward(self, x, return_h=False):
        if x.is_cuda:
            raise RuntimeError("BasicDecoder.forward is a host-side container path; CUDA tensors must go "
                               "through pagnerf_b200.ops.decode (csrc/decoder.cu)")
        h = x
        for i, l in enumerate(self.layers):
            if i == 0:
                h = self.activation(l(x))
            elif i in self.skip:
                h = torch.cat([x, self.activation(l(h))], dim=-1)
            else:
                h = self.activation(l(h))
        out = self.lout(h)
        return (out, h) if return_h else out


class PositionalEmbedder(nn.Module):
    def __init__(self, num_freq, max_freq_log2, log_sampling=True, include_input=True, input_dim=3):
        super().__init__()
        self.num_freq, self.include_input, self.input_dim = num_freq, include_input, input_dim
        if log_sampling:
            bands = 2.0 ** torch.linspace(0.0, max_freq_log2, steps=num_freq)
        else:
            bands = torch.linspace(1, 2.0 ** max_freq_log2, steps=num_freq)
        self.register_buffer('bands', bands, persistent=False)
        self.out_dim = (input_dim if include_input else 0) + num_freq * input_dim * 2

    def forward(self, coords):
        N = coords.shape[0]
        winded = (coords[:, None] * self.bands[None, :, None].to(coords)).reshape(N, -1)
        enc = torch.cat([torch.sin(winded), torch.cos(winded)], dim=-1)
        return torch.cat([coords, enc], dim=-1) if self.include_input else enc


def get_positional_embedder(frequencies, active, input_dim=3):
    if not active:
        return nn.Identity(), input_dim
    enc = PositionalEmbedder(frequencies, frequencies - 1, input_dim=input_dim)
    return enc, enc.out_dim


def _identity(x):
    return x


def get_activation_class(activation_type):
    if activation_type == 'none':      # wisp: Identity() -- PanopticDDensityNeF's delta-density head (pc_nerf/panoptic_dd_nef.py:52)
        return _identity
    if activation_type == 'relu':
        return torch.relu
    if activation_type == 'sin':
        return torch.sin
    raise NotImplementedError(f"activation {activation_type}")


def get_layer_class(layer_type):
    if layer_type in ('none', 'linear', None):
        return nn.Linear
    raise NotImplementedError(f"layer {layer_type}")


class PerfTimer:
    def __init__(self, activate=False, show_memory=False, print_mode=True):
        self.activate = activate

    def reset(self):
        return

    def check(self, name=None):
        return


class Pipeline(nn.Module):
    def __init__(self, nef, tracer=None):
        super().__init__()
        self.nef = nef
        self.tracer = tracer

    def forward(self, *args, **kwargs):
        if self.tracer is not None:
            return self.tracer(self.nef, *args, **kwargs)
        return self.nef(*args, **kwargs)
