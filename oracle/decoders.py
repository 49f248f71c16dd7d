"""Oracle: wisp BasicDecoder + PositionalEmbedder restated in plain torch.  TEST INFRASTRUCTURE ONLY.

kaolin-wisp v0.1.1 (README.md:42) is absent; behaviour as recalled in SURVEY.md Appendix A.8
(parity UNPINNED for the wisp classes; the SHAPES are pinned by the reference's call sites
pc_nerf/panoptic_nef.py:75,114-164 and configs/bup20/best.yaml:66-97).
"""
import torch
from torch import nn

from . import autocast


class BasicDecoderOracle(nn.Module):
    """num_layers hidden Linear+activation (first in->hidden; ReLU, or none for wisp's 'none' = Identity), then `lout`
    Linear(hidden,out) w/o activation."""

    def __init__(self, input_dim, output_dim, num_layers=1, hidden_dim=64, bias=True, activation='relu'):
        super().__init__()
        self.activation = activation
        dims = [input_dim] + [hidden_dim] * num_layers
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1], bias=bias) for i in range(num_layers)])
        self.lout = nn.Linear(hidden_dim, output_dim, bias=bias)

    def forward(self, x):
        h = x
        if autocast.enabled():      # the reference's fp16 autocast numerics (oracle/autocast.py); output fp16
            for l in self.layers:
                h = autocast.linear(h, l.weight, l.bias)
                h = torch.relu(h) if self.activation == 'relu' else h
            return autocast.linear(h, self.lout.weight, self.lout.bias)
        for l in self.layers:
            h = l(h)
            h = torch.relu(h) if self.activation == 'relu' else h
        return self.lout(h)


class PositionalEmbedderOracle(nn.Module):
    """[x, sin(b_k x) (k-major, xyz fastest), cos(b_k x)], bands 2^linspace(0, f-1, f)."""

    def __init__(self, num_freq, input_dim=3):
        super().__init__()
        self.num_freq = num_freq
        self.register_buffer("bands", 2.0 ** torch.linspace(0.0, num_freq - 1, steps=num_freq))
        self.out_dim = input_dim + 2 * num_freq * input_dim

    def forward(self, coords):
        N = coords.shape[0]
        winded = (coords[:, None] * self.bands[None, :, None].to(coords.dtype)).reshape(N, -1)
        return torch.cat([coords, torch.sin(winded), torch.cos(winded)], dim=-1)
