"""wisp.core.Rays / wisp.core.RenderBuffer stand-ins (containers only).

Used by the reference at pc_nerf/ba_pipeline.py:91, pc_nerf/trainer.py:438-443,488,643-648,706-717
and tracers/panoptic_packed_rf_tracer.py:195.
"""
from dataclasses import dataclass, fields
from typing import Optional

import torch


@dataclass
class Rays:
    origins: torch.Tensor
    dirs: torch.Tensor
    dist_min: float = 0.0
    dist_max: float = 6.0

    def __len__(self):
        return self.origins.shape[0]

    @property
    def shape(self):
        return self.origins.shape[:-1]

    def _map(self, fn):
        return Rays(origins=fn(self.origins), dirs=fn(self.dirs), dist_min=self.dist_min, dist_max=self.dist_max)

    def reshape(self, *dims):
        return self._map(lambda t: t.reshape(*dims))

    def squeeze(self, dim):
        return self._map(lambda t: t.squeeze(dim))

    def contiguous(self):
        return self._map(lambda t: t.contiguous())

    def to(self, *args, **kwargs):
        return self._map(lambda t: t.to(*args, **kwargs))

    def __getitem__(self, idx):
        return self._map(lambda t: t[idx])

    def split(self, split_size):
        return [Rays(origins=o, dirs=d, dist_min=self.dist_min, dist_max=self.dist_max)
                for o, d in zip(self.origins.split(split_size), self.dirs.split(split_size))]

    @classmethod
    def cat(cls, rays_list, dim=0):
        return cls(origins=torch.cat([r.origins for r in rays_list], dim), dirs=torch.cat([r.dirs for r in rays_list], dim),
                   dist_min=min(r.dist_min for r in rays_list), dist_max=max(r.dist_max for r in rays_list))


class RenderBuffer:
    """Named per-ray channels; `a + b` concatenates along the ray dimension (pc_nerf/trainer.py:648)."""

    def __init__(self, rgb: Optional[torch.Tensor] = None, alpha: Optional[torch.Tensor] = None,
                 depth: Optional[torch.Tensor] = None, **extra):
        self._channels = {}
        for k, v in dict(rgb=rgb, alpha=alpha, depth=depth, **extra).items():
            if v is not None:
                self._channels[k] = v

    def __getattr__(self, name):
        ch = self.__dict__.get('_channels', {})
        if name in ch:
            return ch[name]
        if name in ('rgb', 'alpha', 'depth'):
            return None
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name == '_channels':
            object.__setattr__(self, name, value)
        else:
            self._channels[name] = value

    @property
    def channels(self):
        return set(self._channels.keys())

    def get_channel(self, name):
        return self._channels[name]

    def _map(self, fn):
        return RenderBuffer(**{k: (fn(v) if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in self._channels.items()})

    def __add__(self, other):
        if other is None:
            return self
        out = {}
        for k in self._channels.keys() | other._channels.keys():
            a, b = self._channels.get(k), other._channels.get(k)
            if a is None or b is None:
                out[k] = a if b is None else b
            elif a.dim() == 0:
                out[k] = a + b
            else:
                out[k] = torch.cat([a, b], dim=0)
        return RenderBuffer(**out)

    def __radd__(self, other):
        return self if other in (None, 0) else other.__add__(self)

    def reshape(self, *dims):
        return self._map(lambda t: t.reshape(*dims))

    def cpu(self):
        return self._map(lambda t: t.cpu())

    def cuda(self):
        return self._map(lambda t: t.cuda())

    def detach(self):
        return self._map(lambda t: t.detach())

    def to(self, *a, **k):
        return self._map(lambda t: t.to(*a, **k))

    def byte(self):
        return self._map(lambda t: (t * 255.0).to(torch.uint8) if t.is_floating_point() else t)

    def image(self):
        return self
