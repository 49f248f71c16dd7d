"""Occtree: feature-less occupancy-octree grid (reference grids/occtree.py:30-91)."""
import torch

from .base import BLASGrid


class Occtree(BLASGrid):
    def __init__(self, blas_level: int = 7, **kwargs):
        super().__init__()
        self.kwargs = kwargs
        self._init_blas(blas_level)
        self.num_lods = 1
        self.active_lods = [0]
        self._register_blas_buffers()

    def freeze(self):
        return
