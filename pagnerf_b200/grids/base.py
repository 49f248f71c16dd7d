"""Common base of the grid plugins: the wisp `BLASGrid` / `HashGrid` constructor + raymarch contract
the reference's grids inherit (reference grids/occtree.py:30-91,
grids/permuto_grid.py:13-45).  Owns the occupancy-octree accel-struct and its checkpoint buffers."""
import numpy as np
import torch
import torch.nn as nn

from .. import spc


class BLASGrid(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def _init_blas(self, blas_level):
        self.blas_level = blas_level
        self.blas = spc.OctreeAS()
        self.blas.init_dense(self.blas_level)
        self.dense_points = spc.unbatched_get_level_points(self.blas.points, self.blas.pyramid, self.blas_level).clone()
        self.num_cells = self.dense_points.shape[0]
        self.occupancy = torch.zeros(self.num_cells)

    def _register_blas_buffers(self):
        # same state_dict keys as the reference (grids/occtree.py:69-74, grids/permuto_grid.py:33-38)
        for name, t in (('blas_octree', self.blas.octree), ('blas_points', self.blas.points),
                        ('blas_prefix', self.blas.prefix), ('blas_pyramid', self.blas.pyramid)):
            if name in self._buffers:
                self._buffers[name] = t
            else:
                self.register_buffer(name, t)

    def blas_init(self, octree):
        self.blas.init(octree)
        self._register_blas_buffers()

    def blas_init_from_mask(self, mask, register=True):
        """prune(): rebuild the accel-struct from the dense occupancy mask (Morton order = `dense_points` order) with the
        device-side builder; `register` re-registers the checkpoint buffers (PermutoGrid) or leaves the state_dict keys alone."""
        self.blas.init_from_mask(mask, self.blas_level)
        if register:
            self._register_blas_buffers()

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # octree size changes with pruning: adopt the checkpoint's accel-struct before the strict shape check
        key = prefix + 'blas_octree'
        if key in state_dict and hasattr(self, 'blas'):
            self.blas_init(state_dict[key])
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    # interpolate() of the grids built on this base looks features up by position only: tracers may ask the marcher to skip the
    # octree point indices (need_pidx=False)
    interpolate_needs_pidx = False

    def raymarch(self, rays, level=None, num_samples=64, raymarch_type='voxel', need_pidx=True):
        """Wrapper over OctreeAS.raymarch at blas_level (reference grids/occtree.py:85-91)."""
        return self.blas.raymarch(rays, level=self.blas_level, num_samples=num_samples, raymarch_type=raymarch_type,
                                  need_pidx=need_pidx)

    def raytrace(self, rays, level=None, with_exit=False):
        return self.blas.raytrace(rays, level=self.blas_level, with_exit=with_exit)


class HashGridBase(BLASGrid):
    """wisp.models.grids.HashGrid constructor contract (feature-less: the tables live in the subclasses)."""

    def __init__(self, feature_dim, interpolation_type='linear', multiscale_type='cat', feature_std=0.0,
                 feature_bias=0.0, codebook_bitwidth=8, blas_level=7, **kwargs):
        super().__init__()
        self.feature_dim = feature_dim
        self.interpolation_type = interpolation_type
        self.multiscale_type = multiscale_type
        self.feature_std = feature_std
        self.feature_bias = feature_bias
        self.codebook_bitwidth = codebook_bitwidth
        self.kwargs = kwargs
        self._init_blas(blas_level)

    def init_from_octree(self, base_lod, num_lods):
        self.init_from_resolutions([2 ** L for L in range(base_lod, base_lod + num_lods)])

    def init_from_geometric(self, min_width, max_width, num_lods):
        b = np.exp((np.log(max_width) - np.log(min_width)) / (num_lods - 1))
        self.init_from_resolutions([int(np.floor(min_width * (b ** l))) for l in range(num_lods)])

    def init_from_resolutions(self, resolutions):
        raise NotImplementedError
