"""PermutoGrid: multi-resolution permutohedral-lattice feature grid (reference grids/permuto_grid.py:13-71).

`PermutoEncoding` is the parameter container + launcher for csrc/permuto.cu with the constructor of
R. A. Rosu's permutohedral_encoding.PermutoEncoding (pos_dim, capacity, nr_levels,
nr_feat_per_level, scale_per_level) and the same state_dict keys
(`lattice_values`, `random_shift_per_level`, `scale_factor`, `anneal_window`)."""
import logging as log

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .base import HashGridBase


class PermutoEncoding(nn.Module):
    def __init__(self, pos_dim, capacity, nr_levels, nr_feat_per_level, scale_per_level,
                 apply_random_shift_per_level=True, agg_scale_threshold=0.05):
        super().__init__()
        if pos_dim != 3 or nr_feat_per_level != 2:
            raise NotImplementedError("csrc/permuto.cu is specialised to pos_dim=3, 2 features per level "
                                      "(configs/bup20/*.yaml); got pos_dim=%d F=%d" % (pos_dim, nr_feat_per_level))
        self.pos_dim, self.capacity, self.nr_levels, self.nr_feat = pos_dim, int(capacity), int(nr_levels), int(nr_feat_per_level)
        scales = np.asarray(scale_per_level, dtype=np.float64)
        assert scales.shape[0] == nr_levels
        self.lattice_values = nn.Parameter(torch.randn(nr_levels, self.capacity, self.nr_feat) * 1e-5)
        shift = torch.randn(nr_levels, 3) * 10.0 if apply_random_shift_per_level else torch.zeros(nr_levels, 3)
        self.register_buffer('random_shift_per_level', shift)
        sf = np.stack([1.0 / np.sqrt((i + 1) * (i + 2)) / scales for i in range(3)], axis=1)
        self.register_buffer('scale_factor', torch.from_numpy(sf.astype(np.float32)))
        self.register_buffer('anneal_window', torch.ones(nr_levels))
        # coarse levels (lattice spacing >= 5 % of the unit cube) use warp-aggregated scatter in backward; measured on the
        # bench workload (tools/permuto_tune.py, 410 k samples): 0 levels 1249 us, 4: 445, 8: 193 (best), 10: 215, 14: 357
        self.n_agg_levels = int((scales >= agg_scale_threshold).sum())

    def output_dims(self):
        return self.nr_levels * self.nr_feat

    def forward(self, positions):
        return ops.permuto_encode(positions, self.lattice_values, self.scale_factor, self.random_shift_per_level,
                                  self.anneal_window, self.n_agg_levels)


class PermutoGrid(HashGridBase):
    def __init__(self, *args, coarsest_scale=1.0, finest_scale=0.001, capacity_log_2=18, num_lods=24, **kwargs):
        super().__init__(*args, **kwargs)
        self._register_blas_buffers()
        self.coarsest_scale = coarsest_scale
        self.finest_scale = finest_scale
        self.capacity = pow(2, capacity_log_2)
        self.num_lods = num_lods
        self.multiscale_type = 'cat'

    def set_capacity(self, capacity_log_2):
        self.capacity = pow(2, capacity_log_2)

    def init_from_scales(self):
        self.active_lods = [x for x in range(self.num_lods)]
        self.max_lod = self.num_lods - 1
        self.resolutions = np.geomspace(self.coarsest_scale, self.finest_scale, num=self.num_lods)
        log.info(f"Active Resolutions: {self.resolutions}")
        self.embedder = PermutoEncoding(3, self.capacity, self.num_lods, self.feature_dim, self.resolutions)

    def interpolate(self, coords, lod_idx=None, pidx=None):
        if coords.numel() == 0:
            return torch.empty([0, 1, self.num_lods * self.feature_dim], device=coords.device)
        pos = coords.reshape(-1, 3)
        if torch.is_autocast_enabled():
            # custom_fwd(cast_inputs=torch.half) then .type(torch.float) (grids/permuto_grid.py:65,71)
            pos = pos.half()
        return self.embedder(pos.type(torch.float))
