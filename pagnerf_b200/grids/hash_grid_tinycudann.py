"""HashGridTinyCudaNN: instant-ngp hash grid with tiny-cuda-nn's level / index conventions
(reference grids/hash_grid_tinycudann.py:8-47), backed by csrc/hashgrid.cu (flavour 0)."""
import logging as log

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .base import HashGridBase


def tcnn_levels(n_levels, log2_T, base_resolution, per_level_scale=2.0):
    """Per-level (scale f32, resolution, entry offset, entries) exactly as tcnn's GridEncoding lays them out."""
    scales, ress, offs, sizes = [], [], [], []
    off = 0
    l2 = np.float32(np.log2(np.float32(per_level_scale)))
    for l in range(n_levels):
        scale = np.float32(np.exp2(np.float32(l) * l2) * np.float32(base_resolution) - np.float32(1.0))
        res = int(np.ceil(scale)) + 1
        n = min(res ** 3, 2 ** 32 - 8)
        n = (n + 7) // 8 * 8
        n = min(n, 1 << log2_T)
        scales.append(scale); ress.append(res); offs.append(off); sizes.append(n)
        off += n
    return np.array(scales, np.float32), np.array(ress, np.int64), np.array(offs, np.int64), np.array(sizes, np.int64), off


class TcnnEncoding(nn.Module):
    """tcnn.Encoding(n_input_dims=3, {otype: HashGrid, ...}) container: flat `params` like upstream."""

    def __init__(self, n_input_dims, encoding_config, agg_resolution_threshold=16):
        super().__init__()
        c = encoding_config
        if n_input_dims != 3 or c.get("otype", "HashGrid") != "HashGrid" or c["n_features_per_level"] != 2:
            raise NotImplementedError("csrc/hashgrid.cu is specialised to 3-D HashGrid with 2 features per level")
        self.n_levels, self.F = int(c["n_levels"]), 2
        sc, rs, of, sz, total = tcnn_levels(self.n_levels, int(c["log2_hashmap_size"]), c["base_resolution"],
                                            c.get("per_level_scale", 2.0))
        self.register_buffer('level_scale', torch.from_numpy(sc), persistent=False)
        self.register_buffer('level_res', torch.from_numpy(rs).to(torch.int32), persistent=False)
        self.register_buffer('level_offset', torch.from_numpy(of).to(torch.int32), persistent=False)
        self.register_buffer('level_size', torch.from_numpy(sz).to(torch.int32), persistent=False)
        self.params = nn.Parameter((torch.rand(total * self.F) * 2 - 1) * 1e-4)
        self.n_output_dims = self.n_levels * self.F
        # warp-aggregated scatter only on the coarsest level: measured on BASELINE configs 1 / 3 (tools/hash_tune.py, B200):
        # 0 / 1 / 3 / 5 aggregated levels = 0.31 / 0.28 / 0.29 / 0.41 ms and 1.51 / 1.19 / 1.79 / 3.31 ms per launch
        self.n_agg_levels = int((rs <= agg_resolution_threshold).sum())
        self.round_half = True  # upstream returns __half; the wrapper casts to float (:41)

    def forward(self, x):
        return ops.hash_encode(x, self.params, 0, self.level_scale, self.level_res, self.level_offset, self.level_size,
                               self.round_half, self.n_agg_levels)


class HashGridTinyCudaNN(HashGridBase):
    def init_from_resolutions(self, resolutions):
        self.resolutions = resolutions
        self.num_lods = len(resolutions)
        self.active_lods = [x for x in range(self.num_lods)]
        self.max_lod = self.num_lods - 1
        log.info(f"Active Resolutions: {self.resolutions}")
        self.embedder = TcnnEncoding(
            n_input_dims=3,
            encoding_config={"otype": "HashGrid", "n_levels": self.num_lods, "n_features_per_level": self.feature_dim,
                             "log2_hashmap_size": self.codebook_bitwidth, "base_resolution": resolutions[0],
                             "per_level_scale": 2})

    def interpolate(self, coords, lod_idx, pidx=None):
        batch, num_samples, _ = coords.shape
        pos = coords.reshape(-1, 3)
        if torch.is_autocast_enabled():
            pos = pos.half()  # custom_fwd(cast_inputs=torch.half), grids/hash_grid_tinycudann.py:36
        feats = self.embedder(pos.float()).type(torch.float)
        if self.multiscale_type == 'cat':
            return feats
        elif self.multiscale_type == 'sum':
            return feats.reshape(batch, num_samples, len(self.resolutions), feats.shape[-1] // len(self.resolutions)).sum(-2)
        raise NotImplementedError
