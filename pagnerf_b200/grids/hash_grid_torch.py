"""HashGridTorch: the reference's own torch hash grid (grids/hash_grid_torch.py:48-141) served by
csrc/hashgrid.cu (flavour 1).  `HashEmbedder` keeps the reference's constructor and state_dict keys
(`embeddings.{l}.weight`) while storing ONE contiguous table [L, 2^T, F] for the kernel."""
import logging as log

import torch
import torch.nn as nn

from .. import ops
from .base import HashGridBase


class HashEmbedder(nn.Module):
    def __init__(self, n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16,
                 finest_resolution=512, agg_resolution_threshold=16):
        super().__init__()
        if n_features_per_level != 2:
            raise NotImplementedError("csrc/hashgrid.cu is specialised to 2 features per level")
        self.n_levels = n_levels
        self.n_features_per_level = n_features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = torch.tensor(base_resolution)
        self.finest_resolution = torch.tensor(finest_resolution)
        self.out_dim = self.n_levels * self.n_features_per_level
        # same float32 torch arithmetic as the reference (:59, :99) -> identical per-level resolutions
        self.b = torch.exp((torch.log(self.finest_resolution) - torch.log(self.base_resolution)) / (n_levels - 1))
        res = torch.stack([torch.floor(self.base_resolution * self.b ** i) for i in range(n_levels)]).float()
        T = 2 ** log2_hashmap_size
        self.register_buffer('level_res', res, persistent=False)
        self.register_buffer('level_offset', (torch.arange(n_levels) * T).to(torch.int32), persistent=False)
        self.register_buffer('level_size', torch.full((n_levels,), T, dtype=torch.int32), persistent=False)
        w = torch.empty(n_levels, T, n_features_per_level)
        nn.init.uniform_(w, a=-0.0001, b=0.0001)
        self.embeddings_weight = nn.Parameter(w)
        # warp-aggregated scatter only on the coarsest level: measured on BASELINE configs 1 / 3 (tools/hash_tune.py, B200):
        # 0 / 1 / 3 / 5 aggregated levels = 0.31 / 0.28 / 0.29 / 0.41 ms and 1.51 / 1.19 / 1.79 / 3.31 ms per launch
        self.n_agg_levels = int((res <= agg_resolution_threshold).sum())
        self._register_state_dict_hook(self._split_hook)
        self._register_load_state_dict_pre_hook(self._merge_hook)

    @staticmethod
    def _split_hook(module, state_dict, prefix, local_metadata):
        w = state_dict.pop(prefix + 'embeddings_weight')
        for l in range(w.shape[0]):
            state_dict[f"{prefix}embeddings.{l}.weight"] = w[l]
        return state_dict

    def _merge_hook(self, state_dict, prefix, *args):
        keys = [f"{prefix}embeddings.{l}.weight" for l in range(self.n_levels)]
        if all(k in state_dict for k in keys):
            state_dict[prefix + 'embeddings_weight'] = torch.stack([state_dict.pop(k) for k in keys])

    def forward(self, x):
        return ops.hash_encode(x, self.embeddings_weight, 1, self.level_res, None, self.level_offset, self.level_size,
                               False, self.n_agg_levels)


class HashGridTorch(HashGridBase):
    def init_from_resolutions(self, resolutions):
        self.resolutions = resolutions
        self.num_lods = len(resolutions)
        self.active_lods = [x for x in range(self.num_lods)]
        self.max_lod = self.num_lods - 1
        log.info(f"Active Resolutions: {self.resolutions}")
        self.embedder = HashEmbedder(n_levels=self.num_lods, n_features_per_level=self.feature_dim,
                                     log2_hashmap_size=self.codebook_bitwidth,
                                     base_resolution=resolutions[0], finest_resolution=resolutions[-1])

    def interpolate(self, coords, lod_idx, pidx=None):
        batch, num_samples, _ = coords.shape
        feats = self.embedder(coords.reshape(-1, 3))
        if self.multiscale_type == 'cat':
            return feats
        elif self.multiscale_type == 'sum':
            return feats.reshape(batch, num_samples, len(self.resolutions), feats.shape[-1] // len(self.resolutions)).sum(-2)
        raise NotImplementedError
