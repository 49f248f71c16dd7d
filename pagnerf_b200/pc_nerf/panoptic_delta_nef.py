"""PanopticDeltaNeF: PanopticNeF + sparse panoptic delta grid with the stop-gradient into the color
branch preserved (reference pc_nerf/panoptic_delta_nef.py:21-259; stop-grad :214-215, fusion :225-236)."""
import copy
import logging as log

import torch

from ..wisp_compat import get_positional_embedder
from .panoptic_nef import PanopticNeF


class PanopticDeltaNeF(PanopticNeF):
    def __init__(self, delta_num_layers: int = 1, delta_hidden_dim: int = 64, inst_soft_temperature: float = 0.0,
                 **kwargs):
        self.delta_num_layers = delta_num_layers
        self.delta_hidden_dim = delta_hidden_dim
        self.inst_soft_temperature = inst_soft_temperature
        super().__init__(**kwargs)

    def init_grid(self):
        super().init_grid()
        if self.panoptic_features_type in ['delta', 'separate'] or self.panoptic_features_type is None:
            self.delta_grid = copy.deepcopy(self.grid)
            if self.grid_type == "PermutoGrid" and self.panoptic_features_type in ['delta', 'separate']:
                self.delta_grid.set_capacity(self.kwargs['delta_capacity_log_2'])

    def init_embedder(self):
        self.pos_embedder, self.pos_embed_dim = get_positional_embedder(self.pos_multires, True)
        log.info(f"Pos Embed Dim: {self.pos_embed_dim}")
        super().init_embedder()

    def get_nef_type(self):
        return 'delta_panoptic_nef'

    def _prune_supported(self):
        return True      # the delta field prunes every grid type (pc_nerf/panoptic_delta_nef.py:63-104)

    def _prune_grids(self):
        return [self.grid] + ([self.delta_grid] if 'delta_grid' in dir(self) else [])

    def register_forward_functions(self):
        self._register_forward_function(self.rgb_semantics, ["density", "rgb", "semantics", "inst_embedding"])

    def _panoptic_inputs(self, feats, coords, lod_idx, rows=None):
        """rows (int64 [K], inference only): packed samples to evaluate -- the delta grid is looked up for those only."""
        pft = self.panoptic_features_type
        feats_detached = feats.detach()          # :214
        if rows is not None:
            feats_detached = feats_detached.index_select(0, rows)
            coords = coords.reshape(-1, 1, 3).index_select(0, rows)
        if pft in ['delta', 'separate'] or pft is None:
            delta_feats = self._encode(self.delta_grid, coords.detach(), lod_idx)   # coords.detach(): :215
        if pft == 'delta' or pft is None:
            return feats_detached, delta_feats    # panop = feats.detach() + delta (:226)
        if pft == 'separate':
            return delta_feats, None
        if pft == 'appearance':
            return feats_detached, None
        raise NotImplementedError(f'Panoptic feature type "{pft}" is not served by the fused decoders '
                                  '(pos_encoding / position change the decoder input width)')

    def rgb_semantics(self, coords, ray_d, compute_channels, pidx=None, lod_idx=None):
        out_dict = {}
        if not compute_channels:
            return out_dict
        if lod_idx is None:
            lod_idx = len(self.grid.active_lods) - 1
        batch, num_samples, _ = coords.shape
        if self.position_input:
            raise NotImplementedError
        feats = self._encode(self.grid, coords, lod_idx)
        # density decoder runs whenever any channel is requested (:182)
        sigma, rgb = self._dc(feats, ray_d, num_samples, 'rgb' in compute_channels)
        if 'density' in compute_channels:
            out_dict['density'] = sigma.reshape(batch, num_samples, 1)
        if 'rgb' in compute_channels:
            out_dict['rgb'] = rgb.reshape(batch, num_samples, 3)
        want_sem, want_inst = 'semantics' in compute_channels, 'inst_embedding' in compute_channels
        if want_sem or want_inst:
            a, b = self._panoptic_inputs(feats, coords, lod_idx)
            sem, inst = self._pan(a, b, want_sem, want_inst, self.inst_soft_temperature)
            if want_sem:
                out_dict['semantics'] = sem
            if want_inst:
                out_dict['inst_embedding'] = inst
        return out_dict
