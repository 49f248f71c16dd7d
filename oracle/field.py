"""Oracle: PanopticNeF / PanopticDeltaNeF forward and PanopticPackedRFTracer.trace on the CPU.

TEST INFRASTRUCTURE ONLY.  Follows the in-tree reference text:
  pc_nerf/panoptic_nef.py:108-164 (decoder shapes), :253-363 (rgb_semantics)
  pc_nerf/panoptic_delta_nef.py:39-44 (delta grid), :116-259 (rgb_semantics; stop-grad :214-215)
  tracers/panoptic_packed_rf_tracer.py:85-205 (filter, two integrations, alpha-on-top, scatter)
The glue is additionally pinned by tests/golden/trace_*.npz, produced by running the
reference's own tracer / nef source unmodified on stub wisp/kaolin modules (make_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import autocast, spc
from .decoders import BasicDecoderOracle, PositionalEmbedderOracle


class FieldOracle(nn.Module):
    """Random-init panoptic field: encoder(s) + 4 decoders, best.yaml shapes by default."""

    def __init__(self, grid, delta_grid=None, feat_dim=48, hidden_dim=64, num_classes=6, num_instances=200,
                 view_multires=4, num_layers=1, sem_num_layers=1, inst_num_layers=2,
                 sem_softmax=True, inst_softmax=True, inst_soft_temperature=0.0, seed=0, delta_density=False):
        super().__init__()
        torch.manual_seed(seed)
        self.grid, self.delta_grid = grid, delta_grid
        self.view_embedder = PositionalEmbedderOracle(view_multires)
        self.decoder_density = BasicDecoderOracle(feat_dim, 16, num_layers, hidden_dim)
        self.decoder_density.lout.bias.data[0] = 1.0  # pc_nerf/panoptic_nef.py:123
        self.decoder_color = BasicDecoderOracle(16 + self.view_embedder.out_dim, 3, num_layers + 1, hidden_dim)
        self.decoder_semantics = BasicDecoderOracle(feat_dim, num_classes, sem_num_layers, hidden_dim)
        self.decoder_inst = BasicDecoderOracle(feat_dim, num_instances, inst_num_layers, hidden_dim)
        # PanopticDDensityNeF (pc_nerf/panoptic_dd_nef.py:41-58): activation-free 1-hidden-layer head feat -> 64 -> 1
        self.decoder_delta_density = BasicDecoderOracle(feat_dim, 1, 1, hidden_dim, activation='none') if delta_density else None
        self.lod_weights = torch.ones(feat_dim)
        self.sem_softmax, self.inst_softmax = sem_softmax, inst_softmax
        self.inst_soft_temperature = inst_soft_temperature

    def forward(self, coords, ray_d, channels):
        """coords [M,S,3], ray_d [M,3] -> dict (shapes as the reference: density [M,S,1], rgb [M,S,3],
        semantics [M*S,C], inst_embedding [M*S,C])."""
        out = {}
        batch, S, _ = coords.shape
        amp = autocast.enabled()      # the reference's autocast step (oracle/autocast.py): fp16-rounded coordinates into the
        ph = getattr(self, 'pos_half', None)      # permutohedral encoders (grids/permuto_grid.py:65), fp16 Linear layers, fp32 softmax
        if amp if ph is None else ph:             # pos_half=True alone: exact fp32 arithmetic on the fp16-rounded coordinates
            coords = coords.half().to(coords.dtype)
        feats = self.grid(coords.reshape(-1, 3)) * self.lod_weights.to(coords.dtype)
        density_feats = self.decoder_density(feats)
        density = torch.relu(density_feats[..., 0:1]).reshape(batch, S, 1)
        if 'density' in channels:
            out['density'] = density
        if 'rgb' in channels:
            ve = self.view_embedder(-ray_d)[:, None].repeat(1, S, 1).view(-1, self.view_embedder.out_dim)
            out['rgb'] = torch.sigmoid(self.decoder_color(torch.cat([density_feats, ve], -1))).reshape(batch, S, 3)
        if any(c in channels for c in ('semantics', 'inst_embedding', 'delta_density', 'panoptic_density')):
            if self.delta_grid is not None:
                dfe = self.delta_grid(coords.detach().reshape(-1, 3)) * self.lod_weights.to(coords.dtype)
                panop = feats.detach() + dfe
            else:  # PanopticNeF with sem_detach / inst_detach = True (defaults)
                panop = feats.detach()
        if 'delta_density' in channels or 'panoptic_density' in channels:      # pc_nerf/panoptic_dd_nef.py:236-247
            dd = self.decoder_delta_density(panop).reshape(batch, S, 1)
            if 'delta_density' in channels:
                out['delta_density'] = dd
            if 'panoptic_density' in channels:
                out['panoptic_density'] = torch.relu(density_feats[..., 0:1].reshape(batch, S, 1).detach() + dd)
        if 'semantics' in channels:
            s = self.decoder_semantics(panop)
            out['semantics'] = F.softmax(s.float() if amp else s, -1) if self.sem_softmax else s
        if 'inst_embedding' in channels:
            e = self.decoder_inst(panop)
            if self.inst_soft_temperature > 0.0:
                e = e / self.inst_soft_temperature
            out['inst_embedding'] = F.softmax(e.float() if amp else e, -1) if self.inst_softmax else e
        return out


def trace_oracle(field, origins, dirs, ridx, samples, depths, deltas, boundary, channels,
                 bg_color='white', dd=False):
    """PanopticPackedRFTracer.trace after the marcher (tracers/...:113-195).

    ridx int64 [M], samples [M,S,3], depths [M,S,1] or [M,1], deltas [M*S,1], boundary bool [M*S].
    Returns dict of dense [N,C] outputs (rgb, depth, alpha, hit, semantics, inst_embedding).
    """
    N = origins.shape[0]
    dt = origins.dtype
    out = {}
    ridx = ridx.long()
    ridx_hit = ridx[spc.mark_pack_boundaries(ridx.int())]
    hit_ray_d = dirs.index_select(0, ridx)
    sample_channels = set(channels) - {'depth', 'alpha', 'hit'} | {'density'}
    pan = [c for c in ('semantics', 'inst_embedding') if c in channels]
    if dd and pan:      # tracers/panoptic_dd_packed_rf_tracer.py:102-103
        sample_channels |= {'panoptic_density'}
    feats = field(samples, hit_ray_d, sample_channels)
    tau = feats['density'].reshape(-1, 1) * deltas
    _, w = spc.exponential_integration(None, tau, boundary, exclusive=True)
    alpha = spc.sum_reduce(w, boundary)
    out_alpha = torch.zeros(N, 1, dtype=dt); out_alpha[ridx_hit] = alpha
    out['alpha'] = out_alpha
    hit = torch.zeros(N, dtype=torch.bool); hit[ridx_hit] = alpha[..., 0] > 0.0
    out['hit'] = hit
    if 'rgb' in channels:
        ray_colors = spc.sum_reduce(feats['rgb'].reshape(-1, 3) * w, boundary)
        if bg_color == 'white':
            rgb = torch.ones(N, 3, dtype=dt); color = (1.0 - alpha) + alpha * ray_colors
        else:
            rgb = torch.zeros(N, 3, dtype=dt); color = alpha * ray_colors
        rgb[ridx_hit] = color
        out['rgb'] = rgb
    if 'depth' in channels:
        rd = spc.sum_reduce(depths.reshape(-1, 1) * w, boundary)
        depth = torch.zeros(N, 1, dtype=dt); depth[ridx_hit] = rd
        out['depth'] = depth
    if pan:
        # PanopticPackedRFTracer: second integration of the DETACHED colour density (:149-155);
        # PanopticDDensityPackedRFTracer: integration of the panoptic density, gradients flowing (:128-137)
        ptau = feats['panoptic_density'].reshape(-1, 1) * deltas.detach() if dd else tau.detach()
        _, pw = spc.exponential_integration(None, ptau, boundary, exclusive=True)
        palpha = spc.sum_reduce(pw, boundary)
        for c in pan:
            f = feats[c]
            rf = spc.sum_reduce(pw * f.view(-1, f.shape[-1]), boundary)
            o = torch.zeros(N, f.shape[-1], dtype=dt); o[ridx_hit] = palpha * rf
            out[c] = o
    return out
