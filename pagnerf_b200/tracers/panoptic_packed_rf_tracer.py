"""PanopticPackedRFTracer: packed radiance-field tracer for the panoptic fields.

Drop-in for the reference class (tracers/panoptic_packed_rf_tracer.py:19-205): same constructor,
`trace()` signature and RenderBuffer fields.  march -> nef -> ONE fused compositing kernel replaces
the reference's 2x exponential_integration + 5x sum_reduce + 6x index_put chain; the compositing
conventions (alpha on top of the weighted sum, detached second integration for the panoptic
channels, white/black background, empty rays) are those of the reference, see csrc/composite.cu.
"""
import torch

from .. import ops
from ..wisp_compat import PackedRFTracer, RenderBuffer


def sigma_sparsity_loss(sigmas):
    """Cauchy sparsity loss (reference loss/regularizers.py:37-39)."""
    return torch.log(1.0 + 2 * torch.pow(sigmas, 2))


class PanopticPackedRFTracer(PackedRFTracer):
    def __init__(self, ray_sparcity_reg=0.0, ray_max_travel=6.0, **kwargs):
        super().__init__(**kwargs)
        self.render_channels = {'depth', 'alpha', 'hit'}
        self.base_channels = {'rgb', 'density'}
        self.panoptic_channels = {'semantics', 'inst_embedding'}
        self.ray_sparcity_reg = ray_sparcity_reg
        self.ray_max_travel = ray_max_travel
        self.allow_fused = True     # sync-free fused trace in training mode ('ray' and 'voxel' marching)

    def get_supported_channels(self):
        return {'depth', 'hit', 'rgb', 'alpha', 'semantics', 'inst_embedding'}

    def get_required_nef_channels(self):
        return {'rgb', 'density'}

    def trace(self, nef, channels, extra_channels, rays, lod_idx=None, raymarch_type='voxel', num_steps=64,
              step_size=1.0, bg_color='white', stage='val'):
        assert nef.grid is not None, "this tracer requires a grid"
        N = rays.origins.shape[0]
        dev = rays.origins.device
        if lod_idx is None:
            lod_idx = nef.grid.num_lods - 1

        plain = not extra_channels and not (self.ray_sparcity_reg > 0.0 and stage == 'train')
        if plain and raymarch_type in ('ray', 'voxel') and self.allow_fused and hasattr(nef, 'fused_trace_cfg'):
            # 'voxel' (the trainer's mode from epoch 201 on, configs/bup20/best.yaml:34): nuggets + max-travel filter (:88-108)
            # stay on the device as well
            cfg = nef.fused_trace_cfg(channels, rays, num_steps, bg_color, raymarch_type,
                                      self.ray_max_travel if raymarch_type == 'voxel' else None)
            if cfg is not None:
                # training mode, sync-free: march -> encode -> decode -> composite as one autograd node
                table, dtable, wts = nef.fused_trace_tensors()
                alpha, hit, rgb, depth, sem, inst, m_dev = ops.FusedTraceFn.apply(rays.origins, rays.dirs, cfg, table, dtable, *wts)
                self.last_num_samples = m_dev            # device scalar (no host sync here)
                outputs = {'alpha': alpha, 'hit': hit}
                for name, val in (('rgb', rgb), ('depth', depth), ('semantics', sem), ('inst_embedding', inst)):
                    if name in channels:
                        outputs[name] = val
                return RenderBuffer(**outputs)

        kw = {} if getattr(nef.grid, 'interpolate_needs_pidx', True) else {'need_pidx': False}
        ridx, pidx, samples, depths, deltas, boundary = nef.grid.raymarch(
            rays, level=nef.grid.active_lods[lod_idx], num_samples=num_steps, raymarch_type=raymarch_type, **kw)
        S = samples.shape[1]

        if raymarch_type == 'voxel' and depths.numel() != 0:
            # drop nuggets further than ray_max_travel behind the ray's first hit (:88-108)
            first = ops.ray_offsets(ridx, N)
            valid_mask = ops.max_travel_mask(ridx, depths, first, self.ray_max_travel)
            deltas = deltas.reshape(depths.shape)[valid_mask].reshape(-1, 1)
            ridx, pidx, samples, depths = ridx[valid_mask], (pidx[valid_mask] if pidx is not None else None), samples[valid_mask], depths[valid_mask]

        offsets = ops.ray_offsets(ridx, N) * S        # packed sample range of every ray (empty rays: lo == hi)
        self.last_num_samples, self._last_ridx = int(ridx.shape[0]) * S, ridx
        hit_ray_d = rays.dirs.index_select(0, ridx)

        outputs = {}
        if (not extra_channels and not (self.ray_sparcity_reg > 0.0 and stage == 'train')
                and hasattr(nef, 'fused_panoptic_ok') and nef.fused_panoptic_ok(channels)):
            # training mode: decode + composite fused, the [M,C] panoptic probabilities never reach HBM
            ridx_rows = ridx if S == 1 else ridx.repeat_interleave(S)
            out = nef.trace_composited(samples, hit_ray_d, ridx_rows, deltas, depths, offsets, N, channels,
                                       bg_color == 'white', lod_idx)
            for c in ('alpha', 'hit', 'rgb', 'depth', 'semantics', 'inst_embedding'):
                if c in out and (c in channels or c in ('alpha', 'hit')):
                    outputs[c] = out[c]
            return RenderBuffer(**outputs)
        sample_channels = set(channels - self.render_channels)
        sample_channels.update(['density'])
        out_feats = nef(coords=samples, ray_d=hit_ray_d, pidx=pidx, lod_idx=lod_idx, channels=sample_channels)

        if self.ray_sparcity_reg > 0.0 and stage == 'train':
            all_rays_loss = sigma_sparsity_loss(out_feats['density'].reshape(ridx.shape[0], -1).sum(-1) if S > 1
                                                else out_feats['density'].squeeze())
            ray_wise_loss = torch.scatter_add(torch.zeros_like(rays.origins[:, 0]), 0, ridx, all_rays_loss)
            outputs['ray_sparcity_loss'] = ray_wise_loss.mean() * self.ray_sparcity_reg

        alpha, hit, rgb, depth, sem, inst, _w = ops.composite(
            out_feats['density'], deltas,
            depths if 'depth' in channels else None,
            out_feats['rgb'] if 'rgb' in channels else None,
            out_feats['semantics'] if 'semantics' in channels else None,
            out_feats['inst_embedding'] if 'inst_embedding' in channels else None,
            offsets, bg_white=(bg_color == 'white'))
        outputs['alpha'] = alpha
        outputs['hit'] = hit
        if 'rgb' in channels:
            outputs['rgb'] = rgb
        if 'depth' in channels:
            outputs['depth'] = depth
        if 'semantics' in channels:
            outputs['semantics'] = sem
        if 'inst_embedding' in channels:
            outputs['inst_embedding'] = inst

        extra_outputs = {}
        if extra_channels:
            _, w = ops.exponential_integration(None, out_feats['density'].reshape(-1, 1) * deltas, boundary_from_offsets(offsets, ridx.shape[0] * S))
            for channel in extra_channels:
                feats = nef(coords=samples, ray_d=hit_ray_d, pidx=pidx, lod_idx=lod_idx, channels=channel)
                extra_outputs[channel] = self._integrate_features(feats, alpha, w, offsets, N)
        return RenderBuffer(**outputs, **extra_outputs)

    def _integrate_features(self, feats, alpha, w, offsets, N):
        num_channels = feats.shape[-1]
        ray_feats = ops.SumReduceFn.apply(w.reshape(-1, 1) * feats.reshape(-1, num_channels), offsets)
        return alpha * ray_feats


def boundary_from_offsets(offsets, M):
    b = torch.zeros(M, dtype=torch.bool, device=offsets.device)
    starts = offsets[:-1][offsets[:-1] < offsets[1:]]
    b[starts] = True
    return b
