"""Stub `wisp`, `kaolin`, `permutohedral_encoding`, `tinycudann` modules so that the REFERENCE's own
source files (/root/reference/grids/*.py, pc_nerf/panoptic_{,delta_}nef.py,
tracers/panoptic_packed_rf_tracer.py) can be imported UNMODIFIED in the build container and run on
the CPU.  Test / golden-generation infrastructure only (needs /root/reference: never used by the
`-m gpu` tests, smoke() or bench.py).

Types come from pagnerf_b200.wisp_compat (pure-Python containers); the arithmetic of the absent
third-party CUDA extensions comes from the CPU oracle.  What this pins is the reference's GLUE:
channel gating, stop-gradients, the two integrations, the alpha-on-top compositing convention,
background / scatter behaviour -- not the third-party kernels themselves (parity unpinned there).
"""
import logging
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    if "wisp" in sys.modules and getattr(sys.modules["wisp"], "_pag_stub", False):
        return
    from oracle import spc as ospc, raymarch as orm
    from oracle.permuto import PermutoEncodingOracle
    from oracle.hashgrid import TcnnHashGridOracle
    from pagnerf_b200.wisp_compat.core import Rays, RenderBuffer
    from pagnerf_b200.wisp_compat.nefs import BaseNeuralField
    from pagnerf_b200.wisp_compat.tracers import BaseTracer, PackedRFTracer
    from pagnerf_b200.wisp_compat import modules as wm

    class OctreeAS:
        """wisp.accelstructs.OctreeAS data holder + oracle-backed query/raytrace/raymarch."""
        jitter_seed = 0

        def __init__(self):
            self.octree = self.points = self.pyramid = self.prefix = None
            self.max_level = None

        def init(self, octree):
            octree = octree.cpu().numpy() if torch.is_tensor(octree) else np.asarray(octree)
            level = 0
            n, acc = octree.shape[0], 0
            # recover max level by walking the popcounts
            cnt = 1
            while acc < n:
                nxt = int(ospc._POPC8[octree[acc:acc + cnt]].sum())
                acc += cnt
                cnt = nxt
                level += 1
            points, pyramid, prefix = ospc.scan_octree(octree, level)
            self.octree = torch.from_numpy(octree.copy())
            self.points = torch.from_numpy(points.copy())
            self.pyramid = torch.from_numpy(pyramid.copy())
            self.prefix = torch.from_numpy(prefix.copy())
            self.max_level = level

        def init_dense(self, level):
            self.init(ospc.dense_octree(level))

        def query(self, coords, level=None):
            return torch.from_numpy(ospc.query(self.octree.numpy(), self.prefix.numpy(),
                                               coords.detach().numpy(), level or self.max_level)).long()

        def raymarch(self, rays, level, num_samples, raymarch_type):
            o, d = rays.origins.detach().numpy(), rays.dirs.detach().numpy()
            if raymarch_type == 'voxel':
                ridx, pidx, s, dp, dl, b = orm.raymarch_voxel(
                    self.octree.numpy(), self.points.numpy(), self.pyramid.numpy(), self.prefix.numpy(),
                    o, d, level, num_samples, seed=self.jitter_seed)
                # samples/depths re-attached to the ray tensors for autograd (addcmul in upstream)
                ridx_t = torch.from_numpy(ridx).long()
                dp_t = torch.from_numpy(dp)
                samples = rays.origins[ridx_t][:, None] + rays.dirs[ridx_t][:, None] * dp_t
                samples = torch.from_numpy(s) + (samples - samples.detach())
                return (ridx_t, torch.from_numpy(pidx).long(), samples, dp_t, torch.from_numpy(dl),
                        torch.from_numpy(b))
            elif raymarch_type == 'ray':
                ridx, pidx, s, dp, dl, b = orm.raymarch_ray(
                    self.octree.numpy(), self.prefix.numpy(), o, d, level, num_samples,
                    rays.dist_min, rays.dist_max, seed=self.jitter_seed)
                ridx_t = torch.from_numpy(ridx).long()
                dp_t = torch.from_numpy(dp)
                samples = rays.origins[ridx_t][:, None] + rays.dirs[ridx_t][:, None] * dp_t[:, :, None]
                samples = torch.from_numpy(s) + (samples - samples.detach())
                return (ridx_t, torch.from_numpy(pidx).long(), samples, dp_t, torch.from_numpy(dl),
                        torch.from_numpy(b))
            raise TypeError(raymarch_type)

    class BLASGrid(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    class HashGrid(BLASGrid):
        def __init__(self, feature_dim, interpolation_type='linear', multiscale_type='cat', feature_std=0.0,
                     feature_bias=0.0, codebook_bitwidth=8, blas_level=7, **kwargs):
            super().__init__()
            self.feature_dim, self.interpolation_type, self.multiscale_type = feature_dim, interpolation_type, multiscale_type
            self.feature_std, self.feature_bias, self.codebook_bitwidth = feature_std, feature_bias, codebook_bitwidth
            self.blas_level = blas_level
            self.kwargs = kwargs
            self.blas = OctreeAS()
            self.blas.init_dense(self.blas_level)
            self.dense_points = torch.from_numpy(
                ospc.level_points(self.blas.points.numpy(), self.blas.pyramid.numpy(), self.blas_level).copy())
            self.num_cells = self.dense_points.shape[0]
            self.occupancy = torch.zeros(self.num_cells)

        def init_from_octree(self, base_lod, num_lods):
            self.init_from_resolutions([2 ** (base_lod + i) for i in range(num_lods)])

        def init_from_geometric(self, min_width, max_width, num_lods):
            b = np.exp((np.log(max_width) - np.log(min_width)) / (num_lods - 1))
            self.init_from_resolutions([int(np.floor(min_width * (b ** l))) for l in range(num_lods)])

        def raymarch(self, rays, level=None, num_samples=64, raymarch_type='voxel'):
            return self.blas.raymarch(rays, level=self.blas_level, num_samples=num_samples, raymarch_type=raymarch_type)

    class _Unavailable(BLASGrid):
        pass

    class PermutoEncoding(PermutoEncodingOracle):
        def __init__(self, pos_dim, capacity, nr_levels, nr_feat_per_level, scale_per_level, **kw):
            assert pos_dim == 3
            super().__init__(capacity, nr_levels, nr_feat_per_level, scale_per_level)

    class TcnnEncoding(nn.Module):
        def __init__(self, n_input_dims, encoding_config, **kw):
            super().__init__()
            c = encoding_config
            self.inner = TcnnHashGridOracle(c["n_levels"], c["n_features_per_level"], c["log2_hashmap_size"],
                                            c["base_resolution"], c["per_level_scale"], out_half=True)
            self.params = self.inner.params

        def forward(self, x):
            return self.inner(x)

    class _SpcRender:
        mark_pack_boundaries = staticmethod(ospc.mark_pack_boundaries)
        sum_reduce = staticmethod(ospc.sum_reduce)
        cumsum = staticmethod(ospc.cumsum)

        @staticmethod
        def exponential_integration(feats, tau, boundary, exclusive=True):
            return ospc.exponential_integration(feats, tau, boundary, exclusive)

    class _SpcOps:
        @staticmethod
        def unbatched_points_to_octree(points, level, sorted=False):
            return torch.from_numpy(ospc.points_to_octree(points.cpu().numpy(), level))

        @staticmethod
        def unbatched_get_level_points(points, pyramid, level):
            return torch.from_numpy(ospc.level_points(points.numpy(), pyramid.numpy(), level).copy())

    wisp = _mod("wisp", _pag_stub=True)
    _mod("wisp.core", Rays=Rays, RenderBuffer=RenderBuffer)
    _mod("wisp.core.rays", Rays=Rays)
    _mod("wisp.utils", PsDebugger=object, PerfTimer=wm.PerfTimer)
    _mod("wisp.tracers", BaseTracer=BaseTracer, PackedRFTracer=PackedRFTracer)
    _mod("wisp.tracers.base_tracer", BaseTracer=BaseTracer)
    _mod("wisp.models")
    _mod("wisp.models.nefs", BaseNeuralField=BaseNeuralField)
    _mod("wisp.models.activations", get_activation_class=wm.get_activation_class)
    _mod("wisp.models.layers", get_layer_class=wm.get_layer_class)
    _mod("wisp.models.embedders", get_positional_embedder=wm.get_positional_embedder,
         PositionalEmbedder=wm.PositionalEmbedder)
    _mod("wisp.models.decoders", BasicDecoder=wm.BasicDecoder)
    _mod("wisp.models.pipeline", Pipeline=wm.Pipeline)
    _mod("wisp.accelstructs", OctreeAS=OctreeAS)
    _mod("wisp.ops")
    _mod("wisp.ops.spc", sample_spc=None)
    _mod("wisp.ops.grid")
    _mod("wisp.ops.geometric", sample_unif_sphere=lambda n: F.normalize(torch.randn(n, 3), dim=-1).numpy())
    kaolin = _mod("kaolin")
    _mod("kaolin.render")
    _mod("kaolin.ops")
    spc_render = _mod("kaolin.render.spc", **{k: getattr(_SpcRender, k) for k in
                                              ("mark_pack_boundaries", "sum_reduce", "cumsum", "exponential_integration")})
    spc_ops = _mod("kaolin.ops.spc", unbatched_points_to_octree=_SpcOps.unbatched_points_to_octree,
                   unbatched_get_level_points=_SpcOps.unbatched_get_level_points)
    # `from wisp.models.grids import *` is how the reference's nefs obtain log/np/F/PerfTimer/BasicDecoder/spc_ops
    _mod("wisp.models.grids", BLASGrid=BLASGrid, HashGrid=HashGrid, OctreeGrid=_Unavailable,
         CodebookOctreeGrid=_Unavailable, TriplanarGrid=_Unavailable, log=logging, np=np, F=F, torch=torch, nn=nn,
         PerfTimer=wm.PerfTimer, BasicDecoder=wm.BasicDecoder, spc_ops=spc_ops, OctreeAS=OctreeAS)
    _mod("permutohedral_encoding", PermutoEncoding=PermutoEncoding)
    _mod("tinycudann", Encoding=TcnnEncoding)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return wisp, kaolin, spc_render


def import_reference_hash_grid_torch():
    """Import /root/reference/grids/hash_grid_torch.py verbatim; its import-time
    `torch.tensor(..., device='cuda')` (BOX_OFFSETS, :10-11) is redirected to the CPU."""
    install()
    import importlib
    real_tensor = torch.tensor

    def cpu_tensor(*a, **k):
        k.pop("device", None)
        return real_tensor(*a, **k)

    torch.tensor = cpu_tensor
    try:
        mod = importlib.import_module("grids.hash_grid_torch")
    finally:
        torch.tensor = real_tensor
    return mod
