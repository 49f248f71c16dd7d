"""Generate the golden fixtures in tests/golden/*.npz.   Run HERE (build container) only:

    python -m tests.golden.make_golden

Needs /root/reference (read-only).  Two kinds of golden:

1. hash_torch.npz -- the reference's own `grids/hash_grid_torch.py` (`hash`, `get_voxel_vertices`,
   `HashEmbedder`) imported VERBATIM and run on CPU: hashed vertex indices, features, gradients.
   This is the only hot-path arithmetic that lives in the reference tree, so it is the only
   "reference binary" golden.

2. trace_*.npz -- the reference's own `tracers/panoptic_packed_rf_tracer.py` and
   `pc_nerf/panoptic_{,delta_}nef.py` source run UNMODIFIED on stub wisp/kaolin modules
   (tests/golden/ref_stubs.py) whose third-party ops are the CPU oracle.  Pins the glue
   (channel gating, detach points, double integration, alpha-on-top, background, scatter).

The committed .npz files carry every input (rays, octree, parameters) and every output, so the
tests that consume them never touch /root/reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden import ref_stubs  # noqa: E402


def golden_hash_torch():
    mod = ref_stubs.import_reference_hash_grid_torch()
    torch.manual_seed(0)
    L, Fdim, T, base, fin = 4, 2, 10, 16, 128
    emb = mod.HashEmbedder(n_levels=L, n_features_per_level=Fdim, log2_hashmap_size=T,
                           base_resolution=base, finest_resolution=fin)
    with torch.no_grad():
        for e in emb.embeddings:
            e.weight.mul_(1e3)  # O(0.1) features so that float comparisons are meaningful
    g = torch.Generator().manual_seed(1)
    x = torch.rand(253, 3, generator=g) * 2 - 1
    edge = torch.tensor([[1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], [0.0, 0.0, 0.0], [1.2, -1.3, 0.5],
                         [0.999999, -0.999999, 0.125], [-1.0, 1.0, 0.0], [0.5, 0.5, 0.5]])
    x = torch.cat([x, edge], 0).requires_grad_(True)
    out = emb(x)
    gout = torch.randn(out.shape, generator=g)
    (out * gout).sum().backward()
    idx, res = [], []
    for i in range(L):
        r = torch.floor(emb.base_resolution * emb.b ** i)
        _, _, h, _ = mod.get_voxel_vertices(x.detach(), r, T)
        idx.append(h.numpy().astype(np.int64)); res.append(float(r))
    np.savez_compressed(os.path.join(HERE, "hash_torch.npz"),
                        x=x.detach().numpy(), out=out.detach().numpy(), gout=gout.numpy(),
                        grad_x=x.grad.numpy(),
                        weights=np.stack([e.weight.detach().numpy() for e in emb.embeddings]),
                        grad_weights=np.stack([e.weight.grad.numpy() for e in emb.embeddings]),
                        idx=np.stack(idx), resolutions=np.array(res, np.float32),
                        cfg=np.array([L, Fdim, T, base, fin]))
    print("hash_torch.npz", out.shape, np.stack(idx).shape, res)


def _scene(level, seed=0):
    """Seeded pruned occupancy: a slab |z|<0.3 plus random blobs, at octree `level`."""
    rng = np.random.default_rng(seed)
    n = 1 << level
    ijk = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 3)
    c = (ijk + 0.5) / n * 2 - 1
    occ = np.abs(c[:, 2]) < 0.3
    for _ in range(6):
        ctr = rng.uniform(-0.8, 0.8, 3)
        occ |= np.linalg.norm(c - ctr, axis=1) < 0.2
    return ijk[occ].astype(np.int16)


def _rays(N, seed=0):
    rng = np.random.default_rng(seed)
    o = np.stack([rng.uniform(-0.6, 0.6, N), rng.uniform(-0.6, 0.6, N), np.full(N, 0.9)], 1).astype(np.float32)
    tgt = np.stack([rng.uniform(-0.9, 0.9, N), rng.uniform(-0.9, 0.9, N), rng.uniform(-0.9, 0.0, N)], 1)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    # a few rays that miss everything (pointing up, out of the cube)
    d[:3] = np.array([0.0, 0.0, 1.0], np.float32)
    return o, d


def golden_trace(name, nef_type, grid_type, raymarch_type, num_steps, level=4, N=48, bg_color='white',
                 ray_max_travel=0.7):
    ref_stubs.install()
    if grid_type == "HashGridTorch":
        ref_stubs.import_reference_hash_grid_torch()
    from pc_nerf.panoptic_nef import PanopticNeF
    from pc_nerf.panoptic_delta_nef import PanopticDeltaNeF
    from tracers.panoptic_packed_rf_tracer import PanopticPackedRFTracer
    if nef_type == 'PanopticDDensityNeF':
        from pc_nerf.panoptic_dd_nef import PanopticDDensityNeF
        from tracers.panoptic_dd_packed_rf_tracer import PanopticDDensityPackedRFTracer
    from wisp.core import Rays
    import kaolin.ops.spc as spc_ops

    torch.manual_seed(0)
    L = 6
    kw = dict(grid_type=grid_type, interpolation_type='linear', multiscale_type='cat', feature_dim=2, num_lods=L,
              base_lod=2, hidden_dim=64, num_layers=1, view_multires=4, pos_multires=4, embedder_type='positional',
              activation_type='relu', layer_type='none', num_classes=6, num_instances=20,
              sem_num_layers=1, sem_hidden_dim=64, inst_num_layers=2, inst_hidden_dim=64,
              sem_softmax=True, inst_softmax=True, sem_detach=True, inst_detach=True,
              panoptic_features_type='delta' if nef_type != 'PanopticNeF' else None,
              blas_level=level, coarsest_scale=1.0, finest_scale=0.01, capacity_log_2=10, delta_capacity_log_2=9,
              codebook_bitwidth=10, inst_direct_pos=False)
    cls = {'PanopticDeltaNeF': PanopticDeltaNeF, 'PanopticNeF': PanopticNeF}.get(nef_type) or PanopticDDensityNeF
    nef = cls(**kw)
    if nef_type == 'PanopticDDensityNeF':
        nef.inst_direct_pos = False  # same latent attribute as the base class; the DD override never reads it
    if nef_type == 'PanopticNeF':
        nef.inst_direct_pos = False  # never assigned by the reference ctor
    grids = [nef.grid] + ([nef.delta_grid] if hasattr(nef, 'delta_grid') else [])
    for g in grids:
        if grid_type == "PermutoGrid":
            g.init_from_scales()
            with torch.no_grad():
                g.embedder.lattice_values.mul_(3e4)
        else:
            g.init_from_resolutions([16 * 2 ** i for i in range(L)])
            with torch.no_grad():
                g.embedder.params.mul_(3e3)
    pts = _scene(level)
    octree = spc_ops.unbatched_points_to_octree(torch.from_numpy(pts), level, sorted=True)
    for g in grids:
        if grid_type == "PermutoGrid":
            g.blas_init(octree)
        else:
            g.blas.init(octree)
    with torch.no_grad():  # make the field reasonably opaque so that alpha is not ~0
        nef.decoder_density.lout.bias[0] = 6.0
    tcls = PanopticDDensityPackedRFTracer if nef_type == 'PanopticDDensityNeF' else PanopticPackedRFTracer
    tracer = tcls(raymarch_type=raymarch_type, num_steps=num_steps, bg_color=bg_color, ray_max_travel=ray_max_travel)
    o, d = _rays(N)
    o_t = torch.from_numpy(o).requires_grad_(True)
    d_t = torch.from_numpy(d).requires_grad_(True)
    rays = Rays(origins=o_t, dirs=d_t, dist_min=0.0, dist_max=2.0)
    channels = ['rgb', 'depth', 'semantics', 'inst_embedding']
    rb = tracer(nef, channels=channels, rays=rays, lod_idx=None, stage='train')
    g = torch.Generator().manual_seed(7)
    gw = {c: torch.randn(getattr(rb, c).shape, generator=g) for c in channels}
    loss = sum((getattr(rb, c) * gw[c]).sum() for c in channels)
    loss.backward()
    out = {"o": o, "d": d, "octree": octree.numpy(), "level": np.array(level), "num_steps": np.array(num_steps),
           "bg_white": np.array(bg_color == 'white'), "ray_max_travel": np.array(ray_max_travel, np.float32),
           "jitter_seed": np.array(0), "grad_o": o_t.grad.numpy(), "grad_d": d_t.grad.numpy()}
    for c in channels + ['alpha']:
        out["out_" + c] = getattr(rb, c).detach().numpy()
    out["out_hit"] = rb.hit.numpy()
    for c in channels:
        out["gw_" + c] = gw[c].numpy()
    for k, v in nef.state_dict().items():
        out["param:" + k] = v.detach().numpy()
    for k, p in nef.named_parameters():
        out["grad:" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {c: tuple(getattr(rb, c).shape) for c in channels}, "hit", int(rb.hit.sum()), "/", N,
          "alpha mean", float(rb.alpha.mean()))


if __name__ == "__main__":
    golden_hash_torch()
    golden_trace("trace_delta_permuto_ray", "PanopticDeltaNeF", "PermutoGrid", "ray", 48)
    golden_trace("trace_delta_permuto_voxel", "PanopticDeltaNeF", "PermutoGrid", "voxel", 3, bg_color='black')
    golden_trace("trace_nef_tcnn_ray", "PanopticNeF", "HashGridTinyCudaNN", "ray", 32)
    golden_trace("trace_dd_permuto_ray", "PanopticDDensityNeF", "PermutoGrid", "ray", 40)
