"""Oracle: wisp v0.1.1 OctreeAS.raymarch ('ray' and 'voxel').  TEST INFRASTRUCTURE ONLY.

Restates SURVEY.md Appendix A.4 (parity UNPINNED: kaolin-wisp is absent).  Reference call
sites: grids/occtree.py:85-91, tracers/panoptic_packed_rf_tracer.py:85-108.

The upstream code draws jitter from torch.rand; here the jitter is an explicit input
(`jitter` array) or the counter-based stream oracle.f32.jitter_u01(seed, flat_index) that
csrc/octree.cu reproduces bit for bit.
"""
import numpy as np
import torch

from . import spc
from .f32 import jitter_u01


def linspace01(S: int) -> np.ndarray:
    """torch.linspace(0, 1, S) in float32 -- the host computes it once, the kernel reads it."""
    return torch.linspace(0.0, 1.0, S, dtype=torch.float32).numpy().copy()


def raymarch_ray(octree, prefix, origins, dirs, level, num_samples, dist_min, dist_max,
                 jitter=None, seed=0):
    """'ray' mode: S jittered steps per ray between dist_min/dist_max, kept where the octree is occupied.

    Returns ridx i32[M], pidx i32[M], samples f32[M,1,3], depths f32[M,1], deltas f32[M,1], boundary bool[M].
    """
    o = np.asarray(origins, dtype=np.float32)
    d = np.asarray(dirs, dtype=np.float32)
    N, S = o.shape[0], int(num_samples)
    if jitter is None:
        jitter = jitter_u01(seed, np.arange(N * S, dtype=np.uint64)).reshape(N, S)
    jitter = np.asarray(jitter, dtype=np.float32)
    lin = linspace01(S)
    depth = (lin[None, :] + jitter / np.float32(S)).astype(np.float32)
    depth = (depth * np.float32(dist_max - dist_min)).astype(np.float32)
    depth = (depth + np.float32(dist_min)).astype(np.float32)
    samples = (o[:, None, :] + (d[:, None, :] * depth[:, :, None]).astype(np.float32)).astype(np.float32)
    prev = np.concatenate([np.full((N, 1), np.float32(dist_min), dtype=np.float32), depth[:, :-1]], axis=1)
    deltas = (depth - prev).astype(np.float32)
    pidx = spc.query(octree, prefix, samples.reshape(-1, 3), level).reshape(N, S)
    mask = pidx > -1
    ridx = np.broadcast_to(np.arange(N, dtype=np.int32)[:, None], (N, S))[mask]
    boundary = spc.mark_pack_boundaries(torch.from_numpy(ridx.copy())).numpy()
    return (ridx.astype(np.int32), pidx[mask].astype(np.int32), samples[mask][:, None, :],
            depth[mask][:, None], deltas[mask].reshape(-1, 1), boundary)


def raymarch_voxel(octree, points, pyramid, prefix, origins, dirs, level, num_samples,
                   jitter=None, seed=0):
    """'voxel' mode: S jittered samples inside every intersected level-`level` cell.

    Returns ridx i32[K], pidx i32[K], samples f32[K,S,3], depths f32[K,S,1], deltas f32[K*S,1],
    boundary bool[K*S] (True at sample 0 of the first nugget of each ray).
    """
    o = np.asarray(origins, dtype=np.float32)
    d = np.asarray(dirs, dtype=np.float32)
    S = int(num_samples)
    ridx, pidx, depth = spc.raytrace(octree, points, pyramid, prefix, o, d, level)
    K = ridx.shape[0]
    if jitter is None:
        jitter = jitter_u01(seed, np.arange(K * S, dtype=np.uint64)).reshape(K, S)
    jitter = np.asarray(jitter, dtype=np.float32)
    steps = ((np.arange(S, dtype=np.float32)[None, :] + jitter).astype(np.float32) / np.float32(S)).astype(np.float32)
    t0, t1 = depth[:, 0:1], depth[:, 1:2]
    ds = (t0 + ((t1 - t0).astype(np.float32) * steps).astype(np.float32)).astype(np.float32)  # [K,S]
    prev = np.concatenate([t0, ds[:, :-1]], axis=1)
    deltas = (ds - prev).astype(np.float32).reshape(-1, 1)
    samples = (o[ridx][:, None, :] + (d[ridx][:, None, :] * ds[:, :, None]).astype(np.float32)).astype(np.float32)
    nb = spc.mark_pack_boundaries(torch.from_numpy(ridx.astype(np.int32))).numpy()
    boundary = np.zeros((K, S), dtype=bool)
    boundary[:, 0] = nb
    return ridx, pidx, samples, ds[:, :, None], deltas, boundary.reshape(-1)


def max_travel_filter(ridx, depths, ray_max_travel):
    """tracers/panoptic_packed_rf_tracer.py:88-99: keep nuggets whose first-sample depth is within
    ray_max_travel of the ray's first nugget's first-sample depth.  Returns bool[K]."""
    ridx = np.asarray(ridx)
    d0 = np.asarray(depths, dtype=np.float32)[:, 0, 0]
    if ridx.shape[0] == 0:
        return np.zeros(0, dtype=bool)
    first = np.ones(ridx.shape[0], dtype=bool)
    first[1:] = ridx[1:] != ridx[:-1]
    start = np.maximum.accumulate(np.where(first, np.arange(ridx.shape[0]), 0))
    travelled = (d0 - d0[start]).astype(np.float32)
    return travelled < np.float32(ray_max_travel)
