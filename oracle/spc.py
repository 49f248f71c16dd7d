"""Oracle: structured-point-cloud (SPC) octree + packed-ray ops.  TEST INFRASTRUCTURE ONLY.

Restates (parity UNPINNED -- kaolin is not in /root/reference nor in this image):
  kaolin.ops.spc.unbatched_points_to_octree / scan_octrees / generate_points / unbatched_query
  kaolin.render.spc.unbatched_raytrace / mark_pack_boundaries / cumsum / sum_reduce /
  exponential_integration
as described in SURVEY.md Appendix A.1-A.3, A.5.  Call sites in the reference:
  grids/occtree.py:59-62,90   pc_nerf/panoptic_delta_nef.py:99
  tracers/panoptic_packed_rf_tracer.py:114,135,138,152,155,161,173,200

Integer results (octree bytes, point hierarchy, pidx, nugget lists) are the
bit-exact definition for csrc/octree.cu.  Float entry/exit depths follow the
exact float32 op order written here (see oracle.f32).
"""
import numpy as np
import torch

from .f32 import fma32

# ----------------------------------------------------------------------------------------------
# octree construction (A.1)
# ----------------------------------------------------------------------------------------------

def morton_encode(p, level):
    """p int[K,3] in [0,2^level) -> uint64 morton; child index j=(xbit<<2)|(ybit<<1)|zbit."""
    p = np.asarray(p).astype(np.uint64)
    code = np.zeros(p.shape[0], dtype=np.uint64)
    for b in range(level):
        sh = np.uint64(b)
        code |= ((p[:, 0] >> sh) & np.uint64(1)) << np.uint64(3 * b + 2)
        code |= ((p[:, 1] >> sh) & np.uint64(1)) << np.uint64(3 * b + 1)
        code |= ((p[:, 2] >> sh) & np.uint64(1)) << np.uint64(3 * b)
    return code


def morton_decode(code, level):
    code = np.asarray(code, dtype=np.uint64)
    p = np.zeros((code.shape[0], 3), dtype=np.int64)
    for b in range(level):
        p[:, 0] |= (((code >> np.uint64(3 * b + 2)) & np.uint64(1)) << np.uint64(b)).astype(np.int64)
        p[:, 1] |= (((code >> np.uint64(3 * b + 1)) & np.uint64(1)) << np.uint64(b)).astype(np.int64)
        p[:, 2] |= (((code >> np.uint64(3 * b)) & np.uint64(1)) << np.uint64(b)).astype(np.int64)
    return p.astype(np.int16)


_POPC8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.int32)


def points_to_octree(points, level):
    """unbatched_points_to_octree: int[K,3] leaf cells at `level` -> octree uint8[n_nodes]."""
    codes = np.unique(morton_encode(points, level))
    per_level = [None] * (level + 1)
    per_level[level] = codes
    for l in range(level, 0, -1):
        per_level[l - 1] = np.unique(per_level[l] >> np.uint64(3))
    octree = []
    for l in range(level):
        parents = per_level[l]
        child = per_level[l + 1]
        pos = np.searchsorted(parents, child >> np.uint64(3))
        byte = np.zeros(parents.shape[0], dtype=np.uint8)
        np.bitwise_or.at(byte, pos, (np.uint8(1) << (child & np.uint64(7)).astype(np.uint8)))
        octree.append(byte)
    return np.concatenate(octree) if octree else np.zeros(0, dtype=np.uint8)


def scan_octree(octree, level):
    """octree bytes -> (points int16[P,3], pyramid int32[2,level+2], prefix int32[n_nodes+1]).

    prefix = exclusive cumsum of popcount(octree) (with the total appended); the child j of
    node p is point index prefix[p] + popc(octree[p] & ((2<<j)-1)); the root is point 0.
    """
    octree = np.asarray(octree, dtype=np.uint8)
    popc = _POPC8[octree]
    prefix = np.zeros(octree.shape[0] + 1, dtype=np.int32)
    np.cumsum(popc, out=prefix[1:])
    codes = [np.zeros(1, dtype=np.uint64)]
    start = 0
    for l in range(level):
        n = codes[l].shape[0]
        byte = octree[start:start + n]
        start += n
        bits = (byte[:, None] >> np.arange(8, dtype=np.uint8)[None, :]) & 1  # [n,8], child j ascending
        parent = np.repeat(codes[l], 8).reshape(n, 8)
        ch = (parent << np.uint64(3)) | np.arange(8, dtype=np.uint64)[None, :]
        codes.append(ch[bits.astype(bool)])
    counts = np.array([c.shape[0] for c in codes], dtype=np.int32)
    pyramid = np.zeros((2, level + 2), dtype=np.int32)
    pyramid[0, : level + 1] = counts
    pyramid[1, 1: level + 2] = np.cumsum(counts)
    points = np.concatenate([morton_decode(c, l) for l, c in enumerate(codes)], axis=0)
    return points, pyramid, prefix


def dense_octree(level):
    """OctreeAS.init_dense(level): every cell occupied (grids/occtree.py:60)."""
    n_nodes = (8 ** level - 1) // 7
    return np.full(n_nodes, 0xFF, dtype=np.uint8)


def level_points(points, pyramid, level):
    """unbatched_get_level_points (grids/occtree.py:61)."""
    s = int(pyramid[1, level])
    return points[s: s + int(pyramid[0, level])]


# ----------------------------------------------------------------------------------------------
# query (A.3)
# ----------------------------------------------------------------------------------------------

def query(octree, prefix, coords, level):
    """coords f32[P,3] in [-1,1] -> point-hierarchy index int32[P] at `level`, -1 if empty/outside.

    q = floor(2^level * (x*0.5 + 0.5)) with separately rounded float32 mul/add.
    """
    x = np.asarray(coords, dtype=np.float32)
    res = np.float32(2 ** level)
    with np.errstate(all="ignore"):
        q = np.floor(res * (x * np.float32(0.5) + np.float32(0.5)))
    inside = np.all((q >= 0) & (q < res), axis=1) & np.all(np.isfinite(x), axis=1)
    qi = np.where(inside[:, None], q, 0).astype(np.int64)
    node = np.zeros(x.shape[0], dtype=np.int64)
    alive = inside.copy()
    octree = np.asarray(octree, dtype=np.uint8)
    for l in range(level):
        s = level - 1 - l
        j = (((qi[:, 0] >> s) & 1) << 2) | (((qi[:, 1] >> s) & 1) << 1) | ((qi[:, 2] >> s) & 1)
        byte = octree[np.where(alive, node, 0)].astype(np.int64)
        has = ((byte >> j) & 1).astype(bool)
        cnt = _POPC8[byte & ((2 << j) - 1)]
        alive &= has
        node = np.where(alive, prefix[np.where(alive, node, 0)].astype(np.int64) + cnt, 0)
    return np.where(alive, node, -1).astype(np.int32)


# ----------------------------------------------------------------------------------------------
# ray / AABB (A.2)
# ----------------------------------------------------------------------------------------------

def _ray_aabb(o, d, inv, sgn, vc, r):
    """Majercik-style efficient-slab test.  Returns (entry f32[K], exit f32[K]).

    entry: -1 if the origin is strictly inside the box, 0 on a miss, else first accepted slab
    distance; exit = min over axes of the far-plane distance (fmin: NaN ignored).
    """
    with np.errstate(all="ignore"):
        oo = (o - vc).astype(np.float32)
        cmax = np.maximum(np.maximum(np.abs(oo[:, 0]), np.abs(oo[:, 1])), np.abs(oo[:, 2]))
        inside = cmax < r
        rr = np.full_like(oo[:, 0], r)
        d0 = (fma32(rr, sgn[:, 0], -oo[:, 0]) * inv[:, 0]).astype(np.float32)
        d1 = (fma32(rr, sgn[:, 1], -oo[:, 1]) * inv[:, 1]).astype(np.float32)
        d2 = (fma32(rr, sgn[:, 2], -oo[:, 2]) * inv[:, 2]).astype(np.float32)
        lt0y = fma32(d[:, 1], d0, oo[:, 1]); lt0z = fma32(d[:, 2], d0, oo[:, 2])
        lt1x = fma32(d[:, 0], d1, oo[:, 0]); lt1z = fma32(d[:, 2], d1, oo[:, 2])
        lt2x = fma32(d[:, 0], d2, oo[:, 0]); lt2y = fma32(d[:, 1], d2, oo[:, 1])
        t0 = (d0 >= 0) & (np.abs(lt0y) < r) & (np.abs(lt0z) < r)
        t1 = (d1 >= 0) & (np.abs(lt1x) < r) & (np.abs(lt1z) < r)
        t2 = (d2 >= 0) & (np.abs(lt2x) < r) & (np.abs(lt2y) < r)
        entry = np.where(t0, d0, np.where(t1, d1, np.where(t2, d2, np.float32(0.0))))
        entry = np.where(inside, np.float32(-1.0), entry).astype(np.float32)
        e0 = (fma32(-rr, sgn[:, 0], -oo[:, 0]) * inv[:, 0]).astype(np.float32)
        e1 = (fma32(-rr, sgn[:, 1], -oo[:, 1]) * inv[:, 1]).astype(np.float32)
        e2 = (fma32(-rr, sgn[:, 2], -oo[:, 2]) * inv[:, 2]).astype(np.float32)
        exit_ = np.fmin(np.fmin(e0, e1), e2).astype(np.float32)
    return entry, exit_


def raytrace(octree, points, pyramid, prefix, origins, dirs, level):
    """unbatched_raytrace(..., return_depth=True, with_exit=True).

    Level-by-level refinement of (ridx, pidx) nuggets; children emitted in the order
    j = code ^ i (i = 0..7), code = octant of the ray ORIGIN w.r.t. the voxel centre.
    Returns ridx int32[K], pidx int32[K], depth f32[K,2] grouped by ray in input order.
    """
    o_all = np.asarray(origins, dtype=np.float32)
    d_all = np.asarray(dirs, dtype=np.float32)
    octree = np.asarray(octree, dtype=np.uint8)
    N = o_all.shape[0]
    with np.errstate(all="ignore"):
        inv_all = (np.float32(1.0) / d_all).astype(np.float32)
    sgn_all = np.where(np.signbit(d_all), np.float32(1.0), np.float32(-1.0)).astype(np.float32)
    ridx = np.arange(N, dtype=np.int64)
    pidx = np.zeros(N, dtype=np.int64)
    for l in range(level + 1):
        r = np.float32(1.0 / (1 << l))
        p = points[pidx].astype(np.float32)
        vc = (r * (np.float32(2.0) * p + np.float32(1.0)) - np.float32(1.0)).astype(np.float32)  # exact
        entry, exit_ = _ray_aabb(o_all[ridx], d_all[ridx], inv_all[ridx], sgn_all[ridx], vc, r)
        if l == level:
            keep = entry > 0
            ridx, pidx = ridx[keep], pidx[keep]
            depth = np.stack([entry[keep], exit_[keep]], axis=1).astype(np.float32)
            return ridx.astype(np.int32), pidx.astype(np.int32), depth
        hit = entry != 0
        ridx, pidx, p = ridx[hit], pidx[hit], p[hit]
        org = o_all[ridx]
        a = (np.float32(0.5) * org + np.float32(0.5)).astype(np.float32)
        b = (r * (p + np.float32(0.5))).astype(np.float32)
        rel = a - b
        code = ((rel[:, 0] > 0).astype(np.int64) << 2) | ((rel[:, 1] > 0).astype(np.int64) << 1) | (rel[:, 2] > 0).astype(np.int64)
        byte = octree[pidx].astype(np.int64)
        j = code[:, None] ^ np.arange(8, dtype=np.int64)[None, :]
        has = ((byte[:, None] >> j) & 1).astype(bool)
        cnt = _POPC8[byte[:, None] & ((2 << j) - 1)]
        child = prefix[pidx].astype(np.int64)[:, None] + cnt
        ridx = np.repeat(ridx, 8).reshape(-1, 8)[has]
        pidx = child[has]
    raise AssertionError


# ----------------------------------------------------------------------------------------------
# packed-ray ops (A.5) -- torch, differentiable
# ----------------------------------------------------------------------------------------------

def mark_pack_boundaries(ids: torch.Tensor) -> torch.Tensor:
    """b[0]=True, b[i] = ids[i] != ids[i-1]  (tracers/panoptic_packed_rf_tracer.py:114)."""
    b = torch.ones_like(ids, dtype=torch.bool)
    if ids.numel() > 1:
        b[1:] = ids[1:] != ids[:-1]
    return b


def _pack_ids(boundary: torch.Tensor) -> torch.Tensor:
    return torch.cumsum(boundary.to(torch.int64), 0) - 1


def cumsum(x: torch.Tensor, boundary: torch.Tensor, exclusive: bool = True) -> torch.Tensor:
    """Per-pack prefix sum of x[M,C]; accumulated in float64 so the oracle is order-independent."""
    if x.shape[0] == 0:
        return x.clone()
    ids = _pack_ids(boundary)
    xd = x.to(torch.float64)
    c = torch.cumsum(xd, 0)
    first = torch.nonzero(boundary).flatten()
    base = (c[first] - xd[first])[ids]
    c = c - base
    if exclusive:
        c = c - xd
    return c.to(x.dtype)


def sum_reduce(x: torch.Tensor, boundary: torch.Tensor) -> torch.Tensor:
    """[M,C] -> [R,C] per-pack sums (R = number of True in boundary)."""
    R = int(boundary.sum())
    out = torch.zeros(R, x.shape[1], dtype=torch.float64, device=x.device)
    if x.shape[0]:
        out = out.index_add(0, _pack_ids(boundary), x.to(torch.float64))
    return out.to(x.dtype)


def exponential_integration(feats, tau, boundary, exclusive=True):
    """kaolin.render.spc.exponential_integration: returns (sum_reduce(feats*w), w), w = T*alpha.

    The in-tree tracer passes an empty `feats` and only uses w
    (tracers/panoptic_packed_rf_tracer.py:135,152).
    """
    alpha = 1.0 - torch.exp(-tau)
    T = torch.exp(-cumsum(tau, boundary, exclusive=exclusive))
    w = T * alpha
    if feats is None or feats.numel() == 0:
        return feats, w
    return sum_reduce(feats * w, boundary), w
