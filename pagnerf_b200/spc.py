"""Structured-point-cloud occupancy octree: data layout + accel-struct holder.

Host-side mirror of wisp.accelstructs.OctreeAS / kaolin.ops.spc as the reference uses them
(grids/occtree.py:59-78, pc_nerf/panoptic_delta_nef.py:99-104).  Layout:
  octree  u8 [n_nodes]        one byte per non-leaf node, bit j = child j, j = (x<<2)|(y<<1)|z,
                              nodes level by level, Morton order inside a level
  points  i16[n_points,3]     integer coords of every node of every level (root first)
  pyramid i32[2, L+2]         row 0: #points per level; row 1: start offset of each level
  prefix  i32[n_nodes+1]      exclusive cumsum of popcount(octree); child j of node p is point
                              prefix[p] + popc(octree[p] & ((2<<j)-1)); the root is point 0
Construction is the cold path (runs at init and every `prune_every` epochs) and is written with
device-agnostic torch integer ops; traversal / marching are the CUDA kernels in csrc/octree.cu.
"""
import torch

from . import ops

_POPC8 = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int32)


def _popc(byte):
    return _POPC8.to(byte.device)[byte.long()]


def morton_encode(points, level):
    p = points.long()
    code = torch.zeros(p.shape[0], dtype=torch.int64, device=p.device)
    for b in range(level):
        code |= ((p[:, 0] >> b) & 1) << (3 * b + 2)
        code |= ((p[:, 1] >> b) & 1) << (3 * b + 1)
        code |= ((p[:, 2] >> b) & 1) << (3 * b)
    return code


def morton_decode(code, level):
    p = torch.zeros(code.shape[0], 3, dtype=torch.int64, device=code.device)
    for b in range(level):
        p[:, 0] |= ((code >> (3 * b + 2)) & 1) << b
        p[:, 1] |= ((code >> (3 * b + 1)) & 1) << b
        p[:, 2] |= ((code >> (3 * b)) & 1) << b
    return p.to(torch.int16)


def unbatched_points_to_octree(points, level, sorted=False):
    """kaolin.ops.spc.unbatched_points_to_octree: leaf cells int[K,3] at `level` -> octree uint8."""
    codes = torch.unique(morton_encode(points, level))
    per_level = [None] * (level + 1)
    per_level[level] = codes
    for l in range(level, 0, -1):
        per_level[l - 1] = torch.unique(per_level[l] >> 3)
    out = []
    for l in range(level):
        parents, child = per_level[l], per_level[l + 1]
        pos = torch.searchsorted(parents, child >> 3)
        byte = torch.zeros(parents.shape[0], dtype=torch.int64, device=points.device)
        byte.index_add_(0, pos, torch.ones_like(child) << (child & 7))  # children unique -> sum == OR
        out.append(byte.to(torch.uint8))
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.uint8, device=points.device)


def scan_octree(octree, level):
    """-> points i16[P,3], pyramid i32[2,level+2] (CPU), prefix i32[n_nodes+1]."""
    dev = octree.device
    popc = _popc(octree)
    prefix = torch.zeros(octree.shape[0] + 1, dtype=torch.int32, device=dev)
    prefix[1:] = torch.cumsum(popc, 0)
    codes = [torch.zeros(1, dtype=torch.int64, device=dev)]
    start = 0
    ar = torch.arange(8, device=dev)
    for l in range(level):
        n = codes[l].shape[0]
        byte = octree[start:start + n].long()
        start += n
        bits = ((byte[:, None] >> ar[None, :]) & 1).bool()
        ch = (codes[l][:, None] << 3) | ar[None, :]
        codes.append(ch[bits])
    counts = torch.tensor([c.shape[0] for c in codes], dtype=torch.int32)
    pyramid = torch.zeros(2, level + 2, dtype=torch.int32)
    pyramid[0, :level + 1] = counts
    pyramid[1, 1:level + 2] = torch.cumsum(counts, 0)
    points = torch.cat([morton_decode(c, l) for l, c in enumerate(codes)], 0)
    return points, pyramid, prefix


def octree_from_mask(mask, level):
    """Dense leaf-occupancy mask bool[8^level] in Morton order (the order of `dense_points`) on a CUDA device ->
    (octree u8[n_nodes], points i16[P,3], pyramid i32[2,level+2] (CPU), prefix i32[n_nodes+1]) with one kernel chain
    (csrc/octree.cu pag_octree_from_mask) instead of unique / searchsorted / ~900 small integer ops; one 4 * (level+2)-word
    D2H read for the sizes (cold path: runs once per prune)."""
    dev = mask.device
    F = (8 ** (level + 1) - 1) // 7
    m = mask.to(torch.uint8).contiguous()
    assert m.numel() == 8 ** level
    i32, i64, u8 = torch.int32, torch.int64, torch.uint8
    exists = torch.empty(F, dtype=i32, device=dev)
    byts = torch.empty(F, dtype=u8, device=dev)
    pos = torch.empty(F + 1, dtype=i64, device=dev)
    popc = torch.empty(F, dtype=i32, device=dev)
    prefix64 = torch.empty(F + 1, dtype=i64, device=dev)
    octree = torch.empty(F, dtype=u8, device=dev)
    points = torch.empty(F, 3, dtype=torch.int16, device=dev)
    prefix = torch.empty(F + 1, dtype=i32, device=dev)
    pyramid = torch.empty(2, level + 2, dtype=i32, device=dev)
    ops.call("pag_octree_from_mask", ops.ptr(m), int(level), ops.ptr(exists), ops.ptr(byts), ops.ptr(pos), ops.ptr(popc), ops.ptr(prefix64),
             ops.ptr(octree), ops.ptr(points), ops.ptr(prefix), ops.ptr(pyramid))
    pyr = pyramid.cpu()
    n_nodes, n_points = int(pyr[1, level]), int(pyr[1, level + 1])
    return octree[:n_nodes].clone(), points[:n_points].clone(), pyr, prefix[:n_nodes + 1].clone()


def unbatched_get_level_points(points, pyramid, level):
    s = int(pyramid[1, level])
    return points[s:s + int(pyramid[0, level])]


def octree_max_level(octree):
    """Recover the depth of an octree byte string by walking the per-level popcounts."""
    n, acc, cnt, level = octree.shape[0], 0, 1, 0
    popc = _popc(octree).cpu()
    while acc < n:
        nxt = int(popc[acc:acc + cnt].sum())
        acc += cnt
        cnt = nxt
        level += 1
    return level


class OctreeAS:
    """Occupancy-octree acceleration structure (wisp.accelstructs.OctreeAS surface)."""

    def __init__(self, device=None):
        self.device = torch.device(device) if device is not None else torch.device(
            'cuda' if torch.cuda.is_available() else 'cpu')
        self.octree = self.points = self.pyramid = self.prefix = None
        self.max_level = None
        self.jitter_seed = 0       # seed of the counter-based jitter stream; bumped per raymarch call
        self.fixed_jitter = False  # tests pin the stream
        self._bits = {}            # level -> occupancy bit field (see level_bits)
        self.version = 0           # bumped by every init(): consumers that cache device pointers (graph.GraphedStep) compare it

    def init(self, octree):
        # the scan is ~40 small integer ops per level: on the tensor's own device (an octree that arrives on the host -- a
        # checkpoint, a golden -- is scanned there and moved once, instead of ~900 tiny launches on the GPU)
        octree = octree.to(torch.uint8).contiguous()
        level = octree_max_level(octree)
        points, self.pyramid, prefix = scan_octree(octree, level)
        self.points = points.contiguous().to(self.device)
        self.prefix = prefix.contiguous().to(self.device)
        octree = octree.to(self.device)
        self.octree = octree
        self.max_level = level
        self.version += 1
        self._refresh_bits()

    def _refresh_bits(self):
        # occupancy bit fields have a fixed size per level: refresh the cached ones IN PLACE, so that a CUDA graph that
        # captured their address (the fused training trace marches against them) sees the pruned octree on its next replay
        octree, level = self.octree, self.max_level
        for lvl, b in list(self._bits.items()):
            if b.device == octree.device and b.is_cuda and lvl <= level:
                ops.call("pag_octree_level_bits", ops.ptr(self.octree), ops.ptr(self.prefix), lvl, ops.ptr(b))
            else:
                del self._bits[lvl]

    def init_from_mask(self, mask, level):
        """Rebuild from the dense leaf-occupancy mask (Morton order) on the device -- the prune() path."""
        self.device = mask.device
        self.octree, self.points, self.pyramid, self.prefix = octree_from_mask(mask, level)
        self.max_level = level
        self.version += 1
        self._refresh_bits()

    def init_dense(self, level):
        n_nodes = (8 ** level - 1) // 7
        self.init(torch.full((n_nodes,), 0xFF, dtype=torch.uint8))      # scanned on the host, moved once

    def to(self, device):
        device = torch.device(device)
        if self.octree is not None and self.octree.device != device:
            self.octree = self.octree.to(device)
            self.points = self.points.to(device)
            self.prefix = self.prefix.to(device)
        self.device = device
        return self

    def level_bits(self, level):
        """Occupancy bit field of `level` (int32 words, bit (ix*res + iy)*res + iz), derived from the octree on first use and
        cached until the octree changes; the fused training trace marches against it (csrc/octree.cu)."""
        level = int(level)
        b = self._bits.get(level)
        if b is None or b.device != self.octree.device:
            b = torch.empty((8 ** level) // 32, dtype=torch.int32, device=self.octree.device)
            ops.call("pag_octree_level_bits", ops.ptr(self.octree), ops.ptr(self.prefix), level, ops.ptr(b))
            self._bits[level] = b
        return b

    def query(self, coords, level=None):
        lvl = self.max_level if level is None else level
        return ops.octree_query(self.octree, self.prefix, coords, lvl).long()

    def raytrace(self, rays, level=None, with_exit=True):
        lvl = self.max_level if level is None else level
        ridx, pidx, depth, _ = ops.raytrace(self.octree, self.prefix, rays.origins, rays.dirs, lvl)
        return ridx, pidx, (depth if with_exit else depth[:, :1])

    def raymarch(self, rays, level=None, num_samples=64, raymarch_type='voxel', need_pidx=True):
        """-> (ridx, pidx, samples, depths, deltas, boundary) with the shapes of wisp v0.1.1
        (SURVEY Appendix A.4).  Sample positions stay attached to rays.origins / rays.dirs for
        pose optimisation (pc_nerf/ba_pipeline.py:49-51).
        need_pidx=False ('ray' mode; callers whose grid does not index features by octree point): pidx is None and the march
        runs against the occupancy bit field -- the same samples from a 4-5x cheaper kernel pair."""
        self.to(rays.origins.device)
        lvl = self.max_level if level is None else level
        seed = self.jitter_seed
        if not self.fixed_jitter:
            self.jitter_seed = (self.jitter_seed + 1) & 0x7FFFFFFF
        dmin = float(rays.dist_min) if not torch.is_tensor(rays.dist_min) else float(rays.dist_min.flatten()[0])
        dmax = float(rays.dist_max) if not torch.is_tensor(rays.dist_max) else float(rays.dist_max.flatten()[0])
        if raymarch_type == 'voxel':
            ridx, pidx, samples, depths, deltas, boundary, offsets = ops.raymarch_voxel(
                self.octree, self.prefix, rays.origins, rays.dirs, lvl, num_samples, seed=seed)
            row_offsets = offsets * int(num_samples)
        elif raymarch_type == 'ray':
            if not need_pidx and 2 <= lvl <= 8:          # bit field of level 8: 2 MB
                ridx, pidx, samples, depths, deltas, boundary, offsets = ops.raymarch_ray_bits(
                    self.level_bits(lvl), rays.origins, rays.dirs, lvl, num_samples, dmin, dmax, seed=seed)
            else:
                ridx, pidx, samples, depths, deltas, boundary, offsets = ops.raymarch_ray(
                    self.octree, self.prefix, rays.origins, rays.dirs, lvl, num_samples, dmin, dmax, seed=seed)
            row_offsets = offsets
        else:
            raise TypeError(f"raymarch type {raymarch_type} is wrong, use 'voxel' or 'ray'")
        if rays.origins.requires_grad or rays.dirs.requires_grad:
            samples = ops.RaySamplesFn.apply(rays.origins, rays.dirs, samples, depths, row_offsets)
        return ridx, pidx, samples, depths, deltas, boundary
