"""Oracle: instant-ngp multiresolution hash grids.  TEST INFRASTRUCTURE ONLY.

Two flavours, both on the reference hot path:

* `HashEmbedderOracle` -- restatement of the reference's own torch hash grid,
  grids/hash_grid_torch.py:13-108 (`hash` :13-24, `get_voxel_vertices` :26-46,
  `HashEmbedder` :48-108).  PINNED: tests/golden/make_golden.py imports that file verbatim and
  tests/test_oracle_golden.py checks this class against it (indices bit-exact, features 1e-6).

* `TcnnHashGridOracle` -- restatement of tiny-cuda-nn's `GridEncoding` (HashGrid, Linear
  interpolation) as the reference configures it at grids/hash_grid_tinycudann.py:24-34 and
  calls it at :41.  tiny-cuda-nn is an unpinned, un-vendored dependency: parity UNPINNED,
  algorithm as recalled in SURVEY.md Appendix A.7 (grid.h `kernel_grid`, `grid_index`,
  `pos_fract`, coherent-prime hash {1, 2654435761, 805459861}).  Coordinates arrive in [-1,1]
  and negative cells wrap through (uint32)(int) exactly like upstream would.
"""
import numpy as np
import torch
from torch import nn

from .f32 import fma32

PRIMES = (1, 2654435761, 805459861)


# ------------------------------------------------------------------------------------------
# HashNeRF / reference torch flavour
# ------------------------------------------------------------------------------------------

def hashnerf_resolutions(n_levels, base_resolution, finest_resolution):
    """floor(base * b**i) exactly as grids/hash_grid_torch.py:55-59,99 computes it (torch float32)."""
    base = torch.tensor(base_resolution)
    fin = torch.tensor(finest_resolution)
    b = torch.exp((torch.log(fin) - torch.log(base)) / (n_levels - 1))
    return [float(torch.floor(base * b ** i)) for i in range(n_levels)]


def hashnerf_level_indices(x, resolution, log2_T):
    """x f32 tensor [M,3] -> (bottom_left int32 [M,3], hashed idx int64 [M,8]); corner order x-major."""
    box_min, box_max = -1.0, 1.0
    xc = torch.clamp(x, min=box_min, max=box_max)
    grid_size = torch.tensor((box_max - box_min), dtype=torch.float32) / torch.tensor(resolution, dtype=torch.float32)
    bl = torch.floor((xc - box_min) / grid_size).int()
    offs = torch.tensor([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], dtype=torch.int32)
    v = (bl[:, None, :] + offs[None]).to(torch.int64) & 0xFFFFFFFF
    h = (v[..., 0] * PRIMES[0]) ^ ((v[..., 1] * PRIMES[1]) & 0xFFFFFFFF) ^ ((v[..., 2] * PRIMES[2]) & 0xFFFFFFFF)
    return bl, h & ((1 << log2_T) - 1), grid_size


class HashEmbedderOracle(nn.Module):
    def __init__(self, n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                 base_resolution=16, finest_resolution=512, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.n_levels, self.F, self.log2_T = n_levels, n_features_per_level, log2_hashmap_size
        self.resolutions = hashnerf_resolutions(n_levels, base_resolution, finest_resolution)
        w = (torch.rand(n_levels, 2 ** log2_hashmap_size, n_features_per_level, generator=g) * 2 - 1) * 1e-4
        self.embeddings = nn.Parameter(w)
        self.out_dim = n_levels * n_features_per_level

    def indices(self, x):
        return torch.stack([hashnerf_level_indices(x.detach().float(), r, self.log2_T)[1] for r in self.resolutions])

    def forward(self, x):
        outs = []
        for l, res in enumerate(self.resolutions):
            bl, idx, grid_size = hashnerf_level_indices(x.detach().float(), res, self.log2_T)
            grid_size = grid_size.to(x.dtype)
            vmin = bl.to(x.dtype) * grid_size + (-1.0)
            vmax = vmin + 1.0 * grid_size
            w = (x - vmin) / (vmax - vmin)
            e = self.embeddings[l].to(x.dtype)[idx]  # [M,8,F]
            wx, wy, wz = w[:, 0:1], w[:, 1:2], w[:, 2:3]
            c00 = e[:, 0] * (1 - wx) + e[:, 4] * wx
            c01 = e[:, 1] * (1 - wx) + e[:, 5] * wx
            c10 = e[:, 2] * (1 - wx) + e[:, 6] * wx
            c11 = e[:, 3] * (1 - wx) + e[:, 7] * wx
            c0 = c00 * (1 - wy) + c10 * wy
            c1 = c01 * (1 - wy) + c11 * wy
            outs.append(c0 * (1 - wz) + c1 * wz)
        return torch.cat(outs, dim=-1)


# ------------------------------------------------------------------------------------------
# tiny-cuda-nn flavour
# ------------------------------------------------------------------------------------------

def tcnn_level_table(n_levels, log2_T, base_resolution, per_level_scale=2.0):
    """Per-level (scale f32, resolution u32, offset u32 [entries], hashmap_size u32)."""
    scales, ress, offs, sizes = [], [], [], []
    off = 0
    l2 = np.float32(np.log2(np.float32(per_level_scale)))
    for l in range(n_levels):
        scale = np.float32(np.exp2(np.float32(l) * l2) * np.float32(base_resolution) - np.float32(1.0))
        res = int(np.ceil(scale)) + 1
        n = res ** 3
        n = min(n, 2 ** 32 - 8)
        n = (n + 7) // 8 * 8
        n = min(n, 1 << log2_T)
        scales.append(scale); ress.append(res); offs.append(off); sizes.append(n)
        off += n
    return (np.array(scales, np.float32), np.array(ress, np.uint32), np.array(offs, np.uint32),
            np.array(sizes, np.uint32), off)


def tcnn_level_indices(x32, scale, res, size, ft=np.float32):
    """x [M,3] -> (w ft[M,3], idx u32[M,8]); corner c: bit dim of c set -> +1 on that dim.
    ft=float32: bit-exact definition (pos = fmaf(scale, x, 0.5)); float64 only to gradcheck the oracle."""
    with np.errstate(all="ignore"):
        pos = fma32(np.float32(scale), x32, np.float32(0.5)) if ft is np.float32 else (float(scale) * x32 + 0.5)
        cell = np.floor(pos)
        w = (pos - cell).astype(ft)
        pg = (cell.astype(np.int64) & 0xFFFFFFFF).astype(np.uint32)  # (uint32_t)(int)floorf(pos)
    M = x32.shape[0]
    idx = np.zeros((M, 8), dtype=np.uint32)
    hashed = int(res) ** 3 > int(size)
    with np.errstate(over="ignore"):
        for c in range(8):
            pl = [pg[:, dmn] + np.uint32((c >> dmn) & 1) for dmn in range(3)]
            if hashed:
                k = (pl[0] * np.uint32(PRIMES[0])) ^ (pl[1] * np.uint32(PRIMES[1])) ^ (pl[2] * np.uint32(PRIMES[2]))
            else:
                k = pl[0] + pl[1] * np.uint32(res) + pl[2] * np.uint32((int(res) * int(res)) & 0xFFFFFFFF)
            idx[:, c] = k % np.uint32(size)
    return w, idx


class TcnnHashGridOracle(nn.Module):
    def __init__(self, n_levels=14, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16,
                 per_level_scale=2.0, seed=0, out_half=False):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.n_levels, self.F = n_levels, n_features_per_level
        (self.scales, self.ress, self.offs, self.sizes, total) = tcnn_level_table(
            n_levels, log2_hashmap_size, base_resolution, per_level_scale)
        self.params = nn.Parameter((torch.rand(total * n_features_per_level, generator=g) * 2 - 1) * 1e-4)
        self.out_half = out_half

    def indices(self, x):
        x32 = x.detach().cpu().numpy().astype(np.float32)
        return [tcnn_level_indices(x32, self.scales[l], self.ress[l], self.sizes[l])[1] for l in range(self.n_levels)]

    def forward(self, x):
        dt = x.dtype
        ft = np.float64 if dt == torch.float64 else np.float32
        x32 = x.detach().cpu().numpy().astype(ft)
        table = self.params.view(-1, self.F).to(dt)
        outs = []
        for l in range(self.n_levels):
            w32, idx = tcnn_level_indices(x32, self.scales[l], self.ress[l], self.sizes[l], ft)
            lin = x * float(self.scales[l])
            w = torch.from_numpy(w32).to(dt) + (lin - lin.detach())
            acc = 0
            for c in range(8):
                wt = 1.0
                for dmn in range(3):
                    wt = wt * (w[:, dmn] if (c >> dmn) & 1 else (1 - w[:, dmn]))
                e = table[torch.from_numpy((idx[:, c].astype(np.int64) + int(self.offs[l])))]
                acc = acc + wt[:, None] * e
            outs.append(acc)
        out = torch.cat(outs, dim=-1)
        if self.out_half:
            out = out.half().to(dt)
        return out
