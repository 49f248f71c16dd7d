"""Oracle: multi-resolution permutohedral-lattice encoding.  TEST INFRASTRUCTURE ONLY.

Restates the published algorithm of R. A. Rosu's `permutohedral_encoding` (unpinned dependency,
README.md:45 of the reference; imported at grids/permuto_grid.py:10, built :57-62, called :71)
as recalled in SURVEY.md Appendix A.6.  Parity UNPINNED (package absent from /root/reference and
from this image).  Definitions fixed here, mirrored by csrc/permuto.cu:

  cf_i        = (pos_i + shift[l,i]) * scale_factor[l,i]              (f32 add, then f32 mul)
  elevated[i] = fma(-i, cf_{i-1}, sm); sm += cf_{i-1}  (i = 3,2,1);  elevated[0] = sm
  rem0_i      = nearest multiple of 4 (ties -> down: `up - e < e - down ? up : down`)
  rank        = pairwise ordering of elevated - rem0 (strict <), += sum(rem0)/4, wrapped to [0,3]
  bary        = Adams et al. 2010 p.10;  key_i = rem0_i + r - (rank_i > 3-r ? 4 : 0), i<3
  hash        = k=0; for i<3: k += key_i; k *= 2531011 (uint32);  idx = k % capacity
  out[m, l*F+f] = anneal[l] * sum_r bary[r] * table[l, idx_r, f]

rem0 / rank / key / idx are the BIT-EXACT targets; bary and outputs are float (1e-4 relative).
"""
import numpy as np
import torch
from torch import nn

from .f32 import fma32

POS_DIM = 3
HASH_MUL = np.uint32(2531011)


def scale_factor_table(scales):
    """scale_factor[l,i] = 1/sqrt((i+1)(i+2)) / scales[l]  (float32 [L,3])."""
    scales = np.asarray(scales, dtype=np.float64)
    sf = np.zeros((scales.shape[0], POS_DIM), dtype=np.float64)
    for i in range(POS_DIM):
        sf[:, i] = 1.0 / np.sqrt((i + 1) * (i + 2)) / scales
    return sf.astype(np.float32)


def lattice_level(pos, sf_l, shift_l, capacity, ft=np.float32):
    """One level.  pos [M,3] -> (elevated ft[M,4], rem0 i32[M,4], rank i32[M,4], idx u32[M,4]).

    ft=float32 is the bit-exact definition (separately rounded ops + one fma per elevated[i]);
    ft=float64 runs the same recipe in double (only used to gradcheck the oracle itself).
    """
    pos = np.asarray(pos, dtype=ft)
    M = pos.shape[0]
    with np.errstate(all="ignore"):
        cf = ((pos + shift_l[None, :].astype(ft)).astype(ft) * sf_l[None, :].astype(ft)).astype(ft)
        elevated = np.zeros((M, 4), dtype=ft)
        sm = np.zeros(M, dtype=ft)
        for i in range(POS_DIM, 0, -1):
            elevated[:, i] = fma32(np.float32(-i), cf[:, i - 1], sm) if ft is np.float32 else (-i * cf[:, i - 1] + sm)
            sm = (sm + cf[:, i - 1]).astype(ft)
        elevated[:, 0] = sm
        v = elevated * ft(0.25)
        up = (np.ceil(v) * ft(4.0)).astype(ft)
        down = (np.floor(v) * ft(4.0)).astype(ft)
        rem0f = np.where((up - elevated).astype(ft) < (elevated - down).astype(ft), up, down)
    rem0 = rem0f.astype(np.int32)
    s = rem0.astype(np.int64).sum(axis=1)
    s = (np.sign(s) * (np.abs(s) // 4)).astype(np.int32)  # C integer division (exact anyway)
    diff = (elevated - rem0.astype(ft)).astype(ft)
    rank = np.zeros((M, 4), dtype=np.int32)
    for i in range(POS_DIM):
        for j in range(i + 1, POS_DIM + 1):
            lt = diff[:, i] < diff[:, j]
            rank[:, i] += lt
            rank[:, j] += ~lt
    rank += s[:, None]
    lo = rank < 0
    hi = rank > POS_DIM
    rank = rank + 4 * lo - 4 * hi
    rem0 = rem0 + 4 * lo - 4 * hi
    idx = np.zeros((M, 4), dtype=np.uint32)
    with np.errstate(over="ignore"):
        for r in range(4):
            k = np.zeros(M, dtype=np.uint32)
            for i in range(POS_DIM):
                key = rem0[:, i] + r - 4 * (rank[:, i] > POS_DIM - r)
                k = (k + key.astype(np.uint32)) * HASH_MUL
            idx[:, r] = k % np.uint32(capacity)
    return elevated, rem0.astype(np.int32), rank.astype(np.int32), idx


class PermutoEncodingOracle(nn.Module):
    """PermutoEncoding(pos_dim=3, capacity, nr_levels, nr_feat_per_level, scale_per_level)."""

    def __init__(self, capacity, nr_levels, nr_feat, scales, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.capacity, self.nr_levels, self.nr_feat = int(capacity), int(nr_levels), int(nr_feat)
        self.lattice_values = nn.Parameter(torch.randn(nr_levels, capacity, nr_feat, generator=g) * 1e-5)
        self.register_buffer("random_shift_per_level", torch.randn(nr_levels, 3, generator=g) * 10.0)
        self.register_buffer("scale_factor", torch.from_numpy(scale_factor_table(scales)))
        self.register_buffer("anneal_window", torch.ones(nr_levels))

    def indices(self, pos, pos_half=False):
        """(rem0, rank, idx) for every level: int32/int32/uint32 arrays [L,M,4].
        pos_half: the reference's autocast step (grids/permuto_grid.py:65 `custom_fwd(cast_inputs=torch.half)`, :71
        `.type(torch.float)`): coordinates rounded to fp16 (round-to-nearest-even) and widened again BEFORE the lattice
        arithmetic -- identical already-rounded positions on both sides, indices stay bit-exact targets."""
        p = pos.detach().cpu().numpy().astype(np.float32)
        if pos_half:
            p = p.astype(np.float16).astype(np.float32)
        sf = self.scale_factor.numpy(); sh = self.random_shift_per_level.numpy()
        out = [lattice_level(p, sf[l], sh[l], self.capacity) for l in range(self.nr_levels)]
        return (np.stack([o[1] for o in out]), np.stack([o[2] for o in out]), np.stack([o[3] for o in out]))

    def forward(self, pos, pos_half=False):
        """pos [M,3] (float32 or float64 leaf) -> [M, L*F], level-major / feature-minor.
        pos_half: round the coordinates to fp16 first (see `indices`); the cast is an autograd op, d/d pos passes through."""
        if pos_half:
            pos = pos.half().to(pos.dtype)
        dt = pos.dtype
        ft = np.float64 if dt == torch.float64 else np.float32
        p32 = pos.detach().cpu().numpy().astype(ft)
        sf = self.scale_factor.numpy(); sh = self.random_shift_per_level.numpy()
        outs = []
        for l in range(self.nr_levels):
            elev, rem0, rank, idx = lattice_level(p32, sf[l], sh[l], self.capacity, ft)
            # differentiable linear map pos -> elevated, re-centred on the exact float32 value
            cf = (pos + self.random_shift_per_level[l].to(dt)) * self.scale_factor[l].to(dt)
            sm = torch.zeros_like(cf[:, 0])
            e = [None] * 4
            for i in range(POS_DIM, 0, -1):
                e[i] = sm - i * cf[:, i - 1]
                sm = sm + cf[:, i - 1]
            e[0] = sm
            e_lin = torch.stack(e, dim=1)
            elevated = torch.from_numpy(elev).to(dt) + (e_lin - e_lin.detach())
            delta = (elevated - torch.from_numpy(rem0).to(dt)) * 0.25
            rk = torch.from_numpy(rank.astype(np.int64))
            bary = torch.zeros(pos.shape[0], 5, dtype=dt)
            bary = bary.scatter_add(1, POS_DIM - rk, delta)
            bary = bary.scatter_add(1, POS_DIM + 1 - rk, -delta)
            b0 = bary[:, 0:1] + 1.0 + bary[:, 4:5]
            bary = torch.cat([b0, bary[:, 1:4]], dim=1)                       # [M,4]
            vals = self.lattice_values[l].to(dt)[torch.from_numpy(idx.astype(np.int64))]  # [M,4,F]
            outs.append((bary[:, :, None] * vals).sum(1) * self.anneal_window[l].to(dt))
        return torch.cat(outs, dim=1)
