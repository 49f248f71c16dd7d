"""Sweep the number of warp-aggregated levels of the hash-grid scatter on the bench workloads (configs 1 and 3)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from pagnerf_b200 import _lib, ops
dev = torch.device('cuda')
for config, rays in ((1, 4096), (3, 16384)):
    wl = bench.Workload(dev, n_rays=rays, seed=0, n_batches=1, config=config)
    grids = [wl.nef.grid] + ([wl.nef.delta_grid] if hasattr(wl.nef, 'delta_grid') else [])
    print("config", config, "levels", wl.nef.grid.embedder.n_levels, "default n_agg", wl.nef.grid.embedder.n_agg_levels, flush=True)
    for k in (0, 1, 2, 3, 4, 5, 6, 8):
        for g in grids:
            g.embedder.n_agg_levels = k
        for _ in range(3):
            wl.forward_backward()
        torch.cuda.synchronize()
        ops.BRANCH_OVERLAP = False
        _lib.timing_reset(True)
        for _ in range(5):
            wl.forward_backward()
        rep = _lib.timing_report()
        _lib.timing_reset(False)
        ops.BRANCH_OVERLAP = True
        t = {n: round(v['ms_per_launch'], 4) for n, v in rep.items() if 'hash' in n}
        print("  n_agg", k, t, flush=True)
