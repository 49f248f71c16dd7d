import sys, torch, time
sys.path.insert(0, '/root/repo')
import bench
from pagnerf_b200 import ops, _lib
from pagnerf_b200._lib import call, ptr
from pagnerf_b200.wisp_compat import Rays
dev = torch.device('cuda:0')
wl = bench.Workload(dev, 16384, seed=0)
o, d = wl.dev[0][0], wl.dev[0][1]
ridx, pidx, samples, depths, deltas, boundary = wl.nef.grid.raymarch(Rays(origins=o, dirs=d, dist_min=0.0, dist_max=2.0), level=0, num_samples=128, raymarch_type='ray')
pos = samples.reshape(-1, 3).contiguous()
M = pos.shape[0]; print("M", M)
e = wl.nef.grid.embedder
tb = e.lattice_values.detach(); L, cap, F = tb.shape
g = torch.randn(M, 2 * L, device=dev) * 1e-3
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
gt = torch.zeros_like(tb)
for n_agg in (0, 1, 2, 3, 4, 6, 8, 10, 14):
    t = timeit(lambda: call("pag_permuto_bwd", ptr(pos), M, ptr(tb), cap, L, F, ptr(e.scale_factor), ptr(e.random_shift_per_level), ptr(e.anneal_window), ptr(g), ptr(gt), None, n_agg))
    print(f"bwd n_agg={n_agg:2d}: {t*1e3:.1f} us")
gp = torch.empty(M, 3, device=dev)
t = timeit(lambda: call("pag_permuto_bwd", ptr(pos), M, ptr(tb), cap, L, F, ptr(e.scale_factor), ptr(e.random_shift_per_level), ptr(e.anneal_window), ptr(g), ptr(gt), ptr(gp), 10))
print(f"bwd +pos n_agg=10: {t*1e3:.1f} us")
out = torch.empty(M, 2 * L, device=dev)
t = timeit(lambda: call("pag_permuto_fwd", ptr(pos), M, ptr(tb), cap, L, F, ptr(e.scale_factor), ptr(e.random_shift_per_level), ptr(e.anneal_window), ptr(out)))
print(f"fwd: {t*1e3:.1f} us  -> {M*972/t/1e6:.0f} GB/s algorithmic")
