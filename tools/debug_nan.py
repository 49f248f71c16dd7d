"""Dirty-allocator check: every torch.empty() of the step comes out of NaN-filled memory, so a read of an unwritten buffer or a
tensor freed while another stream still uses it shows up as non-finite gradients.  python tools/debug_nan.py  (GPU)"""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
dev = torch.device('cuda')
x = torch.full((6_000_000_000 // 4,), float('nan'), device=dev)
del x
ok = True
for rep in range(3):
    wl = bench.Workload(dev, n_rays=16384, seed=0, n_batches=1)
    r = wl.forward_backward()
    torch.cuda.synchronize()
    bad = {n: int((~torch.isfinite(p.grad)).sum()) for n, p in wl.nef.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()}
    print("rep", rep, "loss", float(r['loss'].detach()), "nonfinite", bad, flush=True)
    ok &= not bad
sys.exit(0 if ok else 1)
