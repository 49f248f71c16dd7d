"""How much does the step time vary between ray batches / ranks' seeds?  (synchronous data parallelism pays max over ranks)"""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
dev = torch.device('cuda')
for seed in (0, 1, 2, 3):
    wl = bench.Workload(dev, n_rays=16384, seed=seed, n_batches=4)
    out = []
    for b in range(4):
        for _ in range(3):
            wl.step_idx = b; wl.forward_backward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            wl.step_idx = b; wl.forward_backward()
        e1.record(); torch.cuda.synchronize()
        out.append((int(wl.tracer.last_num_samples), round(e0.elapsed_time(e1) / 10, 3)))
    print("seed", seed, out, flush=True)
