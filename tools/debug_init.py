import sys, torch
sys.path.insert(0, '/root/repo')
import bench
dev = torch.device('cuda')
wl = bench.Workload(dev, n_rays=4096, seed=0, n_batches=1)
r = wl.forward_backward()
torch.cuda.synchronize()
print("done", float(r['loss'].detach()))
