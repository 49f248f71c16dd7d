"""Multi-GPU check (torchrun, NCCL): the fused step's gradients with the 'symm' transport (our multimem / peer-memory all-reduce)
and with the 'fp16' transport against the exact fp32 NCCL all-reduce, on the bench workload; every rank must hold identical results."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pagnerf_b200 import ops
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {}
for tr in ("fp32", "symm", "fp16"):
    wl = bench.Workload(dev, n_rays=4096, seed=rank, n_batches=1)
    for g in (wl.nef.grid, wl.nef.delta_grid):
        g.blas.fixed_jitter, g.blas.jitter_seed = True, 7
    ops.set_grad_sync(True, transport=tr)
    for _ in range(2):      # twice: the persistent symmetric buffers are re-zeroed and re-used
        wl.step_idx = 0
        wl.forward_backward()
    torch.cuda.synchronize()
    res[tr] = {n: p.grad.detach().clone() for n, p in wl.nef.named_parameters()}
ok = True
for tr in ("symm", "fp16"):
    worst = 0.0
    for n, g in res[tr].items():
        r = res["fp32"][n]
        e = float((g - r).norm() / max(float(r.norm()), 1e-30))
        worst = max(worst, e)
        # identical on every rank?
        gg = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(gg, g.contiguous())
        same = all(torch.equal(gg[0], x) for x in gg)
        if not same or e > (2e-3 if tr == 'fp16' else 1e-5):
            ok = False
            print(f"rank {rank} {tr} {n}: rel l2 {e:.3e} same_on_all_ranks {same}", flush=True)
    if rank == 0:
        print(f"transport {tr}: worst rel l2 vs fp32 NCCL {worst:.3e}", flush=True)
if rank == 0:
    print("MG_CHECK", "OK" if ok else "FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
