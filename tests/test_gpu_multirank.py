"""Ray-sharded data parallelism of the fused training trace (SURVEY 8e, section 4 tier 5): N ranks x their shard of the rays,
gradients all-reduced INSIDE the backward (ops.set_grad_sync) == one rank on the concatenated batch.

The driver's GPU box has one GPU for the tests, and NCCL refuses two ranks on one device, so the two ranks share cuda:0 and talk
over gloo (which moves CUDA tensors through host memory): the data path -- which gradients are reduced, in which order, on which
streams, mean vs sum, the fp16 transport with its shared scale -- is exactly the one NCCL runs in bench.py; only the wire differs.
The jitter stream is made identical to the concatenated batch by giving rank 1 the seed whose counter offset equals its first
ray's flat index (the marcher draws u = lowbias32(ray * S + step + seed * 0x9E3779B9))."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
root, port, rank, transport, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
sys.path.insert(0, root)
from tests.util import load_golden, build_cuda_nef
from pagnerf_b200 import ops
from pagnerf_b200.tracers import PanopticPackedRFTracer
from pagnerf_b200.wisp_compat import Rays
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=2)
dev = "cuda:0"
g = load_golden("trace_delta_permuto_ray")
S, seed = int(g["num_steps"]), int(g["jitter_seed"])
N = g["o"].shape[0]
half = N // 2
lo, hi = (0, half) if rank == 0 else (half, 2 * half)
nef = build_cuda_nef(g, dev)
nef.decoder_precision = 'fp16'
inv = pow(0x9E3779B9, -1, 1 << 32)
for gr in (nef.grid, nef.delta_grid):      # rank 1 continues the jitter counter where rank 0's rays end
    gr.blas.fixed_jitter, gr.blas.jitter_seed = True, (seed + lo * S * inv) % (1 << 32)
ops.set_grad_sync(True, transport=transport)
tracer = PanopticPackedRFTracer(raymarch_type='ray', num_steps=S, bg_color='white')
chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
o = torch.from_numpy(g["o"][lo:hi]).to(dev)
d = torch.from_numpy(g["d"][lo:hi]).to(dev)
rb = tracer(nef, channels=chans, rays=Rays(origins=o, dirs=d, dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
assert torch.is_tensor(tracer.last_num_samples)
loss = sum((getattr(rb, c) * torch.from_numpy(g["gw_" + c][lo:hi]).to(dev)).sum() for c in chans)
loss.backward()
torch.cuda.synchronize()
grads = {k: p.grad.detach().cpu().numpy() for k, p in nef.named_parameters() if p.grad is not None}
np.savez(out, **grads)
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
'''


@pytest.mark.parametrize("transport", ["fp32", "fp16"])
def test_sharded_fused_step_with_grad_sync_equals_single_rank(cuda_lib, tmp_path, transport):
    from tests.util import load_golden, build_cuda_nef, assert_close
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    outs = [str(tmp_path / f"grads{r}.npz") for r in range(2)]
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), transport, outs[r]], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    logs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, lg) in enumerate(zip(procs, logs)):
        assert p.returncode == 0 and f"OK {r}" in lg, lg[-3000:]
    # one rank, concatenated batch
    dev = "cuda"
    g = load_golden("trace_delta_permuto_ray")
    N = (g["o"].shape[0] // 2) * 2
    nef = build_cuda_nef(g, dev)
    nef.decoder_precision = 'fp16'
    tracer = PanopticPackedRFTracer(raymarch_type='ray', num_steps=int(g["num_steps"]), bg_color='white')
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    rb = tracer(nef, channels=chans, rays=Rays(origins=torch.from_numpy(g["o"][:N]).to(dev), dirs=torch.from_numpy(g["d"][:N]).to(dev),
                                               dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
    sum((getattr(rb, c) * torch.from_numpy(g["gw_" + c][:N]).to(dev)).sum() for c in chans).backward()
    full = {k: p.grad.detach().cpu() for k, p in nef.named_parameters() if p.grad is not None}
    r0, r1 = np.load(outs[0]), np.load(outs[1])
    assert set(r0.files) == set(full), "every parameter that has a gradient on one rank has the reduced one on the shards"
    for k in full:
        a0, a1 = torch.from_numpy(r0[k]), torch.from_numpy(r1[k])
        assert torch.equal(a0, a1), f"{k}: ranks disagree after the all-reduce"
        # mean over 2 ranks of per-shard sums == half the concatenated batch's gradient
        big = 'lattice_values' in k
        tol = 2e-3 if (transport == 'fp16' and big) else 2e-4      # tables on the wire as halfs: 2^-11 relative per element
        assert_close(2.0 * a0, full[k], rtol=tol, atol_scale=tol, msg=f"{transport} grad {k}")
