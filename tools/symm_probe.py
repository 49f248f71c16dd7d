"""Probe: does torch symmetric memory (peer pointers / NVLS multicast) work on this box?  torchrun --nproc-per-node N tools/symm_probe.py"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; world", hdl.world_size, "multicast", hdl.multicast_ptr != 0,
          "mc_ptr", hex(hdl.multicast_ptr), "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal pad size", hdl.signal_pad_size, flush=True)
    t.fill_(rank + 1)
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, t.shape, t.dtype)
    print(rank, "peer value", float(peer[0]), flush=True)
    hdl.barrier(channel=0)
    # graph capture of barrier
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        hdl.barrier(channel=0)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        hdl.barrier(channel=0)
        t.add_(1.0)
        hdl.barrier(channel=0)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    print(rank, "graph replay ok", float(t[0]), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "FAILED", type(e).__name__, e, flush=True)
dist.barrier()
dist.destroy_process_group()
