"""Ray-sharded data parallelism for the hot path (new in this implementation -- the
reference is single-GPU).  One process per GPU; grid / MLP / pose parameters replicated; each rank
traces its own images' rays; the only exchange is ONE gradient all-reduce per step:
  bucket 0: color-grid table grad   (50.3 MB fp32 at best.yaml)
  bucket 1: delta-grid table grad   (50.3 MB)
  bucket 2: all decoder (+ pose) grads, flattened (~0.15 MB)
issued asynchronously (NCCL over NVLink/NVSwitch; gloo in the CPU tests) and waited on together.
"""
import os

import torch
import torch.distributed as dist

BIG = 1 << 20  # tensors above this many elements get their own bucket


def shard_images(num_images, rank, world_size):
    """Whole images per rank (rank r gets images r::W) so per-image losses stay local
    (reference loss/lin_assignment_things.py:58-80 works per image)."""
    return list(range(rank, num_images, world_size))


def allreduce_grads(params, group=None, average=True):
    """Sum (or average) .grad of `params` across ranks.  Returns the number of collectives issued."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    ws = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    big = [g for g in grads if g.numel() >= BIG]
    small = [g for g in grads if g.numel() < BIG]
    works = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=True) for g in big]
    flat = None
    if small:
        flat = torch.cat([g.reshape(-1) for g in small])
        works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()
    if flat is not None:
        off = 0
        for g in small:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
    if average:
        for g in grads:
            g.div_(ws)
    return len(works)


class SymmetricGradBuffers:
    """The step's big gradients in SYMMETRIC memory (torch.distributed._symmetric_memory: the same allocation mapped into every
    peer of the node and into one NVSwitch multicast address) + their in-place mean all-reduce with csrc/allreduce.cu
    (multimem.ld_reduce / multimem.st -- the sum is formed inside the switch -- or explicit peer loads / stores).

        sg = SymmetricGradBuffers({'table': n0, 'dtable': n1, 'flat': n2}, device, group)     # collective: every rank calls it
        g = sg.view('table')                    # float32 [n0], zero it, scatter into it
        sg.allreduce('table', channel=1)        # barrier -> one kernel -> barrier, on the current stream

    Each concurrently running all-reduce needs its own barrier channel.  The views are persistent: a gradient returned from them
    is valid until the next step overwrites it (ops.set_grad_sync(transport='symm') documents this contract)."""

    def __init__(self, numels, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib = _lib
        self.group = group if group is not None else dist.group.WORLD
        self.offsets, off = {}, 0
        for name, n in numels.items():
            self.offsets[name] = (off, int(n))
            off += (int(n) + 3) // 4 * 4
        self.flag_off = off                      # int32 flags [8 channels][16 ranks] of the cross-rank barrier, inside the same buffer
        self.buf = symm_mem.empty(off + 8 * 16, dtype=torch.float32, device=device)
        self.buf[off:].zero_()
        self.epoch = torch.zeros(8, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.hdl.barrier(channel=0)              # every rank's flags are zero before the first epoch is published
        torch.cuda.synchronize(device)
        # Which path: measured on 2 / 8 B200 -- 50 MB table, 2 ranks: peer loads/stores 0.09 ms, multicast 0.15 ms;
        # 8 ranks: multicast 0.13 ms, peer 0.17 ms.  PAGNERF_SYMM_P2P=1 / 0 forces one of them.
        force = os.environ.get('PAGNERF_SYMM_P2P')
        use_p2p = (force == '1') if force in ('0', '1') else (self.hdl.world_size <= 2)
        self.mc = 0 if use_p2p else int(self.hdl.multicast_ptr)
        self.rank, self.world = self.hdl.rank, self.hdl.world_size
        self.peers = [int(p) for p in self.hdl.buffer_ptrs]
        if self.world > 16:
            raise NotImplementedError("csrc/allreduce.cu: at most 16 ranks per node")

    def view(self, name):
        off, n = self.offsets[name]
        return self.buf[off:off + n]

    def allreduce(self, name, channel=0, mean=True, max_ctas=0, sub=None, upto=None):
        """In-place (mean) all-reduce of one named segment on the current stream.  sub=(start, count): only that element range of
        the segment (multiples of 4); upto=name2: the contiguous span from this segment through the end of segment name2."""
        import ctypes
        off, n = self.offsets[name]
        if upto is not None:
            o2, n2 = self.offsets[upto]
            n = o2 + n2 - off
        if sub is not None:
            off, n = off + int(sub[0]), int(sub[1])
        n4 = (n + 3) // 4 * 4
        peers = (ctypes.c_void_p * self.world)(*self.peers)
        ep = self._lib.ptr(self.epoch)
        # every rank's producers of this segment have finished (stream order + release/acquire flags over NVLink) ...
        self._lib.call("pag_symm_barrier", peers, self.flag_off, ep, self.rank, self.world, int(channel))
        self._lib.call("pag_allreduce_symm", self.mc if self.mc else None, peers, self.rank, self.world, off, n4,
                       (1.0 / self.world) if mean else 1.0, int(max_ctas))
        # ... and every rank's stores have landed before anybody consumes (or re-zeroes) it
        self._lib.call("pag_symm_barrier", peers, self.flag_off, ep, self.rank, self.world, int(channel))
