"""Ray-sharded data parallelism for the hot path (new in this implementation -- the
reference is single-GPU).  One process per GPU; grid / MLP / pose parameters replicated; each rank
traces its own images' rays; the only exchange is ONE gradient all-reduce per step:
  bucket 0: color-grid table grad   (50.3 MB fp32 at best.yaml)
  bucket 1: delta-grid table grad   (50.3 MB)
  bucket 2: all decoder (+ pose) grads, flattened (~0.15 MB)
issued asynchronously (NCCL over NVLink/NVSwitch; gloo in the CPU tests) and waited on together.
"""
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
