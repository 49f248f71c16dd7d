"""CUDA-graph capture of a whole training step (forward + loss + backward) on the fused trace path.

The fused path (ops.FusedTraceFn) has fixed launch geometry: no host synchronisation, worst-case work buffers, the
packed-sample count and the jitter seed live in device memory.  That makes the step capturable once and replayable
with ~10 us of host work, independent of the host CPU (eager PyTorch needs ~2 ms of Python / dispatcher time per
step for the same ~55 launches).
"""
import torch


class GraphedStep:
    """step_fn(*static_inputs) -> loss, run as `loss.backward()`-inclusive CUDA graph.

        g = GraphedStep(step_fn, example_inputs, params, nef)
        loss = g(*new_inputs)        # copies inputs into the static buffers, replays, returns the static loss tensor
    `.grad` of `params` are static tensors refreshed by every replay.  step_fn must not keep references to tensors
    that carry a grad_fn between calls (they would pin AccumulateGrad nodes created on another stream).
    """

    def __init__(self, step_fn, example_inputs, params, nef, warmup=3):
        self.params = list(params)
        self.static_inputs = [x.clone() for x in example_inputs]
        dev = self.static_inputs[0].device
        blas = nef.grid.blas.to(dev)
        if getattr(blas, 'seed_tensor', None) is None:
            blas.seed_tensor = torch.full((1,), int(blas.jitter_seed), dtype=torch.int32, device=dev)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                for p in self.params:
                    p.grad = None
                loss = step_fn(*self.static_inputs)
                loss.backward()
                del loss     # drop the autograd graph: the parameters' AccumulateGrad nodes must be re-created on the capture stream
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = step_fn(*self.static_inputs)
            self.loss.backward()

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
