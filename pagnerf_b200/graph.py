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

    Device pointers baked into the graph: parameters / static inputs (stable), the occupancy bit field and the LOD weights
    (both refreshed IN PLACE by OctreeAS.init / PanopticNeF._lodw, so prune() and LOD annealing need no action), and -- for
    octree levels < 2, where the fused trace marches against the octree itself -- blas.octree / blas.prefix, which prune()
    replaces.  `__call__` compares those addresses with the live ones and re-captures when any changed.
    """

    def __init__(self, step_fn, example_inputs, params, nef, warmup=3, post_backward=None):
        """post_backward(): work captured right after loss.backward() -- gradient all-reduces of parameters outside the fused trace,
        the optimizer step (pagnerf_b200.optim.FusedAdam is capturable: step count and moments live on the device)."""
        self.params = list(params)
        self.step_fn, self.nef, self.warmup, self.post_backward = step_fn, nef, warmup, post_backward
        self.static_inputs = [x.clone() for x in example_inputs]
        self._capture()

    def _signature(self):
        nef = self.nef
        blas = nef.grid.blas
        lvl = nef.grid.blas_level
        bits = blas._bits.get(int(lvl)) if lvl >= 2 else None
        lodw = getattr(nef, '_lodw_dev', None)
        return (bits.data_ptr() if bits is not None else (blas.octree.data_ptr(), blas.prefix.data_ptr()),
                lodw.data_ptr() if lodw is not None else None)

    def _capture(self):
        nef, step_fn, warmup = self.nef, self.step_fn, self.warmup
        dev = self.static_inputs[0].device
        blas = nef.grid.blas.to(dev)
        if getattr(blas, 'seed_tensor', None) is None:
            blas.seed_tensor = torch.full((1,), int(blas.jitter_seed), dtype=torch.int32, device=dev)
        blas.graph_seed_active = True      # fused traces read (and advance) the device-resident jitter seed from here on ...
        try:
            self._capture_inner(step_fn, warmup)
        finally:
            blas.graph_seed_active = False  # ... until capture is over: eager traces keep following blas.jitter_seed
        blas.jitter_seed = int(blas.seed_tensor.item())
        self.blas = blas
        self.sig = self._signature()

    def _capture_inner(self, step_fn, warmup):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                for p in self.params:
                    p.grad = None
                loss = step_fn(*self.static_inputs)
                loss.backward()
                if self.post_backward is not None:
                    self.post_backward()
                del loss     # drop the autograd graph: the parameters' AccumulateGrad nodes must be re-created on the capture stream
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = step_fn(*self.static_inputs)
            self.loss.backward()
            if self.post_backward is not None:
                self.post_backward()

    def __call__(self, *inputs):
        if self._signature() != self.sig:      # prune() on a coarse octree / a re-allocated weight vector: stale addresses
            self.graph = None
            self._capture()
        if not self.blas.fixed_jitter:         # host mirror of the device-side seed the replay advances
            self.blas.jitter_seed = (self.blas.jitter_seed + 1) & 0x7FFFFFFF
        for dst, src in zip(self.static_inputs, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
