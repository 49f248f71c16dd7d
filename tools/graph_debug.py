import sys, torch
sys.path.insert(0, '/root/repo')
import bench
CAP=False
from pagnerf_b200 import _lib
dev = torch.device('cuda:0'); torch.cuda.set_device(0)
wl = bench.Workload(dev, 4096, seed=0); wl.keep_rb = False
for _ in range(3): wl.forward_backward()
torch.cuda.synchronize()
lib = _lib.load()
orig_call = _lib.call
def status():
    return lib.pag_capture_status(_lib.stream())
def dbg_call(name, *args):
    orig_call(name, *args)
    st = status()
    import threading
    if False: print("after", name, "capture status", st, threading.current_thread().name, flush=True)
import pagnerf_b200.ops as ops
ops.call = dbg_call
blas = wl.nef.grid.blas
blas.seed_tensor = torch.zeros(1, dtype=torch.int32, device=dev)
for p in wl.params: p.grad = None
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        for p in wl.params: p.grad = None
        l_ = wl.loss_of(*wl.dev[0]); l_.backward(); del l_
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
for p in wl.params: p.grad = None
g = torch.cuda.CUDAGraph()
CAP=True
try:
    with torch.cuda.graph(g):
        print("start status", status())
        loss = wl.loss_of(*wl.dev[0])
        print("after fwd status", status())
        loss.backward()
        print("after bwd status", status())
    print("captured OK")
    g.replay(); torch.cuda.synchronize(); print("replay ok", float(loss))
except Exception as e:
    import traceback; traceback.print_exc()
    print("FAILED", type(e).__name__)
