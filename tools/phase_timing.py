"""Phase timeline of the instrumented decoder kernels (developer tool, GPU only).

    python -m pagnerf_b200.build --phase-timing
    PAGNERF_B200_LIB=pagnerf_b200/csrc/libpagnerf_b200_dbg.so python tools/phase_timing.py

Runs a few eager steps of the bench workload and prints the clock64 ticks block 0 / thread 0 spent between
consecutive PAG_PHASE marks of the instrumented kernel, per tile."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pagnerf_b200 import _lib


def main():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    wl = bench.Workload(torch.device("cuda:0"))
    for i in range(3):
        wl.forward_backward(wl.batch(i))
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 64)()
    lib.pag_debug_phase_read(buf, 1)
    steps = 5
    for i in range(steps):
        wl.forward_backward(wl.batch(i))
    torch.cuda.synchronize()
    lib.pag_debug_phase_read(buf, 0)
    v = list(buf)
    tot = sum(v)
    print("total ticks/step (block 0):", tot / steps)
    for i, x in enumerate(v):
        if x:
            print(f"phase {i:2d}: {x / steps:10.0f} ticks/step  {100 * x / tot:5.1f}%")


if __name__ == '__main__':
    main()
