"""tcgen05.mma timing probe over the operand images the decoders use (developer tool, GPU only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pagnerf_b200._lib import call, ptr

cyc = torch.zeros(2, dtype=torch.int64, device='cuda')
names = {0: 'fwd  (A K-major, B K-major)', 1: 'dX   (A K-major, B MN-major)', 2: 'dW   (A MN-major, B MN-major, K=128)'}
for mode in (0, 1, 2):
    for N in (16, 64, 208):
        for K in ((64,) if mode < 2 else (128,)):
            for chains in (1, 4, 16):
                reps = 50
                call("pag_tc_mma_bench", mode, N, K, chains, reps, ptr(cyc))
                torch.cuda.synchronize()
                tot, iss = cyc.tolist()
                n = chains * (K // 16 if mode < 2 else 8)
                print(f"{names[mode]:40s} N={N:3d} K={K:3d} chains={chains:2d}: {tot/reps:8.0f} cyc/phase, {tot/reps/n:6.1f} cyc/MMA, issue {iss/reps/n:5.1f} cyc/MMA")
