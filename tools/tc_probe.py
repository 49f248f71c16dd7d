import torch, sys
sys.path.insert(0, '/root/repo')
from pagnerf_b200 import _lib
from pagnerf_b200._lib import call, ptr
torch.manual_seed(0)
dev='cuda'
def run(mode, N, K, FA=0, reps=1, fn="pag_tc_gemm_test16"):
    if mode == 0:
        A = torch.randn(128, K, device=dev); B = torch.randn(N, K, device=dev); ref = A @ B.t()
    elif mode == 1:
        A = torch.randn(128, K, device=dev); B = torch.randn(K, N, device=dev); ref = A @ B
    else:
        A = torch.randn(128, FA, device=dev); B = torch.randn(128, N, device=dev); ref = A.t() @ B
    D = torch.zeros(128, N, device=dev)
    call(fn, mode, ptr(A), ptr(B), ptr(D), N, K, FA, reps)
    torch.cuda.synchronize()
    ref = ref * reps
    rows = ref.shape[0]
    err = (D[:rows] - ref).abs().max().item(); scale = ref.abs().max().item()
    print(f"mode {mode} N={N} K={K} FA={FA} reps={reps}: max err {err:.4e} (ref max {scale:.3f}) rel {err/scale:.2e}", "OK" if err/scale < 3e-3 else "FAIL")
    return err/scale < 3e-3
ok = True
torch.backends.cuda.matmul.allow_tf32 = False
for N,K in [(64,48),(16,64),(64,64),(208,64),(64,16),(48,64)]:
    ok &= run(0,N,K)
for N,K in [(64,64),(48,64),(64,16),(64,208),(48,16)]:
    ok &= run(1,N,K)
for N,FA in [(64,64),(48,64),(64,16),(64,128),(48,8),(64,208)]:
    ok &= run(2,N,0,FA)
ok &= run(0,64,64,reps=3); ok &= run(2,64,0,64,reps=2)
print("ALL OK" if ok else "SOME FAILED")
