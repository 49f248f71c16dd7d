import torch, sys
sys.path.insert(0, '/root/repo')
from pagnerf_b200._lib import call, ptr
dev='cuda'
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
# mode 1: D[m][n] = sum_k A[m][k] B[k][n]; A one-hot on k = m % K -> D[m][n] should be B[m%K][n]
for (N,K) in [(16,8),(16,16),(32,8)]:
    A = torch.zeros(128, K, device=dev); A[torch.arange(128), torch.arange(128) % K] = 1
    B = (torch.arange(K, device=dev)[:,None]*100 + torch.arange(N, device=dev)[None,:]).float() + 1
    D = torch.full((128, N), -7.0, device=dev)
    call("pag_tc_gemm_test", 1, ptr(A), ptr(B), ptr(D), N, K, 0, 1)
    torch.cuda.synchronize()
    print("mode1 N",N,"K",K); print(D[:K+2])
# mode 2: D[j][n] = sum_s A[s][j] B[s][n]; A one-hot: A[s][j]=1 iff j == s % FA ; B[s][n] = s*100+n+1 -> D[j][n] = sum over s≡j of B
for (N,FA) in [(16,8),(16,4)]:
    A = torch.zeros(128, FA, device=dev); A[torch.arange(128), torch.arange(128) % FA] = 1
    B = torch.zeros(128, N, device=dev); B[:FA] = (torch.arange(FA, device=dev)[:,None]*100 + torch.arange(N, device=dev)[None,:]).float() + 1
    D = torch.full((128, N), -7.0, device=dev)
    call("pag_tc_gemm_test", 2, ptr(A), ptr(B), ptr(D), N, 0, FA, 1)
    torch.cuda.synchronize()
    print("mode2 N",N,"FA",FA); print(D[:FA+2]); print("ref"); print((A.t()@B)[:FA])
