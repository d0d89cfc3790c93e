"""Diagnostic (not a pytest): error of every self-test variant, for descriptor bring-up."""
import torch
from keypointnerf_b200 import _lib as L
lib = L.load()
for (N, K) in [(128, 192), (64, 128), (16, 16), (32, 64)]:
    for variant in range(4):
        g = torch.Generator().manual_seed(1)
        A = (torch.randn(128, K, generator=g) * 0.5).half().cuda()
        B = (torch.randn(N, K, generator=g) * 0.5).half().cuda()
        D = torch.full((128, N), float("nan"), device="cuda")
        rc = lib.kpn_selftest_umma(N, K, A.data_ptr(), B.data_ptr(), D.data_ptr(), variant, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = A.float() @ B.float().t()
        print(f"N={N} K={K} variant={variant} rc={rc} maxerr={(D-ref).abs().max().item():.4g} nan={torch.isnan(D).sum().item()}", flush=True)
