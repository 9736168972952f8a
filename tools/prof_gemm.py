"""The cfg4 layer-1 GEMMs in isolation (for timing and ncu captures of gemm_tn_bf16_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparch_b200 import gemm  # noqa: E402

dev = "cuda:0"
M, N, K = 25600, 1024, 1024
g = torch.Generator(device=dev).manual_seed(0)
X = (torch.rand(M, K, device=dev, generator=g) < 0.1).float()
W = torch.randn(N, K, device=dev, generator=g) / 32
dZ = torch.randn(M, N, device=dev, generator=g)
xa, wb = gemm.split_rows(X, 1), gemm.split_rows(W, 3)
ga, wt = gemm.split_rows(dZ, 3), gemm.split_transposed(W, 3)
gt, xt = gemm.split_transposed(dZ, 3), gemm.split_transposed(X, 1)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, fn, passes, flop in [
    ("proj  Z=X W^T   (spikes x fp32, 3 passes)", lambda: gemm.gemm_parts(xa, wb, K), 3, 2.0 * M * N * K),
    ("dgrad dX=dZ W   (fp32 x fp32, 6 passes)", lambda: gemm.gemm_parts(ga, wt, N), 6, 2.0 * M * N * K),
    ("wgrad dW=dZ^T X (fp32 x spikes, 3 passes, split-K)", lambda: gemm.gemm_parts(gt, xt, M), 3, 2.0 * M * N * K),
]:
    ms = timeit(fn)
    print(f"{name}: {ms:.3f} ms  {passes * flop / ms / 1e9:.0f} TFLOP/s bf16-equivalent "
          f"({flop / ms / 1e9:.0f} TFLOP/s of fp32-accurate product)")
for name, fn, nbytes in [
    ("split_rows 3 terms (25600x1024)", lambda: gemm.split_rows(dZ, 3), M * N * 10),
    ("split_transposed 3 terms", lambda: gemm.split_transposed(dZ, 3), M * N * 10),
    ("split_transposed 1 term", lambda: gemm.split_transposed(X, 1), M * K * 6),
]:
    ms = timeit(fn)
    print(f"{name}: {ms * 1e3:.0f} us  {nbytes / ms / 1e6:.0f} GB/s")
