"""Warm timings (CUDA events, 20 repetitions) of the small-K / small-N GEMM shapes of cfg 4."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparch_b200 import gemm  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
M = 25600


def timeit(name, fn, bytes_out):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{name}: {us:.1f} us  ({bytes_out / us / 1e6:.2f} TB/s of output)")


X = torch.randn(M, 40, device=dev, generator=g)
W0 = torch.randn(1024, 40, device=dev, generator=g)
S = (torch.rand(M, 1024, device=dev, generator=g) < 0.1).float()
Wr = torch.randn(35, 1024, device=dev, generator=g)
dZr = torch.randn(M, 35, device=dev, generator=g)
dZ0 = torch.randn(M, 1024, device=dev, generator=g)
xa, w0 = gemm.split_f16(X, 2), gemm.split_f16(W0, 2)
sp, wr = gemm.split_f16(S, 1, scaled=False), gemm.split_f16(Wr, 2)
gr, g0 = gemm.split_f16(dZr, 2), gemm.split_f16(dZ0, 2)
out = torch.empty(M, 1024, device=dev)
timeit("L0 projection  (25600x1024, K=40, 3 passes)", lambda: gemm.gemm_parts(xa, w0, 40, out=out), M * 1024 * 4)
timeit("readout proj   (25600x35, K=1024, 2 passes)", lambda: gemm.gemm_parts(sp, wr, 1024), M * 35 * 4)
timeit("readout dX     (25600x1024, K=35, 3 passes)", lambda: gemm.gemm_parts(gr, wr, 35, b_mn=True, N=1024, out=out), M * 1024 * 4)
timeit("readout dW     (35x1024, K=25600, 2 passes)", lambda: gemm.gemm_parts(gr, sp, M, a_mn=True, b_mn=True, M=35, N=1024), 35 * 1024 * 4)
timeit("L0 dW          (1024x40, K=25600, 3 passes)", lambda: gemm.gemm_parts(g0, xa, M, a_mn=True, b_mn=True, M=1024, N=40), 1024 * 40 * 4)
big = torch.empty(M, 1024, device=dev)
timeit("copy 105 MB (reference for an output-bound kernel)", lambda: big.copy_(out), M * 1024 * 4)
