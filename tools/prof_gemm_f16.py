"""The cfg4 layer-1 GEMMs with the default operand scheme (scaled fp16 hi/lo terms, persistent kernel) in isolation.
   python tools/prof_gemm_f16.py            -> warm CUDA-event timings
   ncu --profile-from-start off ... python tools/prof_gemm_f16.py --once   -> one profiled launch per GEMM"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparch_b200 import gemm  # noqa: E402

dev = "cuda:0"
M, N, K = 25600, 1024, 1024
g = torch.Generator(device=dev).manual_seed(0)
S = (torch.rand(M, K, device=dev, generator=g) < 0.1).float()
W = torch.randn(N, K, device=dev, generator=g) / 32
dZ = torch.randn(M, N, device=dev, generator=g) * 1e-3
sp, wt, gz = gemm.split_f16(S, 1, scaled=False), gemm.split_f16(W, 2), gemm.split_f16(dZ, 2)
out = torch.empty(M, N, device=dev)
outw = torch.empty(N, K, device=dev)
cases = [
    ("proj  Z = S W^T    (spikes x fp32, 2 passes)", lambda: gemm.gemm_parts(sp, wt, K, out=out), 2),
    ("dgrad dX = dZ W    (fp32 x fp32, 3 passes, MN-major B)", lambda: gemm.gemm_parts(gz, wt, N, b_mn=True, N=K, out=out), 3),
    ("wgrad dW = dZ^T S  (fp32 x spikes, 2 passes, MN-major, split-K)",
     lambda: gemm.gemm_parts(gz, sp, M, a_mn=True, b_mn=True, M=N, N=K, out=outw), 2),
    ("dV = S_prev^T dI   (spikes x fp32, 2 passes, frame delay, split-K)",
     lambda: gemm.gemm_parts(sp, gz, M, a_mn=True, b_mn=True, a_koff=-1, M=K, N=N, out=outw), 2),
]
flop = 2.0 * M * N * K
if "--once" in sys.argv:
    for _, fn, _ in cases:
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _, fn, _ in cases:
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
for name, fn, passes in cases:
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms  {passes * flop / ms / 1e9:.0f} TFLOP/s fp16 tensor work "
          f"({flop / ms / 1e9:.0f} TFLOP/s of fp32-accurate product)")
