"""Tapes (S, U, W) of the two forward recurrence kernels on the same input.  Usage: cmp_fwd_tapes.py [T] [Be] [H] [kind]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparch_b200 import functional as F  # noqa: E402

F.W_TAPE_EVERY = 0   # both forward kernels write the full adaptation tape here (the checkpoint tape is tcgen05-only)

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
Be = int(sys.argv[2]) if len(sys.argv) > 2 else 128
H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
kind = sys.argv[4] if len(sys.argv) > 4 else "RLIF"
dev = "cuda:0"
adaptive = kind in ("adLIF", "RadLIF")
gen = torch.Generator(device=dev).manual_seed(1)
r = lambda *s: torch.rand(*s, device=dev, generator=gen)
I = torch.randn(Be, T, H, device=dev, generator=gen) * 3 + 1.2
alpha, beta, a, b = r(H) * 0.14 + 0.82, r(H) * 0.02 + 0.968, r(H), r(H) * 2
V = torch.randn(H, H, device=dev, generator=gen) / H ** 0.5
u0, w0, s0 = r(Be, H), r(Be, H), r(Be, H)
out = {}
for mode in ("mma", "tc"):
    F.RECUR_FWD = mode
    It = I.clone().requires_grad_(True)
    S = F.SpikingCellFunction.apply(It, None, None, alpha, beta if adaptive else None, a if adaptive else None,
                                    b if adaptive else None, V, u0, w0 if adaptive else None, s0, kind, 1.0,
                                    F.NormState("none"))
    sv = S.grad_fn.saved_tensors
    out[mode] = (S.detach().clone(), sv[12].clone(), sv[13].clone() if adaptive else None)
torch.cuda.synchronize()
for name, x, y in zip("SUW", out["mma"], out["tc"]):
    if x is None:
        continue
    d = (x - y).abs()
    bad = (d > 1e-4 * x.abs().max()).nonzero()
    print(name, "max|mma|", float(x.abs().max()), "max diff", float(d.max()), "elements off by > 1e-4 max:", len(bad))
    if len(bad):
        print("   first:", bad[:5].tolist(), " t values:", sorted(set(bad[:, 1].tolist()))[:20], " rows:", sorted(set(bad[:, 0].tolist()))[:20],
              " cols:", sorted(set(bad[:, 2].tolist()))[:20])
