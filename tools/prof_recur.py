"""One RadLIF layer forward + backward at the cfg4 layer shape (for ncu captures of the
recurrence kernels).  Usage: python tools/prof_recur.py [T] [Be] [H]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparch_b200 import functional as F  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 20
Be = int(sys.argv[2]) if len(sys.argv) > 2 else 256
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
kind = sys.argv[4] if len(sys.argv) > 4 else "RadLIF"
dev = "cuda:0"
gen = torch.Generator(device=dev).manual_seed(1)
r = lambda *s: torch.rand(*s, device=dev, generator=gen)
I = (torch.randn(Be, T, H, device=dev, generator=gen) * 4 + 2).requires_grad_(True)
alpha = (r(H) * 0.14 + 0.82).requires_grad_(True)
beta = (r(H) * 0.02 + 0.968).requires_grad_(True)
a = r(H).requires_grad_(True)
b = (r(H) * 2).requires_grad_(True)
V = torch.nn.init.orthogonal_(torch.empty(H, H)).to(dev).requires_grad_(True)
u0, w0, s0 = r(Be, H), r(Be, H), r(Be, H)
g = torch.randn(Be, T, H, device=dev, generator=gen)
adaptive = kind in ("adLIF", "RadLIF")
recurrent = kind in ("RLIF", "RadLIF")
for it in range(3):
    F.timers_enable(True)
    S = F.SpikingCellFunction.apply(I, None, None, alpha, beta if adaptive else None,
                                    a if adaptive else None, b if adaptive else None,
                                    V if recurrent else None, u0, w0 if adaptive else None, s0, kind,
                                    1.0, F.NormState("none"))
    S.backward(g)
    tm = F.timers_collect()
    print(f"iter {it}: rate {float(S.mean()):.3f} fwd {tm['recurrence_fwd']:.3f} ms "
          f"bwd {tm['recurrence_bwd']:.3f} ms  ({tm['recurrence_fwd'] / T * 1e3:.1f} / "
          f"{tm['recurrence_bwd'] / T * 1e3:.1f} us per step)")
