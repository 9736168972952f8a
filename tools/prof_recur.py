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

if os.environ.get("SPARCH_PHASES"):
    from sparch_b200._lib import call, ptr
    dbg = torch.zeros(T, 8, dtype=torch.int64, device=dev)
    call("sparch_recur_debug_clocks", ptr(dbg))
    S = F.SpikingCellFunction.apply(I, None, None, alpha, beta if adaptive else None, a if adaptive else None,
                                    b if adaptive else None, V if recurrent else None, u0,
                                    w0 if adaptive else None, s0, kind, 1.0, F.NormState("none"))
    torch.cuda.synchronize()
    call("sparch_recur_debug_clocks", None)
    c = dbg.cpu().double()
    d, prev_end = c[2:], c[1:-1, 3]
    m = lambda x: float(x.mean())
    print("forward phases, cycles/step (CTA 0,0): spike-word wait+load %.0f | mma %.0f | reduce %.0f | "
          "update+publish %.0f | total %.0f"
          % (m(d[:, 0] - prev_end), m(d[:, 1] - d[:, 0]), m(d[:, 2] - d[:, 1]), m(d[:, 3] - d[:, 2]),
             m(d[:, 3] - prev_end)))
    ov = torch.minimum(c[2:, 1], c[2:, 6]) - torch.maximum(c[2:, 0], c[2:, 5])
    print("team overlap (same SM clock): team1 mma start - team0 mma start %.0f cycles, team1 mma %.0f, overlap of the two MMA "
          "phases %.0f cycles/step" % (m(c[2:, 5] - c[2:, 0]), m(c[2:, 6] - c[2:, 5]), float(ov.clamp(min=0).mean())))
    dbg.zero_()
    call("sparch_recur_debug_clocks", ptr(dbg))
    I2 = I.detach().clone().requires_grad_(True)
    S = F.SpikingCellFunction.apply(I2, None, None, alpha, beta if adaptive else None, a if adaptive else None,
                                    b if adaptive else None, V if recurrent else None, u0,
                                    w0 if adaptive else None, s0, kind, 1.0, F.NormState("none"))
    torch.cuda.synchronize()
    dbg.zero_()
    S.backward(g)
    torch.cuda.synchronize()
    call("sparch_recur_debug_clocks", None)
    c = dbg.cpu().double()[2:-2]
    print("backward phases, cycles/step (CTA 0,0): prefetch issue->wait done %.0f | stream+mma %.0f | "
          "reduce+update+panel %.0f | fence+arrive %.0f | total %.0f"
          % (m(c[:, 1] - c[:, 0]), m(c[:, 2] - c[:, 1]), m(c[:, 3] - c[:, 2]), m(c[:, 4] - c[:, 3]),
             m(c[1:, 0] - c[:-1, 0]) * -1))
