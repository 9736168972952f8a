"""Compare the two reverse-recurrence kernels (mma.sync vs tcgen05) on the same tape.
Usage: python tools/cmp_bwd.py [T] [Be] [H] [kind]"""
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
adaptive = kind in ("adLIF", "RadLIF")
out = {}
for mode in ("mma", "tc"):
    F.RECUR_BWD = mode
    gen = torch.Generator(device=dev).manual_seed(1)
    r = lambda *s: torch.rand(*s, device=dev, generator=gen)
    I = (torch.randn(Be, T, H, device=dev, generator=gen) * 4 + 2).requires_grad_(True)
    alpha = (r(H) * 0.14 + 0.82).requires_grad_(True)
    beta = (r(H) * 0.02 + 0.968).requires_grad_(True)
    a = r(H).requires_grad_(True)
    b = (r(H) * 2).requires_grad_(True)
    V = (torch.randn(H, H, device=dev, generator=gen) / H ** 0.5).requires_grad_(True)
    u0, w0, s0 = r(Be, H), r(Be, H), r(Be, H)
    g = torch.randn(Be, T, H, device=dev, generator=gen)
    for it in range(3):
        for x in (I, alpha, beta, a, b, V):
            x.grad = None
        F.timers_enable(True)
        S = F.SpikingCellFunction.apply(I, None, None, alpha, beta if adaptive else None,
                                        a if adaptive else None, b if adaptive else None, V, u0,
                                        w0 if adaptive else None, s0, kind, 1.0, F.NormState("none"))
        S.backward(g)
        tm = F.timers_collect()
    print(f"{mode}: rate {float(S.mean()):.3f} bwd {tm['recurrence_bwd']:.3f} ms "
          f"({tm['recurrence_bwd'] / T * 1e3:.1f} us per step)")
    out[mode] = [x.grad.clone() for x in ((I, alpha, beta, a, b, V) if adaptive else (I, alpha, V))]
    if mode == "tc" and os.environ.get("SPARCH_PHASES"):
        from sparch_b200._lib import call, ptr
        dbg = torch.zeros(2 * T + 8, 8, dtype=torch.int64, device=dev)
        S = F.SpikingCellFunction.apply(I, None, None, alpha, beta if adaptive else None,
                                        a if adaptive else None, b if adaptive else None, V, u0,
                                        w0 if adaptive else None, s0, kind, 1.0, F.NormState("none"))
        torch.cuda.synchronize()
        call("sparch_recur_debug_clocks", ptr(dbg))
        S.backward(g)
        torch.cuda.synchronize()
        call("sparch_recur_debug_clocks", None)
        e = dbg.cpu().double()[T + 2:2 * T - 3]
        c = dbg.cpu().double()[2:T - 3]   # rows t = 2 .. T-4; step t-1 follows step t in time
        m = lambda x: float(x.mean())
        print("tc phases, cycles/step (CTA 0,0): step start -> counter seen %.0f | -> first k-block landed %.0f | "
              "-> last UMMA issued %.0f | -> D complete %.0f | -> quarters received %.0f | -> released %.0f | total %.0f"
              % (m(c[:, 0] - c[:, 5]), m(c[:, 1] - c[:, 0]), m(c[:, 2] - c[:, 1]), m(c[:, 3] - c[:, 2]),
                 m(c[:, 6] - c[:, 3]), m(c[:, 4] - c[:, 6]), m(c[:-1, 5] - c[1:, 5])))
        print("  D complete -> TMEM read + scatter stores issued %.0f | -> arrivals sent %.0f | -> quarters received %.0f | "
              "-> BPTT update done %.0f | -> chunk max exchanged %.0f | -> panel stores issued %.0f | -> all warps' stores issued %.0f "
              "| -> released %.0f" % (m(e[:, 4] - c[:, 3]), m(e[:, 5] - e[:, 4]), m(c[:, 6] - e[:, 5]), m(e[:, 0] - c[:, 6]),
                                     m(e[:, 1] - e[:, 0]), m(e[:, 2] - e[:, 1]), m(e[:, 3] - e[:, 2]), m(c[:, 4] - e[:, 3])))
names = ("dI", "dalpha", "dbeta", "da", "db", "dV") if adaptive else ("dI", "dalpha", "dV")
for n, x, y in zip(names, out["mma"], out["tc"]):
    den = float(x.abs().max())
    print(f"{n}: max|mma| {den:.4g}  max diff {float((x - y).abs().max()):.4g}  rel {float((x - y).abs().max()) / max(den, 1e-30):.3g}")
