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
        c = dbg.cpu().double()[2:T - 3]            # update thread 64 of CTA (0,0), rows t = 2 .. T-4 (step t-1 follows step t)
        e = dbg.cpu().double()[T + 2:2 * T - 3]    # MMA warp of the same CTA
        m = lambda x: float(x.mean())
        print("tc phases, cycles/step (CTA 0,0): step start (tape loads issued) -> first K block of dI_{t+1} valid %.0f | "
              "-> all K blocks in shared memory %.0f | -> D complete %.0f | -> scatter stores issued %.0f | "
              "-> quarters received %.0f | -> update done + dI_t published %.0f | -> next step start %.0f | total %.0f"
              % (m(c[:, 1] - c[:, 0]), m(c[:, 2] - c[:, 1]), m(c[:, 3] - c[:, 2]), m(c[:, 4] - c[:, 3]),
                 m(c[:, 5] - c[:, 4]), m(c[:, 6] - c[:, 5]), m(c[:-1, 0] - c[1:, 6]), m(c[:-1, 0] - c[1:, 0])))
        print("  all K blocks stored -> quarter maxima sent %.0f | quarters received -> sums + buffers released %.0f | "
              "update done -> dI stores issued %.0f | -> next step's tape loads issued %.0f"
              % (m(c[:, 7] - c[:, 2]), m(e[:, 2] - c[:, 5]), m(e[:, 3] - c[:, 6]), m(c[:-1, 0] - e[1:, 3])))
        print("  first K block valid -> K block 0 / 1 / 2 / 3 stored and signalled %.0f / %.0f / %.0f / %.0f"
              % tuple(m(e[:, 4 + i] - c[:, 1]) for i in range(4)))
        print("  warp 2 of CTA (0,0), all steps: sample rounds entered at K block 0..3", dbg.cpu()[2 * T, :4].tolist(),
              "data re-reads at K block 0..3", dbg.cpu()[2 * T, 4:8].tolist())
        print("  MMA warp: first K block seen -> last UMMA issued %.0f | update warps: last K block stored -> first K block "
              "seen by the MMA warp %.0f | last UMMA issued -> D complete seen %.0f"
              % (m(e[:, 1] - e[:, 0]), m(e[:, 0] - c[:, 2]), m(c[:, 3] - e[:, 1])))
names = ("dI", "dalpha", "dbeta", "da", "db", "dV") if adaptive else ("dI", "dalpha", "dV")
for n, x, y in zip(names, out["mma"], out["tc"]):
    den = float(x.abs().max())
    print(f"{n}: max|mma| {den:.4g}  max diff {float((x - y).abs().max()):.4g}  rel {float((x - y).abs().max()) / max(den, 1e-30):.3g}")
