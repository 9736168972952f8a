"""Forward recurrence of one RadLIF/RLIF layer with both kernels (tcgen05 int8 / mma.sync) on the same input:
CUDA-event time per step and, for the tcgen05 kernel with SPARCH_PHASES=1, the in-kernel phase clocks of CTA (0,0).
Usage: python tools/prof_fwd.py [T] [Be] [H] [kind]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparch_b200 import functional as F  # noqa: E402
from sparch_b200._lib import call, ptr  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
Be = int(sys.argv[2]) if len(sys.argv) > 2 else 256
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
kind = sys.argv[4] if len(sys.argv) > 4 else "RadLIF"
dev = "cuda:0"
adaptive = kind in ("adLIF", "RadLIF")
gen = torch.Generator(device=dev).manual_seed(1)
r = lambda *s: torch.rand(*s, device=dev, generator=gen)
I = torch.randn(Be, T, H, device=dev, generator=gen) * 4 + 2
alpha, beta, a, b = r(H) * 0.14 + 0.82, r(H) * 0.02 + 0.968, r(H), r(H) * 2
V = torch.nn.init.orthogonal_(torch.empty(H, H)).to(dev)
u0, w0, s0 = r(Be, H), r(Be, H), r(Be, H)


def run():
    return F.SpikingCellFunction.apply(I, None, None, alpha, beta if adaptive else None, a if adaptive else None,
                                       b if adaptive else None, V, u0, w0 if adaptive else None, s0, kind, 1.0,
                                       F.NormState("none"))


out = {}
for mode in ("mma", "tc"):
    F.RECUR_FWD = mode
    with torch.no_grad():
        for it in range(4):
            F.timers_enable(True)
            S = run()
            tm = F.timers_collect()
    out[mode] = S
    print(f"{mode}: rate {float(S.mean()):.3f} fwd {tm['recurrence_fwd']:.3f} ms ({tm['recurrence_fwd'] / T * 1e3:.2f} us per step)")
print("spike trains equal:", bool(torch.equal(out["mma"], out["tc"])),
      " flips:", float((out["mma"] != out["tc"]).float().mean()))

if os.environ.get("SPARCH_PHASES"):
    F.RECUR_FWD = "tc"
    dbg = torch.zeros(2 * T + 8, 8, dtype=torch.int64, device=dev)
    call("sparch_recur_debug_clocks", ptr(dbg))
    with torch.no_grad():
        run()
    torch.cuda.synchronize()
    call("sparch_recur_debug_clocks", None)
    c = dbg.cpu().double()[3:T - 1]     # steps t = 3 .. T-2
    prev_pub = dbg.cpu().double()[2:T - 2, 5]
    m = lambda x: float(x.mean())
    print("tc forward phases, cycles/step (CTA 0,0; columns: 0 expansion start, 1 first batch seen, 2 MMA sees batch, "
          "3 commit issued, 4 D complete seen, 5 published, 6 last batch expanded)")
    print("  own publish(t-1) -> expansion start %.0f -> first batch loaded %.0f" % (m(c[:, 0] - prev_pub), m(c[:, 1] - c[:, 0])))
    print("  own publish(t-1) -> first batch of step t seen %.0f | -> MMA wakes %.0f | -> last batch expanded %.0f | "
          "-> commit issued %.0f | -> D seen by update %.0f | -> published %.0f | total %.0f"
          % (m(c[:, 1] - prev_pub), m(c[:, 2] - c[:, 1]), m(c[:, 6] - c[:, 2]), m(c[:, 3] - c[:, 6]),
             m(c[:, 4] - c[:, 3]), m(c[:, 5] - c[:, 4]), m(c[:, 5] - prev_pub)))
    e = dbg.cpu().double()[T + 3:2 * T - 1]
    print("  relative to worker 0's first batch seen: MMA warp sees batches (rotated order) %s | workers r=1..3 done "
          "expanding %s | worker 0 done %.0f"
          % ([round(m(e[:, i] - c[:, 1])) for i in range(4)], [round(m(e[:, 4 + r] - c[:, 1])) for r in (1, 2, 3)],
             m(c[:, 6] - c[:, 1])))
