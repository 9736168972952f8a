import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import snn_oracle as orc
from tests.helpers import rel_err
from tests import test_gpu_parity as tp
from sparch_b200 import functional as F
DEV='cuda:0'
kind, Be, T, H, seed = "RLIF", 128, int(sys.argv[1]) if len(sys.argv)>1 else 100, 512, 3
rng = np.random.default_rng(seed)
I = (rng.standard_normal((Be, T, H)) * 3.0 + 1.2).astype(np.float32)
alpha = rng.uniform(0.80, 0.97, H).astype(np.float32)
beta = rng.uniform(0.96, 0.995, H).astype(np.float32)
a = rng.uniform(0.0, 1.2, H).astype(np.float32)
b = rng.uniform(-0.2, 2.2, H).astype(np.float32)
V = (rng.standard_normal((H, H)) / np.sqrt(H)).astype(np.float32)
u0, w0, s0 = (rng.uniform(0, 1, (Be, H)).astype(np.float32) for _ in range(3))
p = orc.clamp_params(kind, alpha, beta, a, b)
V0 = V.copy(); np.fill_diagonal(V0, 0)
r = orc.cell_forward(kind, I, p["alpha"], None, None, None, V0, u0, None, s0)
gs = rng.standard_normal((Be,T,H)).astype(np.float32)
bw = orc.cell_backward(kind, gs, I, p["alpha"], None, None, None, V0, u0, None, s0, U=r["u"], W=r["w"], S=r["s"])
t = lambda z, g=False: torch.from_numpy(np.ascontiguousarray(z)).to(DEV).requires_grad_(g)
print("rate", r["s"].mean(), "max|dI|", np.abs(bw["dI"]).max(), "max |u|", np.abs(r["u"]).max())
for mode in ("tc", "mma"):
    F.RECUR_BWD = mode
    It, al, Vt = t(I, True), t(alpha, True), t(V, True)
    S = F.SpikingCellFunction.apply(It, None, None, al, None, None, None, Vt, t(u0), None, t(s0), kind, 1.0, F.NormState("none"))
    print(mode, "flips", float((S.detach().cpu().numpy() != r["s"]).mean()))
    S.backward(t(gs))
    d = It.grad.cpu().numpy().astype(np.float64)
    e = np.abs(d - bw["dI"])
    print(mode, "Function dI rel_err", rel_err(d, bw["dI"]), "worst at", np.unravel_index(e.argmax(), e.shape), "per-t max err", [float(e[:, tt].max()) for tt in (0, 1, T//2, T-2, T-1)])
    pt = e.max(axis=(0,2)); print("   err by t (first 10, last 10):", np.round(pt[:10],6), np.round(pt[-10:],6))
F.RECUR_BWD = "tc"
cl = {"alpha": t(p["alpha"])}
dI, pg = tp._bwd_tc_direct(kind, t(gs), t(r["u"]), None, cl, t(V0), t(u0), None, t(s0))
print("direct tc dI rel_err", rel_err(dI.cpu().numpy(), bw["dI"]))
