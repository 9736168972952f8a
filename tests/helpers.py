"""Shared helpers for the parity tests (oracle side only; no product code here)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not p.endswith(("boxcar_kat.npz", "init_contract.npz", "reference_module_run.npz")))


class Golden:
    """One fixture written by oracle/make_golden.py from the untouched reference."""

    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = {k: z[k] for k in z.files}
        meta = json.loads(str(self.z["meta"]))
        self.kwargs = meta["kwargs"]
        self.xshape = tuple(meta["xshape"])
        self.eval = meta["eval"]
        self.input_shape = (self.xshape[0], None) + tuple(self.xshape[2:])

    def t(self, key, device="cpu"):
        return torch.from_numpy(self.z[key]).to(device)

    def state_dict(self, device="cpu"):
        return {k[4:]: torch.from_numpy(v).to(device) for k, v in self.z.items()
                if k.startswith("sd0.")}

    def grads(self):
        return {k[5:]: v for k, v in self.z.items() if k.startswith("grad.")}

    def draws(self):
        n = sum(1 for k in self.z if k.startswith("draw."))
        return [self.z[f"draw.{i}"] for i in range(n)]

    def cur(self, i):
        """Input current of layer i as (Be, T, H) (the norm hook saw it flattened)."""
        c = self.z[f"cur.{i}"]
        return c.reshape(-1, self.xshape[1], c.shape[-1])

    def loss_fn(self, out, y):
        if out.ndim == 2:
            return torch.nn.functional.cross_entropy(out, y)
        return out.sum(1).square().mean()


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / den)


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
