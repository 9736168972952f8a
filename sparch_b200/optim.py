"""Adam for the train step (sparch/exp.py:89 builds ``torch.optim.Adam(net.parameters(), lr)``): same update, one
kernel launch for all parameter tensors (csrc/optim.cu), step count on the device so the step can be captured in a
CUDA graph.  Defaults only: no weight decay, no amsgrad, no maximize.  CUDA fp32 parameters only (no CPU path)."""
import ctypes

import torch

from ._lib import call


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._step_t = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("sparch_b200.optim.Adam updates contiguous CUDA fp32 parameters only")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
            dev = ps[0].device
            if gi not in self._step_t:
                self._step_t[gi] = torch.zeros(1, device=dev, dtype=torch.int64)
            step_t = self._step_t[gi]
            step_t.add_(1)
            stream = torch.cuda.current_stream().cuda_stream
            for i0 in range(0, len(ps), 48):
                chunk = ps[i0:i0 + 48]
                n = len(chunk)
                arr = lambda xs: (ctypes.c_void_p * n)(*xs)
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in chunk]
                call("sparch_adam_step", n, arr([p.data_ptr() for p in chunk]), arr([g.data_ptr() for g in grads]),
                     arr([self.state[p]["exp_avg"].data_ptr() for p in chunk]),
                     arr([self.state[p]["exp_avg_sq"].data_ptr() for p in chunk]),
                     (ctypes.c_int64 * n)(*[p.numel() for p in chunk]), step_t.data_ptr(),
                     float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                     stream)
        return loss
