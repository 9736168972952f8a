"""Adam for the train step (sparch/exp.py:89 builds ``torch.optim.Adam(net.parameters(), lr)``): same update, one
kernel launch for all parameter tensors (csrc/optim.cu).  The step count AND the hyper-parameters live on the device,
so the step can be captured in a CUDA graph and a learning-rate scheduler (exp.py:92-96, ``ReduceLROnPlateau`` edits
``param_groups[i]["lr"]``) still takes effect on replay: ``sync_hyper()`` copies changed values into the device words
(``GraphedTrainStep.step`` calls it before every replay).  The step count is kept in ``state`` (key ``"step"`` of the
group's first parameter) so that it round-trips through ``state_dict()`` / ``load_state_dict()``.
Defaults only: no weight decay, no amsgrad, no maximize.  CUDA fp32 parameters only (no CPU path)."""
import ctypes

import torch

from ._lib import call


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        """grad_scale: factor applied to every gradient inside the kernel (1 / world size when ``.grad`` holds the
        all-reduced SUM of the ranks' gradients: ``parallel.GradSync(average=False)``)."""
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, grad_scale=grad_scale))
        self._hyper = {}        # group index -> (device float[5], tuple of the values it holds)

    @staticmethod
    def _values(group):
        return (float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                float(group.get("grad_scale", 1.0)))

    def sync_hyper(self):
        """Bring the device copies of (lr, betas, eps, grad_scale) up to date with ``param_groups`` (a few bytes, only
        when a value changed).  Stream-ordered: takes effect for every launch / graph replay issued afterwards."""
        for gi, group in enumerate(self.param_groups):
            vals = self._values(group)
            ent = self._hyper.get(gi)
            if ent is None:
                dev = next((p.device for p in group["params"]), None)
                if dev is None or dev.type != "cuda":
                    continue
                ent = [torch.empty(5, device=dev, dtype=torch.float32), None]
                self._hyper[gi] = ent
            if ent[1] != vals:
                ent[0].copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=False)
                ent[1] = vals

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self.sync_hyper()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("sparch_b200.optim.Adam updates contiguous CUDA fp32 parameters only")
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
            if gi not in self._hyper:
                raise RuntimeError("Adam.step() was first called during stream capture: run one eager step (or "
                                   "sync_hyper()) before capturing")
            first = self.state[group["params"][0]]
            if "step" not in first:
                first["step"] = torch.zeros(1, device=ps[0].device, dtype=torch.int64)
            step_t = first["step"]
            if step_t.dtype != torch.int64 or not step_t.is_cuda:     # after load_state_dict from another device / dtype
                step_t = first["step"] = step_t.to(device=ps[0].device, dtype=torch.int64).reshape(1)
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
                     self._hyper[gi][0].data_ptr(), stream)
        return loss
