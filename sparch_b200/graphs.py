"""Whole-train-step CUDA graph (the glue of sparch/exp.py:352-382).

The reference's step is ``net(x)`` -> ``CrossEntropyLoss`` -> ``backward`` -> ``Adam.step`` with two
host syncs per step (exp.py:363, 381).  Small configurations (3x128 LIF/adLIF) launch a few hundred
short kernels per step and are bound by launch overhead, not by the GPU.  ``GraphedTrainStep``
captures forward + loss + backward + optimizer update ONCE into a CUDA graph -- every library
call made by ``sparch_b200`` (including the cooperative persistent kernels, the TMA-fed GEMMs and
the memsets inside the C ABI) is a plain stream operation -- and replays it per batch.

Requirements: ``set_state_init("device")`` (the CPU-generator draws of the default mode are host
work and cannot be captured), an optimizer built with ``capturable=True``, fixed batch shape.
"""
import torch

from . import functional as F
from . import snns


class GraphedTrainStep:
    def __init__(self, net, optimizer, loss_fn, x_example, y_example, warmup=3, sync=None):
        """sync: a ``parallel.GradSync`` (data-parallel runs): its bucketed all-reduces are issued by the
        gradient hooks during the captured backward and become part of the graph (NCCL kernels on the
        process group's stream, joined back before the optimizer step)."""
        if snns._STATE_INIT != "device":
            raise RuntimeError('GraphedTrainStep needs sparch_b200.set_state_init("device")')
        if not x_example.is_cuda:
            raise RuntimeError("GraphedTrainStep runs on CUDA tensors only")
        self.net, self.opt, self.loss_fn, self.sync = net, optimizer, loss_fn, sync
        self._copy_stream = None
        self.x = torch.empty_like(x_example)
        self.y = torch.empty_like(y_example)
        self.x.copy_(x_example)
        self.y.copy_(y_example)
        F.timers_enable(False)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up: one-time attribute calls, allocator pools, Adam state
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        n0 = F.native_launches()
        # thread_local: the NCCL watchdog thread polls its events while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local" if sync is not None else "global"):
            self.out, self.rates, self.loss = self._eager_step()
        self.native_calls_per_step = F.native_launches() - n0

    def _eager_step(self):
        out, rates = self.net(self.x)
        loss = self.loss_fn(out, self.y)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.sync is not None:
            self.sync.finish()
        self.opt.step()
        return out, rates, loss

    def step(self, x, y):
        """Copy the batch into the static buffers and replay; returns the (device) loss tensor."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self._refresh_hyper()
        self.graph.replay()
        return self.loss

    def _refresh_hyper(self):
        """sparch_b200.optim.Adam keeps lr / betas / eps in device words the captured kernel reads: a scheduler's change of
        ``param_groups[i]["lr"]`` (exp.py:92-96) reaches the replay through this copy.  (torch.optim.Adam(capturable=True)
        needs ``lr`` as a tensor for the same effect.)"""
        sync = getattr(self.opt, "sync_hyper", None)
        if sync is not None:
            sync()

    # ---- input double-buffering: the next batch's host->device copy runs on a copy stream under the current step
    def stage(self, x_host, y_host):
        """Start copying the NEXT batch (pinned host tensors) into staging buffers on a side stream."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._sx, self._sy = torch.empty_like(self.x), torch.empty_like(self.y)
            self._staged, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed)      # the previous staged batch has been taken over
            self._sx.copy_(x_host, non_blocking=True)
            self._sy.copy_(y_host, non_blocking=True)
            self._staged.record()

    def step_staged(self):
        """Run one step on the batch handed to stage(); returns the (device) loss tensor."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        self.x.copy_(self._sx, non_blocking=True)             # device-to-device, a few microseconds
        self.y.copy_(self._sy, non_blocking=True)
        self._consumed.record()
        self._refresh_hyper()
        self.graph.replay()
        return self.loss
