"""Data-parallel gradient synchronisation (one process per GPU, torch.distributed).

The reference is single-process (sparch/exp.py:81); batch rows are independent inside every
layer, so the path shards over the batch with ONE exchange per train step: a
sum all-reduce of the parameter gradients, averaged over ranks.  Gradients are bucketed per
layer of ``SNN.snn`` -- two buckets per layer: the projection ``W`` (its gradient is the LAST thing a
layer's backward produces) and everything else (recurrent ``V``, neuron parameters, normalisation:
produced by the cell's backward, before the projection's GEMMs) -- and each bucket's all-reduce is
launched asynchronously from a post-accumulate-grad hook as soon as its last gradient exists, so the
transfer over NVLink overlaps the BPTT of the layers below it and, inside the first layer, the 4 MB of
dV travel under the dW GEMM: only the small dW bucket of the first layer is left exposed at the end of
the backward pass.  ``average=False`` leaves the SUM in ``.grad`` (no division pass): the optimizer
applies 1 / world (``sparch_b200.optim.Adam(grad_scale=1 / world)``).  BatchNorm uses local (per-rank)
batch statistics -- plain data-parallel semantics.

``defer`` (default: on when the module has recurrent spiking layers on CUDA): a bucket whose last gradient exists is
not reduced at once but at the next SAFE POINT -- right after a reverse recurrence kernel has been issued, or in
``finish()``.  The tcgen05 reverse recurrence needs its 32 four-CTA clusters resident at the same time (four per GPC,
exactly what a B200 holds) and lives on L2 latency: with a collective's CTAs on a few SMs underneath it, it measured
14 % slower (1.39 vs 1.22 ms per step at 2 GPUs), more than the collectives take.  Issued after the kernel, a bucket's
all-reduce runs under the BatchNorm-backward passes and the gradient GEMMs that follow, which do not care.

Works with any backend ("nccl" on GPUs; "gloo" in the CPU tests).
"""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, module, group=None, broadcast=True, average=True, defer=None):
        if not dist.is_initialized():
            raise RuntimeError("GradSync needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.average = average
        self.buckets = []
        layers = list(module.snn) if hasattr(module, "snn") else [module]
        seen = set()
        for lay in layers:
            ps = [p for p in lay.parameters() if p.requires_grad and id(p) not in seen]
            seen.update(id(p) for p in ps)
            proj = getattr(lay, "W", None)
            wids = {id(p) for p in proj.parameters()} if isinstance(proj, torch.nn.Module) else set()
            for part in ([p for p in ps if id(p) not in wids], [p for p in ps if id(p) in wids]):
                if part:
                    self._add_bucket(part)
        rest = [p for p in module.parameters() if p.requires_grad and id(p) not in seen]
        if rest:
            self._add_bucket(rest)
        if broadcast:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, src=0, group=group)
        self._handles = []
        self._ready = []
        if defer is None:
            defer = any(getattr(lay, "_recurrent", False) and any(p.is_cuda for p in lay.parameters()) for lay in layers)
        self.defer = bool(defer)
        # reverse recurrences still to come in this backward pass: once the last one has been issued nothing is left to
        # protect, and buckets go out the moment they are complete again
        self._n_bptt = sum(1 for lay in layers if getattr(lay, "_recurrent", False))
        self._bptt_left = self._n_bptt
        if self.defer:
            import weakref
            from . import functional
            ref = weakref.WeakMethod(self._after_bptt)      # the registry must not keep a discarded GradSync alive

            def after_bptt():
                cb = ref()
                if cb is not None:
                    cb()
                elif after_bptt in functional.AFTER_BPTT:
                    functional.AFTER_BPTT.remove(after_bptt)
            functional.AFTER_BPTT.append(after_bptt)

    def _add_bucket(self, params):
        n = sum(p.numel() for p in params)
        flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
        b = {"params": params, "flat": flat, "pending": len(params), "work": None, "views": []}
        off = 0
        for p in params:
            b["views"].append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
            p.register_post_accumulate_grad_hook(self._make_hook(b, len(b["views"]) - 1))
        self.buckets.append(b)

    def _make_hook(self, b, i):
        def hook(p):
            b["views"][i].copy_(p.grad)
            b["pending"] -= 1
            if b["pending"] == 0:
                if self.defer and self._bptt_left > 0:
                    self._ready.append(b)
                else:
                    b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group,
                                                async_op=True)
        return hook

    def _after_bptt(self):
        self._bptt_left -= 1
        self.flush()

    def flush(self):
        """Start the all-reduce of every complete bucket that has not been started (safe point, see ``defer``)."""
        ready, self._ready = self._ready, []
        for b in ready:
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for the outstanding all-reduces, average, and point .grad at the reduced values."""
        self.flush()
        for b in self.buckets:
            if b["pending"] != 0:
                # a parameter received no gradient this step: reduce what there is
                for v, p in zip(b["views"], b["params"]):
                    if p.grad is None:
                        v.zero_()
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group,
                                            async_op=True)
        for b in self.buckets:
            b["work"].wait()
            if self.average:
                b["flat"].div_(self.world)
            for v, p in zip(b["views"], b["params"]):
                p.grad = v
            b["pending"] = len(b["params"])
            b["work"] = None
        self._bptt_left = self._n_bptt
