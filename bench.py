#!/usr/bin/env python
"""Benchmark of the sparch SNN hot path (BASELINE.json metric: train samples/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg4] [--impl ours|reference]

A "step" is one training step of the reference's loop (sparch/exp.py:355-377): forward through
SNN, cross-entropy, backward (BPTT), Adam update, on one batch of synthetic data of the
configured shape.  Default workload: cfg4 = RadLIF [1024,1024,35] + batchnorm + dropout 0.1 on
SC-shaped input (batch 256 per GPU, 100 steps, 40 features, 35 classes) -- BASELINE.json
configs[3], the one the metric is quoted on.

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs resident in HBM;
`e2e` = the same through the public module API with the batch copied from pinned host memory
and the loss read back every step; `roofline` = the membrane-recurrence kernels (the dominant
part of the step) against the measured HBM peak; `cpu_baseline` = the oracle's torch-CPU
restatement of the reference timed on this box's host cores.

`--impl reference` times the reference's CPU implementation of the same step (the oracle's
op-for-op restatement; the reference itself is pure Python/PyTorch and /root/reference does not
exist on the GPU box) with all host threads, on a bounded per-step batch.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg1": dict(desc="LIF [128,128,20], SHD-shaped (B128,T100,F700)", neuron_type="LIF",
                 layer_sizes=[128, 128, 20], dropout=0.0, normalization="none", B=128, T=100, F=700,
                 data="spikes"),
    "cfg2": dict(desc="adLIF [128,128,20] + batchnorm, SHD-shaped (B128,T100,F700)",
                 neuron_type="adLIF", layer_sizes=[128, 128, 20], dropout=0.0,
                 normalization="batchnorm", B=128, T=100, F=700, data="spikes"),
    "cfg3": dict(desc="RLIF [512,512,35], SSC-shaped (B128,T100,F700)", neuron_type="RLIF",
                 layer_sizes=[512, 512, 35], dropout=0.0, normalization="batchnorm", B=128, T=100,
                 F=700, data="spikes"),
    # cfg5's shape in fp32 (bf16 mode is not built): the reference itself is non-finite here with its
    # default init and its tape would need ~30 GB; a stable draw a <- |a| is used.
    "cfg5": dict(desc="bidirectional RadLIF [1024,1024,35] + batchnorm, SC-shaped long sequences (B512,T500,F40), fp32, a<-|a|",
                 neuron_type="RadLIF", layer_sizes=[1024, 1024, 35], dropout=0.1, normalization="batchnorm",
                 B=512, T=500, F=40, data="randn", bidirectional=True, stable_a=True),
    "cfg4": dict(desc="RadLIF [1024,1024,35] + batchnorm + dropout 0.1, SC-shaped (B256,T100,F40)",
                 neuron_type="RadLIF", layer_sizes=[1024, 1024, 35], dropout=0.1,
                 normalization="batchnorm", B=256, T=100, F=40, data="randn"),
}


def make_batch(cfg, B, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    if cfg["data"] == "spikes":
        x = (torch.rand(B, cfg["T"], cfg["F"], generator=g) < 0.03).float()
    else:
        x = torch.randn(B, cfg["T"], cfg["F"], generator=g)
    y = torch.randint(0, cfg["layer_sizes"][-1], (B,), generator=g)
    return x, y


def model_kwargs(cfg):
    return dict(layer_sizes=cfg["layer_sizes"], neuron_type=cfg["neuron_type"], dropout=cfg["dropout"],
                normalization=cfg["normalization"], bidirectional=cfg.get("bidirectional", False))


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.t_mark = 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def mark(self):
        """Start of the loaded window: samples before this moment are ignored."""
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if ts < self.t_mark + 0.05:
                continue
            f = [z.strip() for z in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(cfg, args, world):
    """The `config` object of a bench line: the workload, identical for both arms (--impl ours / reference)."""
    return {"workload": cfg["desc"], "bench_config": args.config, "per_gpu_batch": cfg["B"],
            "global_batch": cfg["B"] * world, "parallelism": f"dp{world}",
            "l2": "per-step working set (>=1 GB of activations/tapes at the headline configuration) exceeds the 126 MB L2"}


# --------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_steps(cfg, B, steps, warmup):
    """Oracle restatement of the reference train step on the host cores; returns samples/s."""
    import torch
    from oracle import snn_oracle as orc
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    net = orc.build_oracle_snn((B, None, cfg["F"]), **model_kwargs(cfg))
    opt = torch.optim.Adam(net.parameters(), 1e-2)
    x, y = make_batch(cfg, B, 1234)
    net.train()
    for _ in range(warmup):
        orc.oracle_train_step(net, opt, x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.oracle_train_step(net, opt, x, y)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    # bounded per-step sample: shrink the per-step batch so K+W steps end within ~2 minutes
    B = min(cfg["B"], 32)
    _, t32 = cpu_reference_steps(cfg, B, 1, 1)
    budget = 120.0 / max(1, args.steps + args.warmup)
    scale = max(1.0, budget / max(t32, 1e-3))
    Bs = int(min(cfg["B"], max(16, (int(B * scale) // 16) * 16)))
    sps, t = cpu_reference_steps(cfg, Bs, args.steps, args.warmup)
    cores = os.cpu_count()
    sample = (f"{args.steps} train steps of the same model at per-step batch {Bs} "
              f"(config batch {cfg['B']}), T={cfg['T']}")
    line = {
        "impl": "reference", "metric": "train samples/sec", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(cfg, args, world),   # (the bounded sample: cpu_baseline.sample)
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm (B200)
def run_ours(args, cfg, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import sparch_b200
    from sparch_b200 import functional as F
    from sparch_b200 import parallel

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # the persistent recurrence kernels occupy 128 of the 148 SMs for most of the step: a collective that asks for
        # more CTAs than the 20 SMs left either waits for them or delays their launch (NCCL's default is up to 32)
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        dist.init_process_group("nccl", device_id=dev)
    B = cfg["B"]
    sparch_b200.set_state_init(args.state_init)
    sparch_b200.set_precision(args.precision)
    torch.manual_seed(0)
    net = sparch_b200.SNN((B, None, cfg["F"]), **model_kwargs(cfg)).to(dev)
    if cfg.get("stable_a"):
        with torch.no_grad():
            for lay in net.snn:
                if hasattr(lay, "a"):
                    lay.a.abs_()
    net.train()
    use_graph = (not args.no_graph) and args.state_init == "device"
    # exp.py:89's Adam: the same update as one launch over all parameter tensors with the step count on the device
    # (sparch_b200/optim.py); --adam torch uses torch.optim.Adam(fused=True)
    if args.adam == "sparch":
        from sparch_b200.optim import Adam
        opt = Adam(net.parameters(), 1e-2, grad_scale=1.0 / world)   # .grad holds the all-reduced SUM (no division pass)
    else:
        opt = torch.optim.Adam(net.parameters(), 1e-2, capturable=use_graph, fused=True)
    sync = parallel.GradSync(net, average=args.adam != "sparch") if world > 1 else None
    x_h, y_h = make_batch(cfg, B, 1234 + rank)
    x_h, y_h = x_h.pin_memory(), y_h.pin_memory()
    x_d, y_d = x_h.to(dev), y_h.to(dev)
    loss_fn = sparch_b200.CrossEntropyLoss()                  # exp.py:83's nn.CrossEntropyLoss() as one launch each way

    def step(x, y):
        out, _ = net(x)                                       # exp.py:359
        loss = loss_fn(out, y)
        opt.zero_grad()
        loss.backward()                                       # exp.py:376
        if sync is not None:
            sync.finish()
        opt.step()
        return loss

    graphed = None
    if use_graph:
        from sparch_b200.graphs import GraphedTrainStep
        graphed = GraphedTrainStep(net, opt, loss_fn, x_d, y_d, sync=sync)
        eager_step = step

        def step(x, y):                                       # noqa: F811
            return graphed.step(x, y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if args.profile:
        for _ in range(args.warmup):
            step(x_d, y_d)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(args.steps):
            step(x_d, y_d)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if rank == 0:
            print(json.dumps({"profile_run": True, "steps": args.steps, "warmup": args.warmup}))
        return
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                       # nvidia-smi needs ~1 s to start sampling
    for _ in range(max(args.warmup, 3)):
        step(x_d, y_d)
    # ---- device-resident number ------------------------------------------------------------
    torch.cuda.synchronize()
    sampler.mark()
    F.timers_enable(True)
    n0 = F.native_launches()
    ms = timed(lambda: step(x_d, y_d), args.steps)
    launches = F.native_launches() - n0
    if graphed is not None:
        launches = graphed.native_calls_per_step * args.steps  # replayed from the graph, not re-issued by the host
    rec_ms = F.timers_collect()                               # {"recurrence_fwd": ms, ...} totals
    F.timers_enable(False)
    value = world * B * args.steps / (ms * 1e-3)
    if graphed is not None:
        # CUDA events cannot be placed inside a replayed graph: the per-kernel durations behind the roofline come
        # from the same steps issued eagerly (same kernels, same buffers) right after the timed region
        F.timers_enable(True)
        for _ in range(args.steps):
            eager_step(x_d, y_d)
        rec_ms = F.timers_collect()
        F.timers_enable(False)

    # ---- end to end: pinned host batch -> device, loss read back every step ------------------
    events = None
    if cfg["data"] == "spikes" and args.input == "events" and graphed is not None:
        # SHD / SSC batches as the files hold them: event lists.  The host ships the events (8 bytes each), the device
        # builds the dense (B, T, 700) count tensor (sparch_b200/data.py = spiking_datasets.py:66-86) straight into the
        # graph's input buffer -- the same input tensor as the dense leg, 1/9 of the host -> device bytes.
        from sparch_b200.data import SpikingBatcher
        bat = SpikingBatcher(nb_steps=cfg["T"], nb_units=cfg["F"], device=dev)
        idx = x_h.nonzero()                                           # synthetic events of the synthetic dense batch
        tb = bat.time_bins
        ev_t = [None] * B
        ev_u = [None] * B
        mid = torch.from_numpy(((tb[:-1] + tb[1:]) / 2).astype("float32"))
        for b_ in range(B):
            sel = idx[idx[:, 0] == b_]
            keep = sel[:, 1] >= 1                                     # bin 0 holds only negative times: never produced
            ev_t[b_] = mid[sel[keep, 1] - 1].numpy()                  # a time inside bin k lies between edges k-1 and k
            ev_u[b_] = sel[keep, 2].numpy()
        ev = bat.pack(ev_t, ev_u, y_h.numpy())
        ev_dev = [torch.empty_like(t, device=dev) for t in ev[:3]]
        nev = int(ev[2][-1])
        bad = torch.empty(1, device=dev, dtype=torch.int32)
        from sparch_b200._lib import call as _call, ptr as _ptr
        events = {"h2d": sum(t.numel() * t.element_size() for t in ev[:3]) + y_h.numel() * 8, "n": nev}

        def e2e_step():
            for d_, h_ in zip(ev_dev, ev[:3]):
                d_.copy_(h_, non_blocking=True)
            graphed.y.copy_(y_h, non_blocking=True)
            _call("sparch_events_to_dense", _ptr(ev_dev[0]), _ptr(ev_dev[1]), _ptr(ev_dev[2]), _ptr(bat._bins), B, cfg["T"],
                  cfg["F"], nev, _ptr(graphed.x), _ptr(bad), torch.cuda.current_stream().cuda_stream)
            graphed._refresh_hyper()
            graphed.graph.replay()
            return float(graphed.loss.item())
    elif graphed is not None:
        # every step: its batch pinned-host -> device (on a copy stream, under the previous step: input double
        # buffering), the step, the loss read back by the host (exp.py:363 loss.item())
        graphed.stage(x_h, y_h)

        def e2e_step():
            loss = graphed.step_staged()
            graphed.stage(x_h, y_h)                           # the NEXT step's batch
            return float(loss.item())
    else:
        def e2e_step():
            x_d.copy_(x_h, non_blocking=True)
            y_d.copy_(y_h, non_blocking=True)
            return float(step(x_d, y_d).item())               # exp.py:363 loss.item()

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    # Short runs end before nvidia-smi has produced enough samples: keep the GPU under the same load
    # for ~1.5 s more.  The step count comes from the rank-reduced timing, so every rank runs the
    # same number of (collective-bearing) steps.
    loaded_s = (ms + ms_e2e) * 1e-3
    if loaded_s < 1.5:
        for _ in range(int((1.5 - loaded_s) / (ms * 1e-3 / args.steps)) + 1):
            step(x_d, y_d)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None            # samples cover the timed + e2e (+ filler) load

    # ---- the drop-in figure: what sparch/exp.py's own loop gets after INTEGRATION.md's one-line switch, nothing else
    # changed -- default (reference-exact, CPU generator) state draws, eager launches, torch.optim.Adam as exp.py:89
    # builds it, nn.CrossEntropyLoss, x.to(device) / y.to(device) from pinned host tensors and loss.item() every step
    # (exp.py:355-363, 375-377).  Single GPU only (the reference loop has no gradient exchange); a few steps.
    dropin = None
    if world == 1 and not args.no_dropin:
        sparch_b200.set_state_init("cpu")
        torch.manual_seed(0)
        net2 = sparch_b200.SNN((B, None, cfg["F"]), **model_kwargs(cfg)).to(dev)
        net2.train()
        opt2 = torch.optim.Adam(net2.parameters(), 1e-2)
        loss2 = torch.nn.CrossEntropyLoss()

        def dropin_step():
            x, y = x_h.to(dev), y_h.to(dev)
            out, _ = net2(x)
            l = loss2(out, y)
            v = l.item()
            opt2.zero_grad()
            l.backward()
            opt2.step()
            return v

        for _ in range(3):
            dropin_step()
        nd = max(3, min(args.steps, 10))
        ms_d = timed(dropin_step, nd)
        dropin = {"value": B * nd / (ms_d * 1e-3), "unit": "samples/s", "ms_per_step": ms_d / nd, "steps": nd,
                  "what": "sparch/exp.py's loop unchanged (reference-exact CPU-generator state draws, eager, "
                          "torch.optim.Adam, nn.CrossEntropyLoss, .to(device) + loss.item() per step)"}
        sparch_b200.set_state_init(args.state_init)
        del net2, opt2

    if rank != 0:
        # Hard exit: tearing down a process group whose collectives live inside a captured CUDA graph was seen to
        # hang at interpreter exit; this rank's measured work is complete and synchronised.
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)
    # ---- roofline of the recurrence kernels (16.5 B per (b,t,h) per spiking layer)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n_spiking = len(cfg["layer_sizes"]) - 1
    Be = B * (2 if cfg.get("bidirectional") else 1)
    elts = sum(Be * cfg["T"] * h for h in cfg["layer_sizes"][:n_spiking])
    alg_bytes = 16.5 * elts
    rec_total_ms = (rec_ms.get("recurrence_fwd", 0.0) + rec_ms.get("recurrence_bwd", 0.0)) / args.steps
    achieved = alg_bytes / (rec_total_ms * 1e-3) / 1e9 if rec_total_ms > 0 else None
    traffic = None
    for name in ("r02_recur_traffic.json", "r01_recur_traffic.json"):   # from the committed ncu --set full capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            if tj.get("config") == args.config:
                traffic = tj["traffic_bytes_per_step"]
                break
        except Exception:
            pass
    # The recurrence is a chain of 2 T dependent steps per layer; a step cannot be shorter than one inter-CTA
    # hand-over through L2 (measured by tools/ubench/exchange_rtt.cu: an EMPTY persistent kernel on the same grid that
    # only publishes and polls tagged words, profiles/r02_ubench_exchange_rtt.txt) nor than its tensor work at the
    # measured dense peak.  floor = T * layers * (t_fwd + t_bwd), t = max(hand-over, tensor time).
    latency = None
    try:
        import re as _re
        rtts = [float(x) for x in _re.findall(r"tagged words \(64 loads in flight\) grid 32 x 4.*?= ([0-9.]+) ns/step",
                                              open(os.path.join(ROOT, "profiles", "r02_ubench_exchange_rtt.txt")).read())]
        rtt_us = min(rtts) * 1e-3
        tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1387.7)))
        hs = [h for h in cfg["layer_sizes"][:n_spiking]]
        rec_layers = [h for h in hs] if cfg["neuron_type"] in ("RLIF", "RadLIF") else []
        if rec_layers:
            t_f = [max(rtt_us, 2.0 * Be * h * h * 2 / (tf * 1e12) * 1e6) for h in rec_layers]   # spikes x (hi, lo) of V0
            t_b = [max(rtt_us, 2.0 * Be * h * h * 3 / (tf * 1e12) * 1e6) for h in rec_layers]   # three fp16 passes
            floor_ms = cfg["T"] * (sum(t_f) + sum(t_b)) * 1e-3
            latency = {"exchange_rtt_us": rtt_us, "t_step_floor_us": {"fwd": t_f, "bwd": t_b}, "floor_ms": floor_ms,
                       "frac_of_floor": floor_ms / rec_total_ms if rec_total_ms else None,
                       "measured_us_per_layer_step": {"fwd": rec_ms.get("recurrence_fwd", 0.0) / args.steps /
                                                      (cfg["T"] * len(rec_layers)) * 1e3,
                                                      "bwd": rec_ms.get("recurrence_bwd", 0.0) / args.steps /
                                                      (cfg["T"] * len(rec_layers)) * 1e3}}
    except Exception:
        latency = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "kernel": "membrane recurrence fwd+bwd (all launches of one train step)",
                "ms_per_step": rec_total_ms,
                "ms_fwd": rec_ms.get("recurrence_fwd", 0.0) / args.steps,
                "ms_bwd": rec_ms.get("recurrence_bwd", 0.0) / args.steps, "algorithmic_bytes_per_step": alg_bytes,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "share_of_step": rec_total_ms / (ms / args.steps) if rec_total_ms else None,
                "timed": ("CUDA events around the recurrence launches of the same steps issued eagerly right after "
                          "the graphed timed region") if graphed is not None else
                         "CUDA events around the recurrence launches inside the timed region",
                "latency": latency,
                "note": "the recurrence is bounded by its 2*T dependent steps per layer (exchange latency + tensor "
                        "issue), not by HBM: `latency` holds the floor from the measured hand-over round trip"}

    # ---- CPU baseline beside it (bounded sample: the full config batch, 2 timed steps) --------
    cpu = None
    if not args.no_cpu_baseline:
        Bc = cfg["B"]
        sps, t = cpu_reference_steps(cfg, Bc, 2, 1)
        cpu = {"value": sps, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"2 timed train steps (1 warm-up) of the full config at batch {Bc}, "
                         f"{t:.2f} s/step"}

    line = {
        "metric": "train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "bf16 products, f32 state/accumulation",
        "data": "synthetic",
        "config": workload_config(cfg, args, world),
        "mode": {"cuda_graph": bool(use_graph),
                 "state_init": args.state_init + (" generator draws of u0/w0/s0 ~ U[0,1) (same distribution "
                                                  "as the reference's CPU draws)" if args.state_init == "device"
                                                  else " generator draws, identical to the reference's")},
        "e2e": {"value": e2e_value, "unit": "samples/s",
                "h2d_bytes_per_step": events["h2d"] if events else x_h.numel() * 4 + y_h.numel() * 8,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                "input": (f"event lists ({events['n']} events per batch), dense (B,T,F) tensor built on the device"
                          if events else "dense fp32 batch from pinned host memory")},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "regions_ms_per_step": {k: v / args.steps for k, v in rec_ms.items()},
    }
    if dropin is not None:
        line["e2e_dropin"] = dropin
    print(json.dumps(line), flush=True)
    if world > 1:
        sys.stderr.flush()
        os._exit(0)                                           # see the note at the other ranks' exit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS),
                    help="cfg4 (default) is the configuration BASELINE.json's metric is quoted on")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--state-init", default="device", choices=["device", "cpu"],
                    help="where the per-forward initial states ~U[0,1) are drawn: 'device' (CUDA generator) "
                         "or 'cpu' (the reference's CPU-generator draws, snns.py:700-702; host-bound)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"],
                    help="fp32 (default, the reference's arithmetic to ~1e-6) or the reduced-precision mode")
    ap.add_argument("--graph", action="store_true",
                    help="(default) replay the whole train step -- forward, loss, backward, gradient all-reduce, "
                         "Adam -- as one CUDA graph (sparch_b200.graphs.GraphedTrainStep; needs device state init)")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly, launch by launch")
    ap.add_argument("--adam", default="sparch", choices=["sparch", "torch"],
                    help="optimizer of the train step: sparch_b200.optim.Adam (one launch) or torch.optim.Adam(fused=True)")
    ap.add_argument("--input", choices=["events", "dense"], default="events",
                    help="spike-shaped configs (cfg1-3): ship the batch as event lists (default) or as the dense tensor")
    ap.add_argument("--no-dropin", action="store_true", help="skip the e2e_dropin leg (exp.py's loop unchanged)")
    ap.add_argument("--profile", action="store_true",
                    help="profiling run (under ncu): exactly --warmup warm-up and --steps timed steps of "
                         "the device-resident loop, no e2e leg, no CPU baseline; prints no bench value")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
               str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    run_ours(args, cfg, rank, local_rank, world)


if __name__ == "__main__":
    main()
