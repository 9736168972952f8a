"""torch.autograd.Function wrappers over the C ABI (include/sparch_b200.h).

These replace the autograd tape the reference builds op by op inside its Python time
loops (sparch/models/snns.py:282-303, 419-445, 554-578, 696-727, 807-825) and its
``SpikeFunctionBoxcar`` (snns.py:20-36).  Inputs must be CUDA fp32 tensors: there is no
CPU path (``RuntimeError`` otherwise).
"""
import ctypes
import functools
import math
import os

import torch

from . import _lib, gemm
from ._lib import call, ptr

KINDS = {"LIF": 0, "adLIF": 1, "RLIF": 2, "RadLIF": 3}

# parameter limits, snns.py:229 / 356-359
ALPHA_LIM = (math.exp(-1 / 5), math.exp(-1 / 25))
BETA_LIM = (math.exp(-1 / 30), math.exp(-1 / 120))
A_LIM = (-1.0, 1.0)
B_LIM = (0.0, 2.0)
_LIMS = (ctypes.c_float * 8)(*ALPHA_LIM, *BETA_LIM, *A_LIM, *B_LIM)   # host array for the C ABI


def _stream():
    return torch.cuda.current_stream().cuda_stream


# Arithmetic mode of the tensor-pipe products.
#   "fp32" (default): fp32-equivalent -- three bf16 terms per fp32 GEMM operand, fp16 hi + lo for the
#           recurrent matrix and the dI panels; matches the reference to ~1e-6.
#   "bf16": reduced precision -- one bf16 term per GEMM operand, hi terms only in the recurrence.
#           State, accumulation, BatchNorm and the neuron update stay fp32.  Stated tolerance
#           (tests/test_gpu_parity.py::test_reduced_precision_mode): teacher-forced single steps flip
#           <= 2e-3 of the spikes, gradients given the masks agree to 3e-2 relative L2.
_PRECISION = "fp32"


def set_precision(mode):
    global _PRECISION
    if mode not in ("fp32", "fp32-bf16x3", "bf16"):
        raise ValueError("precision must be 'fp32', 'fp32-bf16x3' or 'bf16'")
    # "fp32-bf16x3": the fp32-equivalent mode with three bf16 terms per GEMM operand instead of two scaled
    # fp16 terms (6 / 3 tensor-pipe passes instead of 3 / 2): kept for comparison measurements
    gemm.MODE = {"fp32": "f16x2", "fp32-bf16x3": "bf16x3", "bf16": "bf16x1"}[mode]
    _PRECISION = "bf16" if mode == "bf16" else "fp32"


# Optional CUDA-event timers around named regions (bench.py's roofline leg).  Events are recorded
# on the current stream; nothing synchronises until timers_collect().
_TIMERS = {"on": False, "events": []}


def timers_enable(flag):
    _TIMERS["on"] = bool(flag)
    _TIMERS["events"] = []


class _region:
    def __init__(self, name):
        self.name = name
        self.e0 = self.e1 = None

    def __enter__(self):
        if _TIMERS["on"]:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _TIMERS["on"]:
            self.e1.record()
            _TIMERS["events"].append((self.name, self.e0, self.e1))
        return False


def timers_collect():
    """Total milliseconds per region name since timers_enable(True)."""
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in _TIMERS["events"]:
        out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
    _TIMERS["events"] = []
    return out


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sparch_b200 runs on CUDA tensors only (no CPU fallback); "
                               "move the module and its inputs to a B200 device")


def _on_device(fn):
    """Run a Function's forward / backward with the CUDA device of its first CUDA tensor argument current: the C ABI
    launches on the current device's stream and caches per-device attributes, so a model on cuda:1 must not be driven
    while cuda:0 is current (the launch would land on the wrong GPU with foreign pointers)."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        t = next((a for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kw)
        with torch.cuda.device(t.device):
            return fn(*args, **kw)
    return wrapper


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class SpikeFunctionBoxcar(torch.autograd.Function):
    """snns.py:20-36.  forward: x.gt(0).float(); backward: pass-through on -0.5 < x <= 0.5."""

    @staticmethod
    @_on_device
    def forward(ctx, x):
        _require_cuda(x)
        x = _f32c(x)
        ctx.save_for_backward(x)
        s = torch.empty_like(x)
        call("sparch_boxcar_fwd", ptr(x), ptr(s), x.numel(), _stream())
        return s

    @staticmethod
    @_on_device
    def backward(ctx, grad_spikes):
        (x,) = ctx.saved_tensors
        g = _f32c(grad_spikes)
        gx = torch.empty_like(x)
        call("sparch_boxcar_bwd", ptr(x), ptr(g), ptr(gx), x.numel(), _stream())
        return gx


# BatchNorm statistics ride in the projection GEMM's epilogue (gemm_tn_bf16_kernel<true>) when a tile's main loop
# has at least this many (pass, k-block) iterations to hide the butterfly column sums under (the epilogue of tile j
# overlaps the main loop of tile j + 1).  Measured at cfg 4: 4.16 vs 4.20 ms per train step against a separate
# sparch_col_stats pass (layer 1 and the readout qualify; the K = 40 input layer does not).  Before the epilogue lost
# its predicated-off bias loads the fused variant was the slower one (5.01 vs 4.95 ms).
FUSED_STATS_MIN_ITERS = int(os.environ.get("SPARCH_B200_FUSED_STATS_MIN_ITERS", "8"))


class LinearFunction(torch.autograd.Function):
    """Time-parallel projection ``x @ W^T + b`` over all Be*T frames (snns.py:675) and its
    autograd, on the tcgen05 GEMM.  ``in_scale``: None for a general fp32 input (three bf16
    terms); a float c when every input value is exactly 0 or c (spikes, c = 1/(1-p) after
    dropout) so that one exact bf16 term {0,1} suffices and c moves into the epilogue."""

    @staticmethod
    @_on_device
    def forward(ctx, x, weight, bias, in_scale, norm=None, x_terms=None, w_terms=None):
        """norm: the layer's NormState; in 'bn_train' mode the GEMM epilogue can also accumulate the
        BatchNorm column statistics of the output and leave them in ``norm.stats``.
        x_terms: the one-term {0,1} image of x (x = in_scale * x_terms) when the producing layer's
        post pass already wrote it (SpikePost): no split pass over x.
        w_terms: (gemm.Terms of ``weight`` in the current mode, event or None) when SNN.forward split the weights
        ahead on its side stream (parameter-only work)."""
        _require_cuda(x, weight)
        K = x.shape[-1]
        N = weight.shape[0]
        x2d = _f32c(x).reshape(-1, K)
        M = x2d.shape[0]
        with torch.no_grad():
            with _region("gemm_fwd"):
                if x_terms is not None and in_scale is not None:
                    xa, alpha = x_terms, float(in_scale)
                elif in_scale is None:
                    xa, alpha = gemm.split_general(x2d), 1.0
                else:
                    xa, alpha = gemm.split_binary(x2d, prescale=1.0 / in_scale), float(in_scale)
                if w_terms is not None and w_terms[2] == (gemm.MODE, weight.data_ptr(), weight._version):
                    wb = w_terms[0]
                    if w_terms[1] is not None:
                        torch.cuda.current_stream().wait_event(w_terms[1])
                else:
                    wb = gemm.split_general(_f32c(weight))
                # BatchNorm statistics ride in the GEMM epilogue when the tile's main loop is long enough
                # to dwarf it (measured: for the K=40 input layer the epilogue IS the kernel and a
                # separate column-statistics pass over the L2-warm output is cheaper).
                stats = None
                passes = len(gemm.pairs_for(xa.n, wb.n))
                if (norm is not None and norm.mode == "bn_train" and passes * ((K + 63) // 64) >= FUSED_STATS_MIN_ITERS):
                    stats = torch.empty(2, N, device=x2d.device, dtype=torch.float64)
                Z = gemm.gemm_parts(xa, wb, K, alpha=alpha, bias=None if bias is None else _f32c(bias),
                                    stats=stats)
                if norm is not None:
                    norm.stats = stats
        ctx.alpha = alpha
        ctx.norm = norm      # the cell's backward may leave dZ's operand terms there (NormState.dz_terms)
        if norm is not None:
            norm.need_dz32 = bias is not None
        ctx.has_bias = bias is not None
        # the terms serve the backward GEMMs as they are (MN-major operands): no re-split
        ctx.save_for_backward(xa.parts, xa.amax, wb.parts, wb.amax)
        ctx.xshape = x.shape
        ctx.dims = (M, N, K)
        return Z.view(*x.shape[:-1], N)

    @staticmethod
    @_on_device
    def backward(ctx, gZ):
        xa, wb = gemm.Terms(*ctx.saved_tensors[:2]), gemm.Terms(*ctx.saved_tensors[2:])
        M, N, K = ctx.dims
        g2d = _f32c(gZ).reshape(M, N)
        dx = dw = db = None
        with _region("gemm_bwd"):
            # the BatchNorm backward of this layer's cell may have written dZ as operand terms already
            ga = None
            if ctx.norm is not None:
                ga, ctx.norm.dz_terms = ctx.norm.dz_terms, None
            if ga is None:
                ga = gemm.split_general(g2d)
            if ctx.needs_input_grad[0]:
                # dX = dZ @ W: contraction over N; W's terms (N, K) are the MN-major B operand
                dx = gemm.gemm_parts(ga, wb, N, b_mn=True, N=K).view(ctx.xshape)
            if ctx.needs_input_grad[1]:
                # dW = dZ^T @ X: contraction over the Be*T frames, both operands MN-major.  A layer with few
                # outputs (readout: N = 35) computes dW^T = X^T @ dZ instead, so that the small dimension is the
                # narrow UMMA N (tile width 48) and not a 128-row tile that is 73 % padding.
                if N <= 64 and K >= 2 * N:
                    dw = gemm.gemm_parts(xa, ga, M, alpha=ctx.alpha, a_mn=True, b_mn=True, M=K, N=N).t()
                else:
                    dw = gemm.gemm_parts(ga, xa, M, alpha=ctx.alpha, a_mn=True, b_mn=True, M=N, N=K)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = g2d.sum(dim=0)
        return dx, dw, db, None, None, None, None


class NormState:
    """How the (Be*T, H) pre-activations are normalised before the recurrence.

    mode: 'bn_train' (batch statistics, running stats updated in place), 'bn_eval'
    (running statistics), or 'none'.  LayerNorm is applied by the caller.
    """

    def __init__(self, mode="none", running_mean=None, running_var=None, eps=1e-5, momentum=0.05):
        self.mode = mode
        self.running_mean = running_mean
        self.running_var = running_var
        self.eps = eps
        self.momentum = momentum
        self.stats = None   # (2, H) float64 column sum / sum of squares when the projection GEMM fused them
        self.dz_terms = None  # gemm.Terms of dZ left by the BatchNorm backward for the projection's gradient GEMMs
        self.need_dz32 = True  # False: the projection has no bias, nobody reads dZ as an fp32 tensor
        # per-call hand-overs between the cell Function and SpikePostFunction (which runs after it in the
        # forward and before it in the backward):
        self.sterm = None    # gemm.Terms: the {0,1} 16-bit image of S, the S_prev operand of dV
        self.gmax = None     # (Be*T,) row maxima of |dL/dS| for the tcgen05 reverse recurrence
        self.gmax_of = None  # (data_ptr, version) of the tensor those maxima were taken from
        # Packed-plane hand-over (first step): with ``lazy_spikes`` set (by the layer module, which calls
        # spike_post right after the cell) the tcgen05 forward recurrence does NOT write the fp32 spike tensor; the
        # post pass builds the layer's output and the operand terms from the published bit planes (0.25 B/elt).  The
        # tensor the cell Function returns is then uninitialised until spike_post has run (p = 0: it becomes the
        # output itself; p > 0: it is never read).
        self.lazy_spikes = False
        self.bits = None     # the planes of this forward (consumed by spike_post)
        self.s_last = None   # (Be, H) spikes of the last step, for the t = 0 frames of dV
        self.prep = None     # functional.CellPrep made ahead of this call by SNN.forward's side stream (or None)
        # Bidirectional layer without flipped / concatenated copies: B (rows per direction) when the projection ran once on
        # the un-flipped batch -- Z is (B, T, H), the recurrence runs 2B rows (the second half reads Z time-reversed), the
        # post pass writes the merged (B, T, 2H) output, the BatchNorm backward sums a row's two uses.  0 = off.
        self.bidir = 0


def _fold_norm(Z2d, gamma, bn_beta, norm):
    """Returns (scale, shift, mean, rstd); all None when there is no normalisation."""
    M, H = Z2d.shape
    dev = Z2d.device
    if norm.mode == "none":
        return None, None, None, None
    if norm.mode == "bn_eval":
        rstd = torch.rsqrt(norm.running_var + norm.eps)
        scale = rstd if gamma is None else gamma * rstd
        shift = -norm.running_mean * scale
        if bn_beta is not None:
            shift = shift + bn_beta
        return scale.contiguous(), shift.contiguous(), norm.running_mean.contiguous(), rstd
    if norm.mode != "bn_train":
        raise ValueError(f"unknown normalisation mode {norm.mode}")
    if M <= 1:
        raise ValueError("Expected more than 1 value per channel when training")
    sums = norm.stats
    if sums is None:        # projection not done by LinearFunction (direct use of the cell Function)
        sums = torch.empty(2, H, dtype=torch.float64, device=dev)
        call("sparch_col_stats", ptr(Z2d), M, H, ptr(sums[0]), ptr(sums[1]), _stream())
    if norm.bidir:
        # the reference normalises cat([W x, W flip(x)]) (snns.py:666-680): every row twice -- same mean and biased
        # variance, 2M samples in the unbiased variance of the running statistics
        sums, M = sums * 2, 2 * M
    out = torch.empty(4, H, dtype=torch.float32, device=dev)
    call("sparch_bn_fold_train", ptr(sums[0]), ptr(sums[1]), M, ptr(gamma), ptr(bn_beta),
         float(norm.eps), float(norm.momentum), ptr(norm.running_mean), ptr(norm.running_var),
         ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]), H, _stream())
    return out[2], out[3], out[0], out[1]


def _norm_backward_reduce(dI2d, Z2d, norm, mean, rstd):
    """BatchNorm backward, first half: the column reductions sum(dI), sum(dI * xhat) -- and, riding along,
    max|dI| for the fp16 split of dI in the dV GEMM.  Returns (sums, amax) or (None, None) without normalisation."""
    if norm.mode == "none":
        return None, None
    M, H = dI2d.shape
    sums = torch.empty(2, H, dtype=torch.float64, device=dI2d.device)
    amax = torch.empty(1, device=dI2d.device, dtype=torch.int32) if gemm.MODE == "f16x2" else None
    if norm.bidir:
        T = Z2d.shape[0] // norm.bidir
        call("sparch_col_dot_bidir", ptr(dI2d), ptr(Z2d), ptr(mean), ptr(rstd), M, H, T, norm.bidir, ptr(sums[0]),
             ptr(sums[1]), ptr(amax), _stream())
    else:
        call("sparch_col_dot", ptr(dI2d), ptr(Z2d), ptr(mean), ptr(rstd), M, H, ptr(sums[0]),
             ptr(sums[1]), ptr(amax), _stream())
    return sums, amax


def _norm_backward_apply(dI2d, Z2d, gamma, bn_beta, norm, scale, mean, rstd, sums, amax):
    """Second half: dI -> dZ; returns (dgamma, dbn_beta).  In the fp32-equivalent mode with batch statistics dZ
    is written directly as the scaled fp16 terms of the gradient GEMMs (``norm.dz_terms``); the fp32 tensor is
    only produced when the projection has a bias (its gradient sums dZ)."""
    if norm.mode == "none":
        return None, None
    M, H = dI2d.shape
    f16_terms = norm.mode == "bn_train" and gemm.MODE == "f16x2"     # that path leaves the two sums as fp32 itself
    dgamma = sums[1].float() if (gamma is not None and not f16_terms) else None
    dbeta = sums[0].float() if (bn_beta is not None and not f16_terms) else None
    if norm.bidir:
        # dZ (B*T rows) = the sum of a row's two uses; lands in the first half of dI (and / or as operand terms)
        M = Z2d.shape[0]
        T = M // norm.bidir
        if norm.mode == "bn_train" and gemm.MODE == "f16x2":
            ld = (H + 7) // 8 * 8
            parts = torch.empty(2, M, ld, device=dI2d.device, dtype=torch.float16)
            bound = torch.empty(1, device=dI2d.device, dtype=torch.int32)
            coef = torch.empty(4, H, device=dI2d.device, dtype=torch.float32)
            call("sparch_bn_bwd_apply_f16_bidir", ptr(dI2d), ptr(Z2d), ptr(mean), ptr(rstd), ptr(scale), ptr(sums[0]),
                 ptr(sums[1]), M, H, T, ptr(amax), ptr(bound), ptr(coef), ptr(parts[0]), ptr(parts[1]), ld,
                 ptr(dI2d) if norm.need_dz32 else None, _stream())
            norm.dz_terms = gemm.Terms(parts, bound)
            dgamma, dbeta = (coef[3] if gamma is not None else None), (coef[2] if bn_beta is not None else None)
        elif norm.mode == "bn_train":
            call("sparch_bn_bwd_apply_bidir", ptr(dI2d), ptr(Z2d), ptr(mean), ptr(rstd), ptr(scale),
                 ptr(sums[0]), ptr(sums[1]), M, H, T, _stream())
        else:   # running statistics: dZ = scale * (dI_f + flip(dI_b))
            d3 = dI2d.view(2, norm.bidir, T, H)
            d3[0].add_(d3[1].flip(1)).mul_(scale)
        return dgamma, dbeta
    if norm.mode == "bn_train" and gemm.MODE == "f16x2":
        ld = (H + 7) // 8 * 8
        parts = torch.empty(2, M, ld, device=dI2d.device, dtype=torch.float16)
        bound = torch.empty(1, device=dI2d.device, dtype=torch.int32)
        coef = torch.empty(4, H, device=dI2d.device, dtype=torch.float32)
        call("sparch_bn_bwd_apply_f16", ptr(dI2d), ptr(Z2d), ptr(mean), ptr(rstd), ptr(scale), ptr(sums[0]),
             ptr(sums[1]), M, H, ptr(amax), ptr(bound), ptr(coef), ptr(parts[0]), ptr(parts[1]), ld,
             ptr(dI2d) if norm.need_dz32 else None, _stream())
        norm.dz_terms = gemm.Terms(parts, bound)
        dgamma, dbeta = (coef[3] if gamma is not None else None), (coef[2] if bn_beta is not None else None)
    elif norm.mode == "bn_train":
        call("sparch_bn_bwd_apply", ptr(dI2d), ptr(Z2d), ptr(mean), ptr(rstd), ptr(scale),
             ptr(sums[0]), ptr(sums[1]), M, H, None, _stream())
    else:
        dI2d.mul_(scale)
    return dgamma, dbeta


def _norm_backward(dI2d, Z2d, gamma, bn_beta, norm, scale, mean, rstd):
    sums, amax = _norm_backward_reduce(dI2d, Z2d, norm, mean, rstd)
    return _norm_backward_apply(dI2d, Z2d, gamma, bn_beta, norm, scale, mean, rstd, sums, amax)


# Largest hidden size whose 32-column V0 slice (fp16 hi/lo, Hp*128 bytes) fits next to the exchange
# buffers in the persistent kernels' shared memory; beyond it the recurrent kinds take the stepwise path.
RECUR_MAX_H = 1184


# Reverse recurrence kernel: "tc" = tcgen05 kernel (clusters of 4 CTAs split K, TMA-fed fp16 hi/lo panels,
# per-row lagged scale, D in TMEM; H <= RECUR_TC_MAX_H), "mma" = mma.sync kernel (block-floating-point
# panels, per-chunk scales; also serves RECUR_TC_MAX_H < H <= RECUR_MAX_H).
RECUR_BWD = os.environ.get("SPARCH_B200_BWD", "tc")
RECUR_TC_MAX_H = 1024
# Forward recurrence kernel: "tc" = tcgen05 kernel (csrc/recur_fwd_tc.cu: int8 digit planes of V0, spike operand in
# tensor memory, 128 rows x 16 neurons per CTA; H <= 1536), "mma" = mma.sync kernel (csrc/recur.cu).
RECUR_FWD = os.environ.get("SPARCH_B200_FWD", "tc")
RECUR_FWD_TC_MAX_H = 1536
# Adaptation tape of the recurrent adaptive kind when both tcgen05 recurrence kernels serve the layer: C > 0
# keeps w only at the end of every chunk of C steps -- (Be, ceil(T/C), H) instead of (Be, T, H) fp32 -- and the reverse
# kernel recomputes the steps in between by solving snns.py:718 for w_{t-1}; 0 = the full tape.
W_TAPE_EVERY = int(os.environ.get("SPARCH_B200_W_EVERY", "16"))


# Callables run right after a reverse recurrence kernel has been issued (parallel.GradSync registers its ``flush``: the
# data-parallel all-reduces start behind the kernel instead of underneath it).
AFTER_BPTT = []


class CellPrep:
    """What a spiking layer's forward needs that depends on its PARAMETERS (and the initial states) only: the clamped
    neuron parameters (snns.py:706-709), the images of V0 the recurrence kernels read (snns.py:712), rec_0 = s0 @ V0.
    None of it depends on the layer's input, so ``SNN.forward`` computes it for all layers on a side stream at the
    start of the step, under the first projection and the first layer's recurrence (``event``: recorded on that stream
    when this layer's share is complete; None when the work was issued on the consumer's own stream)."""
    __slots__ = ("cl", "V0", "Vc", "img_f", "img_b", "meta", "rec0", "img_i8", "use_tc", "fwd_tc", "path", "states",
                 "event", "key")


def _param_versions(*ts):
    """(storage address, version counter) of every given parameter: a CellPrep made ahead is only used for the very
    tensors, in the very state, it was made from (an in-place change in between -- a hook, an optimizer step -- bumps
    the version and the layer prepares again in line)."""
    return tuple((t.data_ptr(), t._version) for t in ts if t is not None)


def prepare_cell(k, alpha, beta, a, b, V, Be, H, states=None):
    """The parameter-only launches of SpikingCellFunction.forward (same kernels, same order); ``states`` = (u0, w0, s0)
    when they are already drawn (rec_0 needs s0)."""
    adaptive, recurrent = bool(k & 1), bool(k & 2)
    dev = alpha.device
    st = _stream()
    pr = CellPrep()
    pr.event = None
    pr.states = states
    pr.key = (k, Be, H, _PRECISION, RECUR_BWD, RECUR_FWD) + _param_versions(alpha, beta, a, b, V)
    with torch.no_grad():
        alpha, beta, a, b = _f32c(alpha), _f32c(beta), _f32c(a), _f32c(b)
        cl = torch.empty(4 if adaptive else 1, H, device=dev, dtype=torch.float32)   # snns.py:706-709
        call("sparch_neuron_params", ptr(alpha), ptr(beta), ptr(a), ptr(b), _LIMS, cl.shape[0], H, ptr(cl), st)
        pr.cl = cl
        pr.V0 = pr.Vc = pr.img_f = pr.img_b = pr.meta = pr.rec0 = pr.img_i8 = None
        pr.use_tc = pr.fwd_tc = False
        if not recurrent:
            pr.path = "cell"
        elif H > RECUR_MAX_H:                                                   # snns.py:712 (stepwise path only)
            pr.path = "stepwise"
            pr.V0 = torch.empty(H, H, device=dev, dtype=torch.float32)
            call("sparch_recur_v0", ptr(V.detach().contiguous()), H, ptr(pr.V0), st)
        else:
            # persistent tensor-core kernels: s_{t-1} @ V0 from packed spike planes (csrc/recur.cu, recur_fwd_tc.cu)
            pr.path = "persist"
            Hp = _lib.lib().sparch_recur_padded(H)
            use_tc = RECUR_BWD == "tc" and H <= RECUR_TC_MAX_H
            fwd_tc = RECUR_FWD == "tc" and H <= RECUR_FWD_TC_MAX_H
            Vc = V.detach().contiguous()
            img_f = None if fwd_tc else torch.empty(Hp * Hp, device=dev, dtype=torch.int32)
            img_b = None if use_tc else torch.empty(Hp * Hp, device=dev, dtype=torch.int32)
            meta = torch.empty(2, device=dev, dtype=torch.int32)
            # what the forward kernel waits for comes first (the reverse kernel's image is needed a pass later)
            if fwd_tc:
                img_i8 = torch.empty(_lib.lib().sparch_recur_fwd_tc_image_bytes(H), device=dev, dtype=torch.uint8)
                call("sparch_recur_prepare_fwd_tc", ptr(Vc), H, ptr(img_i8), st)
                pr.img_i8 = img_i8
            if states is not None:
                pr.rec0 = _rec0(states[2], Vc, Be, H)
            if img_f is not None or img_b is not None or use_tc:   # (meta: max|V0| for the fp16 images)
                call("sparch_recur_prepare", ptr(Vc), H, ptr(img_f), ptr(img_b), ptr(meta), st)
            if use_tc:  # V0 as swizzled UMMA tiles for the tcgen05 reverse kernel (csrc/recur_tc.cu)
                img_b = torch.empty(_lib.lib().sparch_recur_bwd_tc_image_bytes(H), device=dev, dtype=torch.uint8)
                call("sparch_recur_prepare_tc", ptr(Vc), H, ptr(img_b), ptr(meta), st)
            pr.Vc, pr.img_f, pr.img_b, pr.meta, pr.use_tc, pr.fwd_tc = Vc, img_f, img_b, meta, use_tc, fwd_tc
    return pr


def _rec0(s0, Vc, Be, H):
    """t = 0: s_{-1} is real-valued (snns.py:702): rec_0 = s0 @ V0, V0's zero diagonal applied on the fly."""
    rec0 = torch.empty(Be, H, device=s0.device, dtype=torch.float32)
    with _region("gemm_fwd"):
        call("sparch_small_gemm", ptr(s0), H, 0, ptr(Vc), H, ptr(rec0), H, Be, H, H, 1, _stream())
    return rec0


class SpikingCellFunction(torch.autograd.Function):
    """Normalisation fold + membrane recurrence of one spiking layer.

    forward(Z, gamma, bn_beta, alpha, beta, a, b, V, u0, w0, s0, kind, theta, norm) -> S
      Z (Be,T,H) pre-activations W x (snns.py:675); gamma/bn_beta BatchNorm affine (or None);
      alpha..b raw (unclamped) neuron parameters; V raw recurrent weight (H,H) or None;
      u0,w0,s0 (Be,H) initial states (snns.py:700-702).  S (Be,T,H) fp32 spikes.
    """

    @staticmethod
    @_on_device
    def forward(ctx, Z, gamma, bn_beta, alpha, beta, a, b, V, u0, w0, s0, kind, theta, norm):
        _require_cuda(Z, alpha, u0, s0)
        k = KINDS[kind]
        adaptive, recurrent = bool(k & 1), bool(k & 2)
        Z = _f32c(Z)
        Be, T, H = Z.shape
        if norm.bidir:         # Z is the projection of the un-flipped batch: the recurrence runs both directions
            if norm.bidir != Be:
                raise ValueError("NormState.bidir must equal the batch size of Z")
            Be = 2 * Be
        dev = Z.device
        st = _stream()
        alpha, beta, a, b = _f32c(alpha), _f32c(beta), _f32c(a), _f32c(b)
        with torch.no_grad():
            u0, w0, s0 = _f32c(u0), _f32c(w0) if adaptive else None, _f32c(s0)
            # parameter-only launches: taken from the side stream of SNN.forward when it prepared them for exactly this
            # call (functional.CellPrep), issued here otherwise
            pr = norm.prep
            norm.prep = None
            if pr is not None and (pr.key != (k, Be, H, _PRECISION, RECUR_BWD, RECUR_FWD) +
                                   _param_versions(alpha, beta, a, b, V) or
                                   (pr.states is not None and pr.states[2].data_ptr() != s0.data_ptr())):
                pr = None
            if pr is None:
                pr = prepare_cell(k, alpha, beta, a, b, V, Be, H)
            elif pr.event is not None:
                torch.cuda.current_stream().wait_event(pr.event)
            cl = pr.cl
            al, be, aa, bb = cl[0], (cl[1] if adaptive else None), (cl[2] if adaptive else None), \
                (cl[3] if adaptive else None)
            V0 = pr.V0
            Z2d = Z.view(-1, H)
            scale, shift, mean, rstd = _fold_norm(Z2d, gamma, bn_beta, norm)
            S = torch.empty(Be, T, H, device=dev, dtype=torch.float32)
            U = torch.empty_like(S)
            w_every = W_TAPE_EVERY if (adaptive and recurrent and pr.path == "persist" and pr.fwd_tc and pr.use_tc) else 0
            ctx.w_every = w_every
            Wt = (torch.empty(Be, (T + w_every - 1) // w_every, H, device=dev, dtype=torch.float32) if w_every
                  else torch.empty_like(S)) if adaptive else None
            if norm.bidir and not (recurrent and pr.path == "persist" and pr.fwd_tc and norm.lazy_spikes):
                raise RuntimeError("the copy-free bidirectional path needs the tcgen05 forward recurrence with the "
                                   "packed-plane hand-over (the layer module checks this before choosing it)")
            # (the timed regions wrap the recurrence launches only, event records right next to the call, so that an
            # eager pass measures kernel time and not the host's launch gaps)
            if not recurrent:
                with _region("recurrence_fwd"):
                    call("sparch_cell_fwd", k, ptr(Z), ptr(scale), ptr(shift), ptr(al), ptr(be), ptr(aa),
                         ptr(bb), ptr(u0), ptr(w0), ptr(s0), float(theta), ptr(S), ptr(U), ptr(Wt), Be, T,
                         H, st)
            elif H > RECUR_MAX_H:
                # V0 slice too large for shared memory: general stepwise path, two launches per timestep
                # (s_{t-1} @ V0 in fp32 FFMA, then sparch_cell_step_fwd)
                rec = torch.empty(Be, H, device=dev, dtype=torch.float32)
                region = _region("recurrence_fwd").__enter__()
                for t in range(T):
                    sp_ = s0 if t == 0 else S[:, t - 1, :]                      # snns.py:720
                    call("sparch_small_gemm", ptr(sp_), sp_.stride(0), 0, ptr(V0), H, ptr(rec), H, Be, H, H, 0, st)
                    call("sparch_cell_step_fwd", k, t, ptr(Z), ptr(scale), ptr(shift), ptr(al),
                         ptr(be), ptr(aa), ptr(bb), ptr(rec), ptr(u0), ptr(w0), ptr(s0), float(theta),
                         ptr(S), ptr(U), ptr(Wt), Be, T, H, st)
                region.__exit__()
                ctx.rec = None
            else:
                # persistent tensor-core kernels: s_{t-1} @ V0 from packed spike planes (csrc/recur.cu); images of V0 and
                # rec_0 from prepare_cell
                Hp = _lib.lib().sparch_recur_padded(H)
                use_tc, fwd_tc, img_f, img_b, meta, img_i8 = pr.use_tc, pr.fwd_tc, pr.img_f, pr.img_b, pr.meta, pr.img_i8
                ctx.tc = use_tc
                ctx.rec = (img_b, meta)
                ctx.reduced = int(_PRECISION == "bf16")
                rec0 = pr.rec0 if pr.rec0 is not None else _rec0(s0, pr.Vc, Be, H)
                if fwd_tc:
                    L = _lib.lib()
                    bits = torch.empty(L.sparch_recur_fwd_tc_bits_bytes(Be, T, H) // 4, device=dev, dtype=torch.int32)
                    lazy = bool(norm.lazy_spikes)
                    with _region("recurrence_fwd"):
                        call("sparch_recur_fwd_tc_bidir", k, ptr(Z), ptr(scale), ptr(shift), ptr(al), ptr(be), ptr(aa),
                             ptr(bb), ptr(rec0), ptr(img_i8), ptr(u0), ptr(w0), ptr(s0), float(theta),
                             None if lazy else ptr(S), ptr(U), ptr(Wt), ptr(bits), int(_PRECISION == "bf16"), Be, T, H,
                             int(norm.bidir), w_every, st)
                    if lazy:
                        norm.bits = bits
                else:
                    bits = torch.empty(T, Be, Hp // 32, 2, device=dev, dtype=torch.int32)
                    with _region("recurrence_fwd"):
                        call("sparch_recur_fwd", k, ptr(Z), ptr(scale), ptr(shift), ptr(al), ptr(be), ptr(aa),
                             ptr(bb), ptr(rec0), ptr(img_f), ptr(meta), ptr(u0), ptr(w0), ptr(s0),
                             float(theta), ptr(S), ptr(U), ptr(Wt), ptr(bits), int(_PRECISION == "bf16"), Be, T, H,
                             st)
        if norm.bits is None:
            norm.lazy_spikes = False      # another forward kernel ran: S is an ordinary tensor
        ctx.k, ctx.theta, ctx.norm = k, float(theta), norm
        ctx.has = (gamma is not None, bn_beta is not None)
        ctx.save_for_backward(Z, gamma, bn_beta, alpha, beta, a, b, V0, u0, w0, s0, S, U, Wt, al, be,
                              aa, bb, scale, mean, rstd)
        return S

    @staticmethod
    @_on_device
    def backward(ctx, gS):
        (Z, gamma, bn_beta, alpha, beta, a, b, V0, u0, w0, s0, S, U, Wt, al, be, aa, bb, scale, mean,
         rstd) = ctx.saved_tensors
        k, theta, norm = ctx.k, ctx.theta, ctx.norm
        adaptive, recurrent = bool(k & 1), bool(k & 2)
        Be, T, H = U.shape
        dev = Z.device
        st = _stream()
        G = _f32c(gS)
        dI = torch.empty_like(U)
        npart = 4 if adaptive else 1
        # (the streaming and the tcgen05 kernels write every entry; the stepwise path accumulates, the mma.sync kernel is
        # left with a cleared buffer as before)
        overwrite = (not recurrent) or (ctx.rec is not None and ctx.tc)
        part = (torch.empty if overwrite else torch.zeros)(npart, Be, H, device=dev, dtype=torch.float32)
        pp = [ptr(part[i]) if i < npart else None for i in range(4)]
        ws = sync = None
        if recurrent and ctx.rec is not None:
            if ctx.tc:
                ws = torch.empty(_lib.lib().sparch_recur_bwd_tc_workspace(Be, T, H), device=dev, dtype=torch.uint8)
            else:
                ws = torch.empty(_lib.lib().sparch_recur_bwd_workspace(Be, H), device=dev, dtype=torch.uint8)
                sync = torch.empty(_lib.lib().sparch_recur_sync_words(Be), device=dev, dtype=torch.int32)
        region = _region("recurrence_bwd").__enter__()
        if not recurrent:
            call("sparch_cell_bwd", k, ptr(G), ptr(U), ptr(Wt), ptr(al), ptr(be), ptr(aa), ptr(bb),
                 ptr(u0), ptr(w0), ptr(s0), theta, ptr(dI), pp[0], pp[1], pp[2], pp[3], Be, T, H, st)
            region.__exit__()
            dV = None
        elif ctx.rec is None:
            carry = torch.zeros(2, Be, H, device=dev, dtype=torch.float32)
            recb = torch.empty(Be, H, device=dev, dtype=torch.float32)
            for t in range(T - 1, -1, -1):
                if t < T - 1:
                    call("sparch_small_gemm", ptr(dI[:, t + 1, :]), T * H, 0, ptr(V0), H, ptr(recb), H, Be, H, H, 8, st)
                call("sparch_cell_step_bwd", k, t, ptr(G), ptr(U), ptr(Wt), ptr(al), ptr(be), ptr(aa),
                     ptr(bb), ptr(recb) if t < T - 1 else None, ptr(u0), ptr(w0), ptr(s0), theta,
                     ptr(dI), ptr(carry[0]), ptr(carry[1]) if adaptive else None, pp[0], pp[1], pp[2],
                     pp[3], Be, T, H, st)
        elif ctx.tc:
            img_b, meta = ctx.rec
            # the row maxima left by the dropout backward describe exactly the tensor that pass produced: a hook or a
            # gradient scaling between the two nodes hands over another tensor (or a modified one) -- then the kernel
            # computes the maxima itself (they fix the fp16 scale of the hand-over: a stale bound could overflow it)
            gmax = norm.gmax
            if gmax is not None and getattr(norm, "gmax_of", None) != (G.data_ptr(), G._version):
                gmax = None
            call("sparch_recur_bwd_tc_ck", k, ptr(G), ptr(U), ptr(Wt), ptr(al), ptr(be), ptr(aa), ptr(bb),
                 ptr(img_b), ptr(meta), ptr(u0), ptr(w0), ptr(s0), theta, ptr(dI), pp[0], pp[1], pp[2],
                 pp[3], ptr(ws), ctx.reduced, Be, T, H, ptr(gmax), ctx.w_every, st)
        else:
            img_b, meta = ctx.rec
            call("sparch_recur_bwd", k, ptr(G), ptr(U), ptr(Wt), ptr(al), ptr(be), ptr(aa), ptr(bb),
                 ptr(img_b), ptr(meta), ptr(u0), ptr(w0), ptr(s0), theta, ptr(dI), pp[0], pp[1], pp[2],
                 pp[3], ptr(ws), ptr(sync), ctx.reduced, Be, T, H, st)
        if recurrent:
            region.__exit__()
            for cb in list(AFTER_BPTT):
                cb()
        # BatchNorm backward reductions first: the same pass leaves max|dI| for the dV operand split
        sums, di_amax = _norm_backward_reduce(dI.view(Be * T, H), Z.view(-1, H), norm, mean, rstd)
        if recurrent:
            # dV = sum_t s_{t-1}^T dI_t, diagonal masked (clone().fill_diagonal_(0) backward)
            with _region("gemm_bwd"):
                # frame m of dI pairs with frame m-1 of S (a_koff = -1).  Inside a batch row that is
                # s_{t-1}; across rows it pairs dI[b, 0] with S[b-1, T-1], which is replaced below by
                # the real-valued initial state s0 (snns.py:702).
                first = torch.empty(Be, H, device=dev, dtype=torch.float32)
                if norm.s_last is not None:   # packed-plane forward: S was never written, its last step was kept
                    call("sparch_dv_boundary", ptr(s0), ptr(norm.s_last), Be, 1, H, ptr(first), st)
                else:
                    call("sparch_dv_boundary", ptr(s0), ptr(S), Be, T, H, ptr(first), st)
                dV = torch.empty(H, H, device=dev, dtype=torch.float32)
                if Be * T > 1:
                    if norm.sterm is None and norm.s_last is not None:
                        raise RuntimeError("packed-plane forward without the post pass's operand image: call "
                                           "spike_post(S, p, norm, recurrent=True) after the cell Function")
                    sp = norm.sterm if norm.sterm is not None else gemm.split_binary(S.view(Be * T, H))
                    dit = gemm.split_general(dI.view(Be * T, H), amax=di_amax)
                    gemm.gemm_parts(sp, dit, Be * T, a_mn=True, b_mn=True, a_koff=-1, M=H, N=H, out=dV)
                # + first^T dI[:, 0, :], then the zero diagonal (backward of snns.py:712), in the same epilogue
                call("sparch_small_gemm", ptr(first), H, 1, ptr(dI), T * H, ptr(dV), H, H, H, Be,
                     2 | (4 if Be * T > 1 else 0), st)
        norm.gmax = None     # consumed: belongs to this backward pass only
        pg = torch.empty(npart, H, device=dev, dtype=torch.float32)
        call("sparch_param_grads", ptr(part), ptr(alpha), ptr(beta), ptr(a), ptr(b), _LIMS, npart, Be, H, ptr(pg), st)
        dalpha = pg[0]
        dbeta, da, db = (pg[1], pg[2], pg[3]) if adaptive else (None, None, None)
        dgamma, dbnb = _norm_backward_apply(dI.view(Be * T, H), Z.view(-1, H), gamma, bn_beta, norm,
                                            scale, mean, rstd, sums, di_amax)
        if norm.bidir:
            if norm.mode == "none":      # no normalisation: the gradient of W x is the sum of its two uses
                dI[:norm.bidir].add_(dI[norm.bidir:].flip(1))
            dI = dI[:norm.bidir]         # (B, T, H): the BatchNorm backward left dZ in the first half
        return (dI, dgamma, dbnb, dalpha, dbeta, da, db, dV, None, None, None, None, None, None)


class SpikePost:
    """What the post pass of a spiking layer leaves for its consumers: the next projection's operand
    (``terms``, x = scale * terms) and the per-neuron counts of non-zero outputs."""
    __slots__ = ("terms", "counts", "scale", "rows")

    def __init__(self, terms, counts, scale, rows):
        self.terms, self.counts, self.scale, self.rows = terms, counts, scale, rows

    def rates(self, out):
        """Per-neuron mean of the layer output over (batch, time) (snns.py:174), differentiable in ``out``."""
        return FiringRateFunction.apply(out, self.counts, self.scale / self.rows)


class FiringRateFunction(torch.autograd.Function):
    """mean over (B, T) of a spike tensor whose non-zero values all equal ``scale * rows``, from integer counts."""

    @staticmethod
    @_on_device
    def forward(ctx, out, counts, factor):
        ctx.shape = out.shape
        return counts * factor        # int32 * Python float -> float32: one launch

    @staticmethod
    @_on_device
    def backward(ctx, g):
        B, T, H = ctx.shape
        return (g / (B * T)).expand(B, T, H), None, None


class _DropoutPostFunction(torch.autograd.Function):
    """Dropout of the spike tensor (snns.py:692) fused with the counts / operand terms (csrc/post.cu).  The
    backward regenerates the mask from the seed and also leaves the row maxima of the result for the tcgen05
    reverse recurrence in ``cell_state.gmax``."""

    @staticmethod
    @_on_device
    def forward(ctx, S, p, seed, out_bufs, cell_state, want_gmax):
        Be, T, H = S.shape
        out, term, sterm, counts = out_bufs
        if cell_state.bits is not None:
            _post_from_bits(cell_state, Be, T, H, float(p), seed, out, term, sterm, counts)
        else:
            call("sparch_spike_post_fwd", ptr(S), Be * T, H, float(p), ptr(seed), ptr(out), ptr(term), ptr(sterm),
                 int(term.dtype == torch.float16), ptr(counts), _stream())
        ctx.p, ctx.seed, ctx.cell_state, ctx.want_gmax = float(p), seed, cell_state, want_gmax
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        g = _f32c(g)
        Be, T, H = g.shape
        if ctx.cell_state.bidir:      # g is the gradient of the merged (B, T, 2H) output: un-merge into the recurrence's rows
            B, Be, H = Be, 2 * Be, H // 2
            gS = torch.empty(Be, T, H, device=g.device, dtype=torch.float32)
            gmax = torch.empty(Be * T, device=g.device, dtype=torch.float32) if ctx.want_gmax else None
            call("sparch_spike_post_bwd_bidir", ptr(g), B, T, H, ctx.p, ptr(ctx.seed), ptr(gS), ptr(gmax), _stream())
            ctx.cell_state.gmax = gmax
            ctx.cell_state.gmax_of = (gS.data_ptr(), gS._version) if gmax is not None else None
            return gS, None, None, None, None, None
        gS = torch.empty_like(g)
        gmax = torch.empty(Be * T, device=g.device, dtype=torch.float32) if ctx.want_gmax else None
        call("sparch_spike_post_bwd", ptr(g), Be * T, H, ctx.p, ptr(ctx.seed), ptr(gS), ptr(gmax), _stream())
        ctx.cell_state.gmax = gmax
        ctx.cell_state.gmax_of = (gS.data_ptr(), gS._version) if gmax is not None else None
        return gS, None, None, None, None, None


def _post_from_bits(cell_state, Be, T, H, p, seed, out, term, sterm, counts):
    """The post pass fed by the packed spike planes of the tcgen05 forward (csrc/post.cu: sparch_spike_post_fwd_bits)."""
    s_last = torch.empty(Be, H, device=out.device, dtype=torch.float32)
    call("sparch_spike_post_fwd_bits_bidir", ptr(cell_state.bits), Be, T, H, p, ptr(seed), ptr(out), ptr(term), ptr(sterm),
         int(term.dtype == torch.float16), ptr(counts), ptr(s_last), int(cell_state.bidir), _stream())
    cell_state.bits = None
    cell_state.s_last = s_last


@_on_device
def spike_post(S, p, cell_state, recurrent):
    """Post pass of a spiking layer: returns (out, SpikePost).  p: dropout probability in effect (0 in eval
    mode).  cell_state: the NormState the layer's cell Function was called with."""
    _require_cuda(S)
    S = _f32c(S)
    Be, T, H = S.shape
    M, ld, dev = Be * T, (H + 7) // 8 * 8, S.device
    dt = torch.float16 if gemm.MODE == "f16x2" else torch.bfloat16
    term = torch.empty(1, M, ld, device=dev, dtype=dt)
    counts = torch.empty(H, device=dev, dtype=torch.int32)
    need_grad = torch.is_grad_enabled() and S.requires_grad
    if cell_state.bidir:
        # merged (B, T, 2H) output, operand terms and counts straight from the planes (snns.py:686-692 in one pass); the
        # same node un-merges the gradient, with or without dropout
        if cell_state.bits is None:
            raise RuntimeError("the copy-free bidirectional path needs the packed planes of the tcgen05 forward")
        B = cell_state.bidir
        term = torch.empty(1, B * T, 2 * ld, device=dev, dtype=dt)
        counts = torch.empty(2 * H, device=dev, dtype=torch.int32)
        sterm = torch.empty(1, M, ld, device=dev, dtype=dt) if (recurrent and need_grad) else None
        seed = torch.empty(1, device=dev, dtype=torch.int64).random_() if p > 0.0 else None
        out = torch.empty(B, T, 2 * H, device=dev, dtype=torch.float32)
        want_gmax = recurrent and RECUR_BWD == "tc" and H <= RECUR_TC_MAX_H
        out = _DropoutPostFunction.apply(S, p, seed, (out, term, sterm, counts), cell_state, want_gmax)
        if sterm is not None:
            cell_state.sterm = gemm.Terms(sterm)
        return out, SpikePost(gemm.Terms(term), counts, 1.0 / (1.0 - p), B * T)
    if p > 0.0:
        sterm = torch.empty(1, M, ld, device=dev, dtype=dt) if (recurrent and need_grad) else None
        seed = torch.empty(1, device=dev, dtype=torch.int64).random_()     # device generator, as nn.Dropout
        out = torch.empty_like(S)
        want_gmax = recurrent and RECUR_BWD == "tc" and H <= RECUR_TC_MAX_H
        out = _DropoutPostFunction.apply(S, p, seed, (out, term, sterm, counts), cell_state, want_gmax)
        if sterm is not None:
            cell_state.sterm = gemm.Terms(sterm)
        scale = 1.0 / (1.0 - p)
    else:
        with torch.no_grad():
            if cell_state.bits is not None:     # S is still uninitialised: this pass writes it (it IS the output)
                _post_from_bits(cell_state, Be, T, H, 0.0, None, S, term, None, counts)
            else:
                call("sparch_spike_post_fwd", ptr(S), M, H, 0.0, None, None, ptr(term), None,
                     int(dt == torch.float16), ptr(counts), _stream())
        if recurrent and need_grad:
            cell_state.sterm = gemm.Terms(term)
        out, scale = S, 1.0
    return out, SpikePost(gemm.Terms(term), counts, scale, M)


class ReadoutCellFunction(torch.autograd.Function):
    """Normalisation fold + ReadoutLayer cell (snns.py:807-825): out = sum_t softmax(u_t)."""

    @staticmethod
    @_on_device
    def forward(ctx, Z, gamma, bn_beta, alpha, u0, norm):
        _require_cuda(Z, alpha, u0)
        Z = _f32c(Z)
        B, T, C = Z.shape
        st = _stream()
        alpha = _f32c(alpha)
        with torch.no_grad():
            al = torch.empty(C, device=Z.device, dtype=torch.float32)
            call("sparch_neuron_params", ptr(alpha), None, None, None, _LIMS, 1, C, ptr(al), st)
            u0 = _f32c(u0)
            scale, shift, mean, rstd = _fold_norm(Z.view(B * T, C), gamma, bn_beta, norm)
            out = torch.empty(B, C, device=Z.device, dtype=torch.float32)
            U = torch.empty_like(Z)
            call("sparch_readout_fwd", ptr(Z), ptr(scale), ptr(shift), ptr(al), ptr(u0), ptr(out),
                 ptr(U), B, T, C, st)
        ctx.norm = norm
        ctx.save_for_backward(Z, gamma, bn_beta, alpha, u0, U, al, scale, mean, rstd)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, gout):
        Z, gamma, bn_beta, alpha, u0, U, al, scale, mean, rstd = ctx.saved_tensors
        B, T, C = Z.shape
        st = _stream()
        g = _f32c(gout)
        dI = torch.empty_like(Z)
        part = torch.empty(B, C, device=Z.device, dtype=torch.float32)
        call("sparch_readout_bwd", ptr(g), ptr(U), ptr(al), ptr(u0), ptr(dI), ptr(part), B, T, C, st)
        dalpha = torch.empty(C, device=Z.device, dtype=torch.float32)
        call("sparch_param_grads", ptr(part), ptr(alpha), None, None, None, _LIMS, 1, B, C, ptr(dalpha), st)
        dgamma, dbnb = _norm_backward(dI.view(B * T, C), Z.view(B * T, C), gamma, bn_beta, ctx.norm,
                                      scale, mean, rstd)
        return dI, dgamma, dbnb, dalpha, None, None


class LayerNormFunction(torch.autograd.Function):
    """``nn.LayerNorm(H)`` on the pre-activations W x (normalization="layernorm", snns.py:98-99, 678-680): a row kernel
    each way (csrc/norm.cu) instead of ATen's; parameter gradients from fp64 partial sums in a fixed order."""

    @staticmethod
    @_on_device
    def forward(ctx, x, weight, bias, eps):
        _require_cuda(x)
        x = _f32c(x)
        H = x.shape[-1]
        M = x.numel() // H
        y = torch.empty_like(x)
        stats = torch.empty(2, M, device=x.device, dtype=torch.float32)
        w, b = _f32c(weight), _f32c(bias)
        call("sparch_layernorm_fwd", ptr(x), ptr(w), ptr(b), float(eps), M, H, ptr(y), ptr(stats[0]), ptr(stats[1]),
             _stream())
        ctx.save_for_backward(x, w, stats)
        ctx.has = (weight is not None, bias is not None)
        return y

    @staticmethod
    @_on_device
    def backward(ctx, gy):
        x, w, stats = ctx.saved_tensors
        H = x.shape[-1]
        M = x.numel() // H
        g = _f32c(gy)
        dx = torch.empty_like(x)
        dw = torch.empty(H, device=x.device, dtype=torch.float32) if ctx.has[0] else None
        db = torch.empty(H, device=x.device, dtype=torch.float32) if ctx.has[1] else None
        ws = None
        if dw is not None or db is not None:
            ws = torch.empty(_lib.lib().sparch_layernorm_bwd_workspace(M, H), device=x.device, dtype=torch.uint8)
        call("sparch_layernorm_bwd", ptr(g), ptr(x), ptr(w), ptr(stats[0]), ptr(stats[1]), M, H, ptr(dx), ptr(dw), ptr(db),
             ptr(ws), _stream())
        return dx, dw, db, None


class CrossEntropyFunction(torch.autograd.Function):
    """Mean cross-entropy of (B, C) logits against int64 class targets: what ``nn.CrossEntropyLoss()`` (exp.py:83) computes
    at exp.py:362, as one launch forward and one backward (ATen: log_softmax + nll_loss, twice)."""

    @staticmethod
    @_on_device
    def forward(ctx, logits, target):
        _require_cuda(logits, target)
        if logits.dim() != 2 or target.dim() != 1 or target.dtype != torch.int64 or target.shape[0] != logits.shape[0]:
            raise ValueError("CrossEntropyFunction takes (B, C) logits and (B,) int64 targets")
        x = _f32c(logits.detach())
        y = target.contiguous()
        B, C = x.shape
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        lse = torch.empty(B, device=x.device, dtype=torch.float32)
        call("sparch_ce_fwd", ptr(x), ptr(y), B, C, ptr(loss), ptr(lse), _stream())
        ctx.save_for_backward(x, y, lse)
        return loss

    @staticmethod
    @_on_device
    def backward(ctx, gloss):
        x, y, lse = ctx.saved_tensors
        B, C = x.shape
        g = _f32c(gloss).reshape(1)
        dx = torch.empty_like(x)
        call("sparch_ce_bwd", ptr(x), ptr(y), ptr(lse), ptr(g), B, C, ptr(dx), _stream())
        return dx, None


class CrossEntropyLoss(torch.nn.Module):
    """Drop-in for the ``nn.CrossEntropyLoss()`` of exp.py:83 (default arguments: mean reduction, class-index targets)."""

    def forward(self, logits, target):
        return CrossEntropyFunction.apply(logits, target)


def native_launches():
    """Number of C-ABI calls made so far by this process."""
    return _lib.launches
