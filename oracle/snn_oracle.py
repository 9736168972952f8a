"""CPU oracle for the sparch SNN hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product path
(``sparch_b200``) never imports it and has no CPU fallback.

What it restates (all citations relative to ``/root/reference``):

* ``boxcar_forward`` / ``boxcar_backward``      sparch/models/snns.py:26-36
* ``cell_forward``  (LIF/adLIF/RLIF/RadLIF)      snns.py:282-303, 419-445, 554-578, 696-727
* ``cell_backward`` (explicit BPTT, no autograd) autograd of the above; equations of SURVEY.md 8a
* ``batchnorm_train`` / ``batchnorm_backward``   ``nn.BatchNorm1d(H, momentum=0.05)`` snns.py:678-680
* ``readout_forward`` / ``readout_backward``     snns.py:807-825
* ``OracleSNN``  (torch-CPU autograd module)     snns.py:39-176 + layer classes

Third-party arithmetic: everything in the reference is PyTorch ATen
(requirements.txt:13 pins torch==1.12.0, not vendored).  The explicit
functions below are written in numpy so they can run in float64 (the reference
itself cannot: ``.float()`` is hard-coded at snns.py:29), giving a high
precision truth for the CUDA kernels; ``OracleSNN`` uses the same ATen ops in
the same order as the reference so it reproduces its fp32 rounding.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the
build container by ``oracle/make_golden.py`` (which imports the untouched
``/root/reference``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function here against them.
"""
from __future__ import annotations

import math

import numpy as np

KINDS = ("LIF", "adLIF", "RLIF", "RadLIF")

# parameter limits, snns.py:229 / 356-359
ALPHA_LIM = (math.exp(-1 / 5), math.exp(-1 / 25))
BETA_LIM = (math.exp(-1 / 30), math.exp(-1 / 120))
A_LIM = (-1.0, 1.0)
B_LIM = (0.0, 2.0)


def kind_flags(kind: str):
    """(adaptive, recurrent) for a neuron type name (snns.py:109)."""
    if kind not in KINDS:
        raise ValueError(f"Invalid neuron type {kind}")
    return kind in ("adLIF", "RadLIF"), kind in ("RLIF", "RadLIF")


# --------------------------------------------------------------------------- #
# SpikeFunctionBoxcar                                             snns.py:20-36
# --------------------------------------------------------------------------- #
def boxcar_forward(x):
    """spike iff x > 0 (strict)                                  snns.py:29"""
    return (x > 0).astype(x.dtype)


def boxcar_mask(x):
    """surrogate window: gradient passes iff -0.5 < x <= 0.5     snns.py:33-35"""
    return np.logical_and(x > -0.5, x <= 0.5)


def boxcar_backward(x, grad_spikes):
    return np.where(boxcar_mask(x), grad_spikes, 0).astype(grad_spikes.dtype)


def clamp_params(kind, alpha, beta=None, a=None, b=None, dtype=np.float32):
    """torch.clamp of the neuron parameters (snns.py:290, 428-431, 562, 706-709).
    The limits are python floats, cast to the tensor dtype by ATen."""
    adaptive, _ = kind_flags(kind)
    dt = np.dtype(dtype).type
    out = {"alpha": np.clip(alpha.astype(dtype), dt(ALPHA_LIM[0]), dt(ALPHA_LIM[1]))}
    if adaptive:
        out["beta"] = np.clip(beta.astype(dtype), dt(BETA_LIM[0]), dt(BETA_LIM[1]))
        out["a"] = np.clip(a.astype(dtype), dt(A_LIM[0]), dt(A_LIM[1]))
        out["b"] = np.clip(b.astype(dtype), dt(B_LIM[0]), dt(B_LIM[1]))
    return out


def clamp_grad_mask(p, lim, dtype=np.float32):
    """clamp backward: gradient flows on the closed interval [lo, hi]."""
    dt = np.dtype(dtype).type
    return np.logical_and(p >= dt(lim[0]), p <= dt(lim[1]))


# --------------------------------------------------------------------------- #
# the four cells, forward                      snns.py:282-303/419-445/554-578/696-727
# --------------------------------------------------------------------------- #
def cell_forward(kind, I, alpha, beta=None, a=None, b=None, V0=None,
                 u0=None, w0=None, s0=None, theta=1.0, dtype=np.float32):
    """Run the membrane recurrence over time.

    I      (Be, T, H)  input current after normalisation (``Wx`` in the reference)
    alpha.. (H,)       already clamped
    V0     (H, H)      recurrent matrix with zero diagonal (snns.py:566, 712);
                       the product is ``s @ V0`` (no transpose, snns.py:572, 720)
    u0,w0,s0 (Be, H)   initial states (the reference draws them with torch.rand)

    Returns dict with s, u (Be,T,H) and w (adaptive kinds) -- u/w are the
    post-update states of every step.  Operation order follows the reference
    expression so that float32 rounding is reproduced op by op.
    """
    adaptive, recurrent = kind_flags(kind)
    f = np.dtype(dtype).type
    I = I.astype(dtype)
    Be, T, H = I.shape
    alpha = alpha.astype(dtype)
    one_m_alpha = (f(1) - alpha).astype(dtype)
    u = u0.astype(dtype).copy()
    s = s0.astype(dtype).copy()
    w = w0.astype(dtype).copy() if adaptive else None
    th = f(theta)
    S = np.empty((Be, T, H), dtype)
    U = np.empty((Be, T, H), dtype)
    W = np.empty((Be, T, H), dtype) if adaptive else None
    for t in range(T):
        x = I[:, t, :]
        if adaptive:
            w = (beta * w + a * u + b * s).astype(dtype)          # snns.py:438, 718
        if recurrent:
            x = (x + (s @ V0).astype(dtype)).astype(dtype)        # snns.py:572, 720
        if adaptive:
            x = (x - w).astype(dtype)
        u = (alpha * (u - s) + one_m_alpha * x).astype(dtype)    # snns.py:297, 439
        s = ((u - th) > 0).astype(dtype)                          # snns.py:300 + :29
        S[:, t], U[:, t] = s, u
        if adaptive:
            W[:, t] = w
    return {"s": S, "u": U, "w": W}


def cell_step(kind, I_t, alpha, beta, a, b, V0, u_prev, w_prev, s_prev, theta=1.0,
              dtype=np.float32):
    """One teacher-forced step: all (b, t) rows given their true previous state."""
    r = cell_forward(kind, I_t[:, None, :], alpha, beta, a, b, V0, u_prev, w_prev,
                     s_prev, theta, dtype)
    return {k: (v[:, 0] if v is not None else None) for k, v in r.items()}


# --------------------------------------------------------------------------- #
# explicit BPTT                                           SURVEY.md 8a equations
# --------------------------------------------------------------------------- #
def cell_backward(kind, g, I, alpha, beta=None, a=None, b=None, V0=None,
                  u0=None, w0=None, s0=None, theta=1.0, U=None, W=None, S=None,
                  dtype=np.float64):
    """Reverse-time pass given the forward tapes U, W, S (from cell_forward).

    g (Be,T,H) = dL/ds_t.  Returns dI, dalpha, dbeta, da, db (w.r.t. the CLAMPED
    values; apply clamp_grad_mask afterwards), dV0 (diagonal zeroed).
    Because the spike/surrogate masks are taken from the tapes, this map is
    linear in g -- the 'given-mask backward' of SURVEY.md 7 #1 (ii).
    """
    adaptive, recurrent = kind_flags(kind)
    f = np.dtype(dtype).type
    g = g.astype(dtype)
    I = I.astype(dtype)
    U = U.astype(dtype)
    S = S.astype(dtype)
    Be, T, H = I.shape
    alpha = alpha.astype(dtype)
    oma = f(1) - alpha
    if adaptive:
        beta, a, b, W = (z.astype(dtype) for z in (beta, a, b, W))
    if recurrent:
        V0 = V0.astype(dtype)
    th = np.float32(theta)
    du_n = np.zeros((Be, H), dtype)
    dw_n = np.zeros((Be, H), dtype)
    dI = np.zeros((Be, T, H), dtype)
    dalpha = np.zeros(H, dtype)
    dbeta = np.zeros(H, dtype)
    da = np.zeros(H, dtype)
    db = np.zeros(H, dtype)
    dV = np.zeros((H, H), dtype) if recurrent else None
    for t in range(T - 1, -1, -1):
        u_prev = U[:, t - 1] if t > 0 else u0.astype(dtype)
        s_prev = S[:, t - 1] if t > 0 else s0.astype(dtype)
        if adaptive:
            w_prev = W[:, t - 1] if t > 0 else w0.astype(dtype)
        ds = g[:, t] - alpha * du_n
        if adaptive:
            ds = ds + b * dw_n
        if recurrent:
            ds = ds + (oma * du_n) @ V0.T
        # surrogate window is evaluated on the fp32 value the forward thresholded
        sg = boxcar_mask(U[:, t].astype(np.float32) - th)
        du = ds * sg + alpha * du_n
        if adaptive:
            du = du + a * dw_n
            dw = -oma * du + beta * dw_n
        dI[:, t] = oma * du
        x = I[:, t]
        if recurrent:
            x = x + s_prev @ V0
            dV += s_prev.T @ dI[:, t]
        if adaptive:
            x = x - W[:, t]
            dbeta += (dw * w_prev).sum(0)
            da += (dw * u_prev).sum(0)
            db += (dw * s_prev).sum(0)
        dalpha += (du * ((u_prev - s_prev) - x)).sum(0)
        du_n = du
        if adaptive:
            dw_n = dw
    if recurrent:
        np.fill_diagonal(dV, 0)                                  # clone().fill_diagonal_(0) backward
    return {"dI": dI, "dalpha": dalpha, "dbeta": dbeta if adaptive else None,
            "da": da if adaptive else None, "db": db if adaptive else None, "dV": dV}


# --------------------------------------------------------------------------- #
# BatchNorm1d(H, momentum=0.05, eps=1e-5) over the Be*T rows        snns.py:678-680
# --------------------------------------------------------------------------- #
def batchnorm_train(x2d, gamma, beta, running_mean=None, running_var=None,
                    momentum=0.05, eps=1e-5, dtype=np.float64):
    x = x2d.astype(dtype)
    M = x.shape[0]
    mean = x.mean(0)
    var = x.var(0)                                               # biased, used to normalise
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean) * rstd
    y = xhat * gamma.astype(dtype) + beta.astype(dtype)
    out = {"y": y, "mean": mean, "var": var, "rstd": rstd, "xhat": xhat}
    if running_mean is not None:
        unbiased = var * (M / max(M - 1, 1))
        out["running_mean"] = (1 - momentum) * running_mean + momentum * mean
        out["running_var"] = (1 - momentum) * running_var + momentum * unbiased
    return out


def batchnorm_backward(dy2d, xhat, gamma, rstd, dtype=np.float64):
    dy = dy2d.astype(dtype)
    M = dy.shape[0]
    dgamma = (dy * xhat).sum(0)
    dbeta = dy.sum(0)
    dx = (gamma.astype(dtype) * rstd) * (dy - dbeta / M - xhat * (dgamma / M))
    return {"dx": dx, "dgamma": dgamma, "dbeta": dbeta}


# --------------------------------------------------------------------------- #
# ReadoutLayer                                                    snns.py:807-825
# --------------------------------------------------------------------------- #
def _softmax(z):
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def readout_forward(I, alpha, u0, dtype=np.float64):
    """u_t = alpha u_{t-1} + (1-alpha) I_t ; out = sum_t softmax(u_t)   snns.py:822-823"""
    f = np.dtype(dtype).type
    I = I.astype(dtype)
    Be, T, H = I.shape
    alpha = alpha.astype(dtype)
    u = u0.astype(dtype).copy()
    out = np.zeros((Be, H), dtype)
    U = np.empty((Be, T, H), dtype)
    for t in range(T):
        u = alpha * u + (f(1) - alpha) * I[:, t]
        out = out + _softmax(u)
        U[:, t] = u
    return {"out": out, "u": U}


def readout_backward(gout, I, alpha, u0, U, dtype=np.float64):
    f = np.dtype(dtype).type
    gout = gout.astype(dtype)
    I = I.astype(dtype)
    U = U.astype(dtype)
    alpha = alpha.astype(dtype)
    Be, T, H = I.shape
    du_n = np.zeros((Be, H), dtype)
    dI = np.zeros_like(I)
    dalpha = np.zeros(H, dtype)
    for t in range(T - 1, -1, -1):
        p = _softmax(U[:, t])
        dz = p * (gout - (p * gout).sum(1, keepdims=True))
        du = dz + alpha * du_n
        dI[:, t] = (f(1) - alpha) * du
        u_prev = U[:, t - 1] if t > 0 else u0.astype(dtype)
        dalpha += (du * (u_prev - I[:, t])).sum(0)
        du_n = du
    return {"dI": dI, "dalpha": dalpha}


# --------------------------------------------------------------------------- #
# torch-CPU module with the reference's state_dict layout
# --------------------------------------------------------------------------- #
def _torch():
    import torch
    return torch


def build_oracle_snn(*args, **kwargs):
    """Factory so that importing this file does not import torch."""
    return _make_oracle_classes()["OracleSNN"](*args, **kwargs)


_CLASSES = None


def _make_oracle_classes():
    global _CLASSES
    if _CLASSES is not None:
        return _CLASSES
    torch = _torch()
    nn = torch.nn

    class _Boxcar(torch.autograd.Function):
        """snns.py:20-36"""

        @staticmethod
        def forward(ctx, x):
            ctx.save_for_backward(x)
            return (x > 0).to(torch.float32)

        @staticmethod
        def backward(ctx, gs):
            (x,) = ctx.saved_tensors
            keep = (x > -0.5) & (x <= 0.5)
            return gs * keep.to(gs.dtype)

    def _make_norm(normalization, H):
        if normalization == "batchnorm":
            return nn.BatchNorm1d(H, momentum=0.05)
        if normalization == "layernorm":
            return nn.LayerNorm(H)
        return None

    class OracleSpikingLayer(nn.Module):
        """One class for the four reference layer classes; `kind` selects the
        cell (snns.py:179-727).  Parameter registration order matches the
        reference so state_dict keys line up: alpha[,beta,a,b], W, [V], norm."""

        def __init__(self, kind, input_size, hidden_size, threshold=1.0, dropout=0.0,
                     normalization="batchnorm", use_bias=False, bidirectional=False):
            super().__init__()
            self.kind = kind
            self.adaptive, self.recurrent = kind_flags(kind)
            self.threshold = threshold
            self.bidirectional = bidirectional
            H = int(hidden_size)
            # same RNG draw order as snns.py:638-649 (W, V, alpha, beta, a, b, orthogonal V)
            self.W = nn.Linear(int(input_size), H, bias=use_bias)
            if self.recurrent:
                self.V = nn.Linear(H, H, bias=False)
            self.alpha = nn.Parameter(torch.empty(H).uniform_(*ALPHA_LIM))
            if self.adaptive:
                self.beta = nn.Parameter(torch.empty(H).uniform_(*BETA_LIM))
                self.a = nn.Parameter(torch.empty(H).uniform_(*A_LIM))
                self.b = nn.Parameter(torch.empty(H).uniform_(*B_LIM))
            if self.recurrent:
                nn.init.orthogonal_(self.V.weight)
            self.norm = _make_norm(normalization, H)
            self.drop = nn.Dropout(p=dropout)
            self.capture = None  # set to a dict to record I, u0, w0, s0, s

        def forward(self, x):
            if self.bidirectional:                                # snns.py:666-668
                x = torch.cat([x, x.flip(1)], dim=0)
            Wx = self.W(x)                                       # snns.py:675
            if self.norm is not None:                            # snns.py:678-680
                Be, T, H = Wx.shape
                Wx = self.norm(Wx.reshape(Be * T, H)).reshape(Be, T, H)
            Be, T, H = Wx.shape
            dev = Wx.device
            # CPU-generator draws in the reference's order: ut[, wt], st (snns.py:700-702)
            u = torch.rand(Be, H).to(dev)
            w = torch.rand(Be, H).to(dev) if self.adaptive else None
            s = torch.rand(Be, H).to(dev)
            if self.capture is not None:
                self.capture.update(I=Wx.detach().clone(), u0=u.clone(),
                                    w0=None if w is None else w.clone(), s0=s.clone())
            alpha = torch.clamp(self.alpha, min=ALPHA_LIM[0], max=ALPHA_LIM[1])
            if self.adaptive:
                beta = torch.clamp(self.beta, min=BETA_LIM[0], max=BETA_LIM[1])
                a = torch.clamp(self.a, min=A_LIM[0], max=A_LIM[1])
                b = torch.clamp(self.b, min=B_LIM[0], max=B_LIM[1])
            if self.recurrent:
                V = self.V.weight.clone().fill_diagonal_(0)      # snns.py:712
            spikes, us = [], []
            for t in range(T):
                x_t = Wx[:, t, :]
                if self.adaptive:
                    w = beta * w + a * u + b * s                 # snns.py:718
                if self.recurrent and self.adaptive:
                    u = alpha * (u - s) + (1 - alpha) * (x_t + torch.matmul(s, V) - w)
                elif self.recurrent:
                    u = alpha * (u - s) + (1 - alpha) * (x_t + torch.matmul(s, V))
                elif self.adaptive:
                    u = alpha * (u - s) + (1 - alpha) * (x_t - w)
                else:
                    u = alpha * (u - s) + (1 - alpha) * x_t
                s = _Boxcar.apply(u - self.threshold)            # snns.py:724
                spikes.append(s)
                if self.capture is not None:
                    us.append(u.detach())
            out = torch.stack(spikes, dim=1)
            if self.capture is not None:
                self.capture.update(s=out.detach().clone(), u=torch.stack(us, dim=1))
            if self.bidirectional:                               # snns.py:686-689
                fwd, bwd = out.chunk(2, dim=0)
                out = torch.cat([fwd, bwd.flip(1)], dim=2)
            return self.drop(out)                                # snns.py:692

    class OracleReadout(nn.Module):
        """snns.py:730-825"""

        def __init__(self, input_size, hidden_size, dropout=0.0, normalization="batchnorm",
                     use_bias=False):
            super().__init__()
            H = int(hidden_size)
            self.W = nn.Linear(int(input_size), H, bias=use_bias)
            self.alpha = nn.Parameter(torch.empty(H).uniform_(*ALPHA_LIM))
            self.norm = _make_norm(normalization, H)
            self.drop = nn.Dropout(p=dropout)                    # constructed, never applied
            self.capture = None

        def forward(self, x):
            Wx = self.W(x)
            if self.norm is not None:
                B, T, H = Wx.shape
                Wx = self.norm(Wx.reshape(B * T, H)).reshape(B, T, H)
            B, T, H = Wx.shape
            u = torch.rand(B, H).to(Wx.device)                   # snns.py:812
            if self.capture is not None:
                self.capture.update(I=Wx.detach().clone(), u0=u.clone())
            out = torch.zeros(B, H, device=Wx.device)
            alpha = torch.clamp(self.alpha, min=ALPHA_LIM[0], max=ALPHA_LIM[1])
            for t in range(T):
                u = alpha * u + (1 - alpha) * Wx[:, t, :]        # snns.py:822
                out = out + torch.softmax(u, dim=1)              # snns.py:823
            return out

    class OracleSNN(nn.Module):
        """snns.py:39-176: same constructor, same (out, firing_rates) return."""

        def __init__(self, input_shape, layer_sizes, neuron_type="LIF", threshold=1.0,
                     dropout=0.0, normalization="batchnorm", use_bias=False,
                     bidirectional=False, use_readout_layer=True):
            super().__init__()
            kind_flags(neuron_type)
            self.reshape = len(input_shape) > 3
            self.is_snn = True
            self.use_readout_layer = use_readout_layer
            self.num_layers = len(layer_sizes)
            fin = int(np.prod(input_shape[2:]))
            n_hidden = self.num_layers - 1 if use_readout_layer else self.num_layers
            layers = []
            for i in range(n_hidden):
                layers.append(OracleSpikingLayer(neuron_type, fin, layer_sizes[i], threshold,
                                                 dropout, normalization, use_bias, bidirectional))
                fin = layer_sizes[i] * (2 if bidirectional else 1)
            if use_readout_layer:
                layers.append(OracleReadout(fin, layer_sizes[-1], dropout, normalization, use_bias))
            self.snn = nn.ModuleList(layers)

        def forward(self, x):
            if self.reshape:
                if x.ndim != 4:
                    raise NotImplementedError
                x = x.reshape(x.shape[0], x.shape[1], x.shape[2] * x.shape[3])
            hidden = []
            for i, lay in enumerate(self.snn):
                x = lay(x)
                if not (self.use_readout_layer and i == self.num_layers - 1):
                    hidden.append(x)
            rates = torch.cat(hidden, dim=2).mean(dim=(0, 1))    # snns.py:174
            return x, rates

    _CLASSES = {"OracleSNN": OracleSNN, "OracleSpikingLayer": OracleSpikingLayer,
                "OracleReadout": OracleReadout, "Boxcar": _Boxcar}
    return _CLASSES


def oracle_train_step(model, opt, x, y, seed=42):
    """One reference-style train step (exp.py:355-377): forward, CE, backward, Adam."""
    torch = _torch()
    torch.manual_seed(seed)
    out, rates = model(x)
    loss = torch.nn.functional.cross_entropy(out, y)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return float(loss), out.detach(), rates.detach()
