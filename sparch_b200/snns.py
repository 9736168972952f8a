"""B200-native stand-in for ``sparch.models.snns`` (reference: sparch/models/snns.py).

Same public classes, constructor arguments, attribute names, parameter registration order
(=> identical ``state_dict`` keys and whole-module pickles) and RNG draw order as the
reference, so ``sparch/exp.py:305-314`` can construct ``SNN`` unchanged.  The per-timestep
Python loops of the reference (snns.py:282-303, 419-445, 554-578, 696-727, 807-825) are replaced
by ``torch.autograd.Function`` wrappers over hand-written sm_100a kernels
(``sparch_b200.functional`` -> ``libsparch_b200.so``).  CUDA only: a forward pass on CPU
tensors raises ``RuntimeError``.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from .functional import (LayerNormFunction, LinearFunction, NormState, ReadoutCellFunction, SpikeFunctionBoxcar,
                         SpikingCellFunction, spike_post)
from . import functional as _F
from . import rng as _rng

# SPARCH_B200_LAZY_SPIKES=0: the forward recurrence writes the fp32 spike tensor even where the post pass could read the
# packed planes (comparison / debugging).
_LAZY_SPIKES = os.environ.get("SPARCH_B200_LAZY_SPIKES", "1") != "0"

# Where the per-forward initial states u_{-1}, w_{-1}, s_{-1} ~ U[0,1) are drawn.
#   "cpu"    (default): torch.rand on the CPU default generator, then copied to the device --
#            exactly the reference's draws (snns.py:700-702), so a given torch.manual_seed gives the
#            reference's states.  The serial CPU generator costs ~2 ms per (256, 1024) draw.
#   "device": torch.rand on the CUDA generator in the same order: same distribution, different
#            stream, no host work.
_STATE_INIT = os.environ.get("SPARCH_B200_STATE_INIT", "cpu")


def set_state_init(mode):
    """Select "cpu" (reference-identical draws) or "device" (fast) initial-state generation."""
    global _STATE_INIT
    if mode not in ("cpu", "device"):
        raise ValueError("state init mode must be 'cpu' or 'device'")
    _STATE_INIT = mode


# SPARCH_B200_PREP_AHEAD=0: every layer issues its parameter-only launches (clamps, images of V0, rec_0) itself, on the
# stream of its forward pass, instead of SNN.forward issuing them for all layers on a side stream at the start.
_PREP_AHEAD = os.environ.get("SPARCH_B200_PREP_AHEAD", "1") != "0"
# SPARCH_B200_BIDIR_FUSED=0: bidirectional layers take the reference's flip / cat formulation (copies) everywhere.
_BIDIR_FUSED = os.environ.get("SPARCH_B200_BIDIR_FUSED", "1") != "0"
_SIDE_STREAMS = {}     # device index -> the side stream of SNN.forward's parameter-only work
_PENDING_PREP = {}     # id(layer) -> functional.CellPrep made ahead of the layer's forward (cleared by SNN.forward)


def _side_stream(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = torch.cuda.Stream(device=idx)
    return _SIDE_STREAMS[idx]


# SPARCH_B200_DEVICE_MT=0: the "cpu" mode draws on the host (torch.rand) instead of replaying the CPU generator's MT19937
# stream on the device (sparch_b200/rng.py: same numbers, generator left in the same state).
_DEVICE_MT = os.environ.get("SPARCH_B200_DEVICE_MT", "1") != "0"


def _rand_state(rows, cols, device):
    if _STATE_INIT == "device":
        return torch.rand(rows, cols, device=device)
    if _DEVICE_MT:
        r = _rng.cpu_generator_rand(rows * cols, device)
        if r is not None:
            return r.view(rows, cols)
    return torch.rand(rows, cols).to(device)


__all__ = ["SpikeFunctionBoxcar", "SNN", "LIFLayer", "adLIFLayer", "RLIFLayer", "RadLIFLayer",
           "ReadoutLayer"]


def _norm_args(layer):
    """(Z-transform, gamma, beta, NormState) for the layer's normalisation (snns.py:652-658)."""
    if not layer.normalize:
        return None, None, NormState("none")
    norm = layer.norm
    if isinstance(norm, nn.BatchNorm1d):
        use_batch = norm.training or norm.running_mean is None
        if use_batch:
            if norm.training and norm.num_batches_tracked is not None:
                norm.num_batches_tracked.add_(1)
            track = norm.training and norm.running_mean is not None
            st = NormState("bn_train", norm.running_mean if track else None,
                           norm.running_var if track else None, norm.eps, norm.momentum)
        else:
            st = NormState("bn_eval", norm.running_mean, norm.running_var, norm.eps, norm.momentum)
        return norm.weight, norm.bias, st
    return None, None, None  # LayerNorm: applied by torch before the cell


class _SpikingLayerBase(nn.Module):
    """Shared constructor / forward of the four spiking layers (snns.py:179-727).

    Parameter creation and initialisation follow the reference line by line so that a given
    ``torch.manual_seed`` produces the same weights: W (Linear), [V (Linear)], alpha
    [, beta, a, b] ``uniform_`` in that order, then ``orthogonal_`` on V.
    """

    _kind = None
    _adaptive = False
    _recurrent = False

    def __init__(self, input_size, hidden_size, batch_size, threshold=1.0, dropout=0.0,
                 normalization="batchnorm", use_bias=False, bidirectional=False):
        super().__init__()
        self.input_size = int(input_size)
        self.hidden_size = int(hidden_size)
        self.batch_size = batch_size
        self.threshold = threshold
        self.dropout = dropout
        self.normalization = normalization
        self.use_bias = use_bias
        self.bidirectional = bidirectional
        self.batch_size = self.batch_size * (1 + self.bidirectional)
        self.alpha_lim = [np.exp(-1 / 5), np.exp(-1 / 25)]
        if self._adaptive:
            self.beta_lim = [np.exp(-1 / 30), np.exp(-1 / 120)]
            self.a_lim = [-1.0, 1.0]
            self.b_lim = [0.0, 2.0]
        self.spike_fct = SpikeFunctionBoxcar.apply

        self.W = nn.Linear(self.input_size, self.hidden_size, bias=use_bias)
        if self._recurrent:
            self.V = nn.Linear(self.hidden_size, self.hidden_size, bias=False)
        self.alpha = nn.Parameter(torch.Tensor(self.hidden_size))
        if self._adaptive:
            self.beta = nn.Parameter(torch.Tensor(self.hidden_size))
            self.a = nn.Parameter(torch.Tensor(self.hidden_size))
            self.b = nn.Parameter(torch.Tensor(self.hidden_size))
        nn.init.uniform_(self.alpha, self.alpha_lim[0], self.alpha_lim[1])
        if self._adaptive:
            nn.init.uniform_(self.beta, self.beta_lim[0], self.beta_lim[1])
            nn.init.uniform_(self.a, self.a_lim[0], self.a_lim[1])
            nn.init.uniform_(self.b, self.b_lim[0], self.b_lim[1])
        if self._recurrent:
            nn.init.orthogonal_(self.V.weight)

        self.normalize = False
        if normalization == "batchnorm":
            self.norm = nn.BatchNorm1d(self.hidden_size, momentum=0.05)
            self.normalize = True
        elif normalization == "layernorm":
            self.norm = nn.LayerNorm(self.hidden_size)
            self.normalize = True
        self.drop = nn.Dropout(p=dropout)

    def forward(self, x, in_scale=None, x_terms=None, post_out=None):
        """``in_scale``: None for a general input; c when every input value is exactly 0 or c
        (the previous spiking layer's output), which lets the projection use one exact 16-bit term.
        ``x_terms`` / ``post_out`` (used by SNN.forward): the previous layer's ``SpikePost.terms``, and a list
        that receives this layer's ``SpikePost`` (or None)."""
        out, post = self._forward_post(x, in_scale, x_terms)
        if post_out is not None:
            post_out.append(post)
        return out

    def _forward_post(self, x, in_scale=None, x_terms=None):
        """forward() plus what the layer's post pass leaves for the next layer and for the firing rates
        (``functional.SpikePost`` or None); ``x_terms``: the previous layer's ``SpikePost.terms``."""
        if not x.is_cuda:
            raise RuntimeError("sparch_b200 layers run on CUDA only (no CPU fallback)")
        if self.bidirectional and self._copy_free_bidir(x):
            return self._forward_post_bidir(x, in_scale, x_terms)
        if self.bidirectional:                                   # snns.py:666-668
            if x_terms is not None:   # the same batch-wise flip/cat on the 16-bit operand image instead of a re-split
                n, _, ld = x_terms.parts.shape
                p3 = x_terms.parts.view(n, x.shape[0], x.shape[1], ld)
                x_terms = type(x_terms)(torch.cat([p3, p3.flip(2)], dim=1).view(n, -1, ld), x_terms.amax)
            x = torch.cat([x, x.flip(1)], dim=0)
        if self.batch_size != x.shape[0]:                        # snns.py:671-672
            self.batch_size = x.shape[0]
        gamma, bn_beta, norm = _norm_args(self)
        Wx = LinearFunction.apply(x, self.W.weight, self.W.bias, in_scale, norm, x_terms,
                                  _PENDING_PREP.pop((id(self), "W"), None))                  # snns.py:675
        if norm is None:                                         # layernorm, snns.py:678-680
            if isinstance(self.norm, nn.LayerNorm) and Wx.shape[-1] <= 2048:
                Wx = LayerNormFunction.apply(Wx, self.norm.weight, self.norm.bias, self.norm.eps)
            else:
                Wx = self.norm(Wx)
            norm = NormState("none")
        p = self.drop.p if self.drop.training else 0.0
        # the post pass below runs on the cell's own output: the forward recurrence may hand the spikes over as packed
        # planes instead of an fp32 tensor (functional.NormState.lazy_spikes)
        norm.lazy_spikes = self._recurrent and not self.bidirectional and p < 1.0 and _LAZY_SPIKES
        s = self._cell(Wx, gamma, bn_beta, norm)
        if self.bidirectional:                                   # snns.py:686-689
            s_f, s_b = s.chunk(2, dim=0)
            s = torch.cat([s_f, s_b.flip(1)], dim=2)
        if p >= 1.0:
            return self.drop(s), None                            # snns.py:692
        # dropout (snns.py:692) + spike counts + the next projection's operand in one pass over s; the merged
        # bidirectional tensor is not the cell's own output, so it cannot serve the cell's dV product
        return spike_post(s, p, norm, self._recurrent and not self.bidirectional)

    def _copy_free_bidir(self, x):
        """Whether this bidirectional layer can run without the flipped / concatenated copies of snns.py:666-668 and
        686-689: the projection once on the un-flipped batch, the second half of the recurrence reading it
        time-reversed, the merge written by the post pass (functional.NormState.bidir).  Needs the tcgen05 forward
        recurrence with its packed planes, whole 128-row groups per direction and 16-byte rows."""
        H, p = self.hidden_size, (self.drop.p if self.drop.training else 0.0)
        return (_BIDIR_FUSED and _LAZY_SPIKES and self._recurrent and x.ndim == 3 and x.shape[0] % 128 == 0
                and H % 8 == 0 and H <= _F.RECUR_MAX_H and _F.RECUR_FWD == "tc" and H <= _F.RECUR_FWD_TC_MAX_H
                and p < 1.0 and not (self.normalize and isinstance(self.norm, nn.LayerNorm)))

    def _forward_post_bidir(self, x, in_scale, x_terms):
        B = x.shape[0]
        if self.batch_size != 2 * B:                             # snns.py:671-672 (the reference sees the doubled batch)
            self.batch_size = 2 * B
        gamma, bn_beta, norm = _norm_args(self)
        Wx = LinearFunction.apply(x, self.W.weight, self.W.bias, in_scale, norm, x_terms,
                                  _PENDING_PREP.pop((id(self), "W"), None))                  # snns.py:675, once
        norm.bidir = B
        norm.lazy_spikes = True
        p = self.drop.p if self.drop.training else 0.0
        s = self._cell(Wx, gamma, bn_beta, norm)                 # (2B, T, H) in the recurrence's order, still planes only
        return spike_post(s, p, norm, True)                      # snns.py:686-692: merge + dropout, (B, T, 2H)

    def _cell(self, Wx, gamma, bn_beta, norm):
        device = Wx.device
        Be, H = Wx.shape[0] * (2 if norm.bidir else 1), Wx.shape[2]
        # initial states in the reference's order ut, [wt,] st (snns.py:700-702): from the CPU generator (identical
        # draws), or in "device" mode as one draw of the 2-3 states on the CUDA generator
        prep = _PENDING_PREP.pop(id(self), None)
        if prep is not None and prep.key[1:3] != (Be, H):
            prep = None
        norm.prep = prep
        if prep is not None and prep.states is not None:
            ut, wt, st = prep.states      # drawn on the side stream, in the same order
        elif _STATE_INIT == "device":
            r = torch.rand(3 if self._adaptive else 2, Be, H, device=device)
            ut, wt, st = r[0], (r[1] if self._adaptive else None), r[-1]
        else:
            ut = _rand_state(Be, H, device)
            wt = _rand_state(Be, H, device) if self._adaptive else None
            st = _rand_state(Be, H, device)
        return SpikingCellFunction.apply(
            Wx, gamma, bn_beta, self.alpha, getattr(self, "beta", None), getattr(self, "a", None),
            getattr(self, "b", None), self.V.weight if self._recurrent else None, ut, wt, st,
            self._kind, self.threshold, norm)


class LIFLayer(_SpikingLayerBase):
    """Leaky integrate-and-fire layer (snns.py:179-303)."""
    _kind = "LIF"

    def _lif_cell(self, Wx):
        return self._cell(Wx, None, None, NormState("none"))


class adLIFLayer(_SpikingLayerBase):
    """Adaptive LIF layer (snns.py:306-445)."""
    _kind = "adLIF"
    _adaptive = True

    def _adlif_cell(self, Wx):
        return self._cell(Wx, None, None, NormState("none"))


class RLIFLayer(_SpikingLayerBase):
    """LIF layer with layer-wise recurrent connections (snns.py:448-578)."""
    _kind = "RLIF"
    _recurrent = True

    def _rlif_cell(self, Wx):
        return self._cell(Wx, None, None, NormState("none"))


class RadLIFLayer(_SpikingLayerBase):
    """Adaptive LIF layer with recurrent connections (snns.py:581-727)."""
    _kind = "RadLIF"
    _adaptive = True
    _recurrent = True

    def _radlif_cell(self, Wx):
        return self._cell(Wx, None, None, NormState("none"))


class ReadoutLayer(nn.Module):
    """Non-spiking LIF readout: out = sum_t softmax(u_t) (snns.py:730-825)."""

    def __init__(self, input_size, hidden_size, batch_size, dropout=0.0, normalization="batchnorm",
                 use_bias=False):
        super().__init__()
        self.input_size = int(input_size)
        self.hidden_size = int(hidden_size)
        self.batch_size = batch_size
        self.dropout = dropout
        self.normalization = normalization
        self.use_bias = use_bias
        self.alpha_lim = [np.exp(-1 / 5), np.exp(-1 / 25)]

        self.W = nn.Linear(self.input_size, self.hidden_size, bias=use_bias)
        self.alpha = nn.Parameter(torch.Tensor(self.hidden_size))
        nn.init.uniform_(self.alpha, self.alpha_lim[0], self.alpha_lim[1])

        self.normalize = False
        if normalization == "batchnorm":
            self.norm = nn.BatchNorm1d(self.hidden_size, momentum=0.05)
            self.normalize = True
        elif normalization == "layernorm":
            self.norm = nn.LayerNorm(self.hidden_size)
            self.normalize = True
        self.drop = nn.Dropout(p=dropout)  # constructed but never applied, as in the reference

    def forward(self, x, in_scale=None, x_terms=None):
        if not x.is_cuda:
            raise RuntimeError("sparch_b200 layers run on CUDA only (no CPU fallback)")
        gamma, bn_beta, norm = _norm_args(self)
        Wx = LinearFunction.apply(x, self.W.weight, self.W.bias, in_scale, norm, x_terms,
                                  _PENDING_PREP.pop((id(self), "W"), None))                  # snns.py:796
        if norm is None:
            if isinstance(self.norm, nn.LayerNorm) and Wx.shape[-1] <= 2048:
                Wx = LayerNormFunction.apply(Wx, self.norm.weight, self.norm.bias, self.norm.eps)
            else:
                Wx = self.norm(Wx)
            norm = NormState("none")
        return self._readout_cell(Wx, gamma, bn_beta, norm)

    def _readout_cell(self, Wx, gamma=None, bn_beta=None, norm=None):
        ut = _PENDING_PREP.pop(id(self), None)                   # drawn ahead by SNN.forward, in the reference's order
        if ut is None or ut.shape != (Wx.shape[0], Wx.shape[2]):
            ut = _rand_state(Wx.shape[0], Wx.shape[2], Wx.device)    # snns.py:812
        return ReadoutCellFunction.apply(Wx, gamma, bn_beta, self.alpha, ut,
                                         norm if norm is not None else NormState("none"))


_LAYER_CLASSES = {"LIF": LIFLayer, "adLIF": adLIFLayer, "RLIF": RLIFLayer, "RadLIF": RadLIFLayer}


class SNN(nn.Module):
    """Multi-layered spiking network (snns.py:39-176): same arguments, same
    ``forward(x) -> (outputs, firing_rates)`` contract."""

    def __init__(self, input_shape, layer_sizes, neuron_type="LIF", threshold=1.0, dropout=0.0,
                 normalization="batchnorm", use_bias=False, bidirectional=False,
                 use_readout_layer=True):
        super().__init__()
        self.reshape = True if len(input_shape) > 3 else False
        self.input_size = float(torch.prod(torch.tensor(input_shape[2:])))
        self.batch_size = input_shape[0]
        self.layer_sizes = layer_sizes
        self.num_layers = len(layer_sizes)
        self.num_outputs = layer_sizes[-1]
        self.neuron_type = neuron_type
        self.threshold = threshold
        self.dropout = dropout
        self.normalization = normalization
        self.use_bias = use_bias
        self.bidirectional = bidirectional
        self.use_readout_layer = use_readout_layer
        self.is_snn = True

        if neuron_type not in ["LIF", "adLIF", "RLIF", "RadLIF"]:
            raise ValueError(f"Invalid neuron type {neuron_type}")

        self.snn = self._init_layers()

    def _init_layers(self):
        snn = nn.ModuleList([])
        input_size = self.input_size
        layer_cls = _LAYER_CLASSES[self.neuron_type]
        num_hidden = self.num_layers - 1 if self.use_readout_layer else self.num_layers
        for i in range(num_hidden):
            snn.append(layer_cls(input_size=input_size, hidden_size=self.layer_sizes[i],
                                 batch_size=self.batch_size, threshold=self.threshold,
                                 dropout=self.dropout, normalization=self.normalization,
                                 use_bias=self.use_bias, bidirectional=self.bidirectional))
            input_size = self.layer_sizes[i] * (1 + self.bidirectional)
        if self.use_readout_layer:
            snn.append(ReadoutLayer(input_size=input_size, hidden_size=self.layer_sizes[-1],
                                    batch_size=self.batch_size, dropout=self.dropout,
                                    normalization=self.normalization, use_bias=self.use_bias))
        return snn

    def forward(self, x):
        if self.reshape:                                         # snns.py:160-164
            if x.ndim == 4:
                x = x.reshape(x.shape[0], x.shape[1], x.shape[2] * x.shape[3])
            else:
                raise NotImplementedError
        ahead = self._prepare_ahead(x)
        try:
            return self._forward_layers(x)
        finally:
            if ahead is not None:
                for lay in self.snn:
                    _PENDING_PREP.pop(id(lay), None)
                    _PENDING_PREP.pop((id(lay), "W"), None)
                torch.cuda.current_stream(x.device).wait_stream(ahead)   # (every layer has already waited for its share)

    def _prepare_ahead(self, x):
        """Issue what the spiking layers' forward passes need from their PARAMETERS alone -- the clamps (snns.py:706-709),
        the images of V0 the recurrence kernels read (snns.py:712), and in "device" state mode the initial states
        (snns.py:700-702, same order) with rec_0 = s0 @ V0 -- for all layers at once on a side stream: it runs under the
        first projection and the first layer's recurrence (which leaves 20 SMs idle) instead of in front of every layer's
        recurrence.  Returns the side stream, or None when nothing was issued."""
        if not (_PREP_AHEAD and x.is_cuda and x.ndim == 3):
            return None
        layers = [lay for lay in self.snn if isinstance(lay, _SpikingLayerBase)]
        if not layers:
            return None
        dev = x.device
        with torch.cuda.device(dev):
            Be = x.shape[0] * (2 if self.bidirectional else 1)
            host_draws = None
            if _STATE_INIT == "cpu" and _DEVICE_MT:
                # the reference's CPU-generator draws of ALL layers (ut, [wt,] st per spiking layer, then the readout's ut:
                # the order in which the layers run) as one replay of the generator's stream on the device
                sizes = [(3 if lay._adaptive else 2) * Be * lay.hidden_size for lay in layers]
                ro = self.snn[-1] if isinstance(self.snn[-1], ReadoutLayer) else None
                if ro is not None:
                    sizes.append(x.shape[0] * ro.hidden_size)
                flat = _rng.cpu_generator_rand(sum(sizes), dev)
                if flat is not None:
                    host_draws = list(torch.split(flat, sizes))
                    if ro is not None:
                        _PENDING_PREP[id(ro)] = host_draws.pop().view(x.shape[0], ro.hidden_size)
            main, side = torch.cuda.current_stream(), _side_stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side), torch.no_grad():
                def split_weight(lay):     # the 16-bit operand terms of W (max|W| + split: two launches per layer)
                    w = lay.W.weight
                    if w.is_cuda and w.dtype == torch.float32 and w.is_contiguous():
                        ev = torch.cuda.Event()
                        terms = _F.gemm.split_general(w.detach())
                        ev.record(side)
                        _PENDING_PREP[(id(lay), "W")] = (terms, ev, (_F.gemm.MODE, w.data_ptr(), w._version))
                for li, lay in enumerate(layers):
                    split_weight(lay)
                    H = lay.hidden_size
                    states = None
                    if host_draws is not None:
                        r = host_draws[li].view(-1, Be, H)
                        states = (r[0], (r[1] if lay._adaptive else None), r[-1])
                    elif _STATE_INIT == "device":
                        r = torch.rand(3 if lay._adaptive else 2, Be, H, device=dev)
                        states = (r[0], (r[1] if lay._adaptive else None), r[-1])
                    pr = _F.prepare_cell(_F.KINDS[lay._kind], lay.alpha, getattr(lay, "beta", None),
                                         getattr(lay, "a", None), getattr(lay, "b", None),
                                         lay.V.weight if lay._recurrent else None, Be, H, states)
                    pr.event = torch.cuda.Event()
                    pr.event.record(side)
                    _PENDING_PREP[id(lay)] = pr
                if isinstance(self.snn[-1], ReadoutLayer):
                    split_weight(self.snn[-1])
        return side

    def _forward_layers(self, x):
        rates = []
        in_scale = None          # the network input is a general fp32 tensor
        post = None              # the previous spiking layer's post pass (operand terms, spike counts)
        for i, snn_lay in enumerate(self.snn):
            x_terms = post.terms if post is not None else None
            if not (self.use_readout_layer and i == self.num_layers - 1):
                holder = []
                x = snn_lay(x, in_scale=in_scale, x_terms=x_terms, post_out=holder)   # __call__: hooks fire
                post = holder[0]
                # snns.py:174 takes cat(all_spikes, dim=2).mean(dim=(0, 1)); the per-neuron means are the same
                # without materialising the concatenated (B, T, sum H) tensor -- from the post pass's integer
                # counts when there is one
                rates.append(post.rates(x) if post is not None else x.mean(dim=(0, 1)))
                # a spiking layer emits exactly {0, 1/(1-p)} in training and {0, 1} otherwise
                p = snn_lay.drop.p
                in_scale = 1.0 / (1.0 - p) if (snn_lay.drop.training and 0.0 < p < 1.0) else 1.0
            else:
                x = snn_lay(x, in_scale=in_scale, x_terms=x_terms)
        firing_rates = torch.cat(rates, dim=0)
        return x, firing_rates
