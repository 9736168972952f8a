"""Generate tests/golden/*.npz by running the UNTOUCHED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference ships no golden vectors, so these fixtures -- inputs,
weights, initial-state draws, per-layer currents/spikes, outputs, loss and
autograd gradients produced by ``/root/reference/sparch/models/snns.py`` itself
-- are what pins the oracle (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_*.py).  Nothing at test/bench time reads /root/reference.
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SPARCH_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

CASES = [
    # name, kwargs, (B, T, F), n_classes, stable_a
    ("lif_bn", dict(layer_sizes=[32, 32, 5], neuron_type="LIF"), (4, 12, 7), 5, False),
    ("adlif_bn", dict(layer_sizes=[32, 32, 5], neuron_type="adLIF"), (4, 12, 7), 5, False),
    ("rlif_bn", dict(layer_sizes=[32, 32, 5], neuron_type="RLIF"), (4, 12, 7), 5, False),
    ("radlif_bn", dict(layer_sizes=[32, 32, 5], neuron_type="RadLIF"), (4, 12, 7), 5, False),
    ("radlif_bn_h64", dict(layer_sizes=[64, 64, 6], neuron_type="RadLIF"), (6, 16, 10), 6, True),
    ("radlif_bidir", dict(layer_sizes=[32, 32, 5], neuron_type="RadLIF", bidirectional=True),
     (3, 10, 7), 5, False),
    ("rlif_bidir_bias", dict(layer_sizes=[32, 32, 5], neuron_type="RLIF", bidirectional=True,
                             use_bias=True), (3, 10, 7), 5, False),
    ("adlif_layernorm", dict(layer_sizes=[32, 32, 5], neuron_type="adLIF",
                             normalization="layernorm"), (4, 12, 7), 5, False),
    ("lif_nonorm_bias", dict(layer_sizes=[32, 32, 5], neuron_type="LIF", normalization="none",
                             use_bias=True), (4, 12, 7), 5, False),
    ("radlif_odd_noreadout", dict(layer_sizes=[9, 11, 5], neuron_type="RadLIF",
                                  use_readout_layer=False, threshold=0.7), (5, 12, 7), 5, False),
    ("rlif_4d_input", dict(layer_sizes=[32, 5], neuron_type="RLIF"), (3, 8, 4, 2), 5, False),
]


def run_case(SNN, name, kw, xshape, ncls, stable_a, eval_mode=False):
    torch.manual_seed(0)
    input_shape = (xshape[0], None) + tuple(xshape[2:])
    net = SNN(input_shape=input_shape, **kw)
    if stable_a:
        with torch.no_grad():
            for lay in net.snn:
                if hasattr(lay, "a"):
                    lay.a.abs_()
    # push a few parameters outside their clamp range so clamp fwd/bwd is exercised, and
    # raise the drive (norm gain/offset or W scale) so that neurons actually spike at init
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for lay in net.snn:
            if getattr(lay, "normalize", False):
                lay.norm.weight.copy_(2.0 + 3.0 * torch.rand(lay.norm.weight.shape, generator=gen))
                lay.norm.bias.copy_(0.5 + 1.5 * torch.rand(lay.norm.bias.shape, generator=gen))
            else:
                lay.W.weight.mul_(6.0)
            lay.alpha[0] = 0.5
            lay.alpha[1] = 0.99
            if hasattr(lay, "b"):
                lay.b[2] = -0.3
                lay.a[3] = 1.5
                lay.beta[4] = 0.999
    torch.manual_seed(1234)
    if len(xshape) == 3 and xshape[2] > 20:
        x = (torch.rand(*xshape) < 0.03).float()
    else:
        x = torch.randn(*xshape)
    y = torch.randint(0, ncls, (xshape[0],))
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}

    draws, currents, layer_out = [], {}, {}
    real_rand = torch.rand

    def rec_rand(*a, **k):
        r = real_rand(*a, **k)
        draws.append(r.clone())
        return r

    hooks = []
    for i, lay in enumerate(net.snn):
        tgt = lay.norm if getattr(lay, "normalize", False) else lay.W
        hooks.append(tgt.register_forward_hook(
            lambda m, inp, out, i=i: currents.__setitem__(i, out.detach().clone())))
        hooks.append(lay.register_forward_hook(
            lambda m, inp, out, i=i: layer_out.__setitem__(i, out.detach().clone())))
    if eval_mode:
        net.eval()
    torch.manual_seed(42)
    torch.rand = rec_rand
    try:
        out, rates = net(x)
    finally:
        torch.rand = real_rand
    for h in hooks:
        h.remove()
    if out.ndim == 2:
        loss = torch.nn.functional.cross_entropy(out, y)
    else:
        loss = out.sum(1).square().mean()
    grads = {}
    if not eval_mode:
        loss.backward()
        grads = {k: p.grad.clone() for k, p in net.named_parameters()}
    blob = {"x": x.numpy(), "y": y.numpy(), "out": out.detach().numpy(),
            "rates": rates.detach().numpy(), "loss": np.float64(loss.item()),
            "meta": np.array(json.dumps({"kwargs": kw, "xshape": list(xshape), "ncls": ncls,
                                         "eval": eval_mode}))}
    for k, v in sd0.items():
        blob["sd0." + k] = v.numpy()
    for k, v in net.state_dict().items():
        if "running" in k or "num_batches" in k:
            blob["sd1." + k] = v.numpy()
    for k, v in grads.items():
        blob["grad." + k] = v.numpy()
    for i, d in enumerate(draws):
        blob[f"draw.{i}"] = d.numpy()
    for i, c in currents.items():
        blob[f"cur.{i}"] = c.numpy()
    for i, c in layer_out.items():
        blob[f"lay.{i}"] = c.numpy()
    suffix = "_eval" if eval_mode else ""
    np.savez_compressed(os.path.join(OUT, name + suffix + ".npz"), **blob)
    print(f"{name+suffix:28s} loss={loss.item():.6f} mean rate={rates.mean().item():.4f} "
          f"finite={bool(torch.isfinite(out).all())}")


def boxcar_kat(snns):
    x = torch.tensor([-0.6, -0.5, -0.49999, 0.0, 1e-8, 0.5, 0.50001], requires_grad=True)
    s = snns.SpikeFunctionBoxcar.apply(x)
    s.backward(torch.arange(1.0, 8.0))
    lim = torch.tensor([0.5, float(np.exp(-1 / 5)), 0.9, float(np.exp(-1 / 25)), 0.99],
                       requires_grad=True)
    c = torch.clamp(lim, min=np.exp(-1 / 5), max=np.exp(-1 / 25))
    c.sum().backward()
    np.savez(os.path.join(OUT, "boxcar_kat.npz"), x=x.detach().numpy(), s=s.detach().numpy(),
             gx=x.grad.numpy(), clamp_in=lim.detach().numpy(), clamp_out=c.detach().numpy(),
             clamp_grad=lim.grad.numpy())
    print("boxcar_kat", s.tolist(), x.grad.tolist(), lim.grad.tolist())


INIT_CASES = [
    ("LIF", dict(layer_sizes=[8, 6, 3], neuron_type="LIF")),
    ("adLIF", dict(layer_sizes=[8, 6, 3], neuron_type="adLIF", use_bias=True)),
    ("RLIF", dict(layer_sizes=[8, 6, 3], neuron_type="RLIF", bidirectional=True)),
    ("RadLIF", dict(layer_sizes=[8, 6, 3], neuron_type="RadLIF", normalization="layernorm")),
    ("RadLIF_noreadout", dict(layer_sizes=[8, 6], neuron_type="RadLIF", use_readout_layer=False)),
]


def init_contract(SNN):
    """RNG / state_dict contract: the reference's parameters right after
    construction under torch.manual_seed(0), plus the initial-state draws of one forward."""
    blob = {}
    for name, kw in INIT_CASES:
        torch.manual_seed(0)
        net = SNN(input_shape=(2, None, 5), **kw)
        blob[name + ".keys"] = np.array(json.dumps(list(net.state_dict().keys())))
        blob[name + ".kwargs"] = np.array(json.dumps(kw))
        for k, v in net.state_dict().items():
            blob[name + ".sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "init_contract.npz"), **blob)
    print("init_contract", [n for n, _ in INIT_CASES])


def main():
    sys.path.insert(0, REF)
    from sparch.models import snns  # the untouched reference
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    boxcar_kat(snns)
    init_contract(snns.SNN)
    for name, kw, xshape, ncls, stable in CASES:
        run_case(snns.SNN, name, kw, xshape, ncls, stable)
    run_case(snns.SNN, "radlif_bn", CASES[3][1], CASES[3][2], 5, False, eval_mode=True)


if __name__ == "__main__":
    main()
