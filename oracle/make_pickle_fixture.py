"""Write tests/golden/reference_module.pt: a whole-module checkpoint of the UNTOUCHED reference SNN, saved exactly
as sparch/exp.py:462 does (``torch.save(self.net, path)``), plus the case it was run on.

Run in the build container only (needs /root/reference):

    python oracle/make_pickle_fixture.py

tests/test_host_contract.py unpickles it with ``sparch.models.snns`` re-pointed at ``sparch_b200.snns`` (the
one-line switch of INTEGRATION.md) -- the pickle contract of SURVEY.md 8b; tests/test_gpu_parity.py then runs the
unpickled module on the GPU against the reference's recorded outputs.  Nothing at test time reads /root/reference.
"""
import io
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SPARCH_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    sys.path.insert(0, REF)
    from sparch.models.snns import SNN        # the reference's own class
    torch.manual_seed(0)
    kw = dict(layer_sizes=[24, 24, 5], neuron_type="RadLIF", dropout=0.0, normalization="batchnorm")
    net = SNN(input_shape=(4, None, 7), **kw)
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for lay in net.snn:
            if hasattr(lay, "a"):
                lay.a.abs_()
            lay.norm.weight.copy_(2.0 + 3.0 * torch.rand(lay.norm.weight.shape, generator=gen))
            lay.norm.bias.copy_(0.5 + 1.5 * torch.rand(lay.norm.bias.shape, generator=gen))
    buf = io.BytesIO()
    torch.save(net, buf)                       # exp.py:462
    open(os.path.join(OUT, "reference_module.pt"), "wb").write(buf.getvalue())
    # what the reference computes with this module (train mode, seeds as in the other fixtures)
    x = torch.randn(4, 12, 7, generator=torch.Generator().manual_seed(1234))
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    torch.manual_seed(42)
    out, rates = net(x)
    np.savez(os.path.join(OUT, "reference_module_run.npz"), x=x.numpy(), out=out.detach().numpy(),
             rates=rates.detach().numpy(), **{"sd0." + k: v.numpy() for k, v in sd0.items()})
    print("wrote reference_module.pt (%d bytes), out %s" % (len(buf.getvalue()), tuple(out.shape)))


if __name__ == "__main__":
    main()
