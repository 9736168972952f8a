"""CPU tests of the host side: drop-in contract of sparch_b200.snns (constructor, state_dict
layout, RNG draw order, pickling, error behaviour) and the C-ABI library surface.  No kernel is
launched here (there is no GPU in the build container)."""
import ctypes
import io
import json
import os
import pickle
import re

import numpy as np
import pytest
import torch

from oracle import snn_oracle as orc
from tests.helpers import GOLDEN_DIR, Golden, golden_names

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if fn.endswith(".h"):
            src = open(os.path.join(inc, fn)).read()
            names += re.findall(r"SPARCH_API\s+[\w\s\*]+?\b(sparch_\w+)\s*\(", src)
    return names


def test_library_exports_every_declared_symbol():
    from sparch_b200 import _lib, build
    build.build()
    names = _declared_symbols()
    assert len(names) >= 14
    h = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/ but not exported"
    # every declared entry point has a ctypes prototype (and nothing undeclared is bound)
    bound = set(_lib._PROTOS) | {"sparch_last_error"}
    assert bound == set(names), (sorted(bound ^ set(names)))
    assert _lib.lib().sparch_abi_version() >= 1


def test_ctypes_prototypes_match_the_header():
    """Every ctypes prototype in sparch_b200/_lib.py has the arity and argument classes (pointer / int / int64 /
    float) of the declaration in include/sparch_b200.h: a hand-edited signature cannot drift unnoticed."""
    from sparch_b200 import _lib
    src = ""
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if fn.endswith(".h"):
            src += open(os.path.join(inc, fn)).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    decls = re.findall(r"SPARCH_API\s+[\w\s\*]+?\b(sparch_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    assert len(decls) >= 30

    def cls(param):
        param = " ".join(param.split())
        if param in ("void", ""):
            return ""
        if "*" in param or "sparch_stream_t" in param:
            return "p"
        base = param.rsplit(" ", 1)[0].replace("const ", "").strip()
        return {"int": "i", "int64_t": "l", "size_t": "l", "float": "f", "long long": "l"}[base]

    for name, params in decls:
        sig = "".join(cls(x) for x in params.split(","))
        if name == "sparch_last_error":
            continue
        assert _lib._PROTOS[name] == sig, f"{name}: header says {sig!r}, _lib.py binds {_lib._PROTOS[name]!r}"


def test_init_rng_and_state_dict_contract():
    """Same seed -> same parameters, in the same state_dict order, as the reference."""
    import sparch_b200
    z = np.load(os.path.join(GOLDEN_DIR, "init_contract.npz"))
    cases = sorted({k.split(".")[0] for k in z.files})
    assert len(cases) == 5
    for name in cases:
        kw = json.loads(str(z[name + ".kwargs"]))
        keys = json.loads(str(z[name + ".keys"]))
        torch.manual_seed(0)
        net = sparch_b200.SNN(input_shape=(2, None, 5), **kw)
        sd = net.state_dict()
        assert list(sd.keys()) == keys, name
        for k in keys:
            np.testing.assert_array_equal(sd[k].numpy(), z[f"{name}.sd.{k}"], err_msg=f"{name}:{k}")
        # the oracle module obeys the same contract
        torch.manual_seed(0)
        ref = orc.build_oracle_snn((2, None, 5), **kw)
        assert list(ref.state_dict().keys()) == keys
        for k in keys:
            np.testing.assert_array_equal(ref.state_dict()[k].numpy(), z[f"{name}.sd.{k}"])


@pytest.mark.parametrize("name", golden_names())
def test_state_dict_of_fixtures_loads(name):
    import sparch_b200
    g = Golden(name)
    net = sparch_b200.SNN(g.input_shape, **g.kwargs)
    net.load_state_dict(g.state_dict())  # strict: keys and shapes must match the reference's
    assert net.is_snn and len(net.snn) == len(g.kwargs["layer_sizes"])


def test_constructor_contract_and_errors():
    import sparch_b200
    with pytest.raises(ValueError, match="Invalid neuron type"):
        sparch_b200.SNN((4, None, 7), [8, 3], neuron_type="LSTM")
    net = sparch_b200.SNN((4, None, 7), [8, 8, 3], neuron_type="RadLIF", bidirectional=True,
                          dropout=0.1, use_bias=True)
    lay0, lay1, ro = net.snn
    assert type(lay0).__name__ == "RadLIFLayer" and type(ro).__name__ == "ReadoutLayer"
    assert lay0.W.weight.shape == (8, 7) and lay1.W.weight.shape == (8, 16)   # snns.py:140
    assert ro.W.weight.shape == (3, 16) and lay0.batch_size == 8
    assert lay0.V.weight.shape == (8, 8) and lay0.norm.momentum == 0.05
    assert callable(lay0.spike_fct) and lay0.drop.p == 0.1
    with pytest.raises(RuntimeError, match="CUDA only"):
        net(torch.zeros(4, 5, 7))           # no CPU fallback
    net4 = sparch_b200.SNN((4, None, 7, 2), [8, 3])
    assert net4.reshape
    with pytest.raises(NotImplementedError):
        net4(torch.zeros(4, 5, 14))


def test_whole_module_pickle_round_trip():
    """exp.py:462 saves the whole module with torch.save; it must survive a round trip."""
    import sparch_b200
    torch.manual_seed(3)
    net = sparch_b200.SNN((4, None, 7), [8, 8, 3], neuron_type="adLIF")
    buf = io.BytesIO()
    torch.save(net, buf)
    buf.seek(0)
    net2 = torch.load(buf, weights_only=False)
    for (k1, v1), (k2, v2) in zip(net.state_dict().items(), net2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    assert pickle.loads(pickle.dumps(net.snn[0].spike_fct)) is not None


def _load_reference_checkpoint():
    """Unpickle tests/golden/reference_module.pt -- a torch.save(net) of the UNTOUCHED reference SNN
    (oracle/make_pickle_fixture.py, exp.py:462) -- with sparch.models.snns re-pointed at sparch_b200.snns."""
    import sys
    import types
    import sparch_b200.snns as ours
    saved = {k: sys.modules.get(k) for k in ("sparch", "sparch.models", "sparch.models.snns")}
    try:
        pkg, models = types.ModuleType("sparch"), types.ModuleType("sparch.models")
        pkg.models, models.snns = models, ours
        sys.modules.update({"sparch": pkg, "sparch.models": models, "sparch.models.snns": ours})
        return torch.load(os.path.join(GOLDEN_DIR, "reference_module.pt"), weights_only=False)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_reference_checkpoint_unpickles_into_these_classes():
    """SURVEY.md 8b pickle contract: a whole-module checkpoint written by the reference loads after the one-line
    switch of INTEGRATION.md -- same classes by name, same attributes, same state_dict."""
    import sparch_b200.snns as ours
    net = _load_reference_checkpoint()
    assert type(net) is ours.SNN and net.is_snn and net.neuron_type == "RadLIF"
    assert [type(l) for l in net.snn] == [ours.RadLIFLayer, ours.RadLIFLayer, ours.ReadoutLayer]
    run = np.load(os.path.join(GOLDEN_DIR, "reference_module_run.npz"))
    sd = net.state_dict()
    want = {k[4:]: run[k] for k in run.files if k.startswith("sd0.")}
    assert list(sd) == list(want)
    for k, v in want.items():
        np.testing.assert_array_equal(sd[k].numpy(), v)
    lay = net.snn[0]
    assert lay.threshold == 1.0 and lay.normalize and lay.drop.p == 0.0 and not lay.bidirectional
    with pytest.raises(RuntimeError):          # still no CPU path
        net(torch.zeros(4, 3, 7))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under sparch_b200/ may reference it."""
    pkg = os.path.join(ROOT, "sparch_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|snn_oracle", src, re.M), \
                    os.path.join(dp, fn)


def test_host_side_guards_without_a_gpu():
    """Host logic that needs no kernel: operand-term pairings, precision-mode switch, and the loud failures of the
    CUDA-only entry points on CPU tensors (no silent fallback anywhere)."""
    import sparch_b200
    from sparch_b200 import functional as F, gemm
    from sparch_b200.graphs import GraphedTrainStep
    from sparch_b200.optim import Adam
    # every (general, general) pairing drops only the lowest-order product; spikes need one term
    assert gemm.pairs_for(2, 2) == [(1, 0), (0, 1), (0, 0)] and gemm.pairs_for(1, 2) == [(0, 1), (0, 0)]
    assert len(gemm.pairs_for(3, 3)) == 6 and gemm.pairs_for(1, 1) == [(0, 0)]
    try:
        for mode, gm in (("fp32", "f16x2"), ("fp32-bf16x3", "bf16x3"), ("bf16", "bf16x1")):
            sparch_b200.set_precision(mode)
            assert gemm.MODE == gm
        with pytest.raises(ValueError):
            sparch_b200.set_precision("fp64")
    finally:
        sparch_b200.set_precision("fp32")
    with pytest.raises(ValueError):
        sparch_b200.set_state_init("host")
    x = torch.zeros(2, 3, 4)
    with pytest.raises(RuntimeError):
        F.spike_post(x, 0.1, F.NormState("none"), recurrent=False)
    with pytest.raises(RuntimeError):
        F.SpikeFunctionBoxcar.apply(x)
    p = torch.zeros(3, requires_grad=True)
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError):
        Adam([p], 1e-2).step()
    with pytest.raises(ValueError):
        Adam([p], -1.0)
    net = sparch_b200.SNN((2, None, 4), layer_sizes=[8, 8, 3])
    with pytest.raises(RuntimeError):          # needs device state init and CUDA tensors
        GraphedTrainStep(net, Adam(net.parameters(), 1e-2), torch.nn.CrossEntropyLoss(), x, torch.zeros(2, dtype=torch.long))


def test_restated_training_loop_equals_the_references_own():
    """oracle/exp_loop.py against the reference's unbound Experiment.train_one_epoch / valid_one_epoch (exp.py:341-459) on
    the same CPU model (the torch restatement of the reference SNN), stub loaders, optimizer and scheduler: identical
    per-step losses, learning rate and parameters after two epochs, with and without the firing-rate regularisers.
    Needs /root/reference (build container only)."""
    import copy
    import sys
    if not os.path.isdir("/root/reference/sparch"):
        pytest.skip("the reference is not on this machine")
    import types
    # exp.py:26-27 imports the dataloaders, which need torchaudio / h5py packages this image does not have: the loops
    # under test never touch them, so two empty stand-ins take their place for the duration of the import
    stubs = {}
    for name, fn in (("sparch.dataloaders.nonspiking_datasets", "load_hd_or_sc"),
                     ("sparch.dataloaders.spiking_datasets", "load_shd_or_ssc")):
        m = types.ModuleType(name)
        setattr(m, fn, lambda *a, **k: None)
        stubs[name] = m
    had = {k for k in sys.modules if k == "sparch" or k.startswith("sparch.")}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    sys.path.insert(0, "/root/reference")
    try:
        from sparch.exp import Experiment
    finally:
        sys.path.remove("/root/reference")
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if (k == "sparch" or k.startswith("sparch.")) and k not in had]:
            del sys.modules[k]                                  # (Experiment keeps its own references)
    from oracle import exp_loop
    from oracle import snn_oracle as orc
    for reg in (False, True):
        kw = dict(layer_sizes=[24, 24, 10], neuron_type="RadLIF", normalization="batchnorm", dropout=0.0)
        torch.manual_seed(0)
        net_a = orc.build_oracle_snn((8, None, 40), **kw)
        net_b = copy.deepcopy(net_a)
        exps = [exp_loop.stub_experiment(n, torch.device("cpu"), use_regularizers=reg) for n in (net_a, net_b)]
        losses = []
        for i, ex in enumerate(exps):
            torch.manual_seed(5)                               # the initial-state draws (snns.py:700-702)
            ls = []
            for e in (1, 2):
                if i == 0:
                    before = [b[0].clone() for b in ex.train_loader]
                    Experiment.train_one_epoch(ex, e)          # the reference's own function on the stub
                    best = Experiment.valid_one_epoch(ex, e, 0, 0)
                    assert all(torch.equal(a, b[0]) for a, b in zip(before, ex.train_loader))
                else:
                    l, _ = exp_loop.train_one_epoch(ex, e)
                    ls += l
                    best = exp_loop.valid_one_epoch(ex, e, 0, 0)
            losses.append((ls, best, ex.opt.param_groups[-1]["lr"]))
        assert losses[0][1] == losses[1][1] and losses[0][2] == losses[1][2]
        for pa, pb in zip(net_a.parameters(), net_b.parameters()):
            assert torch.equal(pa, pb)
        assert len(losses[1][0]) == 8 and all(np.isfinite(losses[1][0]))


def test_data_oracle_equals_the_references_spiking_dataset():
    """oracle/data_oracle.py against the reference's own SpikingDataset.__getitem__ / generateBatch
    (spiking_datasets.py:66-86) on synthetic SHD-shaped events (float16 times, duplicates in a bin, events in the first and
    in the last bin): identical dense tensors, lengths and labels.  Needs /root/reference (build container only)."""
    import sys
    import types
    if not os.path.isdir("/root/reference/sparch"):
        pytest.skip("the reference is not on this machine")
    had = {k for k in sys.modules if k == "sparch" or k.startswith("sparch.")}
    fake_h5 = "h5py" not in sys.modules
    try:
        import h5py  # noqa: F401
        fake_h5 = False
    except Exception:
        sys.modules["h5py"] = types.ModuleType("h5py")
    sys.path.insert(0, "/root/reference")
    try:
        from sparch.dataloaders.spiking_datasets import SpikingDataset
    finally:
        sys.path.remove("/root/reference")
        if fake_h5:
            sys.modules.pop("h5py", None)
        for k in [k for k in sys.modules if (k == "sparch" or k.startswith("sparch.")) and k not in had]:
            del sys.modules[k]
    from oracle import data_oracle as dor
    T, U, y = dor.synthetic_events(5, seed=3, rate=3000)
    T[0][:3] = 0.0                                          # first bin edge, three times (duplicates sum)
    U[0][:3] = 7
    T[1][-1] = np.float16(1.39)                             # last bin
    stub = types.SimpleNamespace(firing_times=T, units_fired=U, labels=np.asarray(y), device="cpu", nb_steps=100,
                                 nb_units=700, time_bins=np.linspace(0, 1.4, num=100))
    items = [SpikingDataset.__getitem__(stub, i) for i in range(5)]
    xs, xlens, ys = SpikingDataset.generateBatch(stub, items)
    ox, olens, oy = dor.batch_to_dense(T, U, y)
    assert xs.shape == (5, 100, 700) and np.array_equal(xs.numpy(), ox)
    assert ox[0, 1, 7] >= 3.0                               # digitize puts t = 0 into bin 1; duplicates are counts
    assert np.array_equal(xlens.numpy(), olens) and np.array_equal(ys.numpy(), oy)


def test_spiking_batcher_is_cuda_only():
    from sparch_b200.data import SpikingBatcher
    with pytest.raises(RuntimeError):
        SpikingBatcher(device="cpu")


def test_cpu_generator_state_layout_behind_the_device_draws():
    """sparch_b200/rng.py replays torch's default CPU generator on the device from its state blob.  This pins, on the CPU,
    everything that code assumes about the blob and the stream: 5056 bytes, `left` at byte 8, `next` at 16, 624 state words
    as uint64 from byte 24; torch.rand(n) = one MT19937 word per element, (word & (2^24 - 1)) * 2^-24, in element order;
    and that writing position / state back into the blob moves the generator to exactly that point."""
    import struct
    import numpy as np
    from sparch_b200 import rng

    def mt_block(mt):                      # one regeneration of the 624-word state (Matsumoto & Nishimura 1998)
        mt = mt.copy()
        for i in range(624):
            y = (int(mt[i]) & 0x80000000) | (int(mt[(i + 1) % 624]) & 0x7fffffff)
            mt[i] = int(mt[(i + 397) % 624]) ^ (y >> 1) ^ (0x9908b0df if y & 1 else 0)
        return mt

    def temper(y):
        y ^= y >> 11
        y ^= (y << 7) & 0x9d2c5680
        y ^= (y << 15) & 0xefc60000
        y ^= y >> 18
        return y & 0xffffffff

    torch.manual_seed(1234)
    torch.rand(100)                                            # somewhere inside a block
    blob = torch.get_rng_state()
    assert blob.numel() == rng._STATE_BYTES
    raw = bytearray(blob.numpy().tobytes())
    left, = struct.unpack_from("<i", raw, rng._OFF_LEFT)
    nxt, = struct.unpack_from("<Q", raw, rng._OFF_NEXT)
    pos = rng._MT_N + 1 - left
    assert pos == nxt == 100
    mt = np.frombuffer(raw, dtype=np.uint64, count=rng._MT_N, offset=rng._OFF_STATE).astype(np.uint64)
    n = 700                                                    # crosses into the next block
    want = torch.rand(n)
    got, words, p = [], mt, pos
    while len(got) < n:
        if p == 624:
            words, p = mt_block(words), 0
        got.append((temper(int(words[p])) & 0xffffff) * 2.0 ** -24)
        p += 1
    assert np.array_equal(np.asarray(got, dtype=np.float32), want.numpy())
    # the blob the device path would write back: new state words, position p
    struct.pack_into("<i", raw, rng._OFF_LEFT, rng._MT_N + 1 - p)
    struct.pack_into("<Q", raw, rng._OFF_NEXT, p)
    raw[rng._OFF_STATE:rng._OFF_STATE + 8 * rng._MT_N] = words.astype(np.uint64).tobytes()
    after = torch.get_rng_state()
    assert bytes(raw) == after.numpy().tobytes()
