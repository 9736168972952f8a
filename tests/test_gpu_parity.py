"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden
fixtures produced by the untouched reference.  Protocol of SURVEY.md 7 #1:
 (i)   teacher-forced single steps: u_t within rel 1e-5, spikes exact except straddlers;
 (ii)  given-mask backward: the BPTT is linear given the tapes -> gradients within rel 2e-5;
 (iii) free-running spike trains on short seeded runs: exact for LIF/adLIF, <= 1e-3 flips for
       the recurrent kinds (the recurrent sum is re-associated);
 (iv)  whole-model outputs / loss / gradients / running stats against the reference's own.
"""
import numpy as np
import pytest
import torch

from oracle import snn_oracle as orc
from tests.helpers import Golden, golden_names, rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
FLIP_TOL = 1e-3   # fraction of spike bits allowed to differ on the recurrent kinds
U_RTOL = 1e-5     # membrane potential, relative to max |u|
G_RTOL = 2e-5     # gradients given the masks, relative to max |g|


def _mods():
    import sparch_b200
    from sparch_b200 import functional
    return sparch_b200, functional


def _net(g, device=DEV):
    sp, _ = _mods()
    net = sp.SNN(g.input_shape, **g.kwargs)
    net.load_state_dict(g.state_dict())
    net = net.to(device)
    if g.eval:
        net.eval()
    return net


def _layer_params(lay):
    kind = lay._kind
    alpha = lay.alpha.detach().cpu().numpy()
    p = orc.clamp_params(kind, alpha,
                         lay.beta.detach().cpu().numpy() if lay._adaptive else alpha,
                         lay.a.detach().cpu().numpy() if lay._adaptive else alpha,
                         lay.b.detach().cpu().numpy() if lay._adaptive else alpha)
    V0 = None
    if lay._recurrent:
        V0 = lay.V.weight.detach().cpu().numpy().copy()
        np.fill_diagonal(V0, 0)
    return p, V0


def _hidden_layer_inputs(g, net):
    """Yield (i, layer, I, u0, w0, s0) for every spiking layer from the golden captures."""
    draws = g.draws()
    di = 0
    for i, lay in enumerate(net.snn):
        if not hasattr(lay, "_kind"):
            break
        u0 = draws[di]; di += 1
        w0 = None
        if lay._adaptive:
            w0 = draws[di]; di += 1
        s0 = draws[di]; di += 1
        yield i, lay, g.cur(i), u0, w0, s0


def _run_cell(lay, I, u0, w0, s0, need_grad=False):
    _, F = _mods()
    t = lambda z: None if z is None else torch.from_numpy(np.ascontiguousarray(z)).to(DEV)
    It = t(I)
    if need_grad:
        It.requires_grad_(True)
    S = F.SpikingCellFunction.apply(
        It, None, None, lay.alpha, getattr(lay, "beta", None), getattr(lay, "a", None),
        getattr(lay, "b", None), lay.V.weight if lay._recurrent else None, t(u0), t(w0), t(s0),
        lay._kind, lay.threshold, F.NormState("none"))
    return It, S


def test_boxcar_known_answers_gpu():
    _, F = _mods()
    x = torch.tensor([-0.6, -0.5, -0.49999, 0.0, 1e-8, 0.5, 0.50001], device=DEV, requires_grad=True)
    s = F.SpikeFunctionBoxcar.apply(x)
    s.backward(torch.arange(1.0, 8.0, device=DEV))
    assert s.tolist() == [0, 0, 0, 0, 1, 1, 1]
    assert x.grad.tolist() == [0, 0, 3, 4, 5, 6, 0]
    # a larger random vector against the oracle
    xr = torch.randn(100003, device=DEV, requires_grad=True)
    gr = torch.randn(100003, device=DEV)
    sr = F.SpikeFunctionBoxcar.apply(xr)
    sr.backward(gr)
    xn = xr.detach().cpu().numpy()
    np.testing.assert_array_equal(sr.detach().cpu().numpy(), orc.boxcar_forward(xn))
    np.testing.assert_array_equal(xr.grad.cpu().numpy(), orc.boxcar_backward(xn, gr.cpu().numpy()))


CELL_CASES = ["lif_bn", "adlif_bn", "rlif_bn", "radlif_bn", "radlif_bn_h64", "radlif_odd_noreadout",
              "radlif_bidir", "rlif_bidir_bias", "adlif_layernorm", "lif_nonorm_bias", "rlif_4d_input"]


@pytest.mark.parametrize("name", CELL_CASES)
def test_cell_teacher_forced_single_steps(name):
    """Every (b, t) as its own one-step problem with the oracle's true previous state."""
    g = Golden(name)
    net = _net(g)
    for i, lay, I, u0, w0, s0 in _hidden_layer_inputs(g, net):
        p, V0 = _layer_params(lay)
        r = orc.cell_forward(lay._kind, I, p["alpha"], p.get("beta"), p.get("a"), p.get("b"), V0, u0,
                             w0, s0, theta=lay.threshold)
        Be, T, H = I.shape
        prev = lambda X, x0: np.concatenate([x0[:, None, :], X[:, :-1, :]], axis=1).reshape(Be * T, H)
        up, sp = prev(r["u"], u0), prev(r["s"], s0)
        wp = prev(r["w"], w0) if lay._adaptive else None
        _, S = _run_cell(lay, I.reshape(Be * T, 1, H), up, wp, sp)
        # the CUDA membrane value is not returned by the public Function; recompute it from the
        # oracle's single step and compare spikes, allowing only threshold straddlers to differ
        u_ref = r["u"].reshape(Be * T, H)
        s_ref = r["s"].reshape(Be * T, H)
        s_gpu = S.detach().cpu().numpy().reshape(Be * T, H)
        diff = s_gpu != s_ref
        straddle = np.abs(u_ref - np.float32(lay.threshold)) < 1e-5 * max(1.0, np.abs(u_ref).max())
        assert not (diff & ~straddle).any(), (name, i, int(diff.sum()))
        assert diff.mean() <= 1e-4, (name, i, float(diff.mean()))


@pytest.mark.parametrize("name", CELL_CASES)
def test_cell_free_running_and_given_mask_backward(name):
    g = Golden(name)
    net = _net(g)
    rng = np.random.default_rng(11)
    for i, lay, I, u0, w0, s0 in _hidden_layer_inputs(g, net):
        p, V0 = _layer_params(lay)
        r = orc.cell_forward(lay._kind, I, p["alpha"], p.get("beta"), p.get("a"), p.get("b"), V0, u0,
                             w0, s0, theta=lay.threshold)
        for q in (lay.alpha, getattr(lay, "beta", None), getattr(lay, "a", None),
                  getattr(lay, "b", None), lay.V.weight if lay._recurrent else None):
            if q is not None:
                q.grad = None
        It, S = _run_cell(lay, I, u0, w0, s0, need_grad=True)
        s_gpu = S.detach().cpu().numpy()
        flips = float((s_gpu != r["s"]).mean())
        assert flips <= (FLIP_TOL if lay._recurrent else 0.0), (name, i, flips)
        if flips > 0:
            continue  # the masks differ: gradients are not comparable
        gs = rng.standard_normal(s_gpu.shape).astype(np.float32)
        S.backward(torch.from_numpy(gs).to(DEV))
        bw = orc.cell_backward(lay._kind, gs, I, p["alpha"], p.get("beta"), p.get("a"), p.get("b"),
                               V0, u0, w0, s0, theta=lay.threshold, U=r["u"], W=r["w"], S=r["s"])
        assert rel_err(It.grad.cpu().numpy(), bw["dI"]) < G_RTOL, (name, i, "dI")
        m = orc.clamp_grad_mask(lay.alpha.detach().cpu().numpy(), orc.ALPHA_LIM)
        assert rel_err(lay.alpha.grad.cpu().numpy(), bw["dalpha"] * m) < G_RTOL, (name, i, "dalpha")
        if lay._adaptive:
            for k, lim in (("beta", orc.BETA_LIM), ("a", orc.A_LIM), ("b", orc.B_LIM)):
                m = orc.clamp_grad_mask(getattr(lay, k).detach().cpu().numpy(), lim)
                assert rel_err(getattr(lay, k).grad.cpu().numpy(), bw["d" + k] * m) < G_RTOL, (name, i, k)
        if lay._recurrent:
            assert rel_err(lay.V.weight.grad.cpu().numpy(), bw["dV"]) < G_RTOL, (name, i, "dV")


_FLIPPED_FIXTURES = []


@pytest.mark.parametrize("name", golden_names())
def test_whole_model_against_reference_fixture(name):
    """Same weights, same seeds, same input as the reference run that wrote the fixture."""
    g = Golden(name)
    net = _net(g)
    lay_out = {}
    hooks = [lay.register_forward_hook(lambda m, inp, out, i=i: lay_out.__setitem__(i, out.detach()))
             for i, lay in enumerate(net.snn)]
    torch.manual_seed(42)
    out, rates = net(g.t("x", DEV))
    for h in hooks:
        h.remove()
    total, flipped = 0, 0
    for i, lay in enumerate(net.snn):
        if hasattr(lay, "_kind"):
            ref = g.z[f"lay.{i}"]
            got = lay_out[i].cpu().numpy()
            assert got.shape == ref.shape
            total += ref.size
            flipped += int((got != ref).sum())
    assert flipped <= FLIP_TOL * total, (name, flipped, total)
    if flipped:
        # threshold straddlers (reference self-noise): values downstream of a flipped spike are not
        # comparable element by element; what stays well-defined is checked -- the firing rates move by at most the
        # flipped share -- and the fixture is counted: test_fixture_flip_budget fails if this happens to more than one
        _FLIPPED_FIXTURES.append(name)
        np.testing.assert_allclose(rates.detach().cpu().numpy(), g.z["rates"], rtol=0, atol=2.0 * flipped / total + 1e-6)
        assert np.isfinite(out.detach().cpu().numpy()).all()
        return
    np.testing.assert_allclose(out.detach().cpu().numpy(), g.z["out"], rtol=5e-5, atol=2e-6)
    np.testing.assert_allclose(rates.detach().cpu().numpy(), g.z["rates"], rtol=1e-6, atol=1e-7)
    if g.eval:
        return
    loss = g.loss_fn(out, g.t("y", DEV))
    assert abs(loss.item() - float(g.z["loss"])) <= 5e-5 * max(1.0, abs(float(g.z["loss"])))
    loss.backward()
    params = dict(net.named_parameters())
    for k, ref in g.grads().items():
        got = params[k].grad.cpu().numpy()
        # + 1e-6: a bias in front of BatchNorm has an analytically zero gradient (pure round-off)
        assert np.abs(got - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-6, (name, k)
    for k, v in net.state_dict().items():
        if "running" in k:
            np.testing.assert_allclose(v.cpu().numpy(), g.z["sd1." + k], rtol=2e-5, atol=1e-6)
        if "num_batches" in k:
            assert int(v) == int(g.z["sd1." + k])


def test_fixture_flip_budget():
    """At most one of the reference's fixtures may have taken the threshold-straddler exit above (observed: none -- the
    forward recurrence accumulates s @ V0 exactly in int32)."""
    assert len(_FLIPPED_FIXTURES) <= 1, _FLIPPED_FIXTURES


@pytest.mark.parametrize("kind,H,B,T,F", [("LIF", 128, 16, 50, 70), ("adLIF", 96, 16, 50, 70),
                                         ("RLIF", 256, 16, 40, 70), ("RadLIF", 256, 16, 40, 40)])
def test_medium_model_against_oracle_module_on_gpu(kind, H, B, T, F):
    """Layer-0 spike trains of a medium network against the oracle's torch restatement run on the
    same device (stable adaptation draw a <- |a| so that the comparison is meaningful)."""
    sp, _ = _mods()
    kw = dict(layer_sizes=[H, H, 10], neuron_type=kind, normalization="batchnorm")
    torch.manual_seed(0)
    net = sp.SNN((B, None, F), **kw)
    ref = orc.build_oracle_snn((B, None, F), **kw)
    ref.load_state_dict(net.state_dict())
    for m in (net, ref):
        with torch.no_grad():
            for lay in m.snn:
                if hasattr(lay, "a"):
                    lay.a.abs_()
                if hasattr(lay, "norm") and isinstance(lay.norm, torch.nn.BatchNorm1d):
                    lay.norm.weight.fill_(3.0)
                    lay.norm.bias.fill_(0.8)
    net, ref = net.to(DEV), ref.to(DEV)
    for lay in ref.snn:
        lay.capture = {}
    torch.manual_seed(1234)
    x = torch.randn(B, T, F, device=DEV)
    got = {}
    h = net.snn[0].register_forward_hook(lambda m, i, o: got.__setitem__(0, o.detach()))
    torch.manual_seed(42)
    out, rates = net(x)
    h.remove()
    torch.manual_seed(42)
    out_r, rates_r = ref(x)
    s_ref = ref.snn[0].capture["s"]
    flips = float((got[0] != s_ref).float().mean())
    assert 0.01 < float(s_ref.mean()) < 0.9, "degenerate firing rate: the test would be vacuous"
    assert flips <= 1e-4 if kind in ("LIF", "adLIF") else flips <= FLIP_TOL, (kind, flips)


def test_backward_is_linear_in_upstream_gradient_large():
    """Size-independent property at the north-star layer width: given the forward tapes the
    reverse pass is linear, bwd(g1 + 2 g2) == bwd(g1) + 2 bwd(g2)."""
    _, F = _mods()
    Be, T, H = 32, 24, 1024
    gen = torch.Generator(device=DEV).manual_seed(5)
    r = lambda *s: torch.rand(*s, device=DEV, generator=gen)
    I = (torch.randn(Be, T, H, device=DEV, generator=gen) * 4.0 + 2.0)
    alpha = r(H) * 0.14 + 0.82
    beta = r(H) * 0.02 + 0.968
    a = r(H)
    b = r(H) * 2
    V = torch.nn.init.orthogonal_(torch.empty(H, H)).to(DEV)
    u0, w0, s0 = r(Be, H), r(Be, H), r(Be, H)

    def run(gs):
        leaves = [z.clone().requires_grad_(True) for z in (I, alpha, beta, a, b, V)]
        S = F.SpikingCellFunction.apply(leaves[0], None, None, *leaves[1:], u0, w0, s0, "RadLIF", 1.0,
                                        F.NormState("none"))
        S.backward(gs)
        return S.detach(), [z.grad for z in leaves]

    g1 = torch.randn(Be, T, H, device=DEV, generator=gen)
    g2 = torch.randn(Be, T, H, device=DEV, generator=gen)
    S1, A = run(g1)
    S2, Bq = run(g2)
    S3, C = run(g1 + 2 * g2)
    assert torch.equal(S1, S2) and torch.equal(S1, S3), "forward is not deterministic"
    assert 0.01 < float(S1.mean()) < 0.9
    for x, y, z in zip(A, Bq, C):
        ref = x.double() + 2 * y.double()
        err = float((z.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        assert err < 5e-5, err


def test_errors_and_contract():
    sp, F = _mods()
    with pytest.raises(ValueError):
        sp.SNN((4, None, 7), [8, 8, 3], neuron_type="GRU")
    net = sp.SNN((4, None, 7), [8, 8, 3], neuron_type="LIF")
    with pytest.raises(RuntimeError):
        net(torch.randn(4, 5, 7))  # CPU tensors: no fallback
    net4 = sp.SNN((4, None, 7, 2), [8, 3], neuron_type="LIF").to(DEV)
    with pytest.raises(NotImplementedError):
        net4(torch.randn(4, 5, 14, device=DEV))
    out, fr = net4(torch.randn(4, 5, 7, 2, device=DEV))
    assert out.shape == (4, 3) and fr.shape == (8,)
    # batch size may change between calls (snns.py:671-672); empty time axis is tolerated by the ABI
    net = net.to(DEV)
    o1, _ = net(torch.randn(4, 5, 7, device=DEV))
    o2, _ = net(torch.randn(9, 6, 7, device=DEV))
    assert o1.shape == (4, 3) and o2.shape == (9, 3)
    with pytest.raises(RuntimeError):
        from sparch_b200._lib import call
        call("sparch_cell_fwd", 7, None, None, None, None, None, None, None, None, None, None, 1.0,
             None, None, None, 1, 1, 1, None)


def test_eval_mode_and_dropout_train_mode():
    sp, _ = _mods()
    torch.manual_seed(0)
    net = sp.SNN((8, None, 12), [32, 32, 4], neuron_type="RadLIF", dropout=0.25).to(DEV)
    x = torch.randn(8, 10, 12, device=DEV)
    net.train()
    out, fr = net(x)
    vals = torch.unique(net.snn[0](x))
    # dropout output values are 0 or 1/(1-p) (snns.py:692)
    assert all(abs(float(v)) < 1e-6 or abs(float(v) - 1 / 0.75) < 1e-5 for v in vals)
    out.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    net.eval()
    torch.manual_seed(3)
    o1, _ = net(x)
    torch.manual_seed(3)
    o2, _ = net(x)
    assert torch.equal(o1, o2)


def _full_w_tape(W_gpu, U_gpu, s_gpu, u0, w0, s0, cp):
    """The adaptation tape a reverse pass reads.  When both tcgen05 recurrence kernels serve the layer the forward keeps w
    only at the end of every chunk of ``F.W_TAPE_EVERY`` steps (and at T-1): the steps in between follow from the
    membrane tape and the spikes by snns.py:718 in the kernel's own operation order -- which also checks the stored
    entries bit for bit."""
    from sparch_b200 import functional as F
    Be, T, H = U_gpu.shape
    if W_gpu.shape[1] == T:
        return W_gpu
    C = F.W_TAPE_EVERY
    assert W_gpu.shape[1] == (T + C - 1) // C
    beta, a, b = (np.asarray(cp[k], dtype=np.float32) for k in ("beta", "a", "b"))
    full = np.empty_like(U_gpu)
    w = w0.astype(np.float32)
    for t in range(T):
        up = u0 if t == 0 else U_gpu[:, t - 1]
        sp = s0 if t == 0 else s_gpu[:, t - 1]
        w = ((beta * w).astype(np.float32) + (a * up).astype(np.float32)).astype(np.float32) + (b * sp).astype(np.float32)
        w = w.astype(np.float32)
        full[:, t] = w
        if t % C == C - 1 or t == T - 1:
            assert np.array_equal(W_gpu[:, t // C], w), ("checkpoint tape", t)
    return full


def _oracle_cell_check(kind, Be, T, H, seed, drive=(3.0, 1.2), stable=True):
    """Free-running CUDA cell vs the numpy oracle on random inputs: spikes (flip fraction), then the
    given-mask backward when the spike trains agree exactly."""
    _, F = _mods()
    rng = np.random.default_rng(seed)
    adaptive, recurrent = orc.kind_flags(kind)
    I = (rng.standard_normal((Be, T, H)) * drive[0] + drive[1]).astype(np.float32)
    alpha = rng.uniform(0.80, 0.97, H).astype(np.float32)          # some outside the clamp window
    beta = rng.uniform(0.96, 0.995, H).astype(np.float32)
    a = rng.uniform(0.0 if stable else -1.2, 1.2, H).astype(np.float32)
    b = rng.uniform(-0.2, 2.2, H).astype(np.float32)
    V = (rng.standard_normal((H, H)) / np.sqrt(H)).astype(np.float32)
    u0, w0, s0 = (rng.uniform(0, 1, (Be, H)).astype(np.float32) for _ in range(3))
    p = orc.clamp_params(kind, alpha, beta, a, b)
    V0 = None
    if recurrent:
        V0 = V.copy()
        np.fill_diagonal(V0, 0)
    r = orc.cell_forward(kind, I, p["alpha"], p.get("beta"), p.get("a"), p.get("b"), V0, u0,
                         w0 if adaptive else None, s0)
    t = lambda z, g=False: torch.from_numpy(z).to(DEV).requires_grad_(g)
    It, al, be, aa, bb, Vt = t(I, True), t(alpha, True), t(beta, True), t(a, True), t(b, True), t(V, True)
    S = F.SpikingCellFunction.apply(It, None, None, al, be if adaptive else None, aa if adaptive else None,
                                    bb if adaptive else None, Vt if recurrent else None, t(u0),
                                    t(w0) if adaptive else None, t(s0), kind, 1.0, F.NormState("none"))
    s_gpu = S.detach().cpu().numpy()
    rate = float(r["s"].mean())
    flips = float((s_gpu != r["s"]).mean())
    assert Be * T * H < 1000 or 0.005 < rate < 0.95, rate
    assert flips <= (FLIP_TOL if recurrent else 0.0), (kind, Be, T, H, flips)
    # Given-mask backward on the CUDA forward's OWN tapes: with them the reverse pass is linear and well-posed even
    # when a membrane value straddles the threshold or an edge of the surrogate window by one ulp (at cfg3's 6.5 M
    # elements one does: identical spike trains, one differing window bit, 1 % error in dI through either reverse
    # kernel when the oracle used its own tape).  The tapes themselves are compared with the oracle's first.
    sv = S.grad_fn.saved_tensors          # (..., S, U, W, ...) as saved by SpikingCellFunction.forward
    U_gpu = sv[12].cpu().numpy()
    W_gpu = _full_w_tape(sv[13].cpu().numpy(), U_gpu, s_gpu, u0, w0, s0, p) if adaptive else None
    if flips == 0:
        assert rel_err(U_gpu, r["u"]) < U_RTOL
        if adaptive:
            assert rel_err(W_gpu, r["w"]) < U_RTOL
    gs = rng.standard_normal(s_gpu.shape).astype(np.float32)
    S.backward(torch.from_numpy(gs).to(DEV))
    bw = orc.cell_backward(kind, gs, I, p["alpha"], p.get("beta"), p.get("a"), p.get("b"), V0, u0,
                           w0 if adaptive else None, s0, U=U_gpu, W=W_gpu, S=s_gpu)
    tol = 5e-5  # longer chains than the fixtures: fp32 accumulation over T*Be terms
    assert rel_err(It.grad.cpu().numpy(), bw["dI"]) < tol
    assert rel_err(al.grad.cpu().numpy(), bw["dalpha"] * orc.clamp_grad_mask(alpha, orc.ALPHA_LIM)) < tol
    if adaptive:
        for k, g_, raw, lim in (("beta", be, beta, orc.BETA_LIM), ("a", aa, a, orc.A_LIM), ("b", bb, b, orc.B_LIM)):
            assert rel_err(g_.grad.cpu().numpy(), bw["d" + k] * orc.clamp_grad_mask(raw, lim)) < tol, k
    if recurrent:
        assert rel_err(Vt.grad.cpu().numpy(), bw["dV"]) < tol
    return flips


@pytest.mark.parametrize("kind,Be,T,H", [
    ("RadLIF", 300, 7, 96),     # 5 row groups: ragged last group
    ("RadLIF", 70, 5, 1000),    # H not a multiple of 32: padded spike words and V0 slices
    ("RLIF", 3, 1, 40),         # single timestep, tiny batch
    ("RadLIF", 640, 4, 512),    # 10 groups x 16 slices = 160 CTAs > 148: two cooperative launches
    ("RLIF", 33, 60, 130),      # longer chain, odd sizes
    ("adLIF", 65, 33, 77),      # streaming kernels, odd sizes, T not a multiple of the prefetch depth
    ("LIF", 1, 1, 1),
    ("RadLIF", 6, 4, 1440),     # beyond the resident-V0 limit: stepwise general path
    ("RLIF", 40, 3, 1184),      # the largest hidden size the persistent kernels hold
])
def test_cell_edge_shapes_against_oracle(kind, Be, T, H):
    _oracle_cell_check(kind, Be, T, H, seed=Be + T + H)


@pytest.mark.parametrize("mode", ["tc", "mma"])
@pytest.mark.parametrize("kind,Be,T,H", [
    ("RadLIF", 300, 7, 96),     # ragged last 64-row group, hidden size padded to the K quarters
    ("RadLIF", 70, 5, 1000),    # H not a multiple of 32
    ("RLIF", 3, 1, 40),         # single timestep: no UMMA at all
    ("RLIF", 33, 60, 130),      # longer chain, odd sizes
    ("RadLIF", 640, 4, 512),    # 10 groups x 16 slices: two cooperative launches
    ("RadLIF", 256, 12, 1024),  # the cfg4 layer width
])
def test_both_reverse_recurrence_kernels_against_oracle(kind, Be, T, H, mode):
    """The tcgen05 reverse-recurrence kernel (sparch_recur_bwd_tc, the default up to H = 1024) and the
    mma.sync one (sparch_recur_bwd) meet the same bar."""
    _, F = _mods()
    old = F.RECUR_BWD
    F.RECUR_BWD = mode
    try:
        n0 = F.native_launches()
        _oracle_cell_check(kind, Be, T, H, seed=Be + T + H)
        assert F.native_launches() > n0
    finally:
        F.RECUR_BWD = old


@pytest.mark.parametrize("mode", ["tc", "mma"])
@pytest.mark.parametrize("kind,Be,T,H", [
    ("RadLIF", 300, 7, 96),     # 3 row groups of 128, ragged last one; 6 slices: one partial batch of producers
    ("RadLIF", 70, 5, 1000),    # H not a multiple of 16: dead neurons in the last slice, K padded to 1024
    ("RLIF", 3, 1, 40),         # single timestep: no UMMA at all
    ("RLIF", 33, 60, 130),      # longer chain, odd sizes
    ("RadLIF", 640, 4, 512),    # 5 row groups x 32 slices = 160 CTAs > 148: two cooperative launches
    ("RadLIF", 256, 12, 1024),  # the cfg4 layer width: 4 batches of 16 producers
    ("RLIF", 130, 6, 1100),     # 5 batches, the last one partial (69 slices)
])
def test_both_forward_recurrence_kernels_against_oracle(kind, Be, T, H, mode):
    """The tcgen05 forward recurrence (sparch_recur_fwd_tc: int8 digit planes of V0, spike operand in tensor memory;
    the default up to H = 1536) and the mma.sync one (sparch_recur_fwd) meet the same bar: spike trains against the
    oracle (<= 1e-3 flips, observed 0) and, where they coincide, the gradients of the given-mask backward."""
    _, F = _mods()
    old = F.RECUR_FWD
    F.RECUR_FWD = mode
    try:
        n0 = F.native_launches()
        _oracle_cell_check(kind, Be, T, H, seed=Be + T + H)
        assert F.native_launches() > n0
    finally:
        F.RECUR_FWD = old


def test_empty_batch_and_time_are_tolerated():
    _, F = _mods()
    for Be, T in ((0, 5), (4, 0)):
        Z = torch.zeros(Be, T, 8, device=DEV)
        al = torch.full((8,), 0.9, device=DEV)
        S = F.SpikingCellFunction.apply(Z, None, None, al, None, None, None, None,
                                        torch.zeros(Be, 8, device=DEV), None, torch.zeros(Be, 8, device=DEV),
                                        "LIF", 1.0, F.NormState("none"))
        assert S.shape == (Be, T, 8)


@pytest.mark.parametrize("kind", ["LIF", "adLIF", "RLIF", "RadLIF"])
def test_single_step_abi_against_oracle(kind):
    """sparch_cell_step_fwd / sparch_cell_step_bwd (one timestep, caller supplies s_{t-1} @ V0):
    drive the time loop from Python and compare tapes and gradients with the oracle."""
    from sparch_b200._lib import call, ptr
    from sparch_b200.functional import KINDS
    rng = np.random.default_rng(5)
    Be, T, H = 5, 9, 24
    adaptive, recurrent = orc.kind_flags(kind)
    k = KINDS[kind]
    I = (rng.standard_normal((Be, T, H)) * 3 + 1).astype(np.float32)
    alpha = rng.uniform(0.82, 0.96, H).astype(np.float32)
    beta = rng.uniform(0.97, 0.99, H).astype(np.float32)
    a = rng.uniform(0, 1, H).astype(np.float32)
    b = rng.uniform(0, 2, H).astype(np.float32)
    V0 = (rng.standard_normal((H, H)) / 5).astype(np.float32)
    np.fill_diagonal(V0, 0)
    u0, w0, s0 = (rng.uniform(0, 1, (Be, H)).astype(np.float32) for _ in range(3))
    r = orc.cell_forward(kind, I, alpha, beta, a, b, V0 if recurrent else None, u0, w0 if adaptive else None, s0)
    t_ = lambda z: torch.from_numpy(np.ascontiguousarray(z)).to(DEV)
    It, al, be, aa, bb, Vt, u0t, w0t, s0t = map(t_, (I, alpha, beta, a, b, V0, u0, w0, s0))
    S, U, W = torch.empty_like(It), torch.empty_like(It), torch.empty_like(It)
    st = torch.cuda.current_stream().cuda_stream
    on = lambda z, flag: ptr(z) if flag else None
    for t in range(T):
        rec = (s0t if t == 0 else S[:, t - 1]) @ Vt if recurrent else None
        call("sparch_cell_step_fwd", k, t, ptr(It), None, None, ptr(al), on(be, adaptive), on(aa, adaptive),
             on(bb, adaptive), ptr(rec), ptr(u0t), on(w0t, adaptive), ptr(s0t), 1.0, ptr(S), ptr(U),
             on(W, adaptive), Be, T, H, st)
    np.testing.assert_array_equal(S.cpu().numpy(), r["s"])
    assert rel_err(U.cpu().numpy(), r["u"]) < 1e-5
    gs = rng.standard_normal((Be, T, H)).astype(np.float32)
    G = t_(gs)
    dI = torch.empty_like(It)
    carry = torch.zeros(2, Be, H, device=DEV)
    part = torch.zeros(4, Be, H, device=DEV)
    for t in range(T - 1, -1, -1):
        recb = dI[:, t + 1] @ Vt.t() if (recurrent and t < T - 1) else None
        call("sparch_cell_step_bwd", k, t, ptr(G), ptr(U), on(W, adaptive), ptr(al), on(be, adaptive),
             on(aa, adaptive), on(bb, adaptive), ptr(recb), ptr(u0t), on(w0t, adaptive), ptr(s0t), 1.0,
             ptr(dI), ptr(carry[0]), on(carry[1], adaptive), ptr(part[0]), on(part[1], adaptive),
             on(part[2], adaptive), on(part[3], adaptive), Be, T, H, st)
    bw = orc.cell_backward(kind, gs, I, alpha, beta, a, b, V0 if recurrent else None, u0,
                           w0 if adaptive else None, s0, U=r["u"], W=r["w"], S=r["s"])
    assert rel_err(dI.cpu().numpy(), bw["dI"]) < G_RTOL
    assert rel_err(part[0].sum(0).cpu().numpy(), bw["dalpha"]) < G_RTOL
    if adaptive:
        for i, kk in enumerate(("dbeta", "da", "db")):
            assert rel_err(part[i + 1].sum(0).cpu().numpy(), bw[kk]) < G_RTOL, kk


def test_graphed_train_step_matches_eager_statistics():
    """GraphedTrainStep (forward + loss + backward + Adam in one CUDA graph, cooperative kernels
    included) trains like the eager step: same first loss up to the random initial states, finite
    and decreasing afterwards, parameters updated on replay."""
    import sparch_b200
    from sparch_b200.graphs import GraphedTrainStep
    sparch_b200.set_state_init("device")
    try:
        torch.manual_seed(0)
        kw = dict(layer_sizes=[64, 64, 5], neuron_type="RadLIF", dropout=0.1)
        net = sparch_b200.SNN((16, None, 12), **kw).to(DEV)
        ref = sparch_b200.SNN((16, None, 12), **kw).to(DEV)
        ref.load_state_dict(net.state_dict())
        x = torch.randn(16, 20, 12, device=DEV)
        y = torch.randint(0, 5, (16,), device=DEV)
        loss_fn = torch.nn.CrossEntropyLoss()
        out, _ = ref(x)
        eager_first = float(loss_fn(out, y))
        opt = torch.optim.Adam(net.parameters(), 1e-2, capturable=True)
        g = GraphedTrainStep(net, opt, loss_fn, x, y, warmup=0 + 1)
        w_before = net.snn[0].W.weight.detach().clone()
        losses = [float(g.step(x, y)) for _ in range(40)]
        assert all(np.isfinite(losses))
        assert abs(losses[0] - eager_first) < 0.5 * eager_first      # same model, different random states
        assert np.mean(losses[-5:]) < np.mean(losses[:5])            # it trains
        assert not torch.equal(w_before, net.snn[0].W.weight)        # replays update the parameters
        assert g.native_calls_per_step > 10
    finally:
        sparch_b200.set_state_init("cpu")


def test_reduced_precision_mode():
    """set_precision("bf16"): one bf16 term per GEMM operand, hi terms only in the recurrence.
    Stated tolerance: teacher-forced single steps flip <= 2e-3 of the spikes; free-running short
    trains flip <= 2e-2; where the trains coincide the gradients agree to 3e-2 relative L2; a whole
    model's loss stays within 5 % of the fp32 mode."""
    import sparch_b200
    from tests.helpers import rel_l2
    g = Golden("radlif_bn_h64")
    net = _net(g)
    try:
        sparch_b200.set_precision("bf16")
        for i, lay, I, u0, w0, s0 in _hidden_layer_inputs(g, net):
            p, V0 = _layer_params(lay)
            r = orc.cell_forward(lay._kind, I, p["alpha"], p.get("beta"), p.get("a"), p.get("b"), V0, u0, w0, s0,
                                 theta=lay.threshold)
            Be, T, H = I.shape
            prev = lambda X, x0: np.concatenate([x0[:, None, :], X[:, :-1, :]], axis=1).reshape(Be * T, H)
            _, S1 = _run_cell(lay, I.reshape(Be * T, 1, H), prev(r["u"], u0), prev(r["w"], w0), prev(r["s"], s0))
            flips1 = float((S1.detach().cpu().numpy().reshape(Be * T, H) != r["s"].reshape(Be * T, H)).mean())
            assert flips1 <= 2e-3, (i, flips1)
            for q in (lay.alpha, lay.beta, lay.a, lay.b, lay.V.weight):
                q.grad = None
            It, S = _run_cell(lay, I, u0, w0, s0, need_grad=True)
            flips = float((S.detach().cpu().numpy() != r["s"]).mean())
            assert flips <= 2e-2, (i, flips)
            if flips == 0:
                gs = np.random.default_rng(1).standard_normal(r["s"].shape).astype(np.float32)
                S.backward(torch.from_numpy(gs).to(DEV))
                bw = orc.cell_backward(lay._kind, gs, I, p["alpha"], p["beta"], p["a"], p["b"], V0, u0, w0, s0,
                                       theta=lay.threshold, U=r["u"], W=r["w"], S=r["s"])
                assert rel_l2(It.grad.cpu().numpy(), bw["dI"]) < 3e-2
                assert rel_l2(lay.V.weight.grad.cpu().numpy(), bw["dV"]) < 3e-2
        # whole model: same weights and seeds in both modes
        losses = {}
        for mode in ("fp32", "bf16"):
            sparch_b200.set_precision(mode)
            m = _net(g)
            torch.manual_seed(42)
            out, _ = m(g.t("x", DEV))
            loss = g.loss_fn(out, g.t("y", DEV))
            loss.backward()
            assert all(torch.isfinite(q.grad).all() for q in m.parameters())
            losses[mode] = float(loss)
        assert abs(losses["bf16"] - losses["fp32"]) <= 0.05 * abs(losses["fp32"]), losses
    finally:
        sparch_b200.set_precision("fp32")


def test_full_size_cfg4_layer_against_oracle_on_gpu():
    """BASELINE.json's headline shape (RadLIF, B=256, T=100, H=1024, F=40): the first layer's spike
    train against the oracle's torch restatement run on the same device (stable draw a <- |a|), the
    forward is bit-reproducible run to run, and the loss of the whole model agrees with the oracle."""
    sp, _ = _mods()
    kw = dict(layer_sizes=[1024, 1024, 35], neuron_type="RadLIF", normalization="batchnorm")
    torch.manual_seed(0)
    net = sp.SNN((256, None, 40), **kw)
    ref = orc.build_oracle_snn((256, None, 40), **kw)
    ref.load_state_dict(net.state_dict())
    for m in (net, ref):
        with torch.no_grad():
            for lay in m.snn:
                if hasattr(lay, "a"):
                    lay.a.abs_()
                if isinstance(getattr(lay, "norm", None), torch.nn.BatchNorm1d):
                    lay.norm.weight.fill_(3.0)   # default gamma=1, beta=0 leaves the net almost silent
                    lay.norm.bias.fill_(0.8)
    net, ref = net.to(DEV), ref.to(DEV)
    ref.snn[0].capture = {}
    torch.manual_seed(1234)
    x = torch.randn(256, 100, 40, device=DEV)
    y = torch.randint(0, 35, (256,), device=DEV)
    got = {}
    h = net.snn[0].register_forward_hook(lambda m, i, o: got.__setitem__(0, o.detach()))
    torch.manual_seed(42)
    out, rates = net(x)
    first = got[0].clone()
    torch.manual_seed(42)
    out2, _ = net(x)
    h.remove()
    assert torch.equal(first, got[0]) and torch.equal(out, out2), "forward is not reproducible"
    torch.manual_seed(42)
    out_r, rates_r = ref(x)
    s_ref = ref.snn[0].capture["s"]
    flips = float((first != s_ref).float().mean())
    assert 0.005 < float(s_ref.mean()) < 0.9
    assert flips <= FLIP_TOL, flips
    loss = torch.nn.functional.cross_entropy(out, y)
    loss_r = torch.nn.functional.cross_entropy(out_r, y)
    # a handful of threshold straddlers reshuffle later spikes: compare at 2 %
    assert abs(float(loss) - float(loss_r)) <= 0.02 * abs(float(loss_r)), (float(loss), float(loss_r))
    assert float((rates - rates_r).abs().max()) < 0.02


@pytest.mark.parametrize("seed", range(10))
def test_recurrent_cells_random_shape_sweep(seed):
    """Seeded random shapes (batch not a multiple of the 32-row teams, hidden size not a multiple of the
    32-neuron slices / 128-neuron K-quarters, 1..7 steps) for the persistent recurrent kernels."""
    rng = np.random.default_rng(1000 + seed)
    kind = ("RLIF", "RadLIF")[seed % 2]
    Be = int(rng.choice([1, 2, 31, 32, 33, 63, 64, 65, 100, 257]))
    H = int(rng.choice([1, 8, 31, 32, 33, 48, 96, 100, 130, 260, 512]))
    T = int(rng.integers(1, 8))
    _oracle_cell_check(kind, Be, T, H, seed=seed, drive=(3.0, 1.5))


@pytest.mark.parametrize("B,T,C", [(5, 100, 35), (3, 700, 20), (2, 23, 300), (4, 1, 7), (2, 40, 1024)])
def test_readout_cell_against_oracle(B, T, C):
    """ReadoutLayer cell (snns.py:807-825) through the C ABI: chunked scan + warp-per-step softmax,
    including shapes whose T exceeds one shared-memory chunk and class counts above one warp."""
    _, F = _mods()
    rng = np.random.default_rng(B * 1000 + T + C)
    I = rng.standard_normal((B, T, C)).astype(np.float32) * 2
    alpha = rng.uniform(0.82, 0.96, C).astype(np.float32)
    u0 = rng.uniform(0, 1, (B, C)).astype(np.float32)
    gout = rng.standard_normal((B, C)).astype(np.float32)
    It = torch.from_numpy(I).to(DEV).requires_grad_(True)
    al = torch.from_numpy(alpha).to(DEV).requires_grad_(True)
    out = F.ReadoutCellFunction.apply(It, None, None, al, torch.from_numpy(u0).to(DEV), F.NormState("none"))
    out.backward(torch.from_numpy(gout).to(DEV))
    f = orc.readout_forward(I, alpha, u0)
    bw = orc.readout_backward(gout, I, alpha, u0, f["u"])
    assert rel_err(out.detach().cpu().numpy(), f["out"]) < 1e-5
    assert rel_err(It.grad.cpu().numpy(), bw["dI"]) < G_RTOL
    assert rel_err(al.grad.cpu().numpy(), bw["dalpha"]) < 5e-5


@pytest.mark.parametrize("Be,T,H,p", [(4, 25, 1024, 0.25), (3, 7, 37, 0.5), (2, 5, 256, 0.1), (5, 9, 40, 0.0)])
def test_spike_post_pass(Be, T, H, p):
    """Fused dropout + counts + operand terms (csrc/post.cu): values, mask statistics, the mask the backward
    regenerates, the row maxima it leaves, and the differentiable firing rates (snns.py:174, 692)."""
    _, F = _mods()
    from sparch_b200 import gemm
    g = torch.Generator(device=DEV).manual_seed(Be * H)
    S = (torch.rand(Be, T, H, device=DEV, generator=g) < 0.3).float().requires_grad_(True)
    st = F.NormState("none")
    torch.manual_seed(7)
    out, post = F.spike_post(S, p, st, recurrent=True)
    scale = 1.0 / (1.0 - p)
    o = out.detach()
    assert torch.all((o == 0) | ((o - scale).abs() < 1e-6))
    assert torch.all(o[S.detach() == 0] == 0)
    nspk = float(S.detach().sum())
    kept = float((o != 0).sum()) / nspk
    assert abs(kept - (1 - p)) < 4 * (p * (1 - p) / nspk) ** 0.5 + 1e-9, kept
    ld = (H + 7) // 8 * 8
    one = post.terms.parts.dtype
    assert post.terms.parts.shape == (1, Be * T, ld) and one == (torch.float16 if gemm.MODE == "f16x2" else torch.bfloat16)
    assert torch.equal(post.terms.parts[0, :, :H].float().view(Be, T, H), (o != 0).float())
    assert float(post.terms.parts[0, :, H:].abs().sum()) == 0
    assert torch.equal(st.sterm.parts[0, :, :H].float().view(Be, T, H), S.detach())
    assert torch.equal(post.counts.long(), (o != 0).sum(dim=(0, 1)))
    rates = post.rates(out)
    assert rel_err(rates.detach().cpu().numpy(), o.mean(dim=(0, 1)).cpu().numpy()) < 1e-6
    gup = torch.randn(Be, T, H, device=DEV, generator=g)
    grate = torch.randn(H, device=DEV, generator=g)
    ((out * gup).sum() + (rates * grate).sum()).backward()
    gin = gup + grate / (Be * T)
    mask = (o != 0).float() if p > 0 else torch.ones_like(o)
    # positions where S == 0 carry no information about the mask in the forward output: compare where S == 1
    sel = S.detach() == 1
    want = gin * mask * scale
    assert rel_err(S.grad[sel].cpu().numpy(), want[sel].cpu().numpy()) < 1e-6
    if p > 0:
        zero_frac = float((S.grad[~sel] == 0).float().mean())
        assert abs(zero_frac - p) < 4 * (p * (1 - p) / float((~sel).sum())) ** 0.5, zero_frac
        assert st.gmax is not None
        np.testing.assert_allclose(st.gmax.view(Be, T).cpu().numpy(), S.grad.abs().amax(dim=2).cpu().numpy(), rtol=0, atol=0)
        # a second draw uses a new seed
        out2, _ = F.spike_post(S, p, F.NormState("none"), recurrent=False)
        assert not torch.equal(out2, out)


def test_param_clamp_and_grad_kernels():
    _, F = _mods()
    from sparch_b200._lib import call, ptr
    H, Be = 77, 19
    g = torch.Generator(device=DEV).manual_seed(4)
    ps = [torch.rand(H, device=DEV, generator=g) * 1.4 - 0.2 for _ in range(4)]
    ps[2] = ps[2] * 3 - 1.5
    ps[0][3] = F.ALPHA_LIM[0]        # exactly on the boundary: gradient passes (closed interval)
    out = torch.empty(4, H, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    call("sparch_neuron_params", ptr(ps[0]), ptr(ps[1]), ptr(ps[2]), ptr(ps[3]), F._LIMS, 4, H, ptr(out), st)
    lims = (F.ALPHA_LIM, F.BETA_LIM, F.A_LIM, F.B_LIM)
    for k in range(4):
        assert torch.equal(out[k], ps[k].clamp(*lims[k]))
    part = torch.randn(4, Be, H, device=DEV, generator=g)
    grads = torch.empty(4, H, device=DEV)
    call("sparch_param_grads", ptr(part), ptr(ps[0]), ptr(ps[1]), ptr(ps[2]), ptr(ps[3]), F._LIMS, 4, Be, H, ptr(grads), st)
    for k in range(4):
        ref = orc.clamp_grad_mask(ps[k].cpu().numpy(), lims[k]) * part[k].double().sum(0).cpu().numpy()
        assert rel_err(grads[k].cpu().numpy(), ref) < 1e-6
    assert float(grads[0][3]) != 0.0


def test_single_launch_adam_matches_torch():
    """sparch_b200.optim.Adam (one launch for all tensors, device step count) against torch.optim.Adam."""
    from sparch_b200.optim import Adam
    g = torch.Generator(device=DEV).manual_seed(11)
    shapes = [(1024, 40), (1024,), (35, 1024), (7,), (1, 1), (300, 257)]
    pa = [torch.randn(*s, device=DEV, generator=g).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa, ob = Adam(pa, 1e-2), torch.optim.Adam(pb, 1e-2)
    for it in range(5):
        for x, y in zip(pa, pb):
            gr = torch.randn(*x.shape, device=DEV, generator=g) * (10.0 ** (it - 2))
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    for x, y in zip(pa, pb):
        assert rel_err(x.detach().cpu().numpy(), y.detach().cpu().numpy()) < 2e-6
    cpu_p = torch.zeros(3, requires_grad=True)
    cpu_p.grad = torch.ones(3)
    with pytest.raises(RuntimeError):
        Adam([cpu_p], 1e-2).step()


def test_adam_lr_change_reaches_graph_replays_and_step_round_trips():
    """exp.py:92-96 drives the learning rate with ReduceLROnPlateau, which edits param_groups[i]["lr"].  The captured
    Adam launch reads lr from a device word, so the change takes effect on the next replay (a by-value scalar would be
    frozen into the graph); the step count lives in optimizer.state and survives state_dict() / load_state_dict()."""
    from sparch_b200.optim import Adam
    g = torch.Generator(device=DEV).manual_seed(5)
    pa = [torch.randn(64, 33, device=DEV, generator=g).requires_grad_(True), torch.randn(17, device=DEV, generator=g).requires_grad_(True)]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    grads = [torch.randn(*p.shape, device=DEV, generator=g) for p in pa]
    for x, y, gr in zip(pa, pb, grads):
        x.grad, y.grad = gr.clone(), gr.clone()
    oa, ob = Adam(pa, 1e-2), torch.optim.Adam(pb, 1e-2)
    oa.step(); ob.step()                       # eager warm-up step (creates the state and the device words)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        oa.step(); ob.step()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(graph):
        oa.step()
    ob.step()                                  # the captured launch did not run: torch is one step ahead on purpose
    graph.replay()                             # ... and now level again (3 steps each)
    for lr in (1e-3, 5e-2):
        for grp in oa.param_groups + ob.param_groups:
            grp["lr"] = lr
        before = pa[0].detach().clone()
        oa.sync_hyper()                        # what GraphedTrainStep.step() does before every replay
        graph.replay()
        ob.step()
        upd = float((pa[0].detach() - before).abs().max())
        assert 0.2 * lr < upd < 5 * lr, (lr, upd)          # the update magnitude follows the new learning rate
        for x, y in zip(pa, pb):
            assert rel_err(x.detach().cpu().numpy(), y.detach().cpu().numpy()) < 2e-6
    sd = oa.state_dict()
    assert int(sd["state"][0]["step"]) == 5
    oc = Adam(pa, 1e-2)
    oc.load_state_dict(sd)
    for grp in oc.param_groups:
        grp["lr"] = 5e-2
    oc.step(); ob.step()                       # resumes at step 6 with the right bias correction
    assert int(oc.state[pa[0]]["step"]) == 6
    for x, y in zip(pa, pb):
        assert rel_err(x.detach().cpu().numpy(), y.detach().cpu().numpy()) < 2e-6


@pytest.mark.parametrize("name", ["radlif_bn", "lif_bn"])
def test_three_bf16_term_mode_against_reference_fixture(name):
    """set_precision("fp32-bf16x3") -- the first operand scheme, kept for comparison -- meets the same fixture
    tolerances as the default two-term fp16 scheme (the post pass and BatchNorm backward take their bf16 paths)."""
    import sparch_b200
    try:
        sparch_b200.set_precision("fp32-bf16x3")
        test_whole_model_against_reference_fixture(name)
    finally:
        sparch_b200.set_precision("fp32")


def test_reference_checkpoint_runs_on_the_gpu():
    """The whole-module checkpoint written by the reference (tests/golden/reference_module.pt) unpickled into these
    classes, moved to the GPU and run: outputs and firing rates as the reference recorded them."""
    from tests.test_host_contract import _load_reference_checkpoint
    import os
    from tests.helpers import GOLDEN_DIR
    net = _load_reference_checkpoint().to(DEV)
    run = np.load(os.path.join(GOLDEN_DIR, "reference_module_run.npz"))
    torch.manual_seed(42)
    out, rates = net(torch.from_numpy(run["x"]).to(DEV))
    np.testing.assert_allclose(out.detach().cpu().numpy(), run["out"], rtol=5e-5, atol=2e-6)
    np.testing.assert_allclose(rates.detach().cpu().numpy(), run["rates"], rtol=1e-6, atol=1e-7)


# ---------------------------------------------------------------------------------------------------------------
# Round-2 regression net ("What's weak" 1-4): headline-shape gradients, the reference's default
# (unstable) adaptation draw, dropout > 0 at value level, the cfg3 and cfg5 shapes.
# ---------------------------------------------------------------------------------------------------------------

def _bwd_tc_direct(kind, G, U, W, cl, V, u0, w0, s0, theta=1.0):
    """sparch_recur_bwd_tc through the C ABI on GIVEN tapes (the oracle's): returns dI and the batch-summed raw
    parameter gradients.  The reverse pass is linear given the tapes, so it is comparable to the fp64 oracle whatever
    the forward dynamics are -- including the reference's default draw a ~ U(-1, 1), where free-running trains of two
    implementations diverge."""
    from sparch_b200._lib import call, lib, ptr
    from sparch_b200.functional import KINDS
    L = lib()
    Be, T, H = G.shape
    adaptive = KINDS[kind] & 1
    st = torch.cuda.current_stream().cuda_stream
    meta = torch.empty(2, device=DEV, dtype=torch.int32)
    call("sparch_recur_prepare", ptr(V), H, None, None, ptr(meta), st)
    img = torch.empty(L.sparch_recur_bwd_tc_image_bytes(H), device=DEV, dtype=torch.uint8)
    call("sparch_recur_prepare_tc", ptr(V), H, ptr(img), ptr(meta), st)
    ws = torch.empty(L.sparch_recur_bwd_tc_workspace(Be, T, H), device=DEV, dtype=torch.uint8)
    dI = torch.empty_like(G)
    part = torch.zeros(4, Be, H, device=DEV)
    on = lambda z: ptr(z) if adaptive else None
    call("sparch_recur_bwd_tc", KINDS[kind], ptr(G), ptr(U), on(W), ptr(cl["alpha"]), on(cl.get("beta")),
         on(cl.get("a")), on(cl.get("b")), ptr(img), ptr(meta), ptr(u0), on(w0), ptr(s0), theta, ptr(dI), ptr(part[0]),
         on(part[1]), on(part[2]), on(part[3]), ptr(ws), 0, Be, T, H, None, st)
    torch.cuda.synchronize()
    return dI, part.sum(1)


@pytest.mark.parametrize("kind,Be,T,H,stable", [
    ("RadLIF", 256, 100, 1024, False),   # BASELINE cfg4's layer, the reference's default a ~ U(-1, 1)
    ("RadLIF", 256, 100, 1024, True),
    ("RLIF", 128, 100, 512, True),       # cfg3's layer
])
def test_given_tape_bptt_at_headline_shapes(kind, Be, T, H, stable):
    """Gradient parity at the benchmark shapes through the tcgen05 reverse kernel: oracle forward (fp32, numpy) ->
    tapes -> CUDA reverse pass on those tapes vs the fp64 oracle.  T = 100 is where the kernel's one-step-late
    per-row scale meets growing adjoints (48 % of the default-init neurons are linearly unstable).  Tolerances:
    5e-5 of the tensor maximum; dI additionally 2e-4 of its own (row, step) maximum for every row."""
    rng = np.random.default_rng(Be + T + H + int(stable))
    adaptive, recurrent = orc.kind_flags(kind)
    I = (rng.standard_normal((Be, T, H)) * 3.0 + 1.2).astype(np.float32)
    alpha = rng.uniform(np.exp(-1 / 5), np.exp(-1 / 25), H).astype(np.float32)      # snns.py:229, 644-647
    beta = rng.uniform(np.exp(-1 / 30), np.exp(-1 / 120), H).astype(np.float32)
    a = rng.uniform(0.0 if stable else -1.0, 1.0, H).astype(np.float32)
    b = rng.uniform(0.0, 2.0, H).astype(np.float32)
    V0 = torch.nn.init.orthogonal_(torch.empty(H, H), generator=torch.Generator().manual_seed(3)).numpy().copy()
    np.fill_diagonal(V0, 0)
    u0, w0, s0 = (rng.uniform(0, 1, (Be, H)).astype(np.float32) for _ in range(3))
    r = orc.cell_forward(kind, I, alpha, beta, a, b, V0, u0, w0 if adaptive else None, s0)
    assert np.isfinite(r["u"]).all()
    if not stable:
        assert np.abs(r["u"]).max() > 1e3, "the unstable draw did not grow: the test would not cover it"
    gs = rng.standard_normal((Be, T, H)).astype(np.float32)
    bw = orc.cell_backward(kind, gs, I, alpha, beta, a, b, V0, u0, w0 if adaptive else None, s0,
                           U=r["u"], W=r["w"], S=r["s"])
    t_ = lambda z: None if z is None else torch.from_numpy(np.ascontiguousarray(z)).to(DEV)
    cl = {"alpha": t_(alpha), "beta": t_(beta), "a": t_(a), "b": t_(b)}
    dI, pg = _bwd_tc_direct(kind, t_(gs), t_(r["u"]), t_(r["w"]) if adaptive else None, cl, t_(V0), t_(u0),
                            t_(w0), t_(s0))
    dI = dI.cpu().numpy().astype(np.float64)
    assert np.isfinite(dI).all()
    assert rel_err(dI, bw["dI"]) < 5e-5
    row_max = np.abs(bw["dI"]).max(axis=2, keepdims=True)
    row_err = (np.abs(dI - bw["dI"]) / np.maximum(row_max, 1e-300)).max()
    assert row_err < 2e-4, row_err
    names = ("dalpha", "dbeta", "da", "db") if adaptive else ("dalpha",)
    for i, k in enumerate(names):
        assert rel_err(pg[i].cpu().numpy(), bw[k]) < 5e-5, k


@pytest.mark.parametrize("kind,Be,T,H", [("RadLIF", 70, 16, 1000), ("RadLIF", 300, 7, 96), ("adLIF", 65, 33, 77),
                                         ("RadLIF", 256, 12, 1024), ("RadLIF", 33, 40, 130)])
def test_cells_with_the_default_unstable_adaptation_draw(kind, Be, T, H):
    """`stable=False`: a ~ U(-1.2, 1.2) as the reference's default init draws it (snns.py:647).  Spike trains are
    compared as everywhere; the gradients whenever the trains coincide (they do at these lengths)."""
    _oracle_cell_check(kind, Be, T, H, seed=7 * Be + T + H, stable=False)


@pytest.mark.parametrize("seed", range(6))
def test_recurrent_cells_random_shape_sweep_unstable(seed):
    rng = np.random.default_rng(2000 + seed)
    kind = ("RLIF", "RadLIF", "RadLIF")[seed % 3]
    Be = int(rng.choice([2, 33, 64, 129, 200]))
    H = int(rng.choice([16, 33, 48, 100, 260, 512]))
    T = int(rng.integers(2, 12))
    _oracle_cell_check(kind, Be, T, H, seed=seed, drive=(3.0, 1.5), stable=False)


def test_rlif_cfg3_layer_free_running_against_oracle():
    """BASELINE cfg3's layer shape (RLIF, B = 128, T = 100, H = 512) free-running against the numpy oracle, then
    the gradients of the given-mask backward."""
    _oracle_cell_check("RLIF", 128, 100, 512, seed=3, drive=(3.0, 1.2))


@pytest.mark.parametrize("kind,Be,T,H,p", [("RadLIF", 64, 20, 256, 0.1), ("adLIF", 32, 25, 96, 0.25),
                                           ("RadLIF", 256, 12, 1024, 0.1)])
def test_dropout_gradients_at_value_level(kind, Be, T, H, p):
    """cfg4 trains with dropout 0.1 (snns.py:692).  The mask is not stored (Philox keyed by a seed word and the
    element position), so it is read back by running the same post pass with the same seed over a tensor of ones;
    the oracle then applies out = S * mask / (1 - p) and its adjoint, and every gradient of the cell is compared."""
    _, F = _mods()
    rng = np.random.default_rng(Be + T + H)
    adaptive, recurrent = orc.kind_flags(kind)
    I = (rng.standard_normal((Be, T, H)) * 3.0 + 1.2).astype(np.float32)
    alpha = rng.uniform(0.80, 0.97, H).astype(np.float32)
    beta = rng.uniform(0.96, 0.995, H).astype(np.float32)
    a = rng.uniform(0.0, 1.2, H).astype(np.float32)
    b = rng.uniform(-0.2, 2.2, H).astype(np.float32)
    V = (rng.standard_normal((H, H)) / np.sqrt(H)).astype(np.float32)
    u0, w0, s0 = (rng.uniform(0, 1, (Be, H)).astype(np.float32) for _ in range(3))
    cp = orc.clamp_params(kind, alpha, beta, a, b)
    V0 = None
    if recurrent:
        V0 = V.copy()
        np.fill_diagonal(V0, 0)
    r = orc.cell_forward(kind, I, cp["alpha"], cp.get("beta"), cp.get("a"), cp.get("b"), V0, u0,
                         w0 if adaptive else None, s0)
    t = lambda z, g=False: torch.from_numpy(z).to(DEV).requires_grad_(g)
    It, al, be, aa, bb, Vt = t(I, True), t(alpha, True), t(beta, True), t(a, True), t(b, True), t(V, True)
    norm = F.NormState("none")
    S = F.SpikingCellFunction.apply(It, None, None, al, be if adaptive else None, aa if adaptive else None,
                                    bb if adaptive else None, Vt if recurrent else None, t(u0),
                                    t(w0) if adaptive else None, t(s0), kind, 1.0, norm)
    assert float((S.detach().cpu() != torch.from_numpy(r["s"])).float().mean()) <= (FLIP_TOL if recurrent else 0.0)
    sv = S.grad_fn.saved_tensors          # the CUDA forward's own tapes (see _oracle_cell_check)
    s_gpu, U_gpu = S.detach().cpu().numpy(), sv[12].cpu().numpy()
    W_gpu = _full_w_tape(sv[13].cpu().numpy(), U_gpu, s_gpu, u0, w0, s0, cp) if adaptive else None
    torch.manual_seed(99)
    out, post = F.spike_post(S, p, norm, recurrent)
    torch.manual_seed(99)                                   # same seed word -> same mask
    ones_out, _ = F.spike_post(torch.ones_like(S), p, F.NormState("none"), recurrent)
    mask = (ones_out.detach() != 0).cpu().numpy().astype(np.float64)
    keep = mask.mean()
    assert abs(keep - (1 - p)) < 0.01
    np.testing.assert_allclose(out.detach().cpu().numpy(), s_gpu * mask / (1 - p), rtol=1e-6, atol=0)
    g_out = rng.standard_normal(S.shape).astype(np.float32)
    out.backward(torch.from_numpy(g_out).to(DEV))
    gs = (g_out.astype(np.float64) * mask / (1 - p)).astype(np.float32)      # dropout's adjoint, as ATen does it
    bw = orc.cell_backward(kind, gs, I, cp["alpha"], cp.get("beta"), cp.get("a"), cp.get("b"), V0, u0,
                           w0 if adaptive else None, s0, U=U_gpu, W=W_gpu, S=s_gpu)
    tol = 5e-5
    assert rel_err(It.grad.cpu().numpy(), bw["dI"]) < tol
    assert rel_err(al.grad.cpu().numpy(), bw["dalpha"] * orc.clamp_grad_mask(alpha, orc.ALPHA_LIM)) < tol
    if adaptive:
        for k, g_, raw, lim in (("beta", be, beta, orc.BETA_LIM), ("a", aa, a, orc.A_LIM), ("b", bb, b, orc.B_LIM)):
            assert rel_err(g_.grad.cpu().numpy(), bw["d" + k] * orc.clamp_grad_mask(raw, lim)) < tol, k
    if recurrent:
        assert rel_err(Vt.grad.cpu().numpy(), bw["dV"]) < tol


@pytest.mark.parametrize("precision,flip_tol,loss_tol", [("fp32", FLIP_TOL, 0.10), ("bf16", 2e-2, 0.10)])
def test_bidirectional_long_sequence_cfg5_shape(precision, flip_tol, loss_tol):
    """BASELINE cfg5's structure at a batch the oracle can hold: bidirectional RadLIF 3x1024, T = 500, B = 8
    (Be = 16 after the direction doubling, second layer fed by 2H = 2048 features), a <- |a| (the reference is
    non-finite at T = 500 with its default draw).  Layer-0 spike trains and the loss against the
    oracle's torch restatement on the same device, in the fp32-equivalent mode and in the reduced (bf16) mode with its
    stated tolerance; every gradient of our model is finite.  The loss band is 10 %: over 500 steps a handful of
    threshold straddlers in layer 0 reshuffle layer 1 (through its BatchNorm over only 16 x 500 rows) -- the reference
    differs from ITSELF by 5-8 % in its logits between two thread counts; firing rates agree to 0.02."""
    import sparch_b200
    sp, _ = _mods()
    kw = dict(layer_sizes=[1024, 1024, 35], neuron_type="RadLIF", normalization="batchnorm", bidirectional=True)
    torch.manual_seed(0)
    net = sp.SNN((8, None, 40), **kw)
    ref = orc.build_oracle_snn((8, None, 40), **kw)
    ref.load_state_dict(net.state_dict())
    for m in (net, ref):
        with torch.no_grad():
            for lay in m.snn:
                if hasattr(lay, "a"):
                    lay.a.abs_()
                if isinstance(getattr(lay, "norm", None), torch.nn.BatchNorm1d):
                    lay.norm.weight.fill_(3.0)
                    lay.norm.bias.fill_(0.8)
    net, ref = net.to(DEV), ref.to(DEV)
    ref.snn[0].capture = {}
    torch.manual_seed(1234)
    x = torch.randn(8, 500, 40, device=DEV)
    y = torch.randint(0, 35, (8,), device=DEV)
    got = {}
    h = net.snn[0].register_forward_hook(lambda m, i, o: got.__setitem__(0, o.detach()))
    try:
        sparch_b200.set_precision(precision)
        torch.manual_seed(42)
        out, rates = net(x)
        loss = torch.nn.functional.cross_entropy(out, y)
        loss.backward()
    finally:
        sparch_b200.set_precision("fp32")
        h.remove()
    with torch.no_grad():
        torch.manual_seed(42)
        out_r, rates_r = ref(x)
    assert got[0].shape == (8, 500, 2048)
    s_ref = ref.snn[0].capture["s"]                 # (2B, T, H) before the direction merge (snns.py:686-689)
    s_f, s_b = s_ref[:8], s_ref[8:].flip(1)
    merged = torch.cat([s_f, s_b], dim=2)
    flips = float((got[0] != merged).float().mean())
    assert 0.005 < float(merged.mean()) < 0.9
    assert flips <= flip_tol, flips
    loss_r = torch.nn.functional.cross_entropy(out_r, y)
    assert abs(float(loss) - float(loss_r)) <= loss_tol * abs(float(loss_r)), (float(loss), float(loss_r))
    assert float((rates - rates_r).abs().max()) < 0.02
    assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in net.parameters())


@pytest.mark.parametrize("B,C", [(256, 35), (1, 20), (7, 1), (300, 1000)])
def test_cross_entropy_against_torch(B, C):
    """sparch_b200.CrossEntropyLoss = nn.CrossEntropyLoss() (exp.py:83, 362): loss and gradient to 1e-6 of torch's fp32
    result, through the C ABI (one launch each way)."""
    import sparch_b200
    g = torch.Generator(device=DEV).manual_seed(B + C)
    x = (torch.randn(B, C, device=DEV, generator=g) * 4).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    y = torch.randint(0, C, (B,), device=DEV, generator=g)
    loss = sparch_b200.CrossEntropyLoss()(x, y)
    loss_r = torch.nn.CrossEntropyLoss()(xr, y)
    (loss * 1.7).backward()
    (loss_r * 1.7).backward()
    assert abs(float(loss) - float(loss_r)) <= 2e-6 * max(1.0, abs(float(loss_r)))
    assert rel_err(x.grad.cpu().numpy(), xr.grad.cpu().numpy()) < 2e-6 or float(xr.grad.abs().max()) < 1e-12
    with pytest.raises(RuntimeError):
        sparch_b200.CrossEntropyLoss()(torch.zeros(2, 3), torch.zeros(2, dtype=torch.long))


def test_recurrent_helper_kernels():
    """sparch_recur_v0 (snns.py:712), sparch_dv_boundary, sparch_zero_diag against their torch one-liners."""
    from sparch_b200._lib import call, ptr
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=DEV).manual_seed(3)
    for H, Be, T in ((9, 3, 4), (256, 5, 1), (100, 1, 7)):
        V = torch.randn(H, H, device=DEV, generator=g)
        V0 = torch.empty_like(V)
        call("sparch_recur_v0", ptr(V), H, ptr(V0), st)
        assert torch.equal(V0, V.clone().fill_diagonal_(0))
        s0 = torch.rand(Be, H, device=DEV, generator=g)
        S = (torch.rand(Be, T, H, device=DEV, generator=g) > 0.7).float()
        first = torch.empty_like(s0)
        call("sparch_dv_boundary", ptr(s0), ptr(S), Be, T, H, ptr(first), st)
        ref = s0.clone()
        if Be > 1:
            ref[1:] -= S[:-1, T - 1, :]
        assert torch.equal(first, ref)
        call("sparch_zero_diag", ptr(V), H, st)
        assert torch.equal(V, V0)


@pytest.mark.parametrize("M,N,K", [(256, 1024, 1024), (1024, 1024, 256), (5, 9, 9), (3, 130, 130), (70, 33, 1)])
def test_small_fp32_gemm_against_fp64(M, N, K):
    """sparch_small_gemm (rec_0 = s0 @ V0 with V0's zero diagonal on the fly, the t = 0 frames of dV accumulated into the
    main product with the diagonal masked, the stepwise paths' products): every layout / flag against fp64, 2e-6 of the
    largest entry (plain fp32 FFMA accumulation)."""
    from sparch_b200._lib import call, ptr
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    r = lambda *s: torch.randn(*s, device=DEV, generator=g)
    def check(C, ref):
        ref = ref.double()
        assert float((C.double() - ref).abs().max()) <= 2e-6 * max(float(ref.abs().max()), 1e-30)
    A, B = r(M, K), r(K, N)
    C = torch.full((M, N), 7.0, device=DEV)
    call("sparch_small_gemm", ptr(A), K, 0, ptr(B), N, ptr(C), N, M, N, K, 0, st)
    check(C, A.double() @ B.double())
    if N == K:                                                     # B = V read with a zero diagonal
        call("sparch_small_gemm", ptr(A), K, 0, ptr(B), N, ptr(C), N, M, N, K, 1, st)
        check(C, A.double() @ B.double().clone().fill_diagonal_(0))
    At = A.t().contiguous()                                        # A stored (K, M); accumulate; zero diagonal of C
    C0 = r(M, N)
    C = C0.clone()
    call("sparch_small_gemm", ptr(At), M, 1, ptr(B), N, ptr(C), N, M, N, K, 4, st)
    check(C, C0.double() + A.double() @ B.double())
    if M == N:
        C = C0.clone()
        call("sparch_small_gemm", ptr(At), M, 1, ptr(B), N, ptr(C), N, M, N, K, 2 | 4, st)
        check(C, (C0.double() + A.double() @ B.double()).fill_diagonal_(0))
    Bt = B.t().contiguous()                                        # B stored (N, K)
    call("sparch_small_gemm", ptr(A), K, 0, ptr(Bt), K, ptr(C), N, M, N, K, 8, st)
    check(C, A.double() @ B.double())
    big = r(M, 3, K)                                               # strided rows (a time slice of a (Be, T, H) tensor)
    call("sparch_small_gemm", ptr(big[:, 1, :]), 3 * K, 0, ptr(B), N, ptr(C), N, M, N, K, 0, st)
    check(C, big[:, 1, :].double() @ B.double())


@pytest.mark.parametrize("neuron_type,reg", [("RadLIF", False), ("LIF", True)])
def test_reference_training_loop_drop_in(neuron_type, reg):
    """INTEGRATION.md's claim, end to end: the reference's train / validation loops (sparch/exp.py:341-459, restated in
    oracle/exp_loop.py and pinned there against the reference's own functions) drive ``sparch_b200.SNN`` exactly as
    they drive ``sparch.models.snns.SNN`` -- constructor call of exp.py:305-314, ``self.net(x)`` returning
    (output, firing_rates), ``is_snn``, torch.optim.Adam + ReduceLROnPlateau + nn.CrossEntropyLoss, the optional
    firing-rate regularisers, ``.item()`` every step, default (reference) state initialisation.  Two epochs on a
    learnable synthetic task: finite losses that decrease, a scheduler step, rates in (0, 1)."""
    from oracle import exp_loop
    sp, _ = _mods()
    torch.manual_seed(0)
    net = sp.SNN(input_shape=(8, None, 40), layer_sizes=[64, 64, 10], neuron_type=neuron_type, dropout=0.1,
                 normalization="batchnorm", use_bias=False, bidirectional=False, use_readout_layer=True).to(DEV)
    with torch.no_grad():
        for lay in net.snn:
            if hasattr(lay, "a"):
                lay.a.abs_()
    ex = exp_loop.stub_experiment(net, torch.device(DEV), lr=5e-3, batches=12, use_regularizers=reg)
    first = last = None
    best = (0, 0)
    for e in (1, 2, 3):
        losses, accs = exp_loop.train_one_epoch(ex, e)
        assert all(np.isfinite(losses))
        first = first if first is not None else float(np.mean(losses[:4]))
        last = float(np.mean(losses[-4:]))
        best = exp_loop.valid_one_epoch(ex, e, *best)
    assert last < first, (first, last)
    assert any("train mean act rate" in m for m in ex.log)
    out, rates = net(ex.train_loader[0][0].to(DEV))
    assert out.shape == (8, 10) and 0.0 < float(rates.mean()) < 1.0
    assert best[0] in (1, 2, 3) and 0.0 <= best[1] <= 1.0


@pytest.mark.parametrize("B,rate,nb_steps", [(16, 8000, 100), (1, 10, 100), (5, 3000, 250), (3, 0, 100)])
def test_events_to_dense_against_the_data_oracle(B, rate, nb_steps):
    """sparch_b200.data.SpikingBatcher (csrc/data.cu) = the reference's SpikingDataset.__getitem__ + generateBatch
    (spiking_datasets.py:66-86, restated in oracle/data_oracle.py and pinned there against the reference): bit-exact dense
    count tensors from SHD-shaped event lists, incl. duplicates in a bin, an example without events, events on the first
    bin edge; an event at / beyond the last edge raises as the reference's sparse constructor does."""
    from oracle import data_oracle as dor
    from sparch_b200.data import SpikingBatcher
    if rate:
        T, U, y = dor.synthetic_events(B, seed=B + rate, rate=rate)
        T[0][:2] = 0.0
        U[0][:2] = 3
    else:
        T, U, y = [np.zeros(0, np.float16)] * B, [np.zeros(0, np.uint16)] * B, np.arange(B)
    if B > 2 and rate:
        T[2], U[2] = T[2][:0], U[2][:0]                    # an empty example in the middle
    bat = SpikingBatcher(nb_steps=nb_steps, device=DEV)
    x, xlens, yy = bat(T, U, y)
    ox, olens, oy = dor.batch_to_dense(T, U, y, nb_steps=nb_steps)
    assert x.shape == (B, nb_steps, 700) and x.dtype == torch.float32
    assert np.array_equal(x.cpu().numpy(), ox)
    assert np.array_equal(xlens.numpy(), olens) and np.array_equal(yy.cpu().numpy(), oy)
    if rate:
        T[-1] = np.append(T[-1], np.float16(1.4))           # np.digitize -> bin nb_steps: outside the grid
        U[-1] = np.append(U[-1], np.uint16(0))
        with pytest.raises(ValueError):
            bat(T, U, y)
        with pytest.raises(IndexError):
            dor.batch_to_dense(T, U, y, nb_steps=nb_steps)


@pytest.mark.parametrize("M,H,affine", [(2560, 1024, True), (7, 9, True), (300, 130, False), (64, 2048, True), (1, 35, True)])
def test_layernorm_kernels_against_torch(M, H, affine):
    """LayerNormFunction (csrc/norm.cu) = nn.LayerNorm(H) (normalization="layernorm", snns.py:98-99, 678-680): output
    2e-6, dX / dgamma / dbeta 1e-5 of torch's fp32 results (fp64 reference for the gradients)."""
    _, F = _mods()
    g = torch.Generator(device=DEV).manual_seed(M + H)
    x = (torch.randn(M, H, device=DEV, generator=g) * 3 + 1).requires_grad_(True)
    ln = torch.nn.LayerNorm(H, elementwise_affine=affine).to(DEV)
    if affine:
        with torch.no_grad():
            ln.weight.copy_(torch.randn(H, device=DEV, generator=g))
            ln.bias.copy_(torch.randn(H, device=DEV, generator=g))
    y = F.LayerNormFunction.apply(x, ln.weight, ln.bias, ln.eps)
    x64 = x.detach().double().requires_grad_(True)
    ln64 = torch.nn.LayerNorm(H, elementwise_affine=affine).to(DEV).double()
    if affine:
        ln64.load_state_dict({k: v.double() for k, v in ln.state_dict().items()})
    y64 = ln64(x64)
    assert rel_err(y.detach().cpu().numpy(), y64.detach().cpu().numpy()) < 2e-6
    gy = torch.randn(M, H, device=DEV, generator=g)
    params = [ln.weight, ln.bias] if affine else []
    grads = torch.autograd.grad(y, [x] + params, gy)
    grads64 = torch.autograd.grad(y64, [x64] + (list(ln64.parameters()) if affine else []), gy.double())
    for a, b in zip(grads, grads64):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5


def test_graphed_step_equals_eager_step():
    """One replay of GraphedTrainStep = one eager step from the same parameters, seeds and batch: same loss, same
    parameters afterwards (the graph replays the launches the eager step issues; state draws and dropout seeds come
    from the CUDA generator in both)."""
    import copy
    import sparch_b200
    from sparch_b200.graphs import GraphedTrainStep
    from sparch_b200.optim import Adam
    sp, _ = _mods()
    sparch_b200.set_state_init("device")
    try:
        torch.manual_seed(0)
        net_a = sp.SNN((16, None, 40), layer_sizes=[128, 128, 10], neuron_type="RadLIF", normalization="batchnorm",
                       dropout=0.1).to(DEV)
        with torch.no_grad():
            for lay in net_a.snn:
                if hasattr(lay, "a"):
                    lay.a.abs_()
        net_b = copy.deepcopy(net_a)
        g = torch.Generator(device=DEV).manual_seed(1)
        x = torch.randn(16, 30, 40, device=DEV, generator=g)
        y = torch.randint(0, 10, (16,), device=DEV, generator=g)
        loss_fn = sparch_b200.CrossEntropyLoss()
        opt_a, opt_b = Adam(net_a.parameters(), 1e-2), Adam(net_b.parameters(), 1e-2)
        graphed = GraphedTrainStep(net_a, opt_a, loss_fn, x, y, warmup=3)     # 3 eager warm-up steps + capture
        for _ in range(3):                                                      # the same three steps on the copy
            out, _ = net_b(x)
            l = loss_fn(out, y)
            opt_b.zero_grad()
            l.backward()
            opt_b.step()
        # both nets went through three eager steps with different random draws; level them again
        net_b.load_state_dict(net_a.state_dict())
        opt_b.load_state_dict(copy.deepcopy(opt_a.state_dict()))
        torch.manual_seed(77)
        la = float(graphed.step(x, y))
        torch.manual_seed(77)
        net_b.train()
        out, _ = net_b(x)
        lb = loss_fn(out, y)
        opt_b.zero_grad()
        lb.backward()
        opt_b.step()
        assert abs(la - float(lb)) <= 1e-6 * max(1.0, abs(la)), (la, float(lb))
        for (k, pa), pb in zip(net_a.named_parameters(), net_b.parameters()):
            assert rel_err(pa.detach().cpu().numpy(), pb.detach().cpu().numpy()) < 1e-5, k
    finally:
        sparch_b200.set_state_init("cpu")


@pytest.mark.parametrize("kind,H,p,B,T", [("RadLIF", 256, 0.1, 70, 20), ("RLIF", 100, 0.0, 130, 7), ("RadLIF", 1024, 0.25, 256, 12)])
def test_packed_plane_hand_over_equals_fp32_spike_tensor(kind, H, p, B, T):
    """SURVEY 8 N1, first step: with the layer's post pass reading the PACKED spike planes the tcgen05 forward published
    (no fp32 spike tensor is written: 0.25 instead of 8 B/elt between the two kernels), outputs, firing rates and every
    gradient are bit-identical to the path that writes and re-reads fp32 spikes (same seeds, same kernels otherwise)."""
    import sparch_b200.snns as snns_mod
    sp, _ = _mods()
    res = []
    for lazy in (True, False):
        snns_mod._LAZY_SPIKES = lazy
        try:
            torch.manual_seed(0)
            net = sp.SNN((B, None, 40), layer_sizes=[H, H, 10], neuron_type=kind, normalization="batchnorm",
                         dropout=p).to(DEV)
            with torch.no_grad():
                for lay in net.snn:
                    if hasattr(lay, "a"):
                        lay.a.abs_()
                    if isinstance(getattr(lay, "norm", None), torch.nn.BatchNorm1d):
                        lay.norm.weight.fill_(3.0)
                        lay.norm.bias.fill_(0.8)
            g = torch.Generator(device=DEV).manual_seed(5)
            x = torch.randn(B, T, 40, device=DEV, generator=g)
            y = torch.randint(0, 10, (B,), device=DEV, generator=g)
            torch.manual_seed(11)
            out, rates = net(x)
            loss = torch.nn.functional.cross_entropy(out, y) + 0.1 * rates.sum()
            loss.backward()
            res.append((out.detach().clone(), rates.detach().clone(),
                        {k: q.grad.detach().clone() for k, q in net.named_parameters()}))
        finally:
            snns_mod._LAZY_SPIKES = True
    (oa, ra, ga), (ob, rb, gb) = res
    assert 0.005 < float(ra.mean()) < 0.9
    assert torch.equal(oa, ob) and torch.equal(ra, rb)
    for k in ga:
        assert torch.equal(ga[k], gb[k]), k


@pytest.mark.parametrize("kind,norm,p,bias,precision", [
    ("RadLIF", "batchnorm", 0.1, False, "fp32"), ("RLIF", "batchnorm", 0.0, False, "fp32"),
    ("RadLIF", "batchnorm", 0.25, True, "fp32"), ("RadLIF", "none", 0.0, False, "fp32"),
    ("RadLIF", "batchnorm", 0.1, False, "bf16")])
def test_copy_free_bidirectional_equals_flip_cat(kind, norm, p, bias, precision):
    _bidir_fused_vs_flip_cat(kind, norm, p, bias, precision, 128, 20, 256)


@pytest.mark.parametrize("B,T,H", [(128, 1, 64), (128, 17, 264), (256, 33, 40), (128, 16, 1024), (384, 5, 96)])
def test_copy_free_bidirectional_edge_shapes(B, T, H):
    """The same comparison at the edges: a single timestep, T around the adaptation checkpoint distance (16), hidden
    sizes that are multiples of 8 but not of the kernels' 16 / 32 / 256-wide slices, several 128-row groups per direction."""
    _bidir_fused_vs_flip_cat("RadLIF", "batchnorm", 0.1, False, "fp32", B, T, H)


def _bidir_fused_vs_flip_cat(kind, norm, p, bias, precision, B, T, H):
    """a bidirectional layer WITHOUT the reference's flipped / concatenated copies (snns.py:666-668,
    686-689) -- one projection of the un-flipped batch, the second half of the recurrence reading it time-reversed, the
    merge written by the post pass from the packed planes, the BatchNorm backward summing a row's two uses -- against
    the flip / cat formulation (which the fixtures pin to the reference), same seeds: outputs and firing rates equal
    (the dropout mask is keyed by the merged position in both), gradients to rel 2e-5 of their maximum (the batch-
    contracted GEMMs sum the rows in another order; reduced-precision mode: 2e-2)."""
    import sparch_b200
    import sparch_b200.snns as snns_mod
    sp, _ = _mods()
    res = []
    sparch_b200.set_precision(precision)
    try:
        for fused in (True, False):
            snns_mod._BIDIR_FUSED = fused
            torch.manual_seed(0)
            net = sp.SNN((B, None, 40), layer_sizes=[H, H, 10], neuron_type=kind, normalization=norm, dropout=p,
                         use_bias=bias, bidirectional=True).to(DEV)
            with torch.no_grad():
                for lay in net.snn:
                    if hasattr(lay, "a"):
                        lay.a.abs_()
                    if isinstance(getattr(lay, "norm", None), torch.nn.BatchNorm1d):
                        lay.norm.weight.fill_(3.0)
                        lay.norm.bias.fill_(0.8)
            g = torch.Generator(device=DEV).manual_seed(5)
            x = (torch.randn(B, T, 40, device=DEV, generator=g) * (1.0 if norm == "batchnorm" else 10.0)).requires_grad_(True)
            y = torch.randint(0, 10, (B,), device=DEV, generator=g)
            torch.manual_seed(11)
            out, rates = net(x)
            loss = torch.nn.functional.cross_entropy(out, y) + 0.1 * rates.sum()
            loss.backward()
            grads = {k: q.grad.detach().clone() for k, q in net.named_parameters()}
            grads["x"] = x.grad.detach().clone()
            stats = {k: v.detach().clone() for k, v in net.state_dict().items() if "running" in k}
            net.eval()
            torch.manual_seed(12)
            with torch.no_grad():
                out_e, rates_e = net(x)
            res.append((out.detach().clone(), rates.detach().clone(), grads, stats, out_e.clone(), rates_e.clone()))
    finally:
        snns_mod._BIDIR_FUSED = True
        sparch_b200.set_precision("fp32")
    (oa, ra, ga, sa, ea, rea), (ob, rb, gb, sb, eb, reb) = res
    assert ra.shape == (4 * H,) and 0.003 < float(ra.mean()) < 0.9
    assert torch.equal(ra, rb) and torch.equal(oa, ob)
    assert torch.equal(rea, reb) and torch.equal(ea, eb)
    for k in sa:
        assert torch.allclose(sa[k], sb[k], rtol=1e-6, atol=1e-7), k
    tol = 2e-5 if precision == "fp32" else 2e-2
    for k in ga:
        # (a bias in front of BatchNorm has a gradient of exactly zero: what is there is rounding noise of the
        # projection's gradient, so it is measured on that scale)
        ref = gb[k.replace("W.bias", "W.weight")].abs().max() if k.endswith("W.bias") else gb[k].abs().max()
        err = float((ga[k] - gb[k]).abs().max() / ref.clamp_min(1e-30))
        assert err < tol, (k, err)


@pytest.mark.parametrize("seed,pre,n,post", [(0, 0, 1, 5), (42, 0, 256 * 1024, 700), (7, 100, 623, 3), (7, 524, 100, 1300),
                                             (3, 17, 3 * 256 * 1024 + 35, 40), (11, 624, 624, 624), (5, 1, 2000, 0)])
def test_device_mt19937_equals_torch_rand(seed, pre, n, post):
    """sparch_b200.rng.cpu_generator_rand: the reference's initial-state draws (torch.rand on the default CPU generator,
    snns.py:700-702) replayed on the device -- the same float32 values bit for bit, from any position of the
    generator's stream, and the generator left exactly where n host draws would have left it (what the host draws
    next is the same, the state blobs are equal)."""
    from sparch_b200 import rng
    torch.manual_seed(seed)
    torch.rand(pre)
    want = torch.rand(n)
    want_next = torch.rand(post)
    want_state = torch.get_rng_state().clone()
    torch.manual_seed(seed)
    torch.rand(pre)
    got = rng.cpu_generator_rand(n, torch.device(DEV))
    assert got is not None and got.shape == (n,)
    got_next = torch.rand(post)
    assert torch.equal(got.cpu(), want)
    assert torch.equal(got_next, want_next)
    assert torch.equal(torch.get_rng_state(), want_state)


def test_default_state_mode_draws_equal_host_draws():
    """The drop-in's default ("cpu") state mode with the device replay of the generator against the same mode drawing on
    the host: identical outputs, gradients and final generator state for a whole model (so every fixture written by the
    reference, which the default mode reproduces, is reproduced by the device draws too)."""
    import sparch_b200.snns as snns_mod
    sp, _ = _mods()
    res = []
    for dev_mt in (True, False):
        snns_mod._DEVICE_MT = dev_mt
        try:
            torch.manual_seed(0)
            net = sp.SNN((16, None, 40), layer_sizes=[96, 64, 10], neuron_type="RadLIF", normalization="batchnorm",
                         dropout=0.1).to(DEV)
            x = torch.randn(16, 30, 40, generator=torch.Generator().manual_seed(1)).to(DEV)
            torch.manual_seed(21)
            out, rates = net(x)
            (out.square().sum() + rates.sum()).backward()
            res.append((out.detach().clone(), rates.detach().clone(), [q.grad.clone() for q in net.parameters()],
                        torch.get_rng_state().clone()))
        finally:
            snns_mod._DEVICE_MT = True
    (oa, ra, ga, sa), (ob, rb, gb, sb) = res
    assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(sa, sb)
    for a, b in zip(ga, gb):
        assert torch.equal(a, b)


@pytest.mark.parametrize("kind,bidir", [("RadLIF", False), ("adLIF", False), ("RLIF", True)])
def test_side_stream_preparation_equals_inline_order(kind, bidir):
    """SNN.forward issues the parameter-only launches of all layers (clamps, images of V0, weight operand terms, the
    state draws with rec_0) on a side stream at the start of the step; with SPARCH_B200_PREP_AHEAD=0 every layer issues
    them itself, in line.  Same kernels on the same data either way: outputs, rates, gradients and the CPU generator's
    final state are bit-identical (default state mode; two steps, so that the second step's preparation overlaps the
    first step's tail)."""
    import sparch_b200.snns as snns_mod
    sp, _ = _mods()
    res = []
    for ahead in (True, False):
        snns_mod._PREP_AHEAD = ahead
        try:
            torch.manual_seed(0)
            net = sp.SNN((32, None, 40), layer_sizes=[128, 96, 10], neuron_type=kind, normalization="batchnorm",
                         dropout=0.1, bidirectional=bidir).to(DEV)
            opt = torch.optim.SGD(net.parameters(), lr=0.05)
            x = torch.randn(32, 25, 40, generator=torch.Generator().manual_seed(1)).to(DEV)
            torch.manual_seed(21)
            for _ in range(2):
                opt.zero_grad(set_to_none=True)
                out, rates = net(x)
                (out.square().sum() + rates.sum()).backward()
                opt.step()
            res.append((out.detach().clone(), rates.detach().clone(), [q.grad.clone() for q in net.parameters()],
                        torch.get_rng_state().clone()))
        finally:
            snns_mod._PREP_AHEAD = True
    (oa, ra, ga, sa), (ob, rb, gb, sb) = res
    assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(sa, sb)
    for a, b in zip(ga, gb):
        assert torch.equal(a, b)


def test_copy_free_bidirectional_against_oracle():
    """The copy-free bidirectional path (one projection, reversed TMA coordinates, merged output from the packed planes)
    DIRECTLY against the oracle's torch restatement of snns.py:663-694 on the same device -- not only against this
    library's own flip / cat formulation: bidirectional RadLIF [256, 256, 35] + BatchNorm, B = 128 (the smallest batch
    the path takes: one 128-row group per direction), T = 60, a <- |a|.  Layer-0 merged spike trains (forward hook, so
    the hidden output is materialised) within the flip tolerance, firing rates to 0.01, loss to 2 %, all gradients
    finite and the projection's gradient within 5 % in relative L2 (free-running: a straddler reshuffles what follows)."""
    sp, _ = _mods()
    kw = dict(layer_sizes=[256, 256, 35], neuron_type="RadLIF", normalization="batchnorm", bidirectional=True)
    torch.manual_seed(0)
    net = sp.SNN((128, None, 40), **kw)
    ref = orc.build_oracle_snn((128, None, 40), **kw)
    ref.load_state_dict(net.state_dict())
    for m in (net, ref):
        with torch.no_grad():
            for lay in m.snn:
                if hasattr(lay, "a"):
                    lay.a.abs_()
                if isinstance(getattr(lay, "norm", None), torch.nn.BatchNorm1d):
                    lay.norm.weight.fill_(3.0)
                    lay.norm.bias.fill_(0.8)
    net, ref = net.to(DEV), ref.to(DEV)
    assert net.snn[0]._copy_free_bidir(torch.empty(128, 60, 40, device=DEV))
    ref.snn[0].capture = {}
    torch.manual_seed(1234)
    x = torch.randn(128, 60, 40, device=DEV)
    y = torch.randint(0, 35, (128,), device=DEV)
    got = {}
    h = net.snn[0].register_forward_hook(lambda m, i, o: got.__setitem__(0, o.detach()))
    try:
        torch.manual_seed(42)
        out, rates = net(x)
        loss = torch.nn.functional.cross_entropy(out, y)
        loss.backward()
    finally:
        h.remove()
    torch.manual_seed(42)
    out_r, rates_r = ref(x)
    loss_r = torch.nn.functional.cross_entropy(out_r, y)
    loss_r.backward()
    s_ref = ref.snn[0].capture["s"]                 # (2B, T, H) before the direction merge (snns.py:686-689)
    merged = torch.cat([s_ref[:128], s_ref[128:].flip(1)], dim=2)
    assert got[0].shape == (128, 60, 512) and 0.005 < float(merged.mean()) < 0.9
    flips = float((got[0] != merged).float().mean())
    assert flips <= FLIP_TOL, flips
    lo, lr_ = float(loss.detach()), float(loss_r.detach())
    assert abs(lo - lr_) <= 0.02 * abs(lr_), (lo, lr_)
    assert float((rates - rates_r).abs().max()) < 0.01
    assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in net.parameters())
    gw, gw_r = net.snn[0].W.weight.grad, ref.snn[0].W.weight.grad
    assert float((gw - gw_r).norm() / gw_r.norm()) < 0.05
