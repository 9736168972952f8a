"""GPU tests of the tcgen05 GEMM and the bf16 term splitting against fp64 matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref(A, B):
    return A.double() @ B.double().t()


@pytest.mark.parametrize("M,N,K", [(300, 1024, 40), (256, 35, 1024), (1000, 520, 700),
                                   (1024, 1024, 5000), (128, 256, 64), (130, 257, 72), (35, 1024, 3000),
                                   (4096, 1024, 1024)])
def test_gemm_fp32_accuracy(M, N, K):
    from sparch_b200 import gemm
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, device=DEV, generator=g) * 3
    B = torch.randn(N, K, device=DEV, generator=g)
    bias = torch.randn(N, device=DEV, generator=g)
    C = gemm.gemm_parts(gemm.split_rows(A, 3), gemm.split_rows(B, 3), K, alpha=0.5, bias=bias)
    ref = 0.5 * _ref(A, B) + bias.double()
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-6, err
    # binary A (spikes): one term is exact
    S = (torch.rand(M, K, device=DEV, generator=g) < 0.1).float()
    C = gemm.gemm_parts(gemm.split_rows(S, 1), gemm.split_rows(B, 3), K)
    ref = _ref(S, B)
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-6, err


def test_split_terms_reconstruct():
    from sparch_b200 import gemm
    x = torch.randn(77, 53, device=DEV) * 100
    x[0, 0] = 0.0
    x[1, 1] = 1e-30
    p = gemm.split_rows(x, 3)
    assert p.shape == (3, 77, 56) and float(p[:, :, 53:].abs().max()) == 0
    rec = p[0].double() + p[1].double() + p[2].double()
    assert float((rec[:, :53] - x.double()).abs().max() / 100) < 1e-7
    pt = gemm.split_transposed(x, 3)
    assert pt.shape == (3, 53, 80)
    rect = (pt[0].double() + pt[1].double() + pt[2].double())[:, :77]
    assert torch.equal(rect, rec[:, :53].t())


def test_transposed_gemm_with_time_shift():
    """dV-style product: S_prev^T @ dI with the spike rows delayed by one step inside each batch."""
    from sparch_b200 import gemm
    Be, T, H = 6, 9, 96
    g = torch.Generator(device=DEV).manual_seed(3)
    S = (torch.rand(Be, T, H, device=DEV, generator=g) < 0.2).float()
    dI = torch.randn(Be, T, H, device=DEV, generator=g)
    Sp = gemm.split_transposed(S.view(Be * T, H), 1, T=T, shift=1)
    dIt = gemm.split_transposed(dI.view(Be * T, H), 3)
    dV = gemm.gemm_parts(Sp, dIt, Be * T)
    ref = torch.einsum("bti,btj->ij", S[:, :-1].double(), dI[:, 1:].double())
    err = float((dV.double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-6, err


@pytest.mark.parametrize("Kc,M,N", [(5000, 1024, 1024), (700, 96, 40), (1000, 130, 300), (64, 128, 256)])
def test_gemm_mn_major_operands(Kc, M, N):
    """Contraction-over-rows products C = P^T Q straight from the row-major (Kc, M) / (Kc, N) terms."""
    from sparch_b200 import gemm
    g = torch.Generator(device=DEV).manual_seed(Kc + M)
    P = torch.randn(Kc, M, device=DEV, generator=g)
    Q = torch.randn(Kc, N, device=DEV, generator=g) * 2
    C = gemm.gemm_parts(gemm.split_rows(P, 3), gemm.split_rows(Q, 3), Kc, a_mn=True, b_mn=True, M=M, N=N)
    ref = P.double().t() @ Q.double()
    assert float((C.double() - ref).abs().max() / ref.abs().max()) < 2e-6
    # mixed: K-major A with MN-major B  (dX = dZ W with W (H, Fin) row-major)
    A = torch.randn(M, Kc, device=DEV, generator=g)
    C = gemm.gemm_parts(gemm.split_rows(A, 3), gemm.split_rows(Q, 3), Kc, b_mn=True, N=N)
    ref = A.double() @ Q.double()
    assert float((C.double() - ref).abs().max() / ref.abs().max()) < 2e-6


def test_gemm_mn_major_with_frame_delay():
    """dV = sum_m S[m-1]^T dI[m] via a_koff = -1 (frame -1 reads as zero)."""
    from sparch_b200 import gemm
    Mf, H = 900, 96
    g = torch.Generator(device=DEV).manual_seed(9)
    S = (torch.rand(Mf, H, device=DEV, generator=g) < 0.2).float()
    dI = torch.randn(Mf, H, device=DEV, generator=g)
    C = gemm.gemm_parts(gemm.split_rows(S, 1), gemm.split_rows(dI, 3), Mf, a_mn=True, b_mn=True, a_koff=-1,
                        M=H, N=H)
    ref = S[:-1].double().t() @ dI[1:].double()
    assert float((C.double() - ref).abs().max() / ref.abs().max()) < 2e-6


@pytest.mark.parametrize("M,N,K", [(1000, 520, 72), (25600, 1024, 40), (130, 35, 256)])
def test_gemm_fused_column_statistics(M, N, K):
    """BatchNorm statistics of the projection output accumulated in the GEMM epilogue."""
    from sparch_b200 import gemm
    g = torch.Generator(device=DEV).manual_seed(M + N)
    A = torch.randn(M, K, device=DEV, generator=g)
    B = torch.randn(N, K, device=DEV, generator=g)
    bias = torch.randn(N, device=DEV, generator=g)
    stats = torch.empty(2, N, dtype=torch.float64, device=DEV)
    C = gemm.gemm_parts(gemm.split_rows(A, 3), gemm.split_rows(B, 3), K, alpha=1.5, bias=bias, stats=stats)
    ref = C.double()
    assert float((stats[0] - ref.sum(0)).abs().max() / ref.sum(0).abs().max()) < 1e-6
    assert float((stats[1] - (ref * ref).sum(0)).abs().max() / (ref * ref).sum(0).abs().max()) < 1e-6


# ---- scaled fp16 hi/lo terms (the default of the fp32-equivalent mode) ----------------------------------
@pytest.mark.parametrize("M,N,K", [(300, 1024, 40), (256, 35, 1024), (1000, 520, 700), (1024, 1024, 5000),
                                   (130, 257, 72), (4096, 1024, 1024)])
@pytest.mark.parametrize("sa,sb", [(3.0, 1.0), (1e-12, 1e9), (4e4, 1e-3)])
def test_gemm_f16x2_accuracy(M, N, K, sa, sb):
    """Two fp16 terms with one power-of-two scale per tensor: 22 mantissa bits whatever the magnitudes."""
    from sparch_b200 import gemm
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, device=DEV, generator=g) * sa
    B = torch.randn(N, K, device=DEV, generator=g) * sb
    bias = torch.randn(N, device=DEV, generator=g) * (sa * sb)
    ta, tb = gemm.split_f16(A, 2), gemm.split_f16(B, 2)
    assert ta.parts.dtype == torch.float16 and ta.amax is not None
    C = gemm.gemm_parts(ta, tb, K, alpha=0.5, bias=bias)
    ref = 0.5 * _ref(A, B) + bias.double()
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-6, err
    # binary A (spikes after dropout 0.25): one exact unscaled term, the factor in alpha
    S = (torch.rand(M, K, device=DEV, generator=g) < 0.1).float() / 0.75
    C = gemm.gemm_parts(gemm.split_f16(S, 1, prescale=0.75, scaled=False), tb, K, alpha=1 / 0.75)
    ref = _ref(S, B)
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-6, err


def test_split_f16_terms_reconstruct_and_scale():
    from sparch_b200 import gemm
    x = torch.randn(77, 53, device=DEV) * 1e-7
    x[0, 0] = 0.0
    x[1, 1] = 1e-30
    t = gemm.split_f16(x, 2)
    amax = t.amax.view(torch.float32)
    assert float(amax) == float(x.abs().max())
    p = t.parts.double()
    assert p.shape == (2, 77, 56) and float(p[:, :, 53:].abs().max()) == 0
    assert 2 ** 12 <= float(p[0].abs().max()) < 2 ** 13 + 1
    k = torch.round(torch.log2(p[0].abs().max() / x.abs().max().double()))
    rec = (p[0] + p[1])[:, :53] / 2.0 ** k
    assert float((rec - x.double()).abs().max() / x.abs().max()) < 2.0 ** -22
    # all-zero tensor: unscaled, zero terms
    z = gemm.split_f16(torch.zeros(8, 16, device=DEV), 2)
    assert int(z.amax) == 0 and float(z.parts.abs().max()) == 0


@pytest.mark.parametrize("Kc,M,N", [(5000, 1024, 1024), (700, 96, 40), (1000, 130, 300)])
def test_gemm_f16x2_mn_major_and_frame_delay(Kc, M, N):
    from sparch_b200 import gemm
    g = torch.Generator(device=DEV).manual_seed(Kc + M)
    P = torch.randn(Kc, M, device=DEV, generator=g) * 1e-4
    Q = torch.randn(Kc, N, device=DEV, generator=g) * 2
    tq = gemm.split_f16(Q, 2)
    C = gemm.gemm_parts(gemm.split_f16(P, 2), tq, Kc, a_mn=True, b_mn=True, M=M, N=N)
    ref = P.double().t() @ Q.double()
    assert float((C.double() - ref).abs().max() / ref.abs().max()) < 2e-6
    A = torch.randn(M, Kc, device=DEV, generator=g)
    C = gemm.gemm_parts(gemm.split_f16(A, 2), tq, Kc, b_mn=True, N=N)
    ref = A.double() @ Q.double()
    assert float((C.double() - ref).abs().max() / ref.abs().max()) < 2e-6
    S = (torch.rand(Kc, M, device=DEV, generator=g) < 0.2).float()
    C = gemm.gemm_parts(gemm.split_f16(S, 1, scaled=False), tq, Kc, a_mn=True, b_mn=True, a_koff=-1, M=M, N=N)
    ref = S[:-1].double().t() @ Q[1:].double()
    assert float((C.double() - ref).abs().max() / ref.abs().max()) < 2e-6


def test_gemm_f16x2_fused_statistics_and_modes():
    from sparch_b200 import functional as F, gemm
    g = torch.Generator(device=DEV).manual_seed(5)
    A = torch.randn(1000, 72, device=DEV, generator=g)
    B = torch.randn(520, 72, device=DEV, generator=g)
    stats = torch.empty(2, 520, dtype=torch.float64, device=DEV)
    C = gemm.gemm_parts(gemm.split_f16(A, 2), gemm.split_f16(B, 2), 72, alpha=1.5, stats=stats)
    ref = C.double()
    assert float((stats[0] - ref.sum(0)).abs().max() / ref.sum(0).abs().max()) < 1e-6
    assert float((stats[1] - (ref * ref).sum(0)).abs().max() / (ref * ref).sum(0).abs().max()) < 1e-6
    try:
        for mode, dtype, n in (("fp32", torch.float16, 2), ("fp32-bf16x3", torch.bfloat16, 3),
                               ("bf16", torch.bfloat16, 1)):
            F.set_precision(mode)
            t = gemm.split_general(A)
            assert t.parts.dtype == dtype and t.n == n
    finally:
        F.set_precision("fp32")
