"""Host side of the tcgen05 GEMM (csrc/gemm.cu): bf16 term splitting and the multi-pair call.

``C = alpha * sum_{(i,j) in pairs} A_i @ B_j^T`` with fp32 accumulation in TMEM.  An fp32 operand
is carried as three bf16 terms (24 mantissa bits); an operand whose values are exactly
representable in bf16 (spikes, 0/1 masks) as one.
"""
import ctypes

import torch

from . import _lib
from ._lib import call, ptr

# (a_part, b_part) pairs, smallest contributions first so the fp32 accumulator adds them in
# increasing order of magnitude
PAIRS_33 = [(2, 0), (1, 1), (0, 2), (1, 0), (0, 1), (0, 0)]
PAIRS_13 = [(0, 2), (0, 1), (0, 0)]
PAIRS_31 = [(2, 0), (1, 0), (0, 0)]
PAIRS_11 = [(0, 0)]
# fp16 hi/lo terms: hi.lo + lo.hi + hi.hi (lo.lo is below 2^-22 of the product)
PAIRS_22 = [(1, 0), (0, 1), (0, 0)]
PAIRS_12 = [(0, 1), (0, 0)]
PAIRS_21 = [(1, 0), (0, 0)]


def pairs_for(na, nb):
    return {(3, 3): PAIRS_33, (1, 3): PAIRS_13, (3, 1): PAIRS_31, (1, 1): PAIRS_11,
            (2, 2): PAIRS_22, (1, 2): PAIRS_12, (2, 1): PAIRS_21}[(na, nb)]


# How a general fp32 operand travels to the tensor pipe.
#   "f16x2" (default): two fp16 terms of x * 2^k, one power-of-two scale per tensor taken from max|x| on the
#            device (22 mantissa bits; general x general = 3 passes, spikes x general = 2)
#   "bf16x3": three bf16 terms (24 bits, no scale; 6 resp. 3 passes) -- the first round-1 scheme, kept for comparison
#   "bf16x1": one bf16 term (the reduced-precision mode of sparch_b200.set_precision("bf16"))
MODE = "f16x2"


class Terms:
    """The terms of one split operand: ``parts`` (n, rows, ld) fp16/bf16 and, for scaled fp16 terms,
    ``amax`` (1,) int32 holding the bit pattern of max|x| (the GEMM epilogue undoes the scale)."""
    __slots__ = ("parts", "amax")

    def __init__(self, parts, amax=None):
        self.parts = parts
        self.amax = amax

    @property
    def n(self):
        return self.parts.shape[0]


def absmax(x2d):
    """(1,) int32 device tensor: bit pattern of max|x| (no host synchronisation)."""
    x2d = x2d.contiguous() if x2d.stride(-1) != 1 else x2d
    out = torch.empty(1, device=x2d.device, dtype=torch.int32)
    call("sparch_absmax", ptr(x2d), x2d.stride(0), x2d.shape[0], x2d.shape[1], ptr(out), _stream())
    return out


def split_f16(x2d, nparts, prescale=1.0, amax=None, scaled=True):
    """(M, K) fp32 -> Terms of nparts fp16 tensors (M, ld).  scaled: one power-of-two scale from max|x|
    (computed here unless the producer of x already left it in ``amax``)."""
    x2d = x2d.contiguous() if x2d.stride(-1) != 1 else x2d
    M, K = x2d.shape
    ld = _pad8(K)
    compute = scaled and amax is None
    if compute:
        amax = torch.empty(1, device=x2d.device, dtype=torch.int32)
    parts = torch.empty(nparts, M, ld, device=x2d.device, dtype=torch.float16)
    call("sparch_split_f16", ptr(x2d), x2d.stride(0), M, K, nparts, float(prescale), ptr(amax) if scaled else None,
         int(compute), ptr(parts[0]), ptr(parts[1]) if nparts > 1 else None, ld, _stream())
    return Terms(parts, amax if scaled else None)


def split_general(x2d, amax=None):
    """Terms of a general fp32 operand in the current MODE."""
    if MODE == "f16x2":
        return split_f16(x2d, 2, amax=amax)
    return Terms(split_rows(x2d, 3 if MODE == "bf16x3" else 1))


def split_binary(x2d, prescale=1.0):
    """One exact term of an operand whose values are all 0 or 1/prescale (spikes, dropped spikes)."""
    if MODE == "f16x2":
        return split_f16(x2d, 1, prescale=prescale, scaled=False)
    return Terms(split_rows(x2d, 1, prescale=prescale))


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _pad8(n):
    return (n + 7) // 8 * 8


def split_rows(x2d, nparts, prescale=1.0):
    """(M, K) fp32 -> nparts bf16 tensors (M, ld) with ld = K rounded up to 8 (zero padded)."""
    x2d = x2d.contiguous() if x2d.stride(-1) != 1 else x2d
    M, K = x2d.shape
    ld = _pad8(K)
    parts = torch.empty(nparts, M, ld, device=x2d.device, dtype=torch.bfloat16)
    p = [ptr(parts[i]) if i < nparts else None for i in range(3)]
    call("sparch_split_bf16", ptr(x2d), x2d.stride(0), M, K, nparts, float(prescale), p[0], p[1], p[2], ld, _stream())
    return parts


def split_transposed(x2d, nparts, T=0, shift=0, prescale=1.0):
    """(R, C) fp32 contiguous -> nparts bf16 tensors (C, ld >= R): the transposed terms.  T/shift
    delay the time index of rows laid out as (b, t) (zero-filled), see the C header."""
    x2d = x2d.contiguous()
    R, C = x2d.shape
    ld = _pad8(R)
    parts = torch.empty(nparts, C, ld, device=x2d.device, dtype=torch.bfloat16)
    p = [ptr(parts[i]) if i < nparts else None for i in range(3)]
    call("sparch_split_bf16_transpose", ptr(x2d), R, C, nparts, T, shift, float(prescale), p[0], p[1], p[2], ld,
         _stream())
    return parts


def gemm_parts(A, B, K, alpha=1.0, bias=None, out=None, pairs=None, a_mn=False, b_mn=False, a_koff=0,
               M=None, N=None, stats=None):
    """C (M, N) fp32 = alpha * sum_pairs A_i . B_j^T (+ bias).

    K-major operand: terms of shape (n, rows, ld) with rows = M resp. N.  MN-major operand
    (a_mn / b_mn): terms of shape (n, K, ld) holding the (K, M) resp. (K, N) matrix, M / N given
    explicitly (ld may exceed it).  a_koff shifts A's K index (zero fill), MN-major A only.
    stats: optional (2, N) float64 tensor receiving the per-column sum / sum of squares of C.
    A, B: ``Terms`` or plain (n, rows, ld) bf16 tensors of unscaled terms."""
    amax_a = amax_b = None
    if isinstance(A, Terms):
        A, amax_a = A.parts, A.amax
    if isinstance(B, Terms):
        B, amax_b = B.parts, B.amax
    fp16 = A.dtype == torch.float16
    assert (B.dtype == torch.float16) == fp16, "both operands must use the same 16-bit format"
    na, ra, lda = A.shape
    nb, rb, ldb = B.shape
    M = ra if not a_mn else M
    N = rb if not b_mn else N
    if pairs is None:
        pairs = pairs_for(na, nb)
    dev = A.device
    if out is None:
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
    assert out.stride(-1) == 1
    ap = (ctypes.c_void_p * na)(*[A[i].data_ptr() for i in range(na)])
    bp = (ctypes.c_void_p * nb)(*[B[i].data_ptr() for i in range(nb)])
    pa = (ctypes.c_int * len(pairs))(*[p[0] for p in pairs])
    pb = (ctypes.c_int * len(pairs))(*[p[1] for p in pairs])
    ws = None
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    if tiles < 148 and K > 64 and stats is None:
        ws = torch.empty(_lib.lib().sparch_gemm_workspace(M, N, K), device=dev, dtype=torch.uint8)
    call("sparch_gemm_terms", int(fp16), ap, na, ptr(amax_a), bp, nb, ptr(amax_b), lda, ldb, int(a_mn), int(b_mn),
         int(a_koff), pa, pb,
         len(pairs), M, N, K, float(alpha), ptr(bias), ptr(out), out.stride(0),
         None if stats is None else ptr(stats[0]), None if stats is None else ptr(stats[1]), ptr(ws), _stream())
    return out
