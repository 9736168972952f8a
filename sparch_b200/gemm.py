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


def pairs_for(na, nb):
    return {(3, 3): PAIRS_33, (1, 3): PAIRS_13, (3, 1): PAIRS_31, (1, 1): PAIRS_11}[(na, nb)]


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
    stats: optional (2, N) float64 tensor receiving the per-column sum / sum of squares of C."""
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
    call("sparch_gemm_bf16", ap, na, bp, nb, lda, ldb, int(a_mn), int(b_mn), int(a_koff), pa, pb,
         len(pairs), M, N, K, float(alpha), ptr(bias), ptr(out), out.stride(0),
         None if stats is None else ptr(stats[0]), None if stats is None else ptr(stats[1]), ptr(ws), _stream())
    return out
