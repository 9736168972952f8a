/* sparch_b200.h -- C ABI of libsparch_b200.so (hand-written sm_100a CUDA kernels).
 *
 * This is the drop-in boundary for sparch's surrogate-gradient SNN hot path.  The
 * reference has NO native interface for this path: it is Python calling PyTorch ATen
 * (sparch/models/snns.py, whole file).  Each entry point below therefore cites the
 * reference Python lines whose ATen call sequence it replaces.  The binding a
 * maintainer adds on the reference side is a ctypes stub; the
 * repo's own stub is sparch_b200/_lib.py.
 *
 * Conventions:
 *  - every pointer is a DEVICE pointer into caller-owned memory; the library never
 *    allocates or frees persistent memory; tensors are contiguous row-major fp32
 *    unless a parameter says otherwise; activations are (Be, T, H) exactly as the
 *    reference lays them out (batch, time, neuron);
 *  - work is enqueued on the caller's CUDA stream `st` (a cudaStream_t), no sync;
 *  - return 0 on success, negative sparch_status on failure; sparch_last_error()
 *    gives a thread-local message;
 *  - CUDA only: there is no host fallback in this library.
 *
 * Neuron kinds (snns.py:109): 0 LIF, 1 adLIF, 2 RLIF, 3 RadLIF.
 *   bit0 = adaptive (w state, beta/a/b), bit1 = recurrent (V).
 */
#ifndef SPARCH_B200_H
#define SPARCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sparch_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define SPARCH_API __attribute__((visibility("default")))
#else
#define SPARCH_API
#endif

enum sparch_status {
  SPARCH_OK = 0,
  SPARCH_ERR_ARG = -1,     /* bad argument (shape, null pointer, unsupported size) */
  SPARCH_ERR_CUDA = -2,    /* a CUDA runtime call or launch failed                 */
  SPARCH_ERR_DEVICE = -3,  /* device is not sm_100 class / attribute unavailable   */
  SPARCH_ERR_TIMEOUT = -4  /* a persistent kernel's bounded spin expired           */
};

enum sparch_kind { SPARCH_LIF = 0, SPARCH_ADLIF = 1, SPARCH_RLIF = 2, SPARCH_RADLIF = 3 };

SPARCH_API const char* sparch_last_error(void);
SPARCH_API int sparch_abi_version(void);

/* ---- SpikeFunctionBoxcar (snns.py:20-36) ---------------------------------- */
/* s = (x > 0) as fp32                                         snns.py:29       */
SPARCH_API int sparch_boxcar_fwd(const float* x, float* s, int64_t n, sparch_stream_t st);
/* gx = gs where -0.5 < x <= 0.5 else 0                        snns.py:33-35    */
SPARCH_API int sparch_boxcar_bwd(const float* x, const float* gs, float* gx, int64_t n, sparch_stream_t st);

/* ---- BatchNorm1d(H, momentum=0.05) folded to scale/shift (snns.py:678-680) - */
/* sum[h] = sum_m Z[m,h], sumsq[h] = sum_m Z[m,h]^2 in fp64 (outputs are zeroed here). */
SPARCH_API int sparch_col_stats(const float* Z, int64_t M, int H, double* sum, double* sumsq,
                     sparch_stream_t st);
/* sum1[h] = sum_m A[m,h]; sum2[h] = sum_m A[m,h]*(Zn[m,h]-mean[h])*rstd[h]  (BN backward
 * reductions, autograd of snns.py:679).  mean/rstd may be NULL => sum2 = sum A*Zn.  amax_a (may be
 * NULL): also leaves the bit pattern of max|A| (for the fp16 split of A = dI in the dV GEMM).  */
SPARCH_API int sparch_col_dot(const float* A, const float* Zn, const float* mean, const float* rstd,
                   int64_t M, int H, double* sum1, double* sum2, uint32_t* amax_a, sparch_stream_t st);
/* train-mode fold: mean, biased var -> rstd; scale = gamma*rstd; shift = beta-mean*scale;
 * running stats updated with the unbiased variance (momentum form of ATen). gamma/beta or
 * running_* may be NULL.                                                                 */
SPARCH_API int sparch_bn_fold_train(const double* sum, const double* sumsq, int64_t M, const float* gamma,
                         const float* beta, float eps, float momentum, float* running_mean,
                         float* running_var, float* mean, float* rstd, float* scale,
                         float* shift, int H, sparch_stream_t st);
/* dZ[m,h] = scale[h]*(dI[m,h] - sum1[h]/M - xhat[m,h]*sum2[h]/M), in place on dI.  amax (may be
 * NULL): receives the bit pattern of max|dZ| (zeroed here) for the fp16 split of the gradient GEMMs. */
SPARCH_API int sparch_bn_bwd_apply(float* dI, const float* Z, const float* mean, const float* rstd,
                        const float* scale, const double* sum1, const double* sum2, int64_t M,
                        int H, uint32_t* amax, sparch_stream_t st);

/* Same dZ written directly as the two scaled fp16 terms (sparch_split_f16 layout, row stride ldp) that the
 * weight- and data-gradient GEMMs read, without the fp32 tensor in between (dZ32 may be NULL; it may
 * alias dI).  amax_dI: max|dI| bits from sparch_col_dot; bound (out): bit pattern of the upper bound of
 * max|dZ| that fixed the terms' scale -- pass it to sparch_gemm_terms as the operand's amax word;
 * coef: 4*H floats: the per-column constants sum1/M, sum2/M converted once (scratch), then sum1 and sum2 as fp32
 * (= the gradients of the BatchNorm bias and weight, autograd of snns.py:679).                                      */
SPARCH_API int sparch_bn_bwd_apply_f16(const float* dI, const float* Z, const float* mean, const float* rstd,
                        const float* scale, const double* sum1, const double* sum2, int64_t M, int H,
                        const uint32_t* amax_dI, uint32_t* bound, float* coef, void* P0, void* P1,
                        int64_t ldp, float* dZ32, sparch_stream_t st);

/* BatchNorm backward of a BIDIRECTIONAL layer whose projection ran once on the un-flipped batch (the reference feeds
 * cat([x, flip(x)]) through W and the norm, snns.py:666-680: every row of W x appears twice, so batch mean and biased
 * variance are those of the single copy and the gradient of a row of W x is the sum of its two uses).
 * dI (2*rev_from*T, H) in the recurrence's order; Z (rev_from*T, H).  col_dot_bidir: the column sums over all M =
 * 2*rev_from*T rows, row (b, t), b >= rev_from, pairing with Z row (b - rev_from, T-1-t).  apply: M = rev_from*T
 * OUTPUT rows, dZ[b,t] = scale*((dI[b,t] - c1 - xhat*c2) + (dI[b+rev_from, T-1-t] - c1 - xhat*c2)), c = sums / (2M);
 * the fp32 variant writes in place on the first M rows of dI, the f16 variant writes the operand terms (dZ32 may
 * alias dI's first M rows).  H % 4 == 0.                                                                        */
SPARCH_API int sparch_col_dot_bidir(const float* A, const float* Zn, const float* mean, const float* rstd, int64_t M,
                                    int H, int T, int64_t rev_from, double* sum1, double* sum2, uint32_t* amax_a,
                                    sparch_stream_t st);
SPARCH_API int sparch_bn_bwd_apply_bidir(float* dI, const float* Z, const float* mean, const float* rstd,
                                         const float* scale, const double* sum1, const double* sum2, int64_t M,
                                         int H, int T, sparch_stream_t st);
SPARCH_API int sparch_bn_bwd_apply_f16_bidir(const float* dI, const float* Z, const float* mean, const float* rstd,
                                             const float* scale, const double* sum1, const double* sum2, int64_t M,
                                             int H, int T, const uint32_t* amax_dI, uint32_t* bound, float* coef,
                                             void* P0, void* P1, int64_t ldp, float* dZ32, sparch_stream_t st);

/* ---- membrane recurrence, forward (snns.py:282-303, 419-445, 554-578, 696-727) ------- */
/* Non-recurrent kinds (LIF, adLIF): whole time loop in one streaming kernel, state in
 * registers.  I_t = Z[b,t,h]*scale[h]+shift[h] (scale/shift NULL => I = Z).  alpha..b are the
 * CLAMPED parameters.  Writes spikes S (fp32 {0,1}), membrane tape U and adaptation tape W
 * (W may be NULL for LIF).                                                               */
SPARCH_API int sparch_cell_fwd(int kind, const float* Z, const float* scale, const float* shift,
                    const float* alpha, const float* beta, const float* a, const float* b,
                    const float* u0, const float* w0, const float* s0, float theta, float* S,
                    float* U, float* W, int Be, int T, int H, sparch_stream_t st);
/* One timestep t of any kind; `rec` (Be,H) is s_{t-1}@V0 for recurrent kinds (else NULL).
 * Previous state is read from the tapes at t-1 (from u0/w0/s0 when t == 0).              */
SPARCH_API int sparch_cell_step_fwd(int kind, int t, const float* Z, const float* scale,
                         const float* shift, const float* alpha, const float* beta,
                         const float* a, const float* b, const float* rec, const float* u0,
                         const float* w0, const float* s0, float theta, float* S, float* U,
                         float* W, int Be, int T, int H, sparch_stream_t st);

/* ---- LayerNorm (normalization="layernorm": nn.LayerNorm(H) on W x, snns.py:98-99, 678-680) ------------------- */
/* Y[m] = (X[m] - mean_m) * rstd_m * gamma + beta over rows of H <= 2048 values (biased variance, eps inside the root);
 * mean / rstd (M) are kept for the backward.  gamma / beta may be NULL (elementwise_affine=False).               */
SPARCH_API int sparch_layernorm_fwd(const float* X, const float* gamma, const float* beta, float eps, int64_t M, int H,
                                    float* Y, float* mean, float* rstd, sparch_stream_t st);
/* dX, and (when asked for) dgamma = sum_m dY * xhat, dbeta = sum_m dY (fp64 partial sums, fixed order).            */
SPARCH_API size_t sparch_layernorm_bwd_workspace(int64_t M, int H);
SPARCH_API int sparch_layernorm_bwd(const float* dY, const float* X, const float* gamma, const float* mean,
                                    const float* rstd, int64_t M, int H, float* dX, float* dgamma, float* dbeta,
                                    void* workspace, sparch_stream_t st);

/* ---- membrane recurrence, reverse-time BPTT (autograd of the above) ----- */
/* Non-recurrent kinds: whole reverse loop.  G = dL/dS (Be,T,H).  Writes dI (Be,T,H) and the
 * per-(b,h) partial parameter gradients pa,pb,pc,pd (Be,H each; pb..pd NULL for LIF) which
 * the caller reduces over b (sparch_col_stats-style) and masks by the clamp window.      */
SPARCH_API int sparch_cell_bwd(int kind, const float* G, const float* U, const float* W,
                    const float* alpha, const float* beta, const float* a, const float* b,
                    const float* u0, const float* w0, const float* s0, float theta, float* dI,
                    float* p_alpha, float* p_beta, float* p_a, float* p_b, int Be, int T, int H,
                    sparch_stream_t st);
/* One reverse step t of any kind.  recb (Be,H) = dI_{t+1}@V0^T (NULL for t == T-1 or
 * non-recurrent).  du_next/dw_next (Be,H) carry the adjoint state across calls (zero them
 * before t == T-1); p_* are accumulated (+=).                                            */
SPARCH_API int sparch_cell_step_bwd(int kind, int t, const float* G, const float* U, const float* W,
                         const float* alpha, const float* beta, const float* a, const float* b,
                         const float* recb, const float* u0, const float* w0, const float* s0,
                         float theta, float* dI, float* du_next, float* dw_next, float* p_alpha,
                         float* p_beta, float* p_a, float* p_b, int Be, int T, int H,
                         sparch_stream_t st);

/* ---- time-parallel GEMMs on tcgen05/TMEM/TMA (snns.py:675 and its autograd) -------------- */
/* Split an fp32 matrix X (M, K; row stride ldx) into nparts (1..3) bf16 terms,
 * prescale*x = p0+p1+p2, written row-major with row stride ldp (zero padded; ldp % 8 == 0 for
 * the GEMM).  prescale = (1-p) with nparts = 1 turns dropout-scaled spikes {0, 1/(1-p)} into
 * the exact bf16 values {0, 1} (the factor goes into the GEMM's alpha).                     */
SPARCH_API int sparch_split_bf16(const float* X, int64_t ldx, int M, int K, int nparts,
                                 float prescale, void* P0, void* P1, void* P2, int64_t ldp,
                                 sparch_stream_t st);
/* Transposing split: X (R, C) contiguous -> parts (C, ldp >= R): part[c][r] = term(X[r][c]).
 * With T > 0 and shift > 0 rows are (b, t) and output column (b, t) takes X[b, t-shift, :],
 * zero for t < shift (the S_prev operand of dV, autograd of snns.py:720).                    */
SPARCH_API int sparch_split_bf16_transpose(const float* X, int R, int C, int nparts, int T,
                                           int shift, float prescale, void* P0, void* P1,
                                           void* P2, int64_t ldp, sparch_stream_t st);
/* C[M,N] (fp32, row stride ldc) = alpha * sum_p A[pair_a[p]] . B[pair_b[p]]^T (+ bias[n]).
 * K-major operands (a_mn/b_mn = 0): A parts are (M, K) and B parts (N, K) bf16 row-major.
 * MN-major operands (a_mn/b_mn = 1): the part in memory is the (K, M) resp. (K, N) row-major
 * matrix -- the layout the contraction-over-frames GEMMs (dW, dV) find their operands in, so no
 * transpose is materialised.  Row strides lda/ldb are multiples of 8 elements.  a_koff (MN-major
 * A only) is added to A's K coordinate, out-of-range rows read as zero: a_koff = -1 pairs frame
 * m of B with frame m-1 of A (S_prev in dV).  `workspace` (sparch_gemm_workspace bytes, may be
 * NULL) enables deterministic split-K when the output has fewer tiles than the GPU has SMs.
 * stat_sum / stat_sumsq (N doubles each, may be NULL): the epilogue also accumulates the
 * per-column sum and sum of squares of the finished output -- the BatchNorm1d statistics of the
 * projection (snns.py:678-680) without another pass over Z (disables split-K).                */
SPARCH_API size_t sparch_gemm_workspace(int M, int N, int K);
SPARCH_API int sparch_gemm_bf16(const void* const* A_parts, int na, const void* const* B_parts,
                                int nb, int64_t lda, int64_t ldb, int a_mn, int b_mn, int a_koff,
                                const int* pair_a, const int* pair_b, int npairs, int M, int N,
                                int K, float alpha, const float* bias, float* C, int64_t ldc,
                                double* stat_sum, double* stat_sumsq, void* workspace,
                                sparch_stream_t st);

/* fp16 two-term operands (the default of the fp32-equivalent mode): x * 2^k = hi + lo with hi, lo fp16
 * (22 mantissa bits) and ONE power-of-two scale per tensor, chosen from max|x| so that the largest
 * value lands in [2^12, 2^13).  general x general needs 3 tensor-pipe passes (hi.hi, hi.lo, lo.hi) instead
 * of the 6 of three bf16 terms, spikes x general 2 instead of 3.
 * sparch_absmax: *amax = bit pattern of max|X| (uint32; zeroed here).  sparch_split_f16: nparts 1 or 2
 * fp16 terms of prescale * 2^k * X, k derived from *amax (amax NULL: k = 0, e.g. spikes with
 * prescale = 1-p); compute_amax != 0 runs sparch_absmax into *amax first (one call, no host round trip),
 * 0 takes the value a producer kernel left there.  sparch_gemm_terms: sparch_gemm_bf16 with fp16 != 0 selecting fp16 terms; amax_a /
 * amax_b (NULL = unscaled operand) are the same device words the splits used, the epilogue undoes both
 * scales.  No host synchronisation anywhere: the scale travels through device memory.            */
SPARCH_API int sparch_absmax(const float* X, int64_t ldx, int64_t M, int K, uint32_t* amax,
                             sparch_stream_t st);
SPARCH_API int sparch_split_f16(const float* X, int64_t ldx, int M, int K, int nparts, float prescale,
                                uint32_t* amax, int compute_amax, void* P0, void* P1, int64_t ldp,
                                sparch_stream_t st);
SPARCH_API int sparch_gemm_terms(int fp16, const void* const* A_parts, int na, const uint32_t* amax_a,
                                 const void* const* B_parts, int nb, const uint32_t* amax_b,
                                 int64_t lda, int64_t ldb, int a_mn, int b_mn, int a_koff,
                                 const int* pair_a, const int* pair_b, int npairs, int M, int N,
                                 int K, float alpha, const float* bias, float* C, int64_t ldc,
                                 double* stat_sum, double* stat_sumsq, void* workspace,
                                 sparch_stream_t st);

/* ---- recurrent kinds on the tensor pipe (snns.py:554-578, 696-727) --------------------- */
/* Hidden size rounded up to a multiple of 32 (spike words / V0 slices are padded to it).    */
SPARCH_API int sparch_recur_padded(int H);
/* Build the fragment-ordered fp16 hi/lo images of V0 = V with zero diagonal (snns.py:566, 712):
 * img_fwd for s @ V0, img_bwd for dI @ V0^T (either may be NULL).  Each image is Hp*Hp uint32
 * words; meta is 2 ints (meta[0] = E with max|V0| < 2^E).                                    */
SPARCH_API int sparch_recur_prepare(const float* V, int H, uint32_t* img_fwd, uint32_t* img_bwd,
                                    int* meta, sparch_stream_t st);
/* All T steps of an RLIF/RadLIF layer in ONE persistent cooperative kernel per <= 148-CTA batch
 * chunk: V0 slices stay in shared memory, neuron state in registers; each step's spike words
 * cross CTAs through L2 as tagged 64-bit words.  rec0 (Be,H) = s0 @ V0 (s0 is real-valued,
 * snns.py:702).  Writes the fp32 tapes S, U (and W for RadLIF) and `bits`, the packed spike
 * planes [T][Be][Hp/32] of 8-byte words {32 spikes, t+1} (2 * T*Be*Hp/32 uint32).
 * reduced = 0: fp32-equivalent products (V0 as fp16 hi + lo); reduced = 1: the reduced-precision
 * mode, hi terms only (11 mantissa bits), for sparch_recur_bwd also hi x hi only.               */
SPARCH_API int sparch_recur_fwd(int kind, const float* Z, const float* scale, const float* shift,
                                const float* alpha, const float* beta, const float* a,
                                const float* b, const float* rec0, const uint32_t* img_fwd,
                                const int* meta, const float* u0, const float* w0,
                                const float* s0, float theta, float* S, float* U, float* W,
                                uint32_t* bits, int reduced, int Be, int T, int H,
                                sparch_stream_t st);
/* The same layer pass (snns.py:572, 718-724; one persistent cooperative kernel for all T steps) with
 * s_{t-1} @ V0 on tcgen05 integer tensor cores.  A CTA owns 128 batch rows x 16 neurons; its columns of V0 sit in
 * shared memory as three int8 digit planes of a 23-bit fixed-point image (one power-of-two scale per column);
 * the spikes of step t-1 are expanded from the exchanged bit words into a uint8 A operand in TENSOR MEMORY
 * (tcgen05.st) and multiplied by tcgen05.mma.kind::i8 (M = 128, N = 48, exact int32 sums in TMEM).  H <=
 * sparch_recur_fwd_tc_max_h().  img: sparch_recur_fwd_tc_image_bytes(H) bytes filled by
 * sparch_recur_prepare_fwd_tc (digit planes + column scales).  bits: sparch_recur_fwd_tc_bits_bytes() bytes,
 * receives the packed spike tensor [T][ceil(Be/128)][ceil(H/16)][128] of 32-bit words {0xFF, spikes 8..15, 0xFF,
 * spikes 0..7} (zeroed here; these words are also the inter-CTA exchange).  rec0 = s0 @ V0 as for
 * sparch_recur_fwd.  reduced != 0: two digit planes (15-bit image of V0).  Writes the tapes S, U (W).           */
SPARCH_API int sparch_recur_fwd_tc_max_h(void);
SPARCH_API size_t sparch_recur_fwd_tc_image_bytes(int H);
SPARCH_API size_t sparch_recur_fwd_tc_bits_bytes(int Be, int T, int H);
SPARCH_API int sparch_recur_prepare_fwd_tc(const float* V, int H, void* img, sparch_stream_t st);
SPARCH_API int sparch_recur_fwd_tc(int kind, const float* Z, const float* scale, const float* shift,
                                   const float* alpha, const float* beta, const float* a, const float* b,
                                   const float* rec0, const void* img, const float* u0, const float* w0,
                                   const float* s0, float theta, float* S, float* U, float* W,
                                   uint32_t* bits, int reduced, int Be, int T, int H, sparch_stream_t st);
/* Bidirectional layers (snns.py:666-668, 686-689) WITHOUT the flipped / concatenated copies: Z is the projection of the
 * un-flipped batch, (rev_from, T, H); the kernel runs Be = 2 * rev_from rows, rows b >= rev_from read
 * Z[b - rev_from][T - 1 - t] (the reversed sequence).  Tapes and planes stay in the kernel's own (Be, T, H) order, which
 * is what sparch_recur_bwd_tc reads.  rev_from = 0: one direction.  Needs rev_from % 128 == 0, H % 4 == 0.
 * w_every = C > 0: W is a CHECKPOINT tape (Be, ceil(T/C), H) that receives w_t at the last step of every
 * chunk of C steps (and at T-1) instead of the full (Be, T, H) tape; sparch_recur_bwd_tc_ck recomputes the steps in
 * between.  rev_from = w_every = 0 is sparch_recur_fwd_tc.                                                         */
SPARCH_API int sparch_recur_fwd_tc_bidir(int kind, const float* Z, const float* scale, const float* shift,
                                         const float* alpha, const float* beta, const float* a, const float* b,
                                         const float* rec0, const void* img, const float* u0, const float* w0,
                                         const float* s0, float theta, float* S, float* U, float* W,
                                         uint32_t* bits, int reduced, int Be, int T, int H, int rev_from,
                                         int w_every, sparch_stream_t st);
/* Profiling aid: device buffer of T*4 int64 that receives, per timestep, the SM clock of CTA (0,0)
 * after the spike-word wait, after the MMA loop, after the reduction and at the end of the step
 * for the following sparch_recur_fwd launches (NULL switches it off).                          */
SPARCH_API int sparch_recur_debug_clocks(long long* buf);
/* Reverse pass of an RLIF/RadLIF layer, all T steps in one persistent cooperative kernel: dI
 * (Be,T,H) and the per-(b,h) partial parameter gradients p_* (Be,H).  img_bwd from
 * sparch_recur_prepare; workspace of sparch_recur_bwd_workspace() bytes holds the
 * block-floating-point dI panels handed from step to step; sync_ws is
 * sparch_recur_sync_words() ints (one arrival counter per 64-row group, zeroed here).          */
SPARCH_API int sparch_recur_sync_words(int Be);
SPARCH_API size_t sparch_recur_bwd_workspace(int Be, int H);
SPARCH_API int sparch_recur_bwd(int kind, const float* G, const float* U, const float* W,
                                const float* alpha, const float* beta, const float* a,
                                const float* b, const uint32_t* img_bwd, const int* meta,
                                const float* u0, const float* w0, const float* s0, float theta,
                                float* dI, float* p_alpha, float* p_beta, float* p_a,
                                float* p_b, void* workspace, int* sync_ws, int reduced, int Be,
                                int T, int H, sparch_stream_t st);

/* Reverse pass of an RLIF/RadLIF layer with dI_{t+1} @ V0^T on tcgen05: clusters of 4 CTAs split K,
 * A panels as K-major fp16 hi/lo matrices fetched by TMA, D in TMEM, partial products reduce-scattered
 * through distributed shared memory, one per-row scale derived one step late.  H <= 1024.  Same
 * results contract as sparch_recur_bwd (reduced != 0: hi terms only).  img from sparch_recur_prepare_tc
 * (sparch_recur_bwd_tc_image_bytes bytes), meta from sparch_recur_prepare, workspace of
 * sparch_recur_bwd_tc_workspace bytes.  gmax_in (Be*T floats, may be NULL): row maxima of |G| when the
 * producer of G already computed them (sparch_spike_post_bwd); NULL = computed here.              */
SPARCH_API int sparch_recur_tc_padded(int H);
SPARCH_API size_t sparch_recur_bwd_tc_image_bytes(int H);
SPARCH_API size_t sparch_recur_bwd_tc_workspace(int Be, int T, int H);
SPARCH_API int sparch_recur_prepare_tc(const float* V, int H, void* img, const int* meta,
                                       sparch_stream_t st);
SPARCH_API int sparch_recur_bwd_tc(int kind, const float* G, const float* U, const float* W,
                                   const float* alpha, const float* beta, const float* a,
                                   const float* b, const void* img, const int* meta,
                                   const float* u0, const float* w0, const float* s0, float theta,
                                   float* dI, float* p_alpha, float* p_beta, float* p_a,
                                   float* p_b, void* workspace, int reduced, int Be, int T, int H,
                                   const float* gmax_in, sparch_stream_t st);
/* The same pass reading the CHECKPOINT adaptation tape of sparch_recur_fwd_tc_bidir(w_every = C): W is (Be, ceil(T/C), H),
 * w_{t-1} is loaded where step t-1 closes a chunk and recomputed in between by solving snns.py:718 for w_{t-1} (beta >=
 * e^(-1/30): the rounding error of a step back grows by <= 1.034 and is dropped at the next checkpoint; only d(beta)
 * reads w).  w_every = 0: the full tape, i.e. sparch_recur_bwd_tc.                                              */
SPARCH_API int sparch_recur_bwd_tc_ck(int kind, const float* G, const float* U, const float* W,
                                   const float* alpha, const float* beta, const float* a,
                                   const float* b, const void* img, const int* meta,
                                   const float* u0, const float* w0, const float* s0, float theta,
                                   float* dI, float* p_alpha, float* p_beta, float* p_a,
                                   float* p_b, void* workspace, int reduced, int Be, int T, int H,
                                   const float* gmax_in, int w_every, sparch_stream_t st);

/* ---- around the recurrence: dropout, firing-rate counts, operand terms, parameter clamps ------- */
/* One pass over the spike tensor S (M = Be*T rows, H columns, values exactly 0/1):
 *   out   (M,H) fp32 = dropout(S) with keep probability 1-p_drop, kept values scaled by 1/(1-p_drop)
 *         (snns.py:692; NULL allowed when p_drop == 0: the output is S itself);
 *   term  (M, ld = H rounded up to 8) 16-bit {0,1}: 1 where out != 0 -- the exact one-term operand of the
 *         next layer's projection GEMM (fp16 if fp16_terms else bf16; the factor 1/(1-p) goes in alpha);
 *   sterm same for S itself -- the S_prev operand of this layer's dV GEMM; term/sterm may be NULL;
 *   counts (H) int32: number of non-zero outputs per neuron (zeroed here) -- firing rate = counts *
 *         1/(1-p) / M (snns.py:174), exact and order-independent.
 * The mask comes from Philox4x32-10 keyed by *seed (device uint64) and the element's position, so
 * sparch_spike_post_bwd regenerates it instead of reading a stored mask.                         */
SPARCH_API int sparch_spike_post_fwd(const float* S, int64_t M, int H, float p_drop, const void* seed,
                                     float* out, void* term, void* sterm, int fp16_terms, int* counts,
                                     sparch_stream_t st);
/* The same pass fed by the PACKED spike planes the tcgen05 forward recurrence published (sparch_recur_fwd_tc called with
 * S = NULL: [T][group of 128 rows][slice of 16 neurons][128 rows] words, spike i of a slice at bit 2 i) instead of an
 * fp32 spike tensor: the spikes cross HBM as 0.25 B/elt.  out (Be,T,H) is always written (the layer's output);
 * s_last (Be,H), optional, receives the spikes of the last step (operand of sparch_dv_boundary).              */
SPARCH_API int sparch_spike_post_fwd_bits(const uint32_t* bits, int Be, int T, int H, float p_drop, const void* seed,
                                          float* out, void* term, void* sterm, int fp16_terms, int* counts,
                                          float* s_last, sparch_stream_t st);
/* ... and with the bidirectional merge fused (snns.py:686-689: cat([s_f, flip(s_b, time)], feature)): rows b >= rev_from
 * of the planes are the reversed pass of row b - rev_from; out (rev_from, T, 2H), term (rev_from*T, 2*ld) and counts
 * (2H) are written in the merged layout and the dropout mask is keyed by the merged position; sterm (Be*T, ld) and
 * s_last (Be, H) stay in the recurrence's own order (operands of this layer's dV).  H % 8 == 0.              */
SPARCH_API int sparch_spike_post_fwd_bits_bidir(const uint32_t* bits, int Be, int T, int H, float p_drop,
                                                const void* seed, float* out, void* term, void* sterm,
                                                int fp16_terms, int* counts, float* s_last, int rev_from,
                                                sparch_stream_t st);
/* Backward of that merge + dropout: G (B, T, 2H) -> GS (2B, T, H) in the recurrence's order (columns >= H go to row
 * b + B at time T-1-t), mask regenerated from the merged position (p_drop = 0: no mask, seed may be NULL); gmax
 * (2B*T floats, may be NULL) as for sparch_spike_post_bwd.                                                     */
SPARCH_API int sparch_spike_post_bwd_bidir(const float* G, int B, int T, int H, float p_drop, const void* seed,
                                           float* GS, float* gmax, sparch_stream_t st);
/* GS = G * mask / (1-p_drop) with the same mask; gmax (M floats, may be NULL; zeroed here) receives the
 * row maxima of |GS| (input of sparch_recur_bwd_tc).                                              */
SPARCH_API int sparch_spike_post_bwd(const float* G, int64_t M, int H, float p_drop, const void* seed,
                                     float* GS, float* gmax, sparch_stream_t st);
/* out[k][h] = clamp(p_k[h], lims[2k], lims[2k+1]) for k < nk (nk = 1: alpha only; 4: alpha, beta, a, b;
 * snns.py:706-709).  lims is a HOST array of 8 floats.                                           */
SPARCH_API int sparch_neuron_params(const float* alpha, const float* beta, const float* a, const float* b,
                                    const float* lims, int nk, int H, float* out, sparch_stream_t st);
/* grads[k][h] = (sum_b part[k][b][h]) * [lims[2k] <= p_k[h] <= lims[2k+1]]: batch reduction of the
 * per-(b,h) partial parameter gradients and the clamp's backward.  lims is a HOST array.          */
SPARCH_API int sparch_param_grads(const float* part, const float* alpha, const float* beta, const float* a,
                                  const float* b, const float* lims, int nk, int Be, int H, float* grads,
                                  sparch_stream_t st);

/* C (M,N) = [C +] sum_k A(m,k) B[k][n] in fp32 FFMA for the Be-sized products of a recurrent layer: rec_0 = s0 @ V0
 * (snns.py:702, 720: A = s0 (M = Be, K = H), B = V, flags = 1) and the t = 0 frames of dV (A = sparch_dv_boundary's
 * output stored (K = Be, M = H): a_km = 1, B = dI[:, 0, :] with ldb = T * H, flags = 2 | 4).  a_km: A is (K, M)
 * row-major.  flags: 1 read B with a zero diagonal, 2 zero the diagonal of C, 4 accumulate into C, 8 B is stored
 * (N, K) row-major (product with B^T: the stepwise reverse pass dI_{t+1} @ V0^T of hidden sizes beyond the persistent kernels). */
SPARCH_API int sparch_small_gemm(const float* A, int64_t lda, int a_km, const float* B, int64_t ldb, float* C,
                                 int64_t ldc, int M, int N, int K, int flags, sparch_stream_t st);
/* Small helpers of the recurrent layers.  sparch_recur_v0: V0 = V with a zero diagonal (snns.py:712, 566).
 * sparch_zero_diag: A[i][i] = 0 (the backward of that masking, applied to dV).  sparch_dv_boundary: first[b] = s0[b] -
 * S[b-1][T-1] (b > 0), S (Be,T,H): the dV GEMM pairs frame m of dI with frame m-1 of the (b,t)-flattened spikes, this
 * is the correction operand of the Be frames t = 0, whose true partner is the real-valued s0 (snns.py:702).        */
SPARCH_API int sparch_recur_v0(const float* V, int H, float* V0, sparch_stream_t st);
SPARCH_API int sparch_zero_diag(float* A, int H, sparch_stream_t st);
SPARCH_API int sparch_dv_boundary(const float* s0, const float* S, int Be, int T, int H, float* first,
                                  sparch_stream_t st);

/* ---- train-step glue: Adam, exp.py:89 ------------------------------------------- */
/* One launch updates up to 48 parameter tensors (torch.optim.Adam defaults: no weight decay, no amsgrad):
 * g' = g * hyper[4]; m += (1-b1)(g'-m); v = b2 v + (1-b2) g'^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps),
 * t = *step (device int64, already incremented by the caller).  hyper = DEVICE float[5] {lr, beta1, beta2, eps,
 * gradient scale}: a scheduler (exp.py:92-96) changes lr between replays of a captured step without re-capture; the
 * gradient scale is 1/world_size when the gradients hold an all-reduced sum.  params/grads/exp_avg/exp_avg_sq/numel
 * are HOST arrays of `count` device pointers / element counts.                                              */
SPARCH_API int sparch_adam_step(int count, float* const* params, const float* const* grads,
                                float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel,
                                const int64_t* step, const float* hyper, sparch_stream_t st);

/* ---- ReadoutLayer cell (snns.py:807-825) ---------------------------------------------- */
/* u_t = alpha*u_{t-1} + (1-alpha)*I_t ; out = sum_t softmax(u_t, dim=1).  U (B,T,C) tape.  */
SPARCH_API int sparch_readout_fwd(const float* Z, const float* scale, const float* shift,
                       const float* alpha, const float* u0, float* out, float* U, int B, int T,
                       int C, sparch_stream_t st);
/* gout (B,C) -> dI (B,T,C), p_alpha (B,C) partials.                                       */
SPARCH_API int sparch_readout_bwd(const float* gout, const float* U, const float* alpha, const float* u0,
                       float* dI, float* p_alpha, int B, int T, int C, sparch_stream_t st);

/* ---- loss of the train step (exp.py:83 nn.CrossEntropyLoss(), exp.py:362) -------------------------------- */
/* *loss = mean_b(logsumexp(logits[b]) - logits[b][target[b]]), lse[b] kept for the backward; targets outside
 * [0, C) contribute zero.  dlogits = (softmax - onehot) * *gloss / B.  One launch each instead of ATen's
 * log_softmax + nll_loss pairs.                                                                              */
SPARCH_API int sparch_ce_fwd(const float* logits, const int64_t* target, int B, int C, float* loss, float* lse,
                             sparch_stream_t st);
SPARCH_API int sparch_ce_bwd(const float* logits, const int64_t* target, const float* lse, const float* gloss, int B,
                             int C, float* dlogits, sparch_stream_t st);

/* ---- on-device data path of the spiking datasets (sparch/dataloaders/spiking_datasets.py:66-78) ------------- */
/* ---- initial states (snns.py:286-287, 423-425, 558-559, 700-702, 812) ---------------------------------------------
 * The reference draws them with torch.rand on the default CPU generator: MT19937, one 32-bit word per float32,
 * u = (word & (2^24 - 1)) * 2^-24, in element order.  This produces the same n numbers on the device from the generator's
 * 624-word state (state_in, DEVICE pointer) and position pos (0..624: words of the current block already used), and
 * writes the state afterwards to state_out[0..623] and the new position to state_out[624]; the caller puts them back
 * into the generator (torch.set_rng_state), which is then exactly where n host draws would have left it.      */
SPARCH_API int sparch_mt19937_uniform(const uint32_t* state_in, int pos, int64_t n, float* out, uint32_t* state_out,
                                      sparch_stream_t st);

/* A batch of B event lists, concatenated: times[e] (seconds, fp32), units[e] (int32), offsets[b] .. offsets[b+1] the
 * events of example b (offsets[B] = nev).  dense (B, nb_steps, nb_units) fp32 is zeroed and receives, per example, the
 * reference's x.to_dense(): the NUMBER of events of unit u whose time falls into bin np.digitize(t, bins) (bins:
 * nb_steps increasing float64 edges, np.linspace(0, max_time, nb_steps), compared in float64 as numpy does).  *bad
 * (device int) becomes non-zero if an event lies at or beyond the last edge, has a unit outside [0, nb_units) or a NaN
 * time -- where the reference's sparse-tensor constructor raises.                                                   */
SPARCH_API int sparch_events_to_dense(const float* times, const int* units, const int64_t* offsets, const double* bins,
                                      int B, int nb_steps, int nb_units, int64_t nev, float* dense, int* bad,
                                      sparch_stream_t st);

#ifdef __cplusplus
}
#endif
#endif /* SPARCH_B200_H */
