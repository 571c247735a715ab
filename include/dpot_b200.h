/*
 * dpot_b200.h -- C ABI of libdpot_b200.so: the B200 (sm_100a) implementation of the DPOT
 * autoregressive Fourier-operator hot path (forward / backward / optimizer step).
 *
 * The reference (HaoZhongkai/DPOT) is pure Python on this path and has no FFI of its own;
 * each entry point below therefore names the reference *Python* interface (file:line under
 * /root/reference) whose arithmetic it replaces.  The Python binding a maintainer adds is the
 * ctypes stub shown in INTEGRATION.md (dpot_b200/_lib.py is that stub).
 *
 * Conventions
 *  - plain C: pointers, ints, floats; no torch / C++ types.  `stream` is a cudaStream_t
 *    passed as void* (NULL = legacy default stream).
 *  - every device pointer is BORROWED for the duration of the stream-ordered work enqueued by
 *    the call; the library never allocates or frees device memory and never synchronises,
 *    so every entry point is CUDA-graph capturable and re-entrant.
 *  - return value: 0 = ok; >0 = cudaError_t; <0 = argument error (DPOT_E_*).  The message
 *    of the last failure on the calling thread: dpot_last_error_string().
 *  - all tensors are fp32 and contiguous unless a leading-dimension argument says otherwise.
 *  - latent tensors are token-major ("NHWC"): a[B*n, E], n = h*w latent cells, row = b*n + p*w + q.
 */
#ifndef DPOT_B200_H_
#define DPOT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPOT_ABI_VERSION 2

#if defined(__GNUC__)
#define DPOT_API __attribute__((visibility("default")))
#else
#define DPOT_API
#endif

/* argument-error codes */
#define DPOT_E_BADARG   (-1)
#define DPOT_E_UNSUPPORTED (-2)
#define DPOT_E_ALIGN    (-3)

/* activation ids, order of ACTIVATION in models/dpot.py:19 */
enum {
  DPOT_ACT_NONE = -1,
  DPOT_ACT_GELU = 0, DPOT_ACT_TANH = 1, DPOT_ACT_SIGMOID = 2, DPOT_ACT_RELU = 3,
  DPOT_ACT_LEAKY_RELU = 4, DPOT_ACT_SOFTPLUS = 5, DPOT_ACT_ELU = 6, DPOT_ACT_SILU = 7
};

/* GEMM engines */
enum { DPOT_GEMM_AUTO = 0, DPOT_GEMM_SIMT = 1, DPOT_GEMM_TC = 2, DPOT_GEMM_TC16 = 3 };

/* Operand storage formats.  DPOT_FMT_HL16 ("split fp16") stores an fp32 value x as two halves
 *   hi = fp16_rn(x),  lo = fp16_rn((x - hi) * 2048)        (x ~= hi + lo / 2048 to ~2^-22 relative)
 * in the SAME 4 bytes per element as fp32: element (r, k) of a matrix has hi at half-index
 * r*ld + k and lo at r*ld + k + lo_off (ld, lo_off and batch strides counted in halves).  It is
 * the native operand format of the DPOT_GEMM_TC16 engine (tcgen05 kind::f16, 3 MMAs per product:
 * hi*hi into one TMEM accumulator, hi*lo + lo*hi into a second one scaled by 2^-11), written
 * directly by the kernels that produce activations, so the GEMM needs no conversion pass.
 * Range: |x| <= 65504 (fp16); larger magnitudes become inf/NaN (loud, never silent). */
enum { DPOT_FMT_F32 = 0, DPOT_FMT_HL16 = 1,
       /* output only (f16-split engine): split fp16 interleaved per group of 32 columns -- element (m, n) at half index
          m*ldc + (n/32)*64 + n%32 (hi) and +32 (lo): a 128-byte [hi 32 | lo 32] record per (row, 32-column group), the
          operand layout of the tcgen05 output-tail kernel (one record = one pixel's out_layer_dim = 32 channels) */
       DPOT_FMT_HL16G32 = 2 };

DPOT_API int         dpot_abi_version(void);
DPOT_API const char* dpot_last_error_string(void);
/* 1 if the tcgen05 engine can serve this device (sm_100) */
DPOT_API int         dpot_device_supported(void);
/* number of CUDA kernels this library has launched in this process (all threads) */
DPOT_API long long   dpot_launch_count(void);
/* 1 if the tcgen05 3xTF32 GEMM engine is compiled in and the current device can run it */
DPOT_API int         dpot_tc_available(void);
/* tuning knob of the tcgen05 engine: k-blocks (32 fp32) accumulated in TMEM between round-to-nearest
   flushes into the register accumulators (default 2); returns the previous value, <1 only queries */
DPOT_API int         dpot_tc_set_flush(int kblocks);
/* tuning knob: 1 = the hi operand is the hardware truncation of the raw fp32 tile (converter writes only lo),
   0 = explicit round-to-nearest hi; returns the previous value, <0 only queries */
DPOT_API int         dpot_tc_set_trunc(int on);
/* profiling aid: device buffer of 7*64 int64 that receives clock64() pipeline events of CTA 0 of the
   tcgen05 engine (rows: producer, mma, conv-start, conv-done, flush, tile-acc-done, tile-epilogue-done) */
DPOT_API void        dpot_tc_set_trace(long long* dev_buf);

/* ------------------------------------------------------------------------------------------
 * The dense-contraction engine.  C = epilogue( A' * W^T ), fp32 in / fp32 out.
 *   A'[m,k] = A[m,k] * a_scale[s,k] + a_shift[s,k]   (s = m / a_rows_per_sample; tables optional)
 *   v       = sum_k A'[m,k] W[n,k] + bias[n] + rowbias[(m % rowbias_period), n]
 *   v       = act(v)
 *   v       = v * c_scale[s', n] + c_shift[s', n]      (s' = m / c_rows_per_sample; optional)
 *   C[m,n]  = v + residual[m,n]
 * A is [M,K] row-major (or the im2col view of a field, see a_mode), W is [N,K] row-major --
 * the layout torch keeps Conv2d(1x1)/Linear weights in.  `batch` independent problems are
 * addressed by the stride* fields (block-diagonal AFNO weights).
 * Replaces: torch.einsum / Conv2d(1x1) / Linear / ConvTranspose2d call sites
 * models/dpot.py:72-94 (AFNO block MLP), :157-161 (channel MLP), :198-202 (PatchEmbed),
 * :228-232 (TimeAggregator), :303-309 (cls head), :315-321 (out layer).
 * ---------------------------------------------------------------------------------------- */
enum { DPOT_A_PLAIN = 0, DPOT_A_PATCH = 1 };

typedef struct dpot_gemm_args {
  const float* A;  int64_t lda;          /* DPOT_A_PLAIN: [M,K]; DPOT_A_PATCH: field x[B,X,Y,T,C] */
  const float* W;  int64_t ldw;          /* [N,K] */
  float*       C;  int64_t ldc;
  int32_t M, N, K;
  const float* bias;                     /* [N] or NULL */
  const float* rowbias; int32_t rowbias_period; int64_t ldrb;   /* [period, N] or NULL */
  const float* residual; int64_t ldr;    /* [M,N] or NULL (may alias C) */
  int32_t act;                           /* DPOT_ACT_* */
  const float* a_scale; const float* a_shift; int32_t a_rows_per_sample;  /* [*,K] tables or NULL */
  const float* c_scale; const float* c_shift; int32_t c_rows_per_sample;  /* [*,N] tables or NULL */
  /* output row remap: row m is stored at C + (m / c_group) * c_group_stride + (m % c_group) * ldc;
     c_group = 0 disables it */
  int32_t c_group; int64_t c_group_stride;
  int32_t batch; int64_t strideA, strideW, strideC, strideBias;
  /* DPOT_A_PATCH geometry: row m = (b, p, q, t), k = (u, v, c):
     A[m,k] = x[b, p*P+u, q*P+v, t, c]   (PatchEmbed im2col, models/dpot.py:199,375) */
  int32_t a_mode, pX, pY, pT, pC, pP;
  int32_t engine;                        /* DPOT_GEMM_* */
  /* optional GroupNorm statistics of the stored result: out_stats[s, g, 2] (double) receives
     (sum, sum of squares) over the entries of sample s = m / stats_rows_per_sample and channel
     group g = n / (N / stats_groups).  Zeroed by dpot_gemm itself.  Requires batch == 1. */
  double* out_stats; int32_t stats_groups; int32_t stats_rows_per_sample;
  /* training extras.  C_pre: also store the pre-activation value (after bias/rowbias), same layout as C.
     dact_src/dact: multiply by act'(dact_src[m,n]) of activation `dact` before c_scale/residual (chain rule
     through the activation of the previous layer).  c_mode = DPOT_A_PATCH scatters row m = (b,p,q,t),
     column n = (u,v,c) to the field layout x[b, p*P+u, q*P+v, t, c] (p* geometry; gradient w.r.t. the
     input field, the transpose of the im2col load). */
  float* C_pre;
  const float* dact_src; int32_t dact;
  int32_t c_mode;
  /* storage formats (DPOT_FMT_*).  a_fmt/w_fmt = HL16 select the DPOT_GEMM_TC16 engine (both must be
     HL16; then A/W point to halves and lda/ldw/strideA/strideW/*_lo_off are in halves).  c_fmt = HL16
     makes the epilogue store the result in split form (ldc/strideC/c_group_stride/c_lo_off in halves);
     supported by the SIMT and TC16 engines.  out_stats, C_pre, dact_src, c_mode=PATCH need c_fmt = F32
     except out_stats on TC16. */
  int32_t a_fmt, w_fmt, c_fmt;
  int64_t a_lo_off, w_lo_off, c_lo_off;
  /* ---- ABI version 2: the backward-pass forms of the DPOT_GEMM_TC16 engine (zero = the plain forward form) ----
     a_trans / w_trans = 1: the operand is STORED TRANSPOSED, A as [K, M] (row stride lda) / W as [K, N] (row stride
     ldw); the tensor core reads it as an MN-major tile, so neither the data-gradient (dx = g W: w_trans) nor the
     weight-gradient (dW = g^T x: both) needs a transposed copy.  With a_trans the token tile is 64 or 128 rows.
     k_split > 1: the contraction is cut into k_split chunks of k_chunk (a multiple of 64); chunk c writes its partial
     result at C + c * strideC_split (a tensor-core accumulation chain is kept <= 1024 deep; the chunks
     also fill the SMs when the output is small).  No bias / activation / side inputs with k_split > 1.
     ld_pre / stride_pre, ld_dact / stride_dact: leading dimension and batch stride (floats) of C_pre / dact_src on the
     TC16 engine (there they are allowed with c_fmt = HL16: the fc2 data-gradient leaves the engine already multiplied
     by act'(pre) and split, ready to be the operand of the next two contractions). */
  int32_t a_trans, w_trans;
  int32_t k_split; int64_t k_chunk; int64_t strideC_split;
  int64_t ld_pre, stride_pre, ld_dact, stride_dact;
  /* TC16 engine: out_colsum[b * N + n] += sum_m C[m, n] of problem b (double accumulation of the stored values; NOT
     zeroed by the call) -- the bias gradient of the layer whose data gradient this contraction produces. */
  double* out_colsum;
} dpot_gemm_args;

DPOT_API int dpot_gemm(const dpot_gemm_args* args, void* stream);

/* fp32 [rows, cols] (leading dimension lds floats) -> split fp16 (DPOT_FMT_HL16): dst half-index
   r*ldd + c (hi) and r*ldd + c + lo_off (lo).  Optional per-(sample, column) affine applied first:
   x' = x*scale[s, c] + shift[s, c], s = r / rows_per_sample (GroupNorm-apply fused into the split;
   models/dpot.py:175).  cols must be a multiple of 8, pointers 16-byte aligned. */
DPOT_API int dpot_split_f16(const float* src, int64_t lds, int64_t rows, int32_t cols, const float* scale,
                   const float* shift, int32_t rows_per_sample, void* dst, int64_t ldd, int64_t lo_off, void* stream);
/* 1 if the f16-split tcgen05 engine can serve this device */
DPOT_API int dpot_tc16_available(void);
/* tile-plan knob of the f16-split engine (tests / experiments): -1 = auto (cost model), 0 = single-CTA tiles only,
   1 = CTA pairs (tcgen05 cta_group::2, UMMA M = 256) whenever N > 128 */
DPOT_API void dpot_tc16_set_pair(int32_t mode);
/* programmatic dependent launch between the kernels of the forward chain: 0 = plain stream order (default), 1 = on */
DPOT_API void dpot_set_pdl(int32_t on);
/* Operand precision of the f16-split engine: 0 (default) = fp32-faithful, three MMAs per product on the hi / lo planes;
   1 = HALF-PRECISION OPERAND MODE -- only the hi planes (fp16, 11-bit significand: finer than bf16's 8, same tensor-core
   rate) are loaded and multiplied, one MMA per product, fp32 accumulation, fp32 master weights and epilogues unchanged.
   This is the library's 16-bit mixed-precision mode (the reference's bf16 autocast configs, configs/pretrain_medium.yaml);
   results differ from fp32 at the 1e-3 level.  Returns the previous mode; mode < 0 only queries. */
DPOT_API int dpot_tc16_set_precision(int32_t mode);
/* Accumulation-chain limit of the f16-split engine.  A tcgen05 fp32 accumulator truncates on every step, so the error
   of one launch grows linearly with the contraction depth (1.2e-9 * K rel-L2, measured).  dpot_gemm therefore runs a
   plain forward-form contraction (batch 1, fp32 result) deeper than k_chain halves as ceil(K / k_chain) launches over
   K-ranges whose fp32 partial sums pass through the result buffer (round-to-nearest adds in the epilogue); the last
   launch applies the epilogue of the call.  Default 2048; 0 = never chain; negative = query.  Returns the previous
   value.  Serves the channel MLP's fc2 (models/dpot.py:160) of DPOT-M / L / H (4096 / 6144 / 8192 deep). */
DPOT_API int dpot_tc16_set_chain(int32_t k_chain);
/* dpot_gemm with an explicit chain limit and partial-sum buffer scratch[M, ld_scratch] (fp32, may be args->C when the
   result is plain fp32): for results stored in another format (split fp16).  K <= k_chain or k_chain <= 0: plain dpot_gemm. */
DPOT_API int dpot_gemm_chained(const dpot_gemm_args* args, int32_t k_chain, float* scratch, int64_t ld_scratch, void* stream);
/* pipeline-isolation experiments (results are garbage when non-zero): 1 = no TMA loads, 2 = no MMAs, 4 = no epilogue */
DPOT_API void dpot_tc16_set_debug(int32_t mask);

/* ------------------------------------------------------------------------------------------
 * GroupNorm pieces (torch.nn.GroupNorm(8,E), models/dpot.py:142,152,167,175).
 * stats[B, groups, 2] are double (sum, sum of squares) over the (E/groups)*n values of a group.
 * dpot_gn_stats zeroes `stats` itself.  dpot_gn_finalize turns them into per-(sample, channel)
 * affine tables: scale = rstd*gamma, shift = beta - mean*rstd*gamma, consumed by the
 * a_scale/a_shift prologue of dpot_gemm and by the AFNO spectral kernels.
 * ---------------------------------------------------------------------------------------- */
DPOT_API int dpot_gn_stats(const float* x, int32_t B, int32_t n, int32_t E, int32_t groups, double* stats, void* stream);
DPOT_API int dpot_gn_finalize(const double* stats, const float* gamma, const float* beta, int32_t B, int32_t n,
                     int32_t E, int32_t groups, float eps, float* scale, float* shift, void* stream);

/* ------------------------------------------------------------------------------------------
 * AFNO2D spectral transforms (models/dpot.py:59 rfft2 ortho, :102 irfft2 ortho, :106 skip).
 * fwd:  n1 = a*scale+shift (GroupNorm-1 applied on load); S = rfft2_ortho(n1) restricted to the kept
 *       modes k1<km1, k2<km2, written as S[(b,k1,k2), kappa, {re[bs] | im[bs]}]  (row length 2E).
 * inv:  f = irfft2_ortho(O2 zero-padded to [h, h/2+1]) + n1, c2r ignoring Im of k2 in {0, h/2};
 *       also accumulates GroupNorm-2 statistics of f into stats_out[B,groups,2] (must be zeroed
 *       by the caller; NULL to skip).
 * h must be a power of two in [2, 32].
 * ---------------------------------------------------------------------------------------- */
DPOT_API int dpot_afno_fft_fwd(const float* a, const float* scale, const float* shift, int32_t B, int32_t h,
                      int32_t E, int32_t nb, int32_t km1, int32_t km2, float* S, float interior_weight, void* stream);
DPOT_API int dpot_afno_fft_inv(const float* O2, const float* a, const float* scale, const float* shift, int32_t B,
                      int32_t h, int32_t E, int32_t nb, int32_t km1, int32_t km2, float* f,
                      double* stats_out, int32_t groups, float interior_weight, void* stream);
/* fwd with the spectrum stored as split fp16 (DPOT_FMT_HL16): S16 row (b,k1,k2) = [hi: 2E halves | lo: 2E halves] */
DPOT_API int dpot_afno_fft_fwd16(const float* a, const float* scale, const float* shift, int32_t B, int32_t h,
                        int32_t E, int32_t nb, int32_t km1, int32_t km2, void* S16, void* stream);
/* ------------------------------------------------------------------------------------------
 * The fused AFNO2D mixer: GroupNorm-1 by reference -> rfft2 -> block-diagonal complex MLP (two layers, tcgen05) ->
 * irfft2 -> + skip -> f, with the GroupNorm-2 statistics of f, in ONE kernel per block; the spectrum and the hidden
 * layer stay in shared memory / TMEM.  Replaces AFNO2D.forward, models/dpot.py:51-110, together with the GroupNorm
 * calls around it (:167,175).  Geometry served: latent 16 x 16, block size 128 (E = 128 * nb), no mode truncation --
 * DPOT-Ti/S/M at 128^2 -- query dpot_afno_fused_supported; other geometries run the four-kernel path
 * (dpot_afno_fft_fwd16_gn -> 2 x dpot_gemm -> dpot_afno_fft_inv_gn).
 *   packed: dpot_afno_fused_packed_floats(nb) floats written by dpot_afno_fused_pack from the block's
 *           filter.w1/b1/w2/b2 (three pre-swizzled fp16 planes of s*w per 16 KB chunk, biases, 1/s per layer).
 *   stats1: GroupNorm-1 statistics [B, groups, 2] (double: sum, sum of squares) of `lat`; stats2 (may be NULL):
 *           receives the statistics of f, ACCUMULATED (zero it first).  dbg: NULL (test hook: operand-tile images).
 * ---------------------------------------------------------------------------------------- */
DPOT_API int     dpot_afno_fused_supported(int32_t h, int32_t E, int32_t nb, int32_t km1, int32_t km2, int32_t groups);
DPOT_API int64_t dpot_afno_fused_packed_floats(int32_t nb);
DPOT_API int     dpot_afno_fused_pack(const float* w1, const float* b1, const float* w2, const float* b2, int32_t nb,
                                      int32_t bs, float* packed, void* stream);
DPOT_API int     dpot_afno_fused(const float* lat, const double* stats1, const float* gamma1, const float* beta1,
                                 int32_t groups, float eps, int32_t B, int32_t h, int32_t E, int32_t nb, const float* packed,
                                 int32_t act, float* f, double* stats2, float* dbg, void* stream);
/* The same launch with GroupNorm-2 applied inside: a work unit (sample, channel block of 128) covers whole GroupNorm groups
   (group size 32 / 64 / 128), so once its 16 warps have met the unit normalises its f values -- kept in shared memory, each
   thread's outputs in the slots of the inverse-transform scratch it has just consumed -- and stores n2 = GN2(f) as split
   fp16 (rows [hi E | lo E]): the channel MLP's operand, without the separate dpot_split_f16_gn pass (models/dpot.py:175).
   f may then be NULL (inference: nothing else reads it, so it never goes to global memory).  n2_16 = NULL: exactly
   dpot_afno_fused.  stats2 may be NULL. */
DPOT_API int     dpot_afno_fused_gn2(const float* lat, const double* stats1, const float* gamma1, const float* beta1,
                                     int32_t groups, float eps, int32_t B, int32_t h, int32_t E, int32_t nb, const float* packed,
                                     int32_t act, float* f, double* stats2, float* dbg, void* n2_16, const float* gamma2,
                                     const float* beta2, float eps2, void* stream);
/* knob (tests / A-B measurements): 1 = dpot_forward* lets the fused mixer apply GroupNorm-2 (default: +1.5 % on the DPOT-S
   rollout), 0 = separate dpot_split_f16_gn pass, negative = query.  Returns the previous value. */
DPOT_API int     dpot_afno_set_fused_gn2(int32_t on);
/* knob (tests / A-B measurements): -1 = fused whenever supported (default), 0 = never */
DPOT_API void    dpot_afno_set_fused(int32_t mode);
/* profiling aid: device buffer of 8 int64 per unit of CTA 0 receiving clock64() at the phase boundaries of the fused
   mixer (start, spectrum ready, layer-1 done, hidden ready, layer-2 done, column inverse done, unit done); NULL = off */
DPOT_API void    dpot_afno_fused_set_trace(long long* dev_buf);

/* GroupNorm-by-reference variants (the f16-split inference pipeline): instead of finalised scale/shift tables the
   kernels take the raw statistics [B, groups, 2] (double sum, sum of squares) plus gamma/beta and derive the
   per-channel affine themselves (no dpot_gn_finalize launch). */
DPOT_API int dpot_afno_fft_fwd16_gn(const float* a, const double* stats1, const float* gamma1, const float* beta1,
                           int32_t groups, float eps, int32_t B, int32_t h, int32_t E, int32_t nb, int32_t km1,
                           int32_t km2, void* S16, void* stream);
DPOT_API int dpot_afno_fft_inv_gn(const float* O2, const float* a, const double* stats1, const float* gamma1,
                         const float* beta1, int32_t groups, float eps, int32_t B, int32_t h, int32_t E, int32_t nb,
                         int32_t km1, int32_t km2, float* f, double* stats_out, void* stream);
DPOT_API int dpot_split_f16_gn(const float* src, int64_t lds, int64_t rows, int32_t cols, const double* stats,
                      const float* gamma, const float* beta, int32_t groups, float eps, int32_t rows_per_sample,
                      void* dst, int64_t ldd, int64_t lo_off, void* stream);
/* interior_weight multiplies the spectrum columns 0 < k2 < h/2 (on output of fwd / on input of inv); 1 for the
   forward pass.  The two transforms are each other's adjoint up to that weight (Hermitian packing):
   adjoint(inv) = fwd with weight 2, adjoint(fwd) = inv with weight 1/2 -- this is the whole FFT backward.
   scale/shift may be NULL (identity); inv: `a` may be NULL (no skip term). */

/* ------------------------------------------------------------------------------------------
 * Weight packing (run when weights change, not per step).
 * ---------------------------------------------------------------------------------------- */
/* AFNO2D w[2,nb,bs,bs] ("bio": in,out), b[2,nb,bs]  (models/dpot.py:45-48)  ->
   real block form Wc[nb, 2bs(out), 2bs(in)] with [re|im] halves, bc[nb, 2bs]. */
DPOT_API int dpot_pack_afno(const float* w, const float* b, int32_t nb, int32_t bs, float* Wc, float* bc, void* stream);
/* PatchEmbed conv0 weight [mid, C+3, P, P] (models/dpot.py:199) -> W0p[mid, (u,v,c<C)] and the
   coordinate-channel contribution rowbias0[(p,q,t), mid] = b0 + sum_uv W0[:,C..C+2,u,v]*grid
   (get_grid_3d, models/dpot.py:350-360; gx/gy/gt are the fp32 linspace tables). */
DPOT_API int dpot_pack_patch(const float* w0, const float* b0, const float* gx, const float* gy, const float* gt,
                    int32_t mid, int32_t C, int32_t P, int32_t h, int32_t w, int32_t T,
                    float* W0p, float* rowbias0, void* stream);
/* Fold PatchEmbed conv 1x1 (W2[E,mid], b2[E]) + pos_embed[E,n] + TimeAggregator (w[T,E,E],
   temb[T,E] = cos(t*gamma) or ones) (models/dpot.py:201,378,228-232) into
   WeffT[E, Kp] (Kp >= T*mid, zero padded; column t*mid+m) and bias_eff[n, E].  The bias_eff
   buffer must hold n*E + E*E floats (the tail is scratch for sum_t temb*w).  */
DPOT_API int dpot_fold_timeagg(const float* W2, const float* b2, const float* pos, const float* w, const float* temb,
                      int32_t T, int32_t E, int32_t mid, int32_t n, int32_t Kp,
                      float* WeffT, float* bias_eff, void* stream);
/* ConvTranspose2d weight [E, old, P, P], bias[old] (models/dpot.py:316) ->
   WtT[(u,v,o), E], bias_t[(u,v,o)]. */
DPOT_API int dpot_pack_out(const float* wt, const float* bt, int32_t E, int32_t old, int32_t P,
                  float* WtT, float* bias_t, void* stream);

/* ------------------------------------------------------------------------------------------
 * Backward-pass building blocks (autograd of models/dpot.py:364-403; train_temporal.py:227).
 * ---------------------------------------------------------------------------------------- */
/* weight gradient: dW[n,k] (+)= sum_m X[m,n] * Y'[m,k], Y' = Y*y_scale+y_shift (optional) or the im2col view */
typedef struct dpot_wgrad_args {
  const float* X; int64_t ldx;           /* [M,N] upstream gradient */
  const float* Y; int64_t ldy;           /* [M,K] forward input (y_mode = DPOT_A_PATCH: the field) */
  float* dW; int64_t ldw;                /* [N,K] */
  int32_t M, N, K;
  const float* y_scale; const float* y_shift; int32_t y_rows_per_sample;
  int32_t batch; int64_t strideX, strideY, strideW;
  int32_t y_mode, pX, pY, pT, pC, pP;
  int32_t accumulate;                    /* 0: overwrite dW, 1: add */
} dpot_wgrad_args;
DPOT_API int dpot_wgrad(const dpot_wgrad_args* args, void* stream);
/* out[n] (+)= sum_m X[m,n]   (bias gradients) */
DPOT_API int dpot_colsum(const float* X, int64_t ldx, int32_t M, int32_t N, float* out, int32_t accumulate, void* stream);
/* dst[c,r] = src[r,c] for `batch` matrices (weight transposes for the data-gradient GEMMs) */
DPOT_API int dpot_transpose(const float* src, int64_t lds, float* dst, int64_t ldd, int32_t R, int32_t C, int32_t batch,
                   int64_t stride_src, int64_t stride_dst, void* stream);
/* out = x*scale[b,:]+shift[b,:]  (GroupNorm materialised; training path) */
DPOT_API int dpot_gn_apply(const float* x, const float* scale, const float* shift, int32_t B, int32_t n, int32_t E,
                  float* out, void* stream);
/* GroupNorm backward: dx = rstd*(gamma*dy - mean_g(gamma*dy) - xhat*mean_g(gamma*dy*xhat)) (+ add), dgamma += ,
   dbeta += ; stats = the forward (sum, sumsq) doubles; scratch: 2*B*E + 2*B*groups floats. */
DPOT_API int dpot_gn_bwd(const float* dy, const float* x, const double* stats, const float* gamma, const float* add,
                int32_t B, int32_t n, int32_t E, int32_t groups, float eps, float* scratch, float* dx,
                float* dgamma, float* dbeta, void* stream);
/* out = dy * act'(pre)  (activation backward from the stored pre-activation) */
DPOT_API int dpot_act_bwd(const float* dy, const float* pre, int32_t act, int64_t total, float* out, void* stream);
/* rows (b,p,q,u,v) x C  <->  field [b, p*P+u, q*P+v, C]  (to_field = 1 / 0) */
DPOT_API int dpot_pixel_shuffle(const float* src, float* dst, int32_t B, int32_t h, int32_t w, int32_t P, int32_t C,
                       int32_t to_field, void* stream);
/* transpose of dpot_pack_afno for gradients: dw += unpack(dWc), db += unpack(dbc) */
DPOT_API int dpot_unpack_afno_grad(const float* dWc, const float* dbc, int32_t nb, int32_t bs, float* dw, float* db,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * Output head tail (models/dpot.py:317-321, 397-401): per pixel
 *   y1 = act(Y0[(b,p,q), (u,v,:)]) (act already applied by the GEMM epilogue), y2 = act(W2 y1 + b2),
 *   y3 = W4 y2 + b4 -> out[b, p*P+u, q*P+v, to, c]  (x sigma + mu when denorm tables are given).
 * ---------------------------------------------------------------------------------------- */
DPOT_API int dpot_out_tail(const float* Y1, const float* w2, const float* b2, const float* w4, const float* b4,
                  int32_t B, int32_t h, int32_t w, int32_t P, int32_t old, int32_t nout, int32_t act,
                  const float* mu, const float* sigma, int32_t Co, float* out, void* stream);
/* dpot_out_tail whose result goes straight into the ring window (slots (slot0 + j) % T) and pred (frame step*T_out + j):
   the window advance without a copy kernel.  y_scratch as in dpot_rollout_step. */
DPOT_API int dpot_out_tail_ring(const float* Y1, const float* w2, const float* b2, const float* w4, const float* b4, int32_t B,
                       int32_t h, int32_t w, int32_t P, int32_t old, int32_t nout, int32_t act, const float* mu,
                       const float* sigma, int32_t Co, float* y_scratch, float* ring, float* pred, int32_t T, int32_t slot0,
                       int32_t Ttot, int32_t step, void* stream);
/* Output tail on tcgen05: Y1g = the ConvTranspose GEMM's result stored as DPOT_FMT_HL16G32 (one [hi 32 | lo 32] record
   per pixel; out_layer_dim must be 32).  Writes out[B,X,Y,nout] or, when ring != NULL, the ring window / pred as
   dpot_out_tail_ring does.  dpot_out_tail_tc_supported tells whether a geometry is served. */
DPOT_API int dpot_out_tail_tc(const void* Y1g, const float* w2, const float* b2, const float* w4, const float* b4, int32_t B,
                     int32_t h, int32_t w, int32_t P, int32_t old, int32_t nout, int32_t act, const float* mu,
                     const float* sigma, int32_t Co, float* out, float* ring, float* pred, int32_t T, int32_t slot0,
                     int32_t Ttot, int32_t step, void* stream);
DPOT_API int dpot_out_tail_tc_supported(int32_t old, int32_t nout, int32_t Co);
/* engine knob (tests): 0 = auto, 1 = CUDA cores only, 2 = warp-MMA allowed but no tcgen05 tail */
DPOT_API void dpot_out_tail_set_engine(int32_t engine);
/* spatial mean a[B*n,E] -> tok[B,E]  (models/dpot.py:394) */
DPOT_API int dpot_spatial_mean(const float* a, int32_t B, int32_t n, int32_t E, float* tok, void* stream);
/* the same on a split-fp16 latent (DPOT_FMT_HL16 rows [hi E | lo E], as the last block's fc2 writes it for the output
   GEMM): the cls head then costs no extra fp32 copy of the latent. */
DPOT_API int dpot_spatial_mean16(const void* a16, int32_t B, int32_t n, int32_t E, float* tok, void* stream);
/* the same, also (or only: tok = NULL) storing the token split, tok16[B, hi E | lo E]: the operand of the cls head's first
   contraction on the f16-split engine */
DPOT_API int dpot_spatial_mean16s(const void* a16, int32_t B, int32_t n, int32_t E, float* tok, void* tok16, void* stream);
/* per (sample, channel) mean and unbiased std + 1e-6 over (X,Y,T)  (models/dpot.py:367);
   writes musig[B, 2C] = [mu | sigma] and the im2col prologue tables a_scale/a_shift[B, P*P*C]. */
DPOT_API int dpot_input_stats(const float* x, int32_t B, int64_t per_sample, int32_t C, int32_t PP,
                     float* musig, float* a_scale, float* a_shift, void* stream);

/* ------------------------------------------------------------------------------------------
 * Autoregressive window advance (train_temporal.py:219, evaluate.py:208):
 *   xx_next = cat(xx[..., Tb:, :], im) on xx[B,X,Y,T,C], im[B,X,Y,Tb,C]; also copies im into
 *   pred[..., step*Tb:(step+1)*Tb, :] of pred[B,X,Y,Ttot,C] when pred != NULL.
 * ---------------------------------------------------------------------------------------- */
DPOT_API int dpot_window_advance(const float* xx, const float* im, float* xx_next, float* pred, int64_t npix,
                        int32_t T, int32_t Tb, int32_t C, int32_t Ttot, int32_t step, void* stream);

/* Ring form of the same advance: the window lives in a ring buffer ring[B,X,Y,T,C] whose logical frame t
 * is slot (t + t0) % T.  The Tb new frames overwrite the oldest slots slot0 = t0 .. t0+Tb-1 (mod T) and the
 * caller advances t0 <- (t0 + Tb) % T; `pred` as above.  Moves Tb/T of the bytes of dpot_window_advance. */
DPOT_API int dpot_ring_insert(const float* im, float* ring, float* pred, int64_t npix, int32_t T, int32_t Tb,
                     int32_t C, int32_t Ttot, int32_t slot0, int32_t step, void* stream);

/* ------------------------------------------------------------------------------------------
 * PatchEmbed conv0 + activation from the field layout (models/dpot.py:199-200,375), coordinate channels
 * folded into rowbias0 (dpot_pack_patch):  z1[(b,p,q), t*mid + m], row pitch Kp (fp32) or 2*Kp halves with the
 * lo plane at +Kp (out_fmt = DPOT_FMT_HL16).  x[B,X,Y,T,C] is read as a ring in time: logical frame t = slot
 * (t + t0) % T (t0 = 0: the plain layout).  a_scale/a_shift[B, P*P*C]: optional input normalisation tables.
 * ---------------------------------------------------------------------------------------- */
DPOT_API int dpot_patch_embed(const float* x, int32_t t0, const float* W0p, const float* rowbias0, const float* a_scale,
                     const float* a_shift, int32_t B, int32_t X, int32_t Y, int32_t T, int32_t C, int32_t P,
                     int32_t mid, int32_t act, void* z1, int32_t Kp, int32_t out_fmt, void* stream);
/* engine knob (tests): 0 = auto (warp-MMA on split fp16 when the geometry allows, else CUDA cores), 1 = fp32 CUDA cores,
   2 = warp-MMA only.  (Engine 3, a tcgen05 PatchEmbed, was measured slower and removed: DESIGN 4.3.) */
DPOT_API void dpot_patch_embed_set_engine(int32_t engine);

/* ------------------------------------------------------------------------------------------
 * Optimizer: adam()/adamw() of utils/optimizer.py:9-52 / :170-212 on one flat tensor.
 * step = 1-based count after the increment; decoupled = 0 (Adam, L2-coupled decay) or 1 (AdamW);
 * vmax = amsgrad buffer or NULL; grad_scale multiplies the gradient first (1/world for DDP).
 * Hyper-parameters are doubles: the reference derives 1-beta, lr/(1-beta1^t), sqrt(1-beta2^t) as
 * python floats (doubles) and only then rounds to fp32.
 * ---------------------------------------------------------------------------------------- */
DPOT_API int dpot_adam_step(float* p, const float* g, float* m, float* v, float* vmax, int64_t n, double lr,
                   double beta1, double beta2, double eps, double weight_decay, int32_t step,
                   int32_t decoupled, double grad_scale, void* stream);
/* multi-tensor form: `count` tensors described by parallel host arrays */
DPOT_API int dpot_adam_step_multi(float* const* p, const float* const* g, float* const* m, float* const* v,
                         float* const* vmax, const int64_t* n, int32_t count, double lr, double beta1,
                         double beta2, double eps, double weight_decay, const int32_t* steps,
                         int32_t decoupled, double grad_scale, void* stream);
/* the same with torch.nn.utils.clip_grad_norm_ (train_temporal.py:228) folded in: grad_sqnorm = device double with the
   sum of squares of ALL gradients (dpot_grad_sqnorm), the kernel scales every gradient by
   grad_scale * min(1, max_norm / (grad_scale * sqrt(sum) + 1e-6)) on load -- gradients are read once per step and never
   rewritten.  grad_sqnorm = NULL: no clipping. */
DPOT_API int dpot_adam_step_multi_clip(float* const* p, const float* const* g, float* const* m, float* const* v,
                                       float* const* vmax, const int64_t* n, int32_t count, double lr, double beta1,
                                       double beta2, double eps, double weight_decay, const int32_t* steps,
                                       int32_t decoupled, double grad_scale, const double* grad_sqnorm, double max_norm,
                                       void* stream);
/* Lamb.step (utils/optimizer.py:421-499; the `--opt lamb` branch, evaluate.py:135-136) over `count` tensors in two
   launches per 56 tensors and no host synchronisation: stage 1 updates exp_avg / exp_avg_sq and reduces sum p^2 and
   sum u^2 per tensor (u = exp_avg / (sqrt(exp_avg_sq) + eps) [+ wd p]) into norms[2*count] (double, zeroed by the
   call); stage 2 applies p -= lr * bias_correction * trust * u with weight_norm = min(||p||, clamp_value),
   trust = 1 if either norm is 0 else weight_norm / ||u|| (1 when `adam`), and stores (weight_norm, adam_norm,
   trust_ratio) per tensor in info[3*count] (may be NULL) -- the reference's state entries.  steps[t] >= 1 is the
   tensor's step count AFTER the increment (debias: bias_correction = sqrt(1-b2^t)/(1-b1^t), else 1). */
DPOT_API int dpot_lamb_step_multi(float* const* p, const float* const* g, float* const* m, float* const* v, const int64_t* n,
                         int32_t count, double lr, double beta1, double beta2, double eps, double weight_decay,
                         double clamp_value, const int32_t* steps, int32_t debias, int32_t adam, double* norms, float* info,
                         void* stream);
/* global gradient norm: out_sq[0] = sum over `count` tensors of sum g^2 (double accumulation; zeroed by the call).
   Replaces the norm computation of clip_grad_norm_, train_temporal.py:228. */
DPOT_API int dpot_grad_sqnorm(const float* const* g, const int64_t* n, int32_t count, double* out_sq, void* stream);

/* ------------------------------------------------------------------------------------------
 * The autoregressive training loop around model(xx), train_temporal.py:201-230.
 * Fields are x[B, X, Y, T, C] with the channel innermost; npos = X*Y*T, nxy = X*Y; C <= 16.
 * ---------------------------------------------------------------------------------------- */
/* out[b, c] = sum over (X, Y, T) of x^2   (double; zeroed by the call) */
DPOT_API int dpot_chan_sumsq(const float* x, int32_t B, int64_t npos, int32_t C, double* out, void* stream);
/* noise injection, train_temporal.py:205:  out = x + scale * sqrt(sumsq[b, c]) * N(0, 1), normals from a counter-based
   Philox-4x32-10 stream (seed, offset) -- no generator state; sumsq[B, C] (double) is written by the call and kept for
   the backward pass.  out may alias x. */
DPOT_API int dpot_noise_inject(const float* x, int32_t B, int64_t npos, int32_t C, float scale, uint64_t seed,
                               uint64_t offset, double* sumsq, float* out, void* stream);
/* its backward (the reference differentiates through the norm):  dx = dy + scale * x / ||x||_bc * sum_pos(dy * eps),
   eps regenerated from (seed, offset); dot[B, C] (double) is scratch. */
DPOT_API int dpot_noise_inject_bwd(const float* x, const float* dy, int32_t B, int64_t npos, int32_t C, float scale,
                                   uint64_t seed, uint64_t offset, const double* sumsq, double* dot, float* dx, void* stream);
/* SimpleLpLoss(size_average=False).forward(x, y, mask), utils/criterion.py:38-59: loss[0] (=|+=, `accumulate`)
   sum_b ( sum_c ||(x - y) m||_2 / (||y m||_2 + 1e-8) ) / #active channels of b;  mask[B, X, Y, 1, C] or NULL.
   partial[B, C, 3] (double) is scratch; coef[B, C] receives the backward coefficients. */
DPOT_API int dpot_lp_loss(const float* x, const float* y, const float* mask, int32_t B, int64_t nxy, int32_t T, int32_t C,
                          double* partial, float* coef, float* loss, int32_t accumulate, void* stream);
/* dx = gscale[0] * coef[b, c] * m^2 * (x - y)   (gscale = upstream gradient of the scalar loss on the device, NULL = 1) */
DPOT_API int dpot_lp_loss_bwd(const float* x, const float* y, const float* mask, const float* coef, const float* gscale,
                              int32_t B, int64_t nxy, int32_t T, int32_t C, float* dx, void* stream);

/* fwd16 without GroupNorm and with the interior-column weight (weight 2 = the adjoint of the inverse transform, whose
   result the backward pass feeds straight into the split-fp16 contractions).  colsum (may be NULL): double [2E],
   colsum[c] += sum over the stored rows of column c (the bias gradient of the spectral layer; not zeroed). */
DPOT_API int dpot_afno_fft_fwd16w(const float* a, int32_t B, int32_t h, int32_t E, int32_t nb, int32_t km1, int32_t km2,
                                  void* S16, float interior_weight, double* colsum, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-model inference forward: DPOTNet.forward under no_grad (models/dpot.py:364-403).
 * ---------------------------------------------------------------------------------------- */
typedef struct dpot_config {
  int32_t img_size, patch_size, in_channels, out_channels, in_timesteps, out_timesteps;
  int32_t n_blocks, embed_dim, out_layer_dim, depth, modes, hidden_dim, n_cls;
  int32_t normalize, act, time_agg;      /* time_agg: 0 = 'mlp', 1 = 'exp_mlp' */
} dpot_config;

typedef struct dpot_block_params {
  const float *norm1_w, *norm1_b, *w1, *b1, *w2, *b2, *norm2_w, *norm2_b;
  const float *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} dpot_block_params;

typedef struct dpot_params {             /* reference state_dict tensors, SURVEY.md appendix A */
  const float *pos_embed, *pe0_w, *pe0_b, *pe2_w, *pe2_b;
  const float *tagg_w, *tagg_gamma;
  const float *cls0_w, *cls0_b, *cls2_w, *cls2_b, *cls4_w, *cls4_b;
  const float *out0_w, *out0_b, *out2_w, *out2_b, *out4_w, *out4_b;
  const float *mu_w, *mu_b, *sigma_w, *sigma_b;           /* scale_feats_* (normalize only) */
  const dpot_block_params* blocks;                        /* host array [depth] */
  const float *grid_x, *grid_y, *grid_t;                  /* fp32 linspace(0,1,n) tables on device */
  const float *temb;                                      /* [T,E] cos(t*gamma) (ones for 'mlp') */
} dpot_params;

/* number of floats of the packed-weight arena / the per-call workspace */
DPOT_API int64_t dpot_packed_floats(const dpot_config* cfg);
DPOT_API int64_t dpot_workspace_floats(const dpot_config* cfg, int32_t B);
/* derive every packed weight from the raw parameters (weights changed -> call again) */
DPOT_API int dpot_pack_weights(const dpot_config* cfg, const dpot_params* prm, float* packed, void* stream);
/* x[B,X,Y,T,C] -> y[B,X,Y,To,Co], cls[B,n_cls] (cls may be NULL) */
DPOT_API int dpot_forward(const dpot_config* cfg, const dpot_params* prm, const float* packed, const float* x,
                 int32_t B, float* y, float* cls, float* workspace, int32_t engine, void* stream);

/* knob (tests / A-B measurements): the classification head of the f16-split inference pipeline on 1 = the f16-split
   tensor-core engine (default), 0 = the CUDA-core skinny contractions */
DPOT_API void dpot_set_cls_engine(int32_t tc);

/* SM budget of the persistent kernels (contractions, fused mixer, tail / PatchEmbed backward): n > 0 sizes their grids
   for n SMs instead of the whole device, leaving the rest to kernels that run CONCURRENTLY on other streams (NCCL's
   all-reduce during backward: a persistent grid with a static tile partition is delayed by a whole share of tiles for
   every SM it has to share).  0 = the whole device (default).  Returns the previous value; n < 0 only queries. */
DPOT_API int dpot_set_sm_budget(int32_t n);

/* 1: the classification head of dpot_forward* / dpot_rollout_step runs on a library-owned side stream (one per device)
   concurrently with the output head, joined before the call's work on `stream` ends; 0 (default): everything on `stream`.
   Measured slower on B200 (the side stream takes SMs from the persistent contraction kernels); kept as a knob. */
DPOT_API void dpot_set_cls_overlap(int32_t on);

/* dpot_forward on a time-ring input window (see dpot_ring_insert): logical frame t of x is slot (t + t0) % T */
DPOT_API int dpot_forward_ring(const dpot_config* cfg, const dpot_params* prm, const float* packed, const float* x,
                      int32_t t0, int32_t B, float* y, float* cls, float* workspace, int32_t engine, void* stream);
/* One autoregressive step on a ring window (evaluate.py:192-208, train_temporal.py:262-272): im = model(window read at
   ring offset t0); the out_timesteps new frames overwrite the oldest slots (t0 + j) % T of `ring` and are scattered into
   pred[B,X,Y,pred_frames,C] at frame step*out_timesteps + j (pred may be NULL).  The caller advances t0 by out_timesteps.
   y: scratch of the model-output shape (only touched for geometries the fused tail kernel does not serve). */
DPOT_API int dpot_rollout_step(const dpot_config* cfg, const dpot_params* prm, const float* packed, float* ring, int32_t t0,
                      int32_t B, float* y, float* cls, float* ws, int32_t engine, float* pred, int32_t pred_frames,
                      int32_t step, void* stream);

/* ------------------------------------------------------------------------------------------
 * GPU-side batch assembly of the data path, utils/griddataset.py:88-100 (pad_data: bilinear resize of every
 * (t, c) plane to res x res, torch align_corners=False rule; channels padded with 1.0), :152-157 (training window: T_in
 * frames from t_start[b], the next T_ar as targets) and the masks (:156 ones; mask_mode = 1: get_target_mask :102-116 with
 * pred_channels).  raw[B, H0, W0, T0, C0] (device; same-shaped samples), t_start[B] (device int32, t_start + T_in + T_ar
 * <= T0) -> xx[B, res, res, T_in, C], yy[B, res, res, T_ar, C], msk[B, res, res, 1, C] (may be NULL).
 * ---------------------------------------------------------------------------------------- */
DPOT_API int dpot_assemble_batch(const float* raw, const int32_t* t_start, int32_t B, int32_t H0, int32_t W0, int32_t T0,
                                 int32_t C0, int32_t res, int32_t T_in, int32_t T_ar, int32_t C, int32_t mask_mode,
                                 int32_t pred_channels, float* xx, float* yy, float* msk, void* stream);

/* ------------------------------------------------------------------------------------------
 * The training step: DPOTNet.forward with the activations backward needs kept on a caller-owned tape, and its
 * autograd -- what `im, cls = model(xx)` ... `loss.backward()` run in train_temporal.py:206,227 -- with every dense
 * contraction of forward and backward on the f16-split tcgen05 engine (data gradients read the weights in their
 * forward layout, weight gradients read both token-major activations as stored; no transposed copies).
 *   packed  : dpot_packed_floats(cfg) floats, filled by dpot_train_prepare (the fused-mixer slab is not written)
 *   wprep   : dpot_train_wprep_floats(cfg) floats, fold intermediates kept for backward
 *   tape    : dpot_train_tape_floats(cfg, B) floats per forward call, alive until its backward
 *   scratch : dpot_train_scratch_floats(cfg, B) floats, reusable by every call on the same stream
 *   grads   : a dpot_params whose pointers are the DESTINATIONS of the parameter gradients (same shapes as the
 *             parameters; written, not accumulated; grid/temb/mu/sigma members ignored; the cls members are written
 *             only when dcls != NULL; tagg_gamma only for 'exp_mlp').  dx: gradient w.r.t. x or NULL.
 *   events  : NULL, or depth + 2 cudaEvent_t recorded on `stream` as groups of gradients become final -- [0] the
 *             out_layer parameters, [1 + j] block depth-1-j, [depth + 1] everything -- so that a data-parallel caller can
 *             start exchanging a group (train_temporal_parallel.py:185,244) while the rest of backward still runs.
 * dpot_train_supported: 1 when the configuration is served (normalize = False; patch size 8 geometry of DPOT-Ti/S/M/H;
 * out_layer_dim = 32 on the fused tail kernels, other multiples of 8 -- DPOT-H's 128 -- on contractions batched over the
 * intra-patch positions); other configurations train through the generic per-operator path of the Python binding.
 * ---------------------------------------------------------------------------------------- */
DPOT_API int     dpot_train_supported(const dpot_config* cfg);
DPOT_API int64_t dpot_train_tape_floats(const dpot_config* cfg, int32_t B);
DPOT_API int64_t dpot_train_scratch_floats(const dpot_config* cfg, int32_t B);
DPOT_API int64_t dpot_train_wprep_floats(const dpot_config* cfg);
DPOT_API int dpot_train_prepare(const dpot_config* cfg, const dpot_params* prm, float* packed, float* wprep, float* scratch,
                                void* stream);
DPOT_API int dpot_train_forward(const dpot_config* cfg, const dpot_params* prm, const float* packed, const float* wprep,
                                const float* x, int32_t B, float* y, float* cls, float* tape, float* scratch, void* stream);
DPOT_API int dpot_train_backward(const dpot_config* cfg, const dpot_params* prm, const float* packed, const float* wprep,
                                 const float* x, int32_t B, const float* dy, const float* dcls, const float* tape,
                                 float* scratch, const dpot_params* grads, float* dx, void* const* events, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPOT_B200_H_ */
