// Element-wise / reduction kernels of the training step (train_kernels.cu), used by its host orchestration
// (train_step.cu).  Internal to the library: the C ABI of the training step is dpot_train_* in include/dpot_b200.h.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace dpot {

// gradient scale: scale[0] = S = 2^k with S * max|dy| in [1, 2), scale[1] = 1 / S  (bits: 1 uint scratch, zeroed here)
int tk_grad_scale(const float* dy, int64_t n, unsigned* bits, float* scale, cudaStream_t st);
// out[n] += sum_m X[m, n]  (double accumulation; X fp32 with row stride ld, or split fp16: hi at [m*ld + n], lo at + lo_off)
int tk_colsum(const void* X, bool f16, int64_t ld, int64_t lo_off, int M, int N, double* out, cudaStream_t st);
// GroupNorm backward (torch.nn.GroupNorm(8, E) on token-major x[B*n, E]); stats = forward (sum, sumsq) doubles.
//   dx = rstd * (gamma*dy - mean_g(gamma*dy) - xhat * mean_g(gamma*dy*xhat)) (+ add);  dx16 != NULL: also stored split.
//   dgamma / dbeta (written, scaled by inv_scale[0] when given).  scratch: 5*B*E + 2*B*groups floats.
//   colsum (may be NULL): double [E], colsum[c] += sum over rows of dx (the bias gradient of the layer below).
int tk_gn_bwd(const float* dy, const float* x, const double* stats, const float* gamma, const float* add, int B, int n, int E,
              int groups, float eps, const float* inv_scale, float* scratch, float* dx, __half* dx16, float* dgamma,
              float* dbeta, double* colsum, cudaStream_t st);
// dst[i] = inv_scale * sum_s slabs[s * stride + i]
int tk_slab_reduce(const float* slabs, int nslab, int64_t stride, int64_t count, const float* inv_scale, float* dst, cudaStream_t st);
// dst[i] = inv_scale * src[i]  (double -> float)
int tk_finish_double(const double* src, int64_t count, const float* inv_scale, float* dst, cudaStream_t st);
// transpose of pack_afno on the sum of `nslab` partial dWc[nb, 2bs, 2bs] (+ dbc double [nb, 2bs]) -> dw[2,nb,bs,bs], db[2,nb,bs]
int tk_unpack_afno_grad(const float* dWc, int nslab, int64_t stride, const double* dbc, int nb, int bs, const float* inv_scale,
                        float* dw, float* db, cudaStream_t st);
// transpose of pack_out: sum of partial dWtT[(u,v,o), E] -> dwt[E, old, P, P]; dbias_t double [(u,v,o)] -> db[old]
int tk_unpack_out_grad(const float* dWtT, int nslab, int64_t stride, const double* dbias_t, int E, int old, int P,
                       const float* inv_scale, float* dwt, float* db, cudaStream_t st);
// Output-head tail backward (out_layer[1..4] of models/dpot.py:317-321, out_layer_dim = 32):
//   from Y1pre[Mt, PP*32] (pre-activation of the ConvTranspose GEMM) and dout[B, X, Y, nout] * scale[0] ->
//   g1 (gradient w.r.t. Y1pre) stored split [Mt, 2*NP]; dW2 / db2 / dW4 / db4 and db0 (the ConvTranspose bias gradient:
//   sum of g1 over pixels per channel) written (times inv_scale[0]).
//   part: tk_tail_bwd_part_floats(tk_tail_bwd_blocks(...)) floats of scratch (one partial result per thread block).
int tk_tail_bwd_supported(int old, int nout);
int tk_tail_bwd_blocks(int B, int h, int w, int P);
int64_t tk_tail_bwd_part_floats(int nblk);
int tk_tail_bwd(const float* Y1pre, const float* dout, const float* scale, const float* w2, const float* b2, const float* w4,
                int B, int h, int w, int P, int nout, int act, __half* g1, float* part, const float* inv_scale, float* dW2,
                float* db2, float* dW4, float* db4, float* db0, cudaStream_t st);
// PatchEmbed conv0 backward (models/dpot.py:199-200): gz[(b,pq), t*mid + m] (row pitch Kp; act' already applied),
// x[B,X,Y,T,C] -> dW0p (one partial [mid, P*P*C] per thread block); dx (may be NULL) = inv_scale * gz W0p scattered to the field layout.
int tk_patch_bwd_supported(int mid, int T, int K0);
int tk_patch_bwd_slabs(int B, int X, int Y, int P);     // number of per-block partial results [mid, P*P*C] the kernel writes
int tk_patch_bwd(const float* gz, const float* x, const float* W0p, int B, int X, int Y, int T, int C, int P, int mid, int Kp,
                 const float* inv_scale, float* dW0p, float* dx, cudaStream_t st);
// transpose of pack_patch: dW0p (nslab float partials [mid, (u,v,c)]), drb (double [(p,q), Kp] = sum_b gz) -> dpe0_w[mid, C+3, P, P], dpe0_b[mid]
int tk_unpack_patch_grad(const float* dW0p, int nslab, const double* drb, const float* gx, const float* gy, const float* gt, int mid, int C,
                         int P, int h, int w, int T, int Kp, const float* inv_scale, float* dw0, float* db0, cudaStream_t st);
// generic output-head tail (out_layer_dim != 32): dL/d(out field) -> g3[tok, hi (uv, 8) | lo (uv, 8)], nout channels zero-padded
// to 8, scaled by scale[0], split; and dst[i < keep] = inv * sum_b src[b*n + i] (column sums taken per (u, v) problem)
int tk_unshuffle_pad_split(const float* dout, const float* scale, int B, int h, int w, int P, int nout, __half* g3, cudaStream_t st);
int tk_sum_batches(const double* src, int nbat, int n, int keep, const float* inv_scale, float* dst, cudaStream_t st);
// weight packing of the per-step prepare: AFNO real block form written directly as split fp16; PatchEmbed conv0 im2col
// weight + coordinate-channel bias table with the coordinate sums separated (same results as dpot_pack_afno / dpot_pack_patch)
int tk_pack_afno16(const float* w, const float* b, int nb, int bs, __half* Wc16, float* bc, cudaStream_t st);
int tk_pack_patch(const float* w0, const float* b0, const float* gx, const float* gy, const float* gt, int mid, int C, int P, int h,
                  int w, int T, float* W0p, float* rowbias0, cudaStream_t st);
// time-aggregation fold helpers (models/dpot.py:228-232 folded with PatchEmbed conv 1x1 + pos_embed)
//   wts16[(t,i), j] = split(w[t,i,j]*temb[t,i]) rows of [hi E | lo E];  Wsum16[i, j] = split(sum_t temb[t,i] w[t,i,j])
int tk_tagg_scale16(const float* w, const float* temb, int T, int E, __half* wts16, __half* Wsum16, cudaStream_t st);
//   bp16[i, p] = split(b2[i] + pos[i, p])
int tk_tagg_bp16(const float* b2, const float* pos, int E, int n, __half* bp16, cudaStream_t st);
//   zero-padded split copy: dst16[r, c < colsp] = split(c < cols ? src[r*lds + c] : 0)
int tk_pad_split(const float* src, int64_t lds, int64_t rows, int cols, int colsp, __half* dst, cudaStream_t st);
//   Gp16[t][j][m < midp] = split(m < mid ? dWeffT[j, t*mid + m] : 0)
int tk_tagg_pad_g(const float* dWeffT, int E, int Kp, int T, int mid, int midp, __half* dst, cudaStream_t st);
//   dW2[i, m < mid] = inv * sum_t slabs[t][i][m]   (slab rows of midp)
int tk_tagg_dw2_finish(const float* slabs, int T, int E, int mid, int midp, const float* inv_scale, float* dst, cudaStream_t st);
//   dw[t,i,j] = inv * dwt[t,i,j]*temb[t,i];  dtemb[t,i] = sum_j dwt[t,i,j]*w[t,i,j] (float scratch [T,E])
int tk_tagg_finish(const float* dwt, const float* w, const float* temb, int T, int E, const float* inv_scale, float* dw,
                   float* dtemb, cudaStream_t st);
//   dgamma[i] = inv * sum_t dtemb[t,i] * (-sin(tt_t*gamma_i)) * tt_t,  tt = linspace(0,1,T)
int tk_tagg_gamma_grad(const float* dtemb, const float* gamma, int T, int E, const float* inv_scale, float* dgamma, cudaStream_t st);
//   out[i] = inv * sum_p X[i, p]  (row sums; db2 of the 1x1 conv), dst2[i,p] = inv * X[i,p] (dpos)
int tk_rowsum_scale(const float* X, int rows, int cols, const float* inv_scale, float* rowsum, float* scaled, cudaStream_t st);
// C = A^T (fp32, generic small transposes of the fold backward)
int tk_transpose(const float* src, int64_t lds, float* dst, int64_t ldd, int R, int Cc, cudaStream_t st);
// dst = inv_scale * src (float)
int tk_scale_copy(const float* src, int64_t count, const float* inv_scale, float* dst, cudaStream_t st);
// cls head (models/dpot.py:394): dlat16/dlat += dtok[b, :] / n broadcast over the n cells is folded into the last block's
// gradient by this kernel: g[b*n + p, e] += dtok[b, e] / n
int tk_add_mean_grad(const float* dtok, int B, int n, int E, float* g, cudaStream_t st);
// fp32 -> split fp16 with a device scalar factor (first element of `factor`, NULL = 1)
int tk_split_scaled(const float* src, int64_t rows, int cols, const float* factor, __half* dst, int64_t ldd, int64_t lo_off,
                    cudaStream_t st);

}  // namespace dpot
