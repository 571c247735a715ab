// Device-side view of dpot_gemm_args and the A-prologue / C-epilogue shared by both engines.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace dpot {

struct GemmDev {
  const float* A; int64_t lda;
  const float* W; int64_t ldw;
  float* C; int64_t ldc;
  int M, N, K;
  const float* bias;
  const float* rowbias; int rb_period; int64_t ldrb;
  const float* residual; int64_t ldr;
  int act;
  const float* a_scale; const float* a_shift; int a_rps;
  const float* c_scale; const float* c_shift; int c_rps;
  int c_group; int64_t c_group_stride;
  int64_t sA, sW, sC, sBias;
  int a_mode, pX, pY, pT, pC, pP, ph, pw;
  double* out_stats; int st_groups, st_rps;   // fused GroupNorm statistics (tcgen05 engine only)
  float* C_pre; const float* dact_src; int dact; int c_mode;   // training extras
  int a_fmt, w_fmt, c_fmt; int64_t a_lo, w_lo, c_lo;           // DPOT_FMT_* storage (strides in halves when HL16)
  // backward-pass forms of the f16-split engine (dpot_gemm_args ABI 2): transposed operand storage, contraction split
  int a_tr, w_tr, ksplit; int64_t kchunk, sC2;
  int64_t ldpre, sPre, lddact, sDact;                           // C_pre / dact_src geometry on the f16-split engine
  double* out_colsum;                                           // f16-split engine: column sums of the stored result
};

// split fp16 storage (include/dpot_b200.h, DPOT_FMT_HL16): x ~= hi + lo / 2048
constexpr float HL_SCALE = 2048.0f, HL_INV = 1.0f / 2048.0f;
__device__ __forceinline__ void hl_split(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * HL_SCALE);
}

// im2col address of PatchEmbed conv0 (models/dpot.py:199,375): row m = (b,p,q,t), k = (u,v,c)
__device__ __forceinline__ int64_t patch_offset(const GemmDev& p, int m, int k) {
  const int t = m % p.pT; int r = m / p.pT;
  const int q = r % p.pw; r /= p.pw;
  const int pp = r % p.ph; const int b = r / p.ph;
  const int c = k % p.pC; const int uv = k / p.pC;
  const int v = uv % p.pP, u = uv / p.pP;
  return ((((int64_t)b * p.pX + pp * p.pP + u) * p.pY + q * p.pP + v) * p.pT + t) * p.pC + c;
}

__device__ __forceinline__ float gemm_load_a(const GemmDev& p, const float* __restrict__ A, int m, int k) {
  if (m >= p.M || k >= p.K) return 0.f;
  float v = (p.a_mode == DPOT_A_PLAIN) ? A[(int64_t)m * p.lda + k] : A[patch_offset(p, m, k)];
  if (p.a_scale) {
    const int64_t o = (int64_t)(m / p.a_rps) * p.K + k;
    v = fmaf(v, p.a_scale[o], p.a_shift[o]);
  }
  return v;
}

__device__ __forceinline__ int64_t gemm_c_offset(const GemmDev& p, int m) {
  return p.c_group ? (int64_t)(m / p.c_group) * p.c_group_stride + (int64_t)(m % p.c_group) * p.ldc
                   : (int64_t)m * p.ldc;
}

// boff = batch offset (elements) of this problem inside C-shaped side tensors (C_pre, dact_src)
__device__ __forceinline__ float gemm_epilogue_value(const GemmDev& p, const float* __restrict__ bias, int m, int n,
                                                     float v, int64_t boff = 0) {
  if (bias) v += bias[n];
  if (p.rowbias) v += p.rowbias[(int64_t)(m % p.rb_period) * p.ldrb + n];
  if (p.C_pre) p.C_pre[boff + gemm_c_offset(p, m) + n] = v;
  v = act_apply(v, p.act);
  if (p.dact_src) v *= act_grad(p.dact_src[boff + (int64_t)m * p.ldc + n], p.dact);
  if (p.c_scale) {
    const int64_t o = (int64_t)(m / p.c_rps) * p.N + n;
    v = fmaf(v, p.c_scale[o], p.c_shift[o]);
  }
  if (p.residual) v += p.residual[(int64_t)m * p.ldr + n];
  return v;
}

__device__ __forceinline__ void gemm_epilogue_store(const GemmDev& p, float* __restrict__ C,
                                                    const float* __restrict__ bias, int m, int n, float v,
                                                    int64_t boff = 0) {
  const float r = gemm_epilogue_value(p, bias, m, n, v, boff);
  if (p.c_fmt == DPOT_FMT_HL16) {
    __half* Ch = reinterpret_cast<__half*>(C) + gemm_c_offset(p, m) + n;
    __half hi, lo;
    hl_split(r, hi, lo);
    Ch[0] = hi;
    Ch[p.c_lo] = lo;
  } else if (p.c_mode == DPOT_A_PATCH) C[patch_offset(p, m, n)] = r;
  else C[gemm_c_offset(p, m) + n] = r;
}

int gemm_simt_launch(const GemmDev& p, int batch, cudaStream_t st);
int gemm_tc_launch(const GemmDev& p, int batch, cudaStream_t st);   // tcgen05 engine
bool gemm_tc_supports(const GemmDev& p, int batch);
bool gemm_tc_fuses_stats(const GemmDev& p);   // can the TC epilogue accumulate out_stats itself?
// f16-split tcgen05 engine (gemm_tc16.cu)
int gemm_tc16_launch(const GemmDev& p, int batch, cudaStream_t st);
bool gemm_tc16_supports(const GemmDev& p, int batch);
bool gemm_tc16_fuses_stats(const GemmDev& p);
bool tc_device_ok();
int tc_encode_map_f16(void* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                      uint32_t b0, uint32_t b1, uint32_t b2);

}  // namespace dpot
