// Shared device/host helpers for libdpot_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/dpot_b200.h"

namespace dpot {

// ---- error plumbing (thread-local message, no exceptions across the ABI) -------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);
void count_launch();   // kernels launched by this library since load (dpot_launch_count)

#define DPOT_REQUIRE(cond, code, ...)              \
  do {                                             \
    if (!(cond)) {                                 \
      ::dpot::set_error(__VA_ARGS__);              \
      return (code);                               \
    }                                              \
  } while (0)

#define DPOT_CUDA(call)                                             \
  do {                                                              \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess) return ::dpot::cuda_fail(e__, #call);   \
  } while (0)

#define DPOT_LAUNCH_CHECK(name)                                     \
  do {                                                              \
    ::dpot::count_launch();                                         \
    cudaError_t e__ = cudaGetLastError();                           \
    if (e__ != cudaSuccess) return ::dpot::cuda_fail(e__, name);    \
  } while (0)

#define DPOT_CALL(expr)            \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != 0) return rc__;    \
  } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------
// The kernels of the forward chain are launched with cudaLaunchAttributeProgrammaticStreamSerialization: each one
// signals `launch_dependents` at its very start, so the NEXT kernel's CTAs are placed on an SM as soon as the last CTA
// of this kernel leaves it, run their prologue there (barrier init, TMEM allocation, parameter-only loads such as
// weight staging) while other SMs still finish this kernel, and block in `pdl_wait()` -- placed before the first read
// or write of anything a predecessor produces -- until the predecessor grid has completed and flushed.
// Measured on DPOT-S B=32 (bench.py --no-pdl / --no-graph): +1.1 % when kernels are launched one by one, none (-1 %,
// noise) under CUDA-graph replay, where the launch gaps are already gone -- so it is OFF by default.
extern int g_pdl;      // 0 = plain stream order (default), 1 = programmatic dependent launch (dpot_set_pdl)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_if(bool on, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = on ? 1 : 0;
  lc.attrs = at; lc.numAttrs = 1;
  return cudaLaunchKernelEx(&lc, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  return launch_pdl_if(g_pdl != 0, kern, grid, block, smem, st, args...);
}

// Per-device caches: function attributes (opt-in shared memory) and the SM count are properties of a DEVICE; a process
// that drives several GPUs (train_temporal.py --gpu N, tests on multi-GPU boxes) must not reuse device 0's.
static inline int cur_dev() { int d = 0; cudaGetDevice(&d); return d < 0 ? 0 : (d > 63 ? 63 : d); }
int sm_count_cur();    // SM count of the current device (cached per device ordinal)
struct DevOnce {       // "done once per device" flag set for one kernel instantiation
  unsigned long long mask = 0;
  bool need() const { return !((mask >> cur_dev()) & 1ull); }
  void done() { mask |= 1ull << cur_dev(); }
};

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// nn.GELU() (exact erf form, models/dpot.py:19) as x * Phi(x) with ONE branch-free evaluation:
//   Phi(-t) = 0.5 erfc(t / sqrt 2) = 2^q(t),  t = min(|x|, 5.7),  q = degree-9 minimax fit of log2(0.5 erfc(t/sqrt 2))
//   weighted by the sensitivity t * Phi(-t) of the result;  gelu(x) = x * (x > 0 ? 1 - 2^q : 2^q).
// Max abs error vs float64 over [-8, 8]: 2.1e-7 for |x| < 3 (= fp32 rounding of the exact value), rel-L2 2.0e-8
// (fit + check: tools/fit_gelu.py).  15 instructions instead of ~38 for the two-branch erff form.
__device__ __forceinline__ float gelu_fast(float x) {
  const float t = fminf(fabsf(x), 5.7f);
  float q = 1.3804051945953688e-07f;
  q = fmaf(q, t, -3.4098220567102544e-06f);
  q = fmaf(q, t, 3.314594505354762e-05f);
  q = fmaf(q, t, -0.00013128264981787652f);
  q = fmaf(q, t, -0.00031273943022824824f);
  q = fmaf(q, t, 0.007349733263254166f);
  q = fmaf(q, t, -0.05274663493037224f);
  q = fmaf(q, t, -0.4590948820114136f);
  q = fmaf(q, t, -1.1511284112930298f);
  q = fmaf(q, t, -0.9999985098838806f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  return x * (x > 0.f ? 1.0f - e : e);
}

// ---- activations: ACTIVATION table of models/dpot.py:19 (torch module defaults) --------------
// GELU (every shipped config) is evaluated inline; the other seven activations sit behind ONE out-of-line call so that
// kernels which unroll dozens of activation sites do not inline an eight-way switch of transcendental code at each of
// them (measured: the output-tail backward kernel was 27 500 instructions and instruction-fetch bound).
static __device__ __noinline__ float act_apply_other(float x, int act) {
  switch (act) {
    case DPOT_ACT_TANH: return tanhf(x);
    case DPOT_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case DPOT_ACT_RELU: return fmaxf(x, 0.0f);
    case DPOT_ACT_LEAKY_RELU: return x > 0.0f ? x : 0.1f * x;
    case DPOT_ACT_SOFTPLUS: return x > 20.0f ? x : log1pf(expf(x));
    case DPOT_ACT_ELU: return x > 0.0f ? x : expm1f(x);
    case DPOT_ACT_SILU: return x / (1.0f + expf(-x));
    default: return x;
  }
}
__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == DPOT_ACT_GELU) return gelu_fast(x);       // nn.GELU() exact erf form
  return act_apply_other(x, act);
}

// Branch-free erf (max abs error 5.7e-8 over R, checked against scipy): both minimax branches are
// evaluated and selected, so several independent elements interleave without divergence.
__device__ __forceinline__ float erf_select(float a) {
  const float t = fabsf(a), s = a * a;
  float r = fmaf(-1.72853470e-5f, t, 3.83197126e-4f);
  const float u = fmaf(-3.88396438e-3f, t, 2.42546219e-2f);
  r = fmaf(r, s, u);
  r = fmaf(r, t, -1.06777877e-1f);
  r = fmaf(r, t, -6.34846687e-1f);
  r = fmaf(r, t, -1.28717512e-1f);
  r = fmaf(r, t, -t);
  const float big = copysignf(1.0f - __expf(r), a);
  float q = -5.96761703e-4f;
  q = fmaf(q, s, 4.99119423e-3f);
  q = fmaf(q, s, -2.67681349e-2f);
  q = fmaf(q, s, 1.12819925e-1f);
  q = fmaf(q, s, -3.76125336e-1f);
  q = fmaf(q, s, 1.28379166e-1f);
  q = fmaf(q, a, a);
  return t > 0.927734375f ? big : q;
}
__device__ __forceinline__ float gelu_select(float x) { return gelu_fast(x); }

// derivative d act(x)/dx, used by the backward kernels (out of line: see act_apply_other)
static __device__ __noinline__ float act_grad(float x, int act) {
  switch (act) {
    case DPOT_ACT_GELU: {
      const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
      return cdf + x * pdf;
    }
    case DPOT_ACT_TANH: { float t = tanhf(x); return 1.0f - t * t; }
    case DPOT_ACT_SIGMOID: { float s = 1.0f / (1.0f + expf(-x)); return s * (1.0f - s); }
    case DPOT_ACT_RELU: return x > 0.0f ? 1.0f : 0.0f;
    case DPOT_ACT_LEAKY_RELU: return x > 0.0f ? 1.0f : 0.1f;
    case DPOT_ACT_SOFTPLUS: return x > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-x));
    case DPOT_ACT_ELU: return x > 0.0f ? 1.0f : expf(x);
    case DPOT_ACT_SILU: { float s = 1.0f / (1.0f + expf(-x)); return s * (1.0f + x * (1.0f - s)); }
    default: return 1.0f;
  }
}

// d act(x)/dx for the hot backward epilogues: GELU' = Phi(x) + x phi(x) with the branch-free erf above and one
// __expf (abs error ~1e-7, ~30 instructions instead of ~70 for erff + expf); other activations as act_grad.
__device__ __forceinline__ float act_grad_fast(float x, int act) {
  if (act == DPOT_ACT_GELU) {
    const float cdf = 0.5f * (1.0f + erf_select(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return fmaf(x, pdf, cdf);
  }
  return act_grad(x, act);
}

// GroupNorm by reference: a consumer kernel gets the raw statistics (double sum / sum of squares per (sample,
// group), as accumulated by the producer's epilogue) plus gamma/beta and derives its channel's affine itself --
// x' = x*sc + sh with sc = rstd*gamma, sh = beta - mean*sc -- instead of reading tables written by a separate
// finalize launch.  Mean/variance in double (cancellation), the reciprocal square root in fp32 with one Newton
// step (~1e-7 relative): ~25 instructions per thread, no fp64 division or sqrt.
struct GnRef {
  const double* stats; const float* gamma; const float* beta; int groups; int gs; float eps; double inv_cnt;
};
__device__ __forceinline__ void gn_affine_ref(const GnRef& r, int b, int ch, float& sc, float& sh) {
  const double* s = r.stats + ((int64_t)b * r.groups + ch / r.gs) * 2;
  const double mean = s[0] * r.inv_cnt;
  const double var = fma(-mean, mean, s[1] * r.inv_cnt);
  const float v = fmaxf((float)var, 0.f) + r.eps;
  float rstd = rsqrtf(v);
  rstd = rstd * fmaf(-0.5f * v * rstd, rstd, 1.5f);
  sc = rstd * __ldg(r.gamma + ch);
  sh = fmaf(-(float)mean, sc, __ldg(r.beta + ch));
}
static inline GnRef make_gn_ref(const double* stats, const float* gamma, const float* beta, int groups, float eps, int E,
                                int64_t n) {
  GnRef r;
  r.stats = stats; r.gamma = gamma; r.beta = beta; r.groups = groups; r.gs = E / groups; r.eps = eps;
  r.inv_cnt = 1.0 / ((double)(E / groups) * (double)n);
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dpot
