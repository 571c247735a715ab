// Fused Adam / AdamW step: adam() utils/optimizer.py:9-52, adamw() :170-212.
// HBM-bound: reads p,g,m,v and writes p,m,v = 28 B/param (32 B with amsgrad).  One pass,
// float4 vectorised main body + scalar tail; the multi-tensor form packs up to 32 tensors per launch.
#include "common.cuh"

namespace dpot {
namespace {

struct AdamConst {
  float beta1, beta2, one_m_beta1, one_m_beta2, eps, wd, decay_mul, grad_scale;
  int decoupled;
  // clip_grad_norm_ folded into the step (train_temporal.py:228): sqnorm = device double holding sum g^2 over ALL
  // gradients (dpot_grad_sqnorm), max_norm the clip threshold; nullptr = no clipping
  const double* sqnorm; float max_norm;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float* vmax, const AdamConst& c,
                                          float step_size, float inv_sqrt_bc2_dummy, float sqrt_bc2) {
  (void)inv_sqrt_bc2_dummy;
  g *= c.grad_scale;
  if (c.decoupled) p *= c.decay_mul;                 // param.mul_(1 - lr*wd)            :193
  else if (c.wd != 0.f) g = fmaf(c.wd, p, g);        // grad.add(param, alpha=wd)         :36-37
  m = fmaf(c.one_m_beta1, g, m * c.beta1);           // exp_avg.mul_(b1).add_(g, 1-b1)    :40
  v = fmaf(c.one_m_beta2, g * g, v * c.beta2);       // exp_avg_sq.mul_(b2).addcmul_(g,g) :41
  float vv = v;
  if (vmax) { vv = fmaxf(*vmax, v); *vmax = vv; }    // amsgrad                            :44
  const float denom = sqrtf(vv) / sqrt_bc2 + c.eps;  //                                    :46-48
  p = p - step_size * (m / denom);                   // addcdiv_(m, denom, -step_size)    :50-52
}

constexpr int MT_MAX = 56;     // 56 x 68 B of tables = 3.9 KB of kernel parameters (limit 4 KB)
struct MultiArgs {
  float* p[MT_MAX]; const float* g[MT_MAX]; float* m[MT_MAX]; float* v[MT_MAX]; float* vmax[MT_MAX];
  int64_t n[MT_MAX];
  int blk_start[MT_MAX + 1];
  float step_size[MT_MAX], sqrt_bc2[MT_MAX];
  int count;
};

constexpr int ADAM_NT = 256, ADAM_PER_BLOCK = ADAM_NT * 4 * 4;  // 4 float4 per thread

__global__ void __launch_bounds__(ADAM_NT) adam_multi_kernel(const MultiArgs a, AdamConst c) {
  if (c.sqnorm) {
    // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1, in fp32; the norm is
    // that of the gradients AFTER the data-parallel average (grad_scale)
    const float total = (float)sqrt(*c.sqnorm) * c.grad_scale;
    c.grad_scale *= fminf(1.0f, c.max_norm / (total + 1e-6f));
  }
  // locate the tensor of this block (<= 32 entries: linear scan)
  int ti = 0;
  while (ti + 1 < a.count && (int)blockIdx.x >= a.blk_start[ti + 1]) ++ti;
  const int64_t base = (int64_t)(blockIdx.x - a.blk_start[ti]) * ADAM_PER_BLOCK;
  float* __restrict__ p = a.p[ti]; const float* __restrict__ g = a.g[ti];
  float* __restrict__ m = a.m[ti]; float* __restrict__ v = a.v[ti]; float* __restrict__ vm = a.vmax[ti];
  const int64_t n = a.n[ti];
  const float ss = a.step_size[ti], sb = a.sqrt_bc2[ti];
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(vm)) % 16 == 0);
  if (vec && !vm && base + ADAM_PER_BLOCK <= n) {
    // full block, no amsgrad: all 16 loads of the thread are issued before the first use (bytes in flight, HBM-bound)
    float4 P[4], Gv[4], M[4], V[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int64_t i = base + ((int64_t)it * ADAM_NT + threadIdx.x) * 4;
      P[it] = *reinterpret_cast<const float4*>(p + i);
      Gv[it] = __ldcs(reinterpret_cast<const float4*>(g + i));
      M[it] = *reinterpret_cast<const float4*>(m + i);
      V[it] = *reinterpret_cast<const float4*>(v + i);
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int64_t i = base + ((int64_t)it * ADAM_NT + threadIdx.x) * 4;
      adam_elem(P[it].x, Gv[it].x, M[it].x, V[it].x, nullptr, c, ss, 0.f, sb);
      adam_elem(P[it].y, Gv[it].y, M[it].y, V[it].y, nullptr, c, ss, 0.f, sb);
      adam_elem(P[it].z, Gv[it].z, M[it].z, V[it].z, nullptr, c, ss, 0.f, sb);
      adam_elem(P[it].w, Gv[it].w, M[it].w, V[it].w, nullptr, c, ss, 0.f, sb);
      *reinterpret_cast<float4*>(p + i) = P[it];
      *reinterpret_cast<float4*>(m + i) = M[it];
      *reinterpret_cast<float4*>(v + i) = V[it];
    }
    return;
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int64_t i = base + ((int64_t)it * ADAM_NT + threadIdx.x) * 4;
    if (i >= n) break;
    if (vec && i + 4 <= n) {
      float4 P = *reinterpret_cast<float4*>(p + i);
      const float4 G = *reinterpret_cast<const float4*>(g + i);
      float4 M = *reinterpret_cast<float4*>(m + i);
      float4 V = *reinterpret_cast<float4*>(v + i);
      float4 X = vm ? *reinterpret_cast<float4*>(vm + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      adam_elem(P.x, G.x, M.x, V.x, vm ? &X.x : nullptr, c, ss, 0.f, sb);
      adam_elem(P.y, G.y, M.y, V.y, vm ? &X.y : nullptr, c, ss, 0.f, sb);
      adam_elem(P.z, G.z, M.z, V.z, vm ? &X.z : nullptr, c, ss, 0.f, sb);
      adam_elem(P.w, G.w, M.w, V.w, vm ? &X.w : nullptr, c, ss, 0.f, sb);
      *reinterpret_cast<float4*>(p + i) = P;
      *reinterpret_cast<float4*>(m + i) = M;
      *reinterpret_cast<float4*>(v + i) = V;
      if (vm) *reinterpret_cast<float4*>(vm + i) = X;
    } else {
      for (int64_t k = i; k < n && k < i + 4; ++k) {
        float P = p[k], M = m[k], V = v[k];
        float X = vm ? vm[k] : 0.f;
        adam_elem(P, g[k], M, V, vm ? &X : nullptr, c, ss, 0.f, sb);
        p[k] = P; m[k] = M; v[k] = V;
        if (vm) vm[k] = X;
      }
    }
  }
}

}  // namespace
}  // namespace dpot

using namespace dpot;

extern "C" int dpot_adam_step_multi(float* const* p, const float* const* g, float* const* m, float* const* v,
                                    float* const* vmax, const int64_t* n, int32_t count, double lr, double beta1,
                                    double beta2, double eps, double weight_decay, const int32_t* steps,
                                    int32_t decoupled, double grad_scale, void* stream) {
  return dpot_adam_step_multi_clip(p, g, m, v, vmax, n, count, lr, beta1, beta2, eps, weight_decay, steps, decoupled,
                                   grad_scale, nullptr, 0.0, stream);
}

extern "C" int dpot_adam_step_multi_clip(float* const* p, const float* const* g, float* const* m, float* const* v,
                                         float* const* vmax, const int64_t* n, int32_t count, double lr, double beta1,
                                         double beta2, double eps, double weight_decay, const int32_t* steps,
                                         int32_t decoupled, double grad_scale, const double* grad_sqnorm, double max_norm,
                                         void* stream) {
  DPOT_REQUIRE(count >= 0, DPOT_E_BADARG, "dpot_adam_step_multi: negative count");
  if (count == 0) return 0;
  DPOT_REQUIRE(p && g && m && v && n && steps, DPOT_E_BADARG, "dpot_adam_step_multi: null array");
  AdamConst c;
  c.beta1 = (float)beta1; c.beta2 = (float)beta2;
  c.one_m_beta1 = (float)(1.0 - beta1); c.one_m_beta2 = (float)(1.0 - beta2);  // python-float arithmetic, then fp32
  c.eps = (float)eps; c.wd = (float)weight_decay; c.decay_mul = (float)(1.0 - lr * weight_decay);
  c.grad_scale = (float)grad_scale; c.decoupled = decoupled;
  c.sqnorm = grad_sqnorm; c.max_norm = (float)max_norm;
  cudaStream_t st = as_stream(stream);
  int t = 0;
  while (t < count) {
    MultiArgs a;
    a.count = 0;
    int blk = 0;
    for (; t < count && a.count < MT_MAX; ++t) {
      DPOT_REQUIRE(n[t] >= 0 && steps[t] >= 1, DPOT_E_BADARG, "dpot_adam_step_multi: bad size/step of tensor %d", t);
      if (n[t] == 0) continue;                      // empty tensors (NULL data pointers) are legal no-ops
      DPOT_REQUIRE(p[t] && g[t] && m[t] && v[t], DPOT_E_BADARG, "dpot_adam_step_multi: null pointer in tensor %d", t);
      const int k = a.count++;
      a.p[k] = p[t]; a.g[k] = g[t]; a.m[k] = m[t]; a.v[k] = v[t]; a.vmax[k] = vmax ? vmax[t] : nullptr;
      a.n[k] = n[t];
      a.blk_start[k] = blk;
      blk += (int)ceil_div(n[t], ADAM_PER_BLOCK);
      // bias corrections in double like the python floats of the reference (:33-34,50)
      const double bc1 = 1.0 - pow(beta1, (double)steps[t]);
      const double bc2 = 1.0 - pow(beta2, (double)steps[t]);
      a.step_size[k] = (float)(lr / bc1);
      a.sqrt_bc2[k] = (float)sqrt(bc2);
    }
    a.blk_start[a.count] = blk;
    if (blk == 0) continue;
    adam_multi_kernel<<<(unsigned)blk, ADAM_NT, 0, st>>>(a, c);
    DPOT_LAUNCH_CHECK("adam_multi_kernel");
  }
  return 0;
}

extern "C" int dpot_adam_step(float* p, const float* g, float* m, float* v, float* vmax, int64_t n, double lr,
                              double beta1, double beta2, double eps, double weight_decay, int32_t step,
                              int32_t decoupled, double grad_scale, void* stream) {
  float* pp[1] = {p}; const float* gg[1] = {g}; float* mm[1] = {m}; float* vv[1] = {v}; float* xx[1] = {vmax};
  int64_t nn[1] = {n}; int32_t ss[1] = {step};
  return dpot_adam_step_multi(pp, gg, mm, vv, vmax ? xx : nullptr, nn, 1, lr, beta1, beta2, eps, weight_decay, ss,
                              decoupled, grad_scale, stream);
}
