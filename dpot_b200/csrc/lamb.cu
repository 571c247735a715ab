// Fused multi-tensor Lamb step: Lamb.step, utils/optimizer.py:421-499 (the `--opt lamb` branch of train_temporal.py /
// evaluate.py:135-136).  The reference walks the parameters one by one with ~12 elementwise kernels, two norm
// reductions and two host synchronisations (`weight_norm == 0`, the tensor-valued `alpha`) per tensor; here the step is
// two launches per 56 tensors and no host synchronisation:
//
//   stage 1   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  u = m / (sqrt(v) + eps) [+ wd p]       :459-476
//             per-tensor sum p^2 and sum u^2 (fp32 per thread, double across threads / blocks)     :474, 479
//   stage 2   weight_norm = min(||p||, clamp);  trust = 1 if either norm is 0 else weight_norm / ||u||;
//             p -= lr * bias_correction * (adam ? 1 : trust) * u      (u recomputed from m, v, p: bit-identical)  :480-491
//
// HBM: stage 1 reads p, g, m, v and writes m, v (24 B / parameter), stage 2 reads p, m, v and writes p (16 B).
#include "common.cuh"

namespace dpot {
namespace {

constexpr int LT_MAX = 56;
constexpr int LAMB_NT = 256, LAMB_PER_BLOCK = LAMB_NT * 4 * 4;      // 4 float4 per thread

struct LambArgs {
  float* p[LT_MAX]; const float* g[LT_MAX]; float* m[LT_MAX]; float* v[LT_MAX];
  int64_t n[LT_MAX];
  int blk_start[LT_MAX + 1];
  int tidx[LT_MAX];             // index of the tensor in the caller's list (norms / info slots)
  float step_size[LT_MAX];      // lr * bias_correction of the tensor's own step count
  int count;
};
struct LambConst {
  float beta1, beta2, one_m_beta1, one_m_beta2, eps, wd, clamp;
  int adam;
  double* norms;                // [2 * tensors]: sum p^2, sum u^2
  float* info;                  // [3 * tensors]: weight_norm, adam_norm, trust_ratio (the reference's state entries)
};

// the update direction of one element; the same instruction sequence in both stages (explicit roundings: no contraction
// differences between the two kernels)
__device__ __forceinline__ float lamb_dir(float p, float m, float v, const LambConst& c) {
  float u = __fdiv_rn(m, __fadd_rn(__fsqrt_rn(v), c.eps));          // exp_avg / exp_avg_sq.sqrt().add(eps)   :476
  if (c.wd != 0.f) u = __fmaf_rn(c.wd, p, u);                        // adam_step.add_(p, alpha=wd)            :477-478
  return u;
}

__device__ __forceinline__ int tensor_of_block(const LambArgs& a) {
  int ti = 0;
  while (ti + 1 < a.count && (int)blockIdx.x >= a.blk_start[ti + 1]) ++ti;
  return ti;
}

__global__ void __launch_bounds__(LAMB_NT) lamb_stage1_kernel(const LambArgs a, const LambConst c) {
  const int ti = tensor_of_block(a);
  const int64_t base = (int64_t)(blockIdx.x - a.blk_start[ti]) * LAMB_PER_BLOCK;
  const float* __restrict__ p = a.p[ti]; const float* __restrict__ g = a.g[ti];
  float* __restrict__ m = a.m[ti]; float* __restrict__ v = a.v[ti];
  const int64_t n = a.n[ti];
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) % 16 == 0);
  float sp = 0.f, su = 0.f;
  auto elem = [&](float P, float G, float& M, float& V) {
    M = __fmaf_rn(c.one_m_beta1, G, M * c.beta1);                    // :463
    V = __fmaf_rn(c.one_m_beta2, G * G, V * c.beta2);                // :465
    const float u = lamb_dir(P, M, V, c);
    sp = __fmaf_rn(P, P, sp);
    su = __fmaf_rn(u, u, su);
  };
  if (vec && base + LAMB_PER_BLOCK <= n) {
    float4 P[4], G[4], M[4], V[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int64_t i = base + ((int64_t)it * LAMB_NT + threadIdx.x) * 4;
      P[it] = *reinterpret_cast<const float4*>(p + i);
      G[it] = __ldcs(reinterpret_cast<const float4*>(g + i));
      M[it] = *reinterpret_cast<const float4*>(m + i);
      V[it] = *reinterpret_cast<const float4*>(v + i);
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int64_t i = base + ((int64_t)it * LAMB_NT + threadIdx.x) * 4;
      elem(P[it].x, G[it].x, M[it].x, V[it].x);
      elem(P[it].y, G[it].y, M[it].y, V[it].y);
      elem(P[it].z, G[it].z, M[it].z, V[it].z);
      elem(P[it].w, G[it].w, M[it].w, V[it].w);
      *reinterpret_cast<float4*>(m + i) = M[it];
      *reinterpret_cast<float4*>(v + i) = V[it];
    }
  } else {
    for (int it = 0; it < 16; ++it) {
      const int64_t k = base + (int64_t)it * LAMB_NT + threadIdx.x;
      if (k >= n) break;
      float M = m[k], V = v[k];
      elem(p[k], g[k], M, V);
      m[k] = M; v[k] = V;
    }
  }
  // block reduction: fp32 within the warp's 16-64 values per thread is already summed; double from here on
  double dp = (double)sp, du = (double)su;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dp += __shfl_xor_sync(0xffffffffu, dp, o);
    du += __shfl_xor_sync(0xffffffffu, du, o);
  }
  __shared__ double red[2][LAMB_NT / 32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = dp; red[1][w] = du; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tp = 0.0, tu = 0.0;
#pragma unroll
    for (int i = 0; i < LAMB_NT / 32; ++i) { tp += red[0][i]; tu += red[1][i]; }
    atomicAdd(c.norms + 2 * a.tidx[ti], tp);
    atomicAdd(c.norms + 2 * a.tidx[ti] + 1, tu);
  }
}

__global__ void __launch_bounds__(LAMB_NT) lamb_stage2_kernel(const LambArgs a, const LambConst c) {
  const int ti = tensor_of_block(a);
  const int64_t base = (int64_t)(blockIdx.x - a.blk_start[ti]) * LAMB_PER_BLOCK;
  float* __restrict__ p = a.p[ti];
  const float* __restrict__ m = a.m[ti]; const float* __restrict__ v = a.v[ti];
  const int64_t n = a.n[ti];
  const int t = a.tidx[ti];
  // torch.norm(p).clamp(0, clamp_value), torch.norm(adam_step), their ratio: fp32 like the reference's 0-dim tensors
  const float wn = fminf((float)sqrt(c.norms[2 * t]), c.clamp);
  const float un = (float)sqrt(c.norms[2 * t + 1]);
  const float trust = (wn == 0.f || un == 0.f) ? 1.f : __fdiv_rn(wn, un);                 // :481-484
  if (blockIdx.x == (unsigned)a.blk_start[ti] && threadIdx.x == 0 && c.info) {
    c.info[3 * t] = wn; c.info[3 * t + 1] = un; c.info[3 * t + 2] = trust;                 // :485-487
  }
  const float alpha = -a.step_size[ti] * (c.adam ? 1.f : trust);                              // :488-491
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16 == 0);
  if (vec && base + LAMB_PER_BLOCK <= n) {
    float4 P[4], M[4], V[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int64_t i = base + ((int64_t)it * LAMB_NT + threadIdx.x) * 4;
      P[it] = *reinterpret_cast<const float4*>(p + i);
      M[it] = __ldcs(reinterpret_cast<const float4*>(m + i));
      V[it] = __ldcs(reinterpret_cast<const float4*>(v + i));
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int64_t i = base + ((int64_t)it * LAMB_NT + threadIdx.x) * 4;
      P[it].x = __fmaf_rn(alpha, lamb_dir(P[it].x, M[it].x, V[it].x, c), P[it].x);
      P[it].y = __fmaf_rn(alpha, lamb_dir(P[it].y, M[it].y, V[it].y, c), P[it].y);
      P[it].z = __fmaf_rn(alpha, lamb_dir(P[it].z, M[it].z, V[it].z, c), P[it].z);
      P[it].w = __fmaf_rn(alpha, lamb_dir(P[it].w, M[it].w, V[it].w, c), P[it].w);
      *reinterpret_cast<float4*>(p + i) = P[it];
    }
  } else {
    for (int it = 0; it < 16; ++it) {
      const int64_t k = base + (int64_t)it * LAMB_NT + threadIdx.x;
      if (k >= n) break;
      const float P = p[k];
      p[k] = __fmaf_rn(alpha, lamb_dir(P, m[k], v[k], c), P);
    }
  }
}

}  // namespace
}  // namespace dpot

using namespace dpot;

extern "C" int dpot_lamb_step_multi(float* const* p, const float* const* g, float* const* m, float* const* v, const int64_t* n,
                                    int32_t count, double lr, double beta1, double beta2, double eps, double weight_decay,
                                    double clamp_value, const int32_t* steps, int32_t debias, int32_t adam, double* norms, float* info,
                                    void* stream) {
  DPOT_REQUIRE(count >= 0, DPOT_E_BADARG, "dpot_lamb_step_multi: negative count");
  if (count == 0) return 0;
  DPOT_REQUIRE(p && g && m && v && n && norms && steps, DPOT_E_BADARG, "dpot_lamb_step_multi: null array");
  LambConst c;
  c.beta1 = (float)beta1; c.beta2 = (float)beta2;
  c.one_m_beta1 = (float)(1.0 - beta1); c.one_m_beta2 = (float)(1.0 - beta2);
  c.eps = (float)eps; c.wd = (float)weight_decay; c.clamp = (float)clamp_value; c.adam = adam;
  c.norms = norms; c.info = info;
  cudaStream_t st = as_stream(stream);
  DPOT_CUDA(cudaMemsetAsync(norms, 0, sizeof(double) * 2 * (size_t)count, st));
  if (info) DPOT_CUDA(cudaMemsetAsync(info, 0, sizeof(float) * 3 * (size_t)count, st));
  for (int stage = 1; stage <= 2; ++stage) {
    int t = 0;
    while (t < count) {
      LambArgs a;
      a.count = 0;
      int blk = 0;
      for (; t < count && a.count < LT_MAX; ++t) {
        DPOT_REQUIRE(n[t] >= 0 && steps[t] >= 1, DPOT_E_BADARG, "dpot_lamb_step_multi: bad size / step of tensor %d", t);
        if (n[t] == 0) continue;
        DPOT_REQUIRE(p[t] && g[t] && m[t] && v[t], DPOT_E_BADARG, "dpot_lamb_step_multi: null pointer in tensor %d", t);
        const int k = a.count++;
        a.p[k] = p[t]; a.g[k] = g[t]; a.m[k] = m[t]; a.v[k] = v[t]; a.n[k] = n[t]; a.tidx[k] = t;
        // bias_correction in python floats (:468-472), folded into the step size (:475)
        const double corr = debias ? sqrt(1.0 - pow(beta2, (double)steps[t])) / (1.0 - pow(beta1, (double)steps[t])) : 1.0;
        a.step_size[k] = (float)(lr * corr);
        a.blk_start[k] = blk;
        blk += (int)ceil_div(n[t], LAMB_PER_BLOCK);
      }
      a.blk_start[a.count] = blk;
      if (blk == 0) continue;
      if (stage == 1) {
        lamb_stage1_kernel<<<(unsigned)blk, LAMB_NT, 0, st>>>(a, c);
        DPOT_LAUNCH_CHECK("lamb_stage1_kernel");
      } else {
        lamb_stage2_kernel<<<(unsigned)blk, LAMB_NT, 0, st>>>(a, c);
        DPOT_LAUNCH_CHECK("lamb_stage2_kernel");
      }
    }
  }
  return 0;
}
