// The steps either side of model(xx) in the autoregressive training loop (train_temporal.py:201-230), as kernels:
//   * global gradient norm for clip_grad_norm_ (:228) -- multi-tensor sum of squares into ONE device double; the clip
//     coefficient itself is applied inside the fused Adam kernel (adam.cu), so gradients are read once per step;
//   * noise injection  xx += s * ||xx||_{(X,Y,T)} * randn  (:205) with a counter-based Philox-4x32-10 generator:
//     no state tensor, the backward pass regenerates the same normals from (seed, offset);
//   * SimpleLpLoss(size_average=False)  (utils/criterion.py:38-59): masked per-channel relative L2, summed over the
//     batch and divided by the number of active channels, forward + backward.
#include "common.cuh"

namespace dpot {
namespace {

// ------------------------------------------------------------------------------------------ gradient norm
constexpr int SQ_MAX = 64, SQ_NT = 256, SQ_PER_BLOCK = SQ_NT * 4 * 8;
struct SqArgs {
  const float* g[SQ_MAX]; int64_t n[SQ_MAX]; int blk_start[SQ_MAX + 1]; int count;
};

__global__ void __launch_bounds__(SQ_NT) grad_sqnorm_kernel(const SqArgs a, double* __restrict__ out) {
  __shared__ double red[SQ_NT / 32];
  int ti = 0;
  while (ti + 1 < a.count && (int)blockIdx.x >= a.blk_start[ti + 1]) ++ti;
  const int64_t base = (int64_t)(blockIdx.x - a.blk_start[ti]) * SQ_PER_BLOCK;
  const float* __restrict__ g = a.g[ti];
  const int64_t n = a.n[ti];
  const bool vec = reinterpret_cast<uintptr_t>(g) % 16 == 0;
  float s = 0.f;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int64_t i = base + ((int64_t)it * SQ_NT + threadIdx.x) * 4;
    if (i >= n) break;
    if (vec && i + 4 <= n) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g + i));
      s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
    } else {
      for (int64_t k = i; k < n && k < i + 4; ++k) s = fmaf(g[k], g[k], s);
    }
  }
  double d = warp_sum((double)s);       // <= 32 fp32 terms per thread, everything above in double
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < SQ_NT / 32; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

// ------------------------------------------------------------------------------------------ Philox-4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
// four standard normals from one Philox block (Box-Muller on two uniform pairs)
__device__ __forceinline__ void normal4(uint64_t seed, uint64_t offset, uint64_t idx, float (&z)[4]) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float u0 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)r.z + 0.5f) * 2.3283064365386963e-10f, u3 = ((float)r.w + 0.5f) * 2.3283064365386963e-10f;
  const float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
  float s, c;
  __sincosf(6.283185307179586f * u1, &s, &c);
  z[0] = ra * c; z[1] = ra * s;
  __sincosf(6.283185307179586f * u3, &s, &c);
  z[2] = rb * c; z[3] = rb * s;
}

// per (sample, channel) sum of squares over (X, Y, T):  x[b, pos, c], pos = 0..npos-1, channel innermost
__global__ void __launch_bounds__(256) chan_sumsq_kernel(const float* __restrict__ x, int64_t npos, int C, int nsplit,
                                                         double* __restrict__ out /* [B, C] */) {
  __shared__ double red[256];
  const int b = blockIdx.y, sp = blockIdx.x;
  const int64_t total = npos * C, lo = total * sp / nsplit, hi = total * (sp + 1) / nsplit;
  const float* xp = x + (int64_t)b * total;
  // thread t walks elements congruent to its start modulo blockDim*... keep the channel of a thread fixed:
  // stride = lcm-free choice: step 256*C would be wasteful; instead accumulate per thread for channel (i % C) in a small array
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
    const float v = xp[i];
    const int c = (int)(i % C);
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = (k == c) ? fmaf(v, v, acc[k]) : acc[k];
  }
  for (int c = 0; c < C; ++c) {
    red[threadIdx.x] = (double)acc[c];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out + (int64_t)b * C + c, red[0]);
    __syncthreads();
  }
}

// y = x + scale * sqrt(sumsq[b,c]) * N(0,1)      (forward of the noise injection)
// BWD: dx = dy + scale * x / sqrt(sumsq[b,c]) * dot[b,c],  dot[b,c] = sum_pos dy * eps  (the norm is differentiated, as autograd does)
template <int MODE>   // 0 forward, 1 = accumulate dot[b,c] = sum dy*eps, 2 = dx = dy + scale * x / norm * dot
__global__ void __launch_bounds__(256) noise_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                    const double* __restrict__ sumsq, double* __restrict__ dot, int64_t per_sample,
                                                    int C, float scale, uint64_t seed, uint64_t offset, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;          // group of 4 consecutive elements of the sample
  const int64_t i0 = q * 4;
  float dsum[16];
  if (MODE == 1) {
#pragma unroll
    for (int c = 0; c < 16; ++c) dsum[c] = 0.f;
  }
  if (i0 < per_sample) {
    float z[4];
    if (MODE != 2) normal4(seed, offset, (uint64_t)b * (uint64_t)((per_sample + 3) / 4) + (uint64_t)q, z);
    const int64_t g0 = (int64_t)b * per_sample + i0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k >= per_sample) break;
      const int c = (int)((i0 + k) % C);
      if (MODE == 0) {
        const float nrm = (float)sqrt(sumsq[(int64_t)b * C + c]);
        out[g0 + k] = fmaf(scale * nrm, z[k], x[g0 + k]);
      } else if (MODE == 1) {
        const float v = dy[g0 + k] * z[k];
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) dsum[cc] += (cc == c) ? v : 0.f;
      } else {
        const double ss = sumsq[(int64_t)b * C + c];
        const float coef = ss > 0.0 ? scale * (float)(dot[(int64_t)b * C + c] / sqrt(ss)) : 0.f;
        out[g0 + k] = fmaf(coef, x[g0 + k], dy[g0 + k]);
      }
    }
  }
  if (MODE == 1) {
    for (int c = 0; c < C; ++c) {
      const float w = warp_sum(dsum[c]);
      if ((threadIdx.x & 31) == 0 && w != 0.f) atomicAdd(dot + (int64_t)b * C + c, (double)w);
    }
  }
}

// ------------------------------------------------------------------------------------------ SimpleLpLoss
// partial[b, c, 0..2] += sum_pos ((x - y) m)^2, sum_pos (y m)^2, sum_pos m      (mask m[b, pos', c] broadcast over T)
__global__ void __launch_bounds__(256) lp_partial_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         const float* __restrict__ mask, int64_t nxy, int T, int C, int nsplit,
                                                         double* __restrict__ partial) {
  __shared__ double red[3][256];
  const int b = blockIdx.y, sp = blockIdx.x;
  const int64_t total = nxy * T * C, lo = total * sp / nsplit, hi = total * (sp + 1) / nsplit;
  float d2[16], y2[16], ms[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) d2[c] = y2[c] = ms[c] = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
    const int c = (int)(i % C);
    const int64_t pos = i / ((int64_t)T * C);
    const float m = mask ? mask[((int64_t)b * nxy + pos) * C + c] : 1.f;
    const float xv = x[(int64_t)b * total + i] * m, yv = y[(int64_t)b * total + i] * m;
    const float d = xv - yv;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const bool h = k == c;
      d2[k] = h ? fmaf(d, d, d2[k]) : d2[k];
      y2[k] = h ? fmaf(yv, yv, y2[k]) : y2[k];
      ms[k] = h ? ms[k] + m : ms[k];
    }
  }
  for (int c = 0; c < C; ++c) {
    red[0][threadIdx.x] = (double)d2[c]; red[1][threadIdx.x] = (double)y2[c]; red[2][threadIdx.x] = (double)ms[c];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o)
        for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x < 3) atomicAdd(partial + ((int64_t)b * C + c) * 3 + threadIdx.x, red[threadIdx.x][0]);
    __syncthreads();
  }
}

// loss = sum_b ( sum_c d_bc / (yn_bc + 1e-8) ) / nch_b ;  coef[b,c] = 1 / (d_bc * (yn_bc + 1e-8) * nch_b)  (0 where d = 0)
__global__ void lp_finalize_kernel(const double* __restrict__ partial, int B, int C, int has_mask, float* __restrict__ loss,
                                   float* __restrict__ coef, int accumulate) {
  double total = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int nch = 0;
    for (int c = 0; c < C; ++c) nch += (!has_mask || partial[((int64_t)b * C + c) * 3 + 2] != 0.0) ? 1 : 0;
    double s = 0.0;
    for (int c = 0; c < C; ++c) {
      // fp32 norms like torch.norm on fp32 tensors, then the division in fp32 (utils/criterion.py:52-59)
      const float d = sqrtf((float)partial[((int64_t)b * C + c) * 3]);
      const float yn = sqrtf((float)partial[((int64_t)b * C + c) * 3 + 1]) + 1e-8f;
      s += (double)(d / yn);
      coef[(int64_t)b * C + c] = (d > 0.f && nch > 0) ? 1.0f / (d * yn * (float)nch) : 0.f;
    }
    total += nch > 0 ? s / (double)nch : (s > 0.0 ? INFINITY : NAN);      // 0/0 -> nan, x/0 -> inf like torch
  }
  __shared__ double red[32];
  total = warp_sum(total);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) / 32); ++w) t += red[w];
    loss[0] = accumulate ? loss[0] + (float)t : (float)t;
  }
}

// dx = gscale * coef[b,c] * m^2 * (x - y)
__global__ void __launch_bounds__(256) lp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                     const float* __restrict__ mask, const float* __restrict__ coef,
                                                     const float* __restrict__ gscale, int64_t nxy, int T, int C,
                                                     float* __restrict__ dx) {
  const int b = blockIdx.y;
  const int64_t total = nxy * T * C;
  const float gs = gscale ? gscale[0] : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % C);
    const int64_t pos = i / ((int64_t)T * C);
    const float m = mask ? mask[((int64_t)b * nxy + pos) * C + c] : 1.f;
    dx[(int64_t)b * total + i] = gs * coef[(int64_t)b * C + c] * m * m * (x[(int64_t)b * total + i] - y[(int64_t)b * total + i]);
  }
}

}  // namespace
}  // namespace dpot

using namespace dpot;

extern "C" int dpot_grad_sqnorm(const float* const* g, const int64_t* n, int32_t count, double* out_sq, void* stream) {
  DPOT_REQUIRE(count >= 0 && out_sq, DPOT_E_BADARG, "dpot_grad_sqnorm: bad arguments");
  cudaStream_t st = as_stream(stream);
  DPOT_CUDA(cudaMemsetAsync(out_sq, 0, sizeof(double), st));
  int t = 0;
  while (t < count) {
    SqArgs a;
    a.count = 0;
    int blk = 0;
    for (; t < count && a.count < SQ_MAX; ++t) {
      DPOT_REQUIRE(n[t] >= 0, DPOT_E_BADARG, "dpot_grad_sqnorm: negative size of tensor %d", t);
      if (n[t] == 0) continue;
      DPOT_REQUIRE(g[t], DPOT_E_BADARG, "dpot_grad_sqnorm: null pointer in tensor %d", t);
      const int k = a.count++;
      a.g[k] = g[t]; a.n[k] = n[t]; a.blk_start[k] = blk;
      blk += (int)ceil_div(n[t], SQ_PER_BLOCK);
    }
    a.blk_start[a.count] = blk;
    if (blk == 0) continue;
    grad_sqnorm_kernel<<<(unsigned)blk, SQ_NT, 0, st>>>(a, out_sq);
    DPOT_LAUNCH_CHECK("grad_sqnorm_kernel");
  }
  return 0;
}

extern "C" int dpot_chan_sumsq(const float* x, int32_t B, int64_t npos, int32_t C, double* out, void* stream) {
  DPOT_REQUIRE(x && out && B > 0 && npos > 0 && C >= 1 && C <= 16, DPOT_E_BADARG, "dpot_chan_sumsq: bad arguments (C <= 16)");
  cudaStream_t st = as_stream(stream);
  DPOT_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)B * C, st));
  const int nsplit = (int)(npos * C >= (1 << 16) ? 32 : 1);
  chan_sumsq_kernel<<<dim3(nsplit, B), 256, 0, st>>>(x, npos, C, nsplit, out);
  DPOT_LAUNCH_CHECK("chan_sumsq_kernel");
  return 0;
}

extern "C" int dpot_noise_inject(const float* x, int32_t B, int64_t npos, int32_t C, float scale, uint64_t seed,
                                 uint64_t offset, double* sumsq, float* out, void* stream) {
  DPOT_REQUIRE(x && out && sumsq, DPOT_E_BADARG, "dpot_noise_inject: null pointer");
  DPOT_CALL(dpot_chan_sumsq(x, B, npos, C, sumsq, stream));
  const int64_t per = npos * C;
  noise_kernel<0><<<dim3((unsigned)ceil_div(ceil_div(per, 4), 256), B), 256, 0, as_stream(stream)>>>(
      x, nullptr, sumsq, nullptr, per, C, scale, seed, offset, out);
  DPOT_LAUNCH_CHECK("noise_kernel");
  return 0;
}

extern "C" int dpot_noise_inject_bwd(const float* x, const float* dy, int32_t B, int64_t npos, int32_t C, float scale,
                                     uint64_t seed, uint64_t offset, const double* sumsq, double* dot, float* dx,
                                     void* stream) {
  DPOT_REQUIRE(x && dy && sumsq && dot && dx && C >= 1 && C <= 16, DPOT_E_BADARG, "dpot_noise_inject_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  const int64_t per = npos * C;
  DPOT_CUDA(cudaMemsetAsync(dot, 0, sizeof(double) * (size_t)B * C, st));
  const dim3 grid((unsigned)ceil_div(ceil_div(per, 4), 256), B);
  noise_kernel<1><<<grid, 256, 0, st>>>(x, dy, sumsq, dot, per, C, scale, seed, offset, nullptr);
  DPOT_LAUNCH_CHECK("noise_kernel");
  noise_kernel<2><<<grid, 256, 0, st>>>(x, dy, sumsq, dot, per, C, scale, seed, offset, dx);
  DPOT_LAUNCH_CHECK("noise_kernel");
  return 0;
}

extern "C" int dpot_lp_loss(const float* x, const float* y, const float* mask, int32_t B, int64_t nxy, int32_t T, int32_t C,
                            double* partial, float* coef, float* loss, int32_t accumulate, void* stream) {
  DPOT_REQUIRE(x && y && partial && coef && loss && B > 0 && nxy > 0 && T > 0 && C >= 1 && C <= 16, DPOT_E_BADARG,
               "dpot_lp_loss: bad arguments (C <= 16)");
  cudaStream_t st = as_stream(stream);
  DPOT_CUDA(cudaMemsetAsync(partial, 0, sizeof(double) * 3 * (size_t)B * C, st));
  const int nsplit = (int)(nxy * T * C >= (1 << 16) ? 16 : 1);
  lp_partial_kernel<<<dim3(nsplit, B), 256, 0, st>>>(x, y, mask, nxy, T, C, nsplit, partial);
  DPOT_LAUNCH_CHECK("lp_partial_kernel");
  lp_finalize_kernel<<<1, 256, 0, st>>>(partial, B, C, mask ? 1 : 0, loss, coef, accumulate);
  DPOT_LAUNCH_CHECK("lp_finalize_kernel");
  return 0;
}

extern "C" int dpot_lp_loss_bwd(const float* x, const float* y, const float* mask, const float* coef, const float* gscale,
                                int32_t B, int64_t nxy, int32_t T, int32_t C, float* dx, void* stream) {
  DPOT_REQUIRE(x && y && coef && dx && B > 0 && C >= 1, DPOT_E_BADARG, "dpot_lp_loss_bwd: bad arguments");
  const int64_t total = nxy * T * C;
  const unsigned gx = (unsigned)(ceil_div(total, 256) < 1024 ? ceil_div(total, 256) : 1024);
  lp_bwd_kernel<<<dim3(gx, B), 256, 0, as_stream(stream)>>>(x, y, mask, coef, gscale, nxy, T, C, dx);
  DPOT_LAUNCH_CHECK("lp_bwd_kernel");
  return 0;
}
