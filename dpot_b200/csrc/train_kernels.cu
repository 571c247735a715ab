// Kernels of the training step that are not dense contractions: gradient scaling, column sums, GroupNorm backward,
// partial-result reduction + un-packing of weight gradients, the output-head tail backward, the PatchEmbed conv0
// backward and the time-aggregation fold helpers.  Host orchestration: train_step.cu.  Reference: autograd of
// models/dpot.py:364-403 (train_temporal.py:227 loss.backward()).
#include "train_kernels.cuh"
#include "gemm_common.cuh"

namespace dpot {
namespace {

__device__ __forceinline__ float inv_of(const float* inv_scale) { return inv_scale ? __ldg(inv_scale) : 1.0f; }

// ------------------------------------------------------------------------------------------------ gradient scale
__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, unsigned* __restrict__ bits) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = fabsf(x[i]);
    m = (v == v) ? fmaxf(m, v) : INFINITY;          // NaN -> inf: the scale falls back to 1 and the NaN stays loud
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(bits, __float_as_uint(m));
}
__global__ void scale_from_max_kernel(const unsigned* __restrict__ bits, float* __restrict__ scale) {
  const float m = __uint_as_float(*bits);
  float S = 1.f;
  if (m > 0.f && m < INFINITY) {
    int ex;
    frexpf(m, &ex);                                  // m = f * 2^ex, f in [0.5, 1)  ->  m * 2^(1 - ex) in [1, 2)
    int k = 1 - ex;
    k = k < -60 ? -60 : (k > 60 ? 60 : k);
    S = ldexpf(1.f, k);
  }
  scale[0] = S;
  scale[1] = 1.f / S;
}

// ------------------------------------------------------------------------------------------------ column sums
template <bool F16>
__global__ void __launch_bounds__(128) colsum2_kernel(const void* __restrict__ Xv, int64_t ld, int64_t lo_off, int M, int N,
                                                      int rows_per_block, double* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc = 0.f;
  if (F16) {
    const __half* X = reinterpret_cast<const __half*>(Xv);
#pragma unroll 4
    for (int m = m0; m < m1; ++m)
      acc += fmaf(__half2float(X[(int64_t)m * ld + n + lo_off]), HL_INV, __half2float(X[(int64_t)m * ld + n]));
  } else {
    const float* X = reinterpret_cast<const float*>(Xv);
#pragma unroll 4
    for (int m = m0; m < m1; ++m) acc += X[(int64_t)m * ld + n];
  }
  atomicAdd(out + n, (double)acc);
}

// ------------------------------------------------------------------------------------------------ GroupNorm backward
// per (sample, channel) and position half z: A1[z] = sum_pos dy, A2[z] = sum_pos dy * x (raw x; the centring happens in
// the group kernel, which adds the two halves).  grid (E/64, B, 2): 3.5 waves of blocks on 148 SMs instead of 1.7.
__global__ void __launch_bounds__(256) gn_bwd_reduce2_kernel(const float* __restrict__ dy, const float* __restrict__ x, int n, int E,
                                                             int B, float* __restrict__ A1, float* __restrict__ A2) {
  __shared__ float2 r1[8][32], r2[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y, c = blockIdx.x * 64 + lane * 2, z = blockIdx.z;
  const int p0 = z * (n / 2), p1 = z ? n : n / 2;
  float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
  if (c < E) {
    const int64_t base = (int64_t)b * n * E + c;
#pragma unroll 4
    for (int pos = p0 + w; pos < p1; pos += 8) {
      const float2 g = __ldcs(reinterpret_cast<const float2*>(dy + base + (int64_t)pos * E));
      const float2 xv = __ldcs(reinterpret_cast<const float2*>(x + base + (int64_t)pos * E));
      s1.x += g.x; s1.y += g.y;
      s2.x = fmaf(g.x, xv.x, s2.x); s2.y = fmaf(g.y, xv.y, s2.y);
    }
  }
  r1[w][lane] = s1; r2[w][lane] = s2;
  __syncthreads();
  if (w == 0 && c < E) {
    float2 t1 = make_float2(0.f, 0.f), t2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { t1.x += r1[k][lane].x; t1.y += r1[k][lane].y; t2.x += r2[k][lane].x; t2.y += r2[k][lane].y; }
    const int64_t o = ((int64_t)z * B + b) * E + c;
    *reinterpret_cast<float2*>(A1 + o) = t1;
    *reinterpret_cast<float2*>(A2 + o) = t2;
  }
}

// blocks [0, gblocks): one warp per (sample, group) -> dx coefficient tables; blocks after: one thread per channel -> dgamma/dbeta
//   dx = coefA[b,c]*dy + coefBC[b,g,0]*x + coefBC[b,g,1]
__global__ void __launch_bounds__(128) gn_bwd_group2_kernel(const float* __restrict__ A1, const float* __restrict__ A2,
                                                            const double* __restrict__ stats, const float* __restrict__ gamma,
                                                            int B, int n, int E, int groups, float eps, int gblocks,
                                                            const float* __restrict__ inv_scale, float* __restrict__ coefA,
                                                            float* __restrict__ coefBC, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta) {
  const int gs = E / groups;
  const double cnt = (double)gs * n;
  const int64_t half = (int64_t)B * E;            // A1 / A2 hold two position halves
  if ((int)blockIdx.x < gblocks) {
    const int wid = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (wid >= B * groups) return;
    const int b = wid / groups, g = wid % groups;
    const double* st = stats + (int64_t)wid * 2;
    const double mean = st[0] / cnt;
    double var = st[1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    double g1 = 0.0, g2 = 0.0;
    for (int c = g * gs + lane; c < (g + 1) * gs; c += 32) {
      const int64_t o = (int64_t)b * E + c;
      const double gm = (double)gamma[c], a1 = (double)A1[o] + (double)A1[half + o], a2 = (double)A2[o] + (double)A2[half + o];
      g1 += gm * a1;
      g2 += gm * (a2 - mean * a1) * rstd;            // sum gamma * dy * xhat
      coefA[o] = (float)(rstd * gm);
    }
    g1 = warp_sum(g1); g2 = warp_sum(g2);
    if (lane == 0) {
      coefBC[2 * wid] = (float)(-rstd * rstd * g2 / cnt);
      coefBC[2 * wid + 1] = (float)((rstd * rstd * g2 * mean - rstd * g1) / cnt);
    }
  } else {
    // dgamma / dbeta: the (sample, group) mean and rstd once per block in shared memory (fp64 sqrt / divide are slow)
    extern __shared__ float ms[];                  // [B*groups][2] = (mean, rstd)
    for (int i = threadIdx.x; i < B * groups; i += blockDim.x) {
      const double mean = stats[2 * (int64_t)i] / cnt;
      double var = stats[2 * (int64_t)i + 1] / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      ms[2 * i] = (float)mean;
      ms[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const int c = ((int)blockIdx.x - gblocks) * blockDim.x + threadIdx.x;
    if (c >= E) return;
    const int g = c / gs;
    double d1 = 0.0, d2 = 0.0;
    for (int b = 0; b < B; ++b) {
      const int64_t o = (int64_t)b * E + c;
      const float a1 = A1[o] + A1[half + o], a2 = A2[o] + A2[half + o];
      const float mean = ms[2 * (b * groups + g)], rstd = ms[2 * (b * groups + g) + 1];
      d1 += (double)a1;
      d2 += (double)(fmaf(-mean, a1, a2) * rstd);
    }
    const double inv = (double)inv_of(inv_scale);
    dbeta[c] = (float)(d1 * inv);
    dgamma[c] = (float)(d2 * inv);
  }
}

constexpr int GNA_ROWS = 16;
// thread: 4 channels x GNA_ROWS consecutive rows of one sample (coefficients fetched once); grid (E/4/128, rows/GNA_ROWS)
template <bool OUT16>
__global__ void __launch_bounds__(128) gn_bwd_apply2_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ coefA, const float* __restrict__ coefBC,
                                                            const float* __restrict__ add, int64_t rows, int n, int E, int gs,
                                                            int groups, float* __restrict__ dx, __half* __restrict__ dx16,
                                                            int64_t ld16, int64_t lo16, double* __restrict__ colsum) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= E) return;
  const int64_t r0 = (int64_t)blockIdx.y * GNA_ROWS;
  const int b = (int)(r0 / n), g = c / gs;
  const float4 ca = __ldg(reinterpret_cast<const float4*>(coefA + (int64_t)b * E + c));
  const float2 bc = __ldg(reinterpret_cast<const float2*>(coefBC + 2 * ((int64_t)b * groups + g)));
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const int64_t rend = r0 + GNA_ROWS < rows ? r0 + GNA_ROWS : rows;
#pragma unroll 4
  for (int64_t r = r0; r < rend; ++r) {
    const int64_t i = r * E + c;
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(dy + i));
    const float4 xv = __ldcs(reinterpret_cast<const float4*>(x + i));
    float4 o;
    o.x = fmaf(ca.x, gv.x, fmaf(bc.x, xv.x, bc.y));
    o.y = fmaf(ca.y, gv.y, fmaf(bc.x, xv.y, bc.y));
    o.z = fmaf(ca.z, gv.z, fmaf(bc.x, xv.z, bc.y));
    o.w = fmaf(ca.w, gv.w, fmaf(bc.x, xv.w, bc.y));
    if (add) {
      const float4 av = __ldcs(reinterpret_cast<const float4*>(add + i));
      o.x += av.x; o.y += av.y; o.z += av.z; o.w += av.w;
    }
    *reinterpret_cast<float4*>(dx + i) = o;
    s0 += o.x; s1 += o.y; s2 += o.z; s3 += o.w;
    if (OUT16) {
      const float v[4] = {o.x, o.y, o.z, o.w};
      alignas(8) __half hi[4];
      alignas(8) __half lo[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) hl_split(v[u], hi[u], lo[u]);
      *reinterpret_cast<uint2*>(dx16 + r * ld16 + c) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(dx16 + r * ld16 + c + lo16) = *reinterpret_cast<const uint2*>(lo);
    }
  }
  if (colsum) {
    atomicAdd(colsum + c, (double)s0); atomicAdd(colsum + c + 1, (double)s1);
    atomicAdd(colsum + c + 2, (double)s2); atomicAdd(colsum + c + 3, (double)s3);
  }
}

// ------------------------------------------------------------------------------------------------ gradient finishing
__global__ void slab_reduce_kernel(const float* __restrict__ slabs, int nslab, int64_t stride, int64_t count4,
                                   const float* __restrict__ inv_scale, float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count4) return;
  float4 a = *reinterpret_cast<const float4*>(slabs + i * 4);
  for (int s = 1; s < nslab; ++s) {
    const float4 b = *reinterpret_cast<const float4*>(slabs + (int64_t)s * stride + i * 4);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  const float inv = inv_of(inv_scale);
  a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
  *reinterpret_cast<float4*>(dst + i * 4) = a;
}
__global__ void slab_reduce1_kernel(const float* __restrict__ slabs, int nslab, int64_t stride, int64_t count,
                                    const float* __restrict__ inv_scale, float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float a = slabs[i];
  for (int s = 1; s < nslab; ++s) a += slabs[(int64_t)s * stride + i];
  dst[i] = a * inv_of(inv_scale);
}
__global__ void finish_double_kernel(const double* __restrict__ src, int64_t count, const float* __restrict__ inv_scale,
                                     float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = (float)(src[i] * (double)inv_of(inv_scale));
}
__global__ void scale_copy_kernel(const float* __restrict__ src, int64_t count, const float* __restrict__ inv_scale,
                                  float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = src[i] * inv_of(inv_scale);
}

// dWc[nb, 2bs(out), 2bs(in)] partials -> dw[2, nb, bs(in), bs(out)], dbc -> db[2, nb, bs]   (transpose of pack_afno_kernel)
// 32 x 32 (out, in) tiles through shared memory: reads run along `in`, writes along `out`.  grid (bs/32, bs/32, nb), block (32, 16)
__global__ void __launch_bounds__(512) unpack_afno_grad2_kernel(const float* __restrict__ dWc, int nslab, int64_t stride, const double* __restrict__ dbc,
                                         int nb, int bs, const float* __restrict__ inv_scale, float* __restrict__ dw,
                                         float* __restrict__ db) {
  __shared__ float re[32][33], im[32][33];
  const int o0 = blockIdx.x * 32, k0 = blockIdx.y * 32, kap = blockIdx.z;
  const int64_t L = 2 * bs, per = (int64_t)nb * bs * bs;
  const float inv = inv_of(inv_scale);
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int o = o0 + i, ki = k0 + threadIdx.x;
    float r = 0.f, m = 0.f;
    if (o < bs && ki < bs) {
      for (int s = 0; s < nslab; ++s) {
        const float* Wk = dWc + (int64_t)s * stride + (int64_t)kap * 4 * bs * bs;
        // Wc[n=o][k=ki] = wr, Wc[o][ki+bs] = -wi, Wc[o+bs][ki] = wi, Wc[o+bs][ki+bs] = wr
        r += Wk[(int64_t)o * L + ki] + Wk[(int64_t)(o + bs) * L + ki + bs];
        m += -Wk[(int64_t)o * L + ki + bs] + Wk[(int64_t)(o + bs) * L + ki];
      }
    }
    re[i][threadIdx.x] = r; im[i][threadIdx.x] = m;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ki = k0 + i, o = o0 + threadIdx.x;
    if (o < bs && ki < bs) {
      const int64_t d = ((int64_t)kap * bs + ki) * bs + o;
      dw[d] = re[threadIdx.x][i] * inv;
      dw[per + d] = im[threadIdx.x][i] * inv;
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    for (int o = threadIdx.y * 32 + threadIdx.x; o < bs; o += 32 * blockDim.y) {
      db[(int64_t)kap * bs + o] = (float)(dbc[(int64_t)kap * 2 * bs + o] * (double)inv);
      db[(int64_t)nb * bs + (int64_t)kap * bs + o] = (float)(dbc[(int64_t)kap * 2 * bs + bs + o] * (double)inv);
    }
  }
}

// dWtT[(uv, o), e] partials -> dwt[e, o, uv]: a 32 x 32 (uv, e) tile transpose per o;  grid (E/32, ceil(PP/32), old)
__global__ void unpack_out_grad_kernel(const float* __restrict__ dWtT, int nslab, int64_t stride, int E, int old, int PP,
                                       const float* __restrict__ inv_scale, float* __restrict__ dwt) {
  __shared__ float tile[32][33];
  const int e0 = blockIdx.x * 32, uv0 = blockIdx.y * 32, o = blockIdx.z;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int uv = uv0 + i, e = e0 + threadIdx.x;
    float a = 0.f;
    if (uv < PP && e < E)
      for (int s = 0; s < nslab; ++s) a += dWtT[(int64_t)s * stride + ((int64_t)uv * old + o) * E + e];
    tile[i][threadIdx.x] = a;
  }
  __syncthreads();
  const float inv = inv_of(inv_scale);
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int e = e0 + i, uv = uv0 + threadIdx.x;
    if (e < E && uv < PP) dwt[((int64_t)e * old + o) * PP + uv] = tile[threadIdx.x][i] * inv;
  }
}
__global__ void out_bias_grad_kernel(const double* __restrict__ dbias_t, int old, int PP, const float* __restrict__ inv_scale,
                                     float* __restrict__ db) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= old) return;
  double a = 0.0;
  for (int uv = 0; uv < PP; ++uv) a += dbias_t[(int64_t)uv * old + o];
  db[o] = (float)(a * (double)inv_of(inv_scale));
}

// ------------------------------------------------------------------------------------------------ output tail backward
// A tile of 128 pixels per pass of a persistent block; the two 32 x 32 products are register-tiled (4 pixels x 8
// channels per thread) against transposed shared-memory copies ([channel][pixel]: one LDS.128 feeds 4 pixels), so the
// inner loops run 32 FMAs per 3 shared loads.  The parameter gradients (outer products summed over pixels) are taken
// from the same shared copies by a second mapping, accumulated in registers across the tiles of the block and written
// as ONE partial result per block (summed in double by tail_finish_kernel).
constexpr int TB_PIX = 128, TB_OLD = 32, TB_NOUT_MAX = 8, TB_LD = 132;
constexpr int TB_ACC = 1024 + 32 + TB_NOUT_MAX * 32 + TB_NOUT_MAX + 32;  // [dW2 | db2 | dW4 | db4 | db0] floats per block partial
struct TailBwdArgs {
  const float* Y1pre; const float* dout; const float* scale; const float* w2; const float* b2; const float* w4;
  __half* g1; float* part;
  int B, h, w, P, nout, act; int64_t ntiles;
};
__global__ void __launch_bounds__(TB_PIX, 3) tail_bwd_kernel(const TailBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* Y1T = sm;                           // [32 k][TB_LD] act(y1pre), transposed
  float* G2T = Y1T + 32 * TB_LD;             // [32 n][TB_LD] gradient w.r.t. pre2, transposed
  float* Y2s = G2T + 32 * TB_LD;             // [128 px][32 n]
  float* G3s = Y2s + TB_PIX * 32;            // [128 px][8]
  float* W2s = G3s + TB_PIX * 8;             // [n][k]
  float* W2Ts = W2s + 32 * 32;               // [k][n]
  float* W4s = W2Ts + 32 * 32;               // [8 j][32 n] (rows >= nout zero)
  float* b2s = W4s + TB_NOUT_MAX * 32;       // [32]
  const int tid = threadIdx.x, tx = tid & 3, ty = tid >> 2;
  const int PP = a.P * a.P, NP = PP * TB_OLD, nout = a.nout, act = a.act;
  for (int i = tid; i < 32 * 32; i += TB_PIX) { const float v = a.w2[i]; W2s[i] = v; W2Ts[(i & 31) * 32 + (i >> 5)] = v; }
  for (int i = tid; i < TB_NOUT_MAX * 32; i += TB_PIX) W4s[i] = i < nout * 32 ? a.w4[i] : 0.f;
  if (tid < 32) b2s[tid] = a.b2[tid];
  const float S = a.scale ? __ldg(a.scale) : 1.f;
  float dW2r[8], db2r = 0.f, dW4r[2] = {0.f, 0.f}, db4r = 0.f, db0r[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) { dW2r[u] = 0.f; db0r[u] = 0.f; }
  const int R = a.h * a.P, Wd = a.w * a.P;
  const int64_t npix = (int64_t)a.B * a.h * a.w * PP;
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int64_t pix0 = tile * TB_PIX;
    {   // ---- phase A: this thread's pixel -> act(y1pre) column of Y1T, scaled dout row of G3s
      const int64_t pix = pix0 + tid;
      const bool live = pix < npix;
      const int64_t tok = pix / PP; const int uv = (int)(pix % PP);
      const float* yp = a.Y1pre + tok * NP + (int64_t)uv * 32;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        float4 v = live ? __ldg(reinterpret_cast<const float4*>(yp + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        Y1T[(k + 0) * TB_LD + tid] = act_apply(v.x, act);
        Y1T[(k + 1) * TB_LD + tid] = act_apply(v.y, act);
        Y1T[(k + 2) * TB_LD + tid] = act_apply(v.z, act);
        Y1T[(k + 3) * TB_LD + tid] = act_apply(v.w, act);
      }
      const int u_ = uv / a.P, v_ = uv % a.P;
      const int q = (int)(tok % a.w); const int64_t r_ = tok / a.w;
      const int p = (int)(r_ % a.h); const int64_t b = r_ / a.h;
      const float* dp = a.dout + (((b * R + p * a.P + u_) * Wd) + q * a.P + v_) * nout;
#pragma unroll
      for (int j = 0; j < TB_NOUT_MAX; ++j) G3s[tid * 8 + j] = (live && j < nout) ? dp[j] * S : 0.f;
    }
    __syncthreads();
    float acc[4][8];
    // ---- phase B: pre2[px, n] = b2[n] + sum_k y1[px, k] W2[n, k]   (4 px x 8 n per thread)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = b2s[8 * tx + j];
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float4 y = *reinterpret_cast<const float4*>(Y1T + k * TB_LD + 4 * ty);
      const float4 w0 = *reinterpret_cast<const float4*>(W2Ts + k * 32 + 8 * tx);
      const float4 w1 = *reinterpret_cast<const float4*>(W2Ts + k * 32 + 8 * tx + 4);
      const float yv[4] = {y.x, y.y, y.z, y.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(yv[i], wv[j], acc[i][j]);
    }
    // y2 = act(pre2) -> Y2s; g2 = (W4^T g3) * act'(pre2) -> G2T
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int px = 4 * ty + i;
      const float4 ga = *reinterpret_cast<const float4*>(G3s + px * 8);
      const float4 gb = *reinterpret_cast<const float4*>(G3s + px * 8 + 4);
      const float g3[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      float y2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pre = acc[i][j];
        y2[j] = act_apply(pre, act);
        float s = 0.f;
#pragma unroll
        for (int jj = 0; jj < TB_NOUT_MAX; ++jj) s = fmaf(W4s[jj * 32 + 8 * tx + j], g3[jj], s);
        acc[i][j] = s * act_grad_fast(pre, act);
      }
      *reinterpret_cast<float4*>(Y2s + px * 32 + 8 * tx) = make_float4(y2[0], y2[1], y2[2], y2[3]);
      *reinterpret_cast<float4*>(Y2s + px * 32 + 8 * tx + 4) = make_float4(y2[4], y2[5], y2[6], y2[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(G2T + (8 * tx + j) * TB_LD + 4 * ty) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
    __syncthreads();
    // ---- phase C: g1[px, k] = (sum_n g2[px, n] W2[n, k]) * act'(y1pre[px, k]) -> split store
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int n = 0; n < 32; ++n) {
      const float4 g = *reinterpret_cast<const float4*>(G2T + n * TB_LD + 4 * ty);
      const float4 w0 = *reinterpret_cast<const float4*>(W2s + n * 32 + 8 * tx);
      const float4 w1 = *reinterpret_cast<const float4*>(W2s + n * 32 + 8 * tx + 4);
      const float gv[4] = {g.x, g.y, g.z, g.w};
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(gv[i], wv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t pix = pix0 + 4 * ty + i;
      if (pix < npix) {
        const int64_t tok = pix / PP; const int uv = (int)(pix % PP);
        const float* yp = a.Y1pre + tok * NP + (int64_t)uv * 32 + 8 * tx;
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(yp)), p1 = __ldg(reinterpret_cast<const float4*>(yp + 4));
        const float pre[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        alignas(16) __half hi[8];
        alignas(16) __half lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gv = acc[i][j] * act_grad_fast(pre[j], act);
          db0r[j] += gv;                         // bias gradient of the ConvTranspose: sum over pixels per channel
          hl_split(gv, hi[j], lo[j]);
        }
        __half* gp = a.g1 + tok * (2 * (int64_t)NP) + (int64_t)uv * 32 + 8 * tx;
        *reinterpret_cast<uint4*>(gp) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(gp + NP) = *reinterpret_cast<const uint4*>(lo);
      }
    }
    // ---- phase D: parameter gradients of this tile.  dW2[n, k] for n = ty, k = tx + 4 kk
#pragma unroll 2
    for (int p4 = 0; p4 < TB_PIX; p4 += 4) {
      const float4 g = *reinterpret_cast<const float4*>(G2T + ty * TB_LD + p4);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const float4 y = *reinterpret_cast<const float4*>(Y1T + (tx + 4 * kk) * TB_LD + p4);
        dW2r[kk] = fmaf(g.x, y.x, fmaf(g.y, y.y, fmaf(g.z, y.z, fmaf(g.w, y.w, dW2r[kk]))));
      }
    }
    if (tid < 32) {
      float s = 0.f;
      for (int p4 = 0; p4 < TB_PIX; p4 += 4) {
        const float4 g = *reinterpret_cast<const float4*>(G2T + tid * TB_LD + p4);
        s += (g.x + g.y) + (g.z + g.w);
      }
      db2r += s;
    } else if (tid < 32 + TB_NOUT_MAX) {
      float s = 0.f;
      for (int pp = 0; pp < TB_PIX; ++pp) s += G3s[pp * 8 + (tid - 32)];
      db4r += s;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int e = tid + half * TB_PIX;           // entry (j, n) of dW4
      if (e < nout * 32) {
        const int j = e >> 5, n = e & 31;
        float s = 0.f;
#pragma unroll 4
        for (int pp = 0; pp < TB_PIX; ++pp) s = fmaf(G3s[pp * 8 + j], Y2s[pp * 32 + n], s);
        dW4r[half] += s;
      }
    }
    __syncthreads();
  }
  float* part = a.part + (int64_t)blockIdx.x * TB_ACC;
  if (tid < 32) G3s[tid] = 0.f;                // (the tile loop ended on a barrier: G3s is free)
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(G3s + 8 * tx + j, db0r[j]);
  __syncthreads();
  if (tid < 32) part[1024 + 32 + TB_NOUT_MAX * 32 + TB_NOUT_MAX + tid] = G3s[tid];
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) part[ty * 32 + tx + 4 * kk] = dW2r[kk];
  if (tid < 32) part[1024 + tid] = db2r;
  else if (tid < 32 + TB_NOUT_MAX) part[1024 + 32 + TB_NOUT_MAX * 32 + (tid - 32)] = db4r;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int e = tid + half * TB_PIX;
    if (e < TB_NOUT_MAX * 32) part[1024 + 32 + e] = e < nout * 32 ? dW4r[half] : 0.f;
  }
}
// sums the per-block partials and writes the four parameter gradients (scaled by inv_scale)
__global__ void tail_finish_kernel(const float* __restrict__ part, int nblk, int nout, const float* __restrict__ inv_scale,
                                   float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ dW4, float* __restrict__ db4,
                                   float* __restrict__ db0) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= TB_ACC) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += (double)part[(int64_t)b * TB_ACC + e];
  const float v = (float)(s * (double)inv_of(inv_scale));
  if (e < 1024) dW2[e] = v;
  else if (e < 1056) db2[e - 1024] = v;
  else if (e < 1056 + TB_NOUT_MAX * 32) { if (e - 1056 < nout * 32) dW4[e - 1056] = v; }
  else if (e < 1056 + TB_NOUT_MAX * 32 + TB_NOUT_MAX) { if (e - 1056 - TB_NOUT_MAX * 32 < nout) db4[e - 1056 - TB_NOUT_MAX * 32] = v; }
  else db0[e - 1056 - TB_NOUT_MAX * 32 - TB_NOUT_MAX] = v;
}

// ------------------------------------------------------------------------------------------------ PatchEmbed conv0 backward
// One patch (b, p, q) per pass of a persistent block; thread k owns column k = (u, v, c) of the im2col matrix: its slice
// of the weight gradient dW0p[:, k] lives in registers across the block's patches, the patch's pixels and the gradient
// rows gz[t, :] are staged in shared memory.  MIDC / TC > 0: compile-time mid / T (DPOT's Co*P+3 = 35, T = 10: no
// predicates, exact register arrays); 0: runtime values up to PB_MID / PB_T.  One partial dW0p per block.
constexpr int PB_MID = 40, PB_T = 10, PB_NT = 256;
struct PatchBwdArgs {
  const float* gz; const float* x; const float* W0p; const float* inv_scale; float* dW0p; float* dx;   // dW0p: [gridDim.x][mid*K0] partials
  int B, X, Y, T, C, P, mid, Kp, K0, h, w;
  int koff, moff, mcnt, dx_accum;   // this launch: im2col columns [koff, koff + 256), conv0 channels [moff, moff + mcnt)
};
template <int MIDC, int TC, bool DX>
__global__ void __launch_bounds__(PB_NT, MIDC ? 2 : 1) patch_bwd_kernel(const PatchBwdArgs a) {
  constexpr int MB = MIDC ? MIDC : PB_MID, TB = TC ? TC : PB_T;
  extern __shared__ __align__(16) float sm[];
  float* gzs = sm;                              // [T][mid]  (<= 400 floats)
  float* xs = sm + PB_T * PB_MID;               // patch tile [u][(v, t, c)] = P * (P*T*C) floats; reused for dx
  const int tid = threadIdx.x, K0 = a.K0, T = TC ? TC : a.T, mid = MIDC ? MIDC : a.mcnt, C = a.C, P = a.P;
  const int midt = a.mid, moff = a.moff;        // gz rows hold all midt channels; this launch handles [moff, moff + mid)
  const int run = P * T * C;                    // contiguous floats of one patch row u
  const int k = a.koff + tid;
  const bool kon = k < K0;
  const int c = k % C, uv = k / C, v = uv % P, u = uv / P;
  float wcol[DX ? MB : 1], dW[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    if (DX) wcol[m] = (kon && m < mid) ? a.W0p[(int64_t)(moff + m) * K0 + k] : 0.f;
    dW[m] = 0.f;
  }
  const float inv = inv_of(a.inv_scale);
  const int64_t npatch = (int64_t)a.B * a.h * a.w;
  for (int64_t pt = blockIdx.x; pt < npatch; pt += gridDim.x) {
    const int q = (int)(pt % a.w); const int64_t r_ = pt / a.w;
    const int p = (int)(r_ % a.h); const int b = (int)(r_ / a.h);
    for (int i = tid; i < T * mid; i += PB_NT) gzs[i] = a.gz[pt * a.Kp + (i / mid) * midt + moff + (i % mid)];
    for (int i = tid; i < P * run; i += PB_NT) {
      const int uu = i / run, rr = i % run;
      xs[i] = a.x[(((int64_t)b * a.X + p * P + uu) * a.Y + (int64_t)q * P) * T * C + rr];
    }
    __syncthreads();
    float xk[TB], dxk[DX ? TB : 1];
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      xk[t] = (kon && t < T) ? xs[u * run + (v * T + t) * C + c] : 0.f;
      if (DX) dxk[t] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < TB; ++t) {
      if (TC || t < T) {
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          if (MIDC || m < mid) {
            const float gv = gzs[t * mid + m];
            dW[m] = fmaf(gv, xk[t], dW[m]);
            if (DX) dxk[t] = fmaf(gv, wcol[m], dxk[t]);
          }
        }
      }
    }
    if (DX) {
      if (K0 <= PB_NT && !a.dx_accum) {
        // the launch covers every column of the patch: stage dx in the (dead) pixel tile, store whole image-row runs
        __syncthreads();
        if (kon) {
#pragma unroll
          for (int t = 0; t < TB; ++t)
            if (TC || t < T) xs[u * run + (v * T + t) * C + c] = dxk[t] * inv;
        }
        __syncthreads();
        for (int i = tid; i < P * run; i += PB_NT) {
          const int uu = i / run, rr = i % run;
          a.dx[(((int64_t)b * a.X + p * P + uu) * a.Y + (int64_t)q * P) * T * C + rr] = xs[i];
        }
      } else if (kon) {
        // column / channel chunks (patch 16): each launch owns its columns; a later channel chunk adds to the first one's
#pragma unroll
        for (int t = 0; t < TB; ++t)
          if (TC || t < T) {
            float* dp = a.dx + (((int64_t)b * a.X + p * P + u) * a.Y + (int64_t)q * P + v) * T * C + t * C + c;
            *dp = a.dx_accum ? *dp + dxk[t] * inv : dxk[t] * inv;
          }
      }
    }
    __syncthreads();
  }
  if (kon) {
    float* slab = a.dW0p + (int64_t)blockIdx.x * midt * K0;
#pragma unroll
    for (int m = 0; m < MB; ++m)
      if (MIDC || m < mid) slab[(int64_t)(moff + m) * K0 + k] = dW[m];
  }
}

// dpe0_w[m, c < C, u, v] = inv * sum over the per-block partials dW0p[s][m, (u,v,c)]
__global__ void __launch_bounds__(256) patch_w_reduce_kernel(const float* __restrict__ dW0p, int nslab, int mid, int C, int PP,
                                                             const float* __restrict__ inv_scale, float* __restrict__ dw0) {
  const int K0 = PP * C;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= mid * K0) return;
  const int m = idx / K0, i = idx % K0;
  double acc = 0.0;
  for (int sl = 0; sl < nslab; ++sl) acc += (double)dW0p[(int64_t)sl * mid * K0 + idx];
  const int c = i % C, uv = i / C;
  dw0[((int64_t)m * (C + 3) + c) * PP + uv] = (float)(acc * (double)inv_of(inv_scale));
}
// coordinate channels and bias of conv0 from drb[(p,q), t*mid + m] (row pitch Kp, summed over the batch):
// grid (mid, 2P + 2): which < P: x-channel row u = which; < 2P: y-channel column v; == 2P: t-channel; 2P + 1: bias
__global__ void __launch_bounds__(256) patch_coord_grad_kernel(const double* __restrict__ drb, const float* __restrict__ gx,
                                                               const float* __restrict__ gy, const float* __restrict__ gt, int mid,
                                                               int C, int P, int h, int w, int T, int Kp,
                                                               const float* __restrict__ inv_scale, float* __restrict__ dw0,
                                                               float* __restrict__ db0) {
  __shared__ double red[256];
  const int m = blockIdx.x, which = blockIdx.y, tid = threadIdx.x;
  const int PP = P * P, nitem = h * w * T;
  double acc = 0.0;
  for (int i = tid; i < nitem; i += blockDim.x) {
    const int t = i % T; const int pq = i / T; const int q = pq % w, p = pq / w;
    const double g = drb[(int64_t)pq * Kp + t * mid + m];
    double f;
    if (which < P) f = (double)gx[p * P + which];
    else if (which < 2 * P) f = (double)gy[q * P + (which - P)];
    else if (which == 2 * P) f = (double)gt[t];
    else f = 1.0;
    acc += g * f;
  }
  red[tid] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  const float val = (float)(red[0] * (double)inv_of(inv_scale));
  float* wx = dw0 + ((int64_t)m * (C + 3) + C) * PP;
  if (which < P) { if (tid < P) wx[which * P + tid] = val; }                       // x channel: row u, every v
  else if (which < 2 * P) { if (tid < P) wx[PP + tid * P + (which - P)] = val; }   // y channel: column v, every u
  else if (which == 2 * P) { if (tid < PP) wx[2 * PP + tid] = val; }               // t channel: every (u, v)
  else if (tid == 0) db0[m] = val;
}

// ------------------------------------------------------------------------------------------------ time aggregation fold
// wts16[(t,i), j] = split(w[t,i,j] * temb[t,i])   (rows of [hi E | lo E] halves)
__global__ void __launch_bounds__(256) tagg_scale16_kernel(const float* __restrict__ w, const float* __restrict__ temb, int64_t rows,
                                                           int E8, __half* __restrict__ dst) {
  const int64_t total = rows * E8;
  const int E = E8 * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / E8;
    const int c = (int)(i % E8) * 8;
    const float te = __ldg(temb + r);
    const float4 a = __ldcs(reinterpret_cast<const float4*>(w + r * E + c));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(w + r * E + c + 4));
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    alignas(16) __half hi[8];
    alignas(16) __half lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) hl_split(v[u] * te, hi[u], lo[u]);
    *reinterpret_cast<uint4*>(dst + r * 2 * E + c) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + r * 2 * E + c + E) = *reinterpret_cast<const uint4*>(lo);
  }
}
// Wsum16[i, j] = split(sum_t temb[t,i] * w[t,i,j])
__global__ void tagg_wsum16_kernel(const float* __restrict__ w, const float* __restrict__ temb, int T, int E, __half* __restrict__ dst) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)E * E) return;
  const int i = (int)(idx / E), j = (int)(idx % E);
  double acc = 0.0;
  for (int t = 0; t < T; ++t) acc += (double)(temb[(int64_t)t * E + i] * w[(int64_t)t * E * E + idx]);
  __half hi, lo;
  hl_split((float)acc, hi, lo);
  dst[(int64_t)i * 2 * E + j] = hi;
  dst[(int64_t)i * 2 * E + E + j] = lo;
}
// bp16[i, p] = split(b2[i] + pos[i, p])
__global__ void tagg_bp16_kernel(const float* __restrict__ b2, const float* __restrict__ pos, int E, int n, __half* __restrict__ dst) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)E * n) return;
  const int i = (int)(idx / n), p = (int)(idx % n);
  __half hi, lo;
  hl_split(b2[i] + pos[idx], hi, lo);
  dst[(int64_t)i * 2 * n + p] = hi;
  dst[(int64_t)i * 2 * n + n + p] = lo;
}
// dst16[r, c] = split(c < cols ? src[r*lds + c] : 0), c < colsp   (zero-padded split copy; rows of [hi colsp | lo colsp])
__global__ void pad_split_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int cols, int colsp, __half* __restrict__ dst) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * colsp) return;
  const int64_t r = idx / colsp; const int c = (int)(idx % colsp);
  __half hi, lo;
  hl_split(c < cols ? src[r * lds + c] : 0.f, hi, lo);
  dst[r * 2 * colsp + c] = hi;
  dst[r * 2 * colsp + colsp + c] = lo;
}
// Gp16[t][j][m] = split(m < mid ? dWeffT[j, t*mid + m] : 0), m < midp
__global__ void tagg_pad_g_kernel(const float* __restrict__ dWeffT, int E, int Kp, int T, int mid, int midp, __half* __restrict__ dst) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)T * E * midp) return;
  const int m = (int)(idx % midp); const int64_t r = idx / midp;     // r = t*E + j
  const int j = (int)(r % E), t = (int)(r / E);
  __half hi, lo;
  hl_split(m < mid ? dWeffT[(int64_t)j * Kp + t * mid + m] : 0.f, hi, lo);
  dst[r * 2 * midp + m] = hi;
  dst[r * 2 * midp + midp + m] = lo;
}
// dW2[i, m] = inv * sum_t slabs[t][i][m], m < mid (slab rows of midp)
__global__ void tagg_dw2_finish_kernel(const float* __restrict__ slabs, int T, int E, int mid, int midp, const float* __restrict__ inv_scale,
                                       float* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= E * mid) return;
  const int i = idx / mid, m = idx % mid;
  double acc = 0.0;
  for (int t = 0; t < T; ++t) acc += (double)slabs[((int64_t)t * E + i) * midp + m];
  dst[idx] = (float)(acc * (double)inv_of(inv_scale));
}
// one warp per (t, i) row
__global__ void __launch_bounds__(256) tagg_finish_kernel(const float* __restrict__ dwt, const float* __restrict__ w,
                                                          const float* __restrict__ temb, int64_t rows, int E,
                                                          const float* __restrict__ inv_scale, float* __restrict__ dw,
                                                          float* __restrict__ dtemb) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float te = temb[row] * inv_of(inv_scale);
  double acc = 0.0;
  for (int j = lane; j < E; j += 32) {
    const float g = dwt[row * E + j];
    acc += (double)g * (double)w[row * E + j];
    dw[row * E + j] = g * te;
  }
  acc = warp_sum(acc);
  if (lane == 0) dtemb[row] = (float)acc;
}
__global__ void tagg_gamma_grad_kernel(const float* __restrict__ dtemb, const float* __restrict__ gamma, int T, int E,
                                       const float* __restrict__ inv_scale, float* __restrict__ dgamma) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E) return;
  double acc = 0.0;
  for (int t = 0; t < T; ++t) {
    const double tt = T > 1 ? (double)t / (double)(T - 1) : 0.0;     // torch.linspace(0, 1, T)
    const float arg = (float)tt * gamma[i];                          // the forward's fp32 product tt * gamma
    acc += (double)dtemb[(int64_t)t * E + i] * (-sin((double)arg)) * tt;
  }
  dgamma[i] = (float)(acc * (double)inv_of(inv_scale));
}
__global__ void __launch_bounds__(256) rowsum_scale_kernel(const float* __restrict__ X, int rows, int cols,
                                                           const float* __restrict__ inv_scale, float* __restrict__ rowsum,
                                                           float* __restrict__ scaled) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float inv = inv_of(inv_scale);
  double acc = 0.0;
  for (int j = lane; j < cols; j += 32) {
    const float v = X[(int64_t)row * cols + j];
    acc += (double)v;
    scaled[(int64_t)row * cols + j] = v * inv;
  }
  acc = warp_sum(acc);
  if (lane == 0) rowsum[row] = (float)(acc * (double)inv);
}
__global__ void transpose2_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int R, int Cc) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? src[(int64_t)r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < R) dst[(int64_t)c * ldd + r] = tile[threadIdx.x][i];
  }
}
__global__ void add_mean_grad_kernel(const float* __restrict__ dtok, int64_t total4, int n, int E, float* __restrict__ g) {
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const int64_t i = i4 * 4;
  const int64_t row = i / E; const int c = (int)(i % E);
  const int b = (int)(row / n);
  const float s = 1.0f / (float)n;
  const float4 d = *reinterpret_cast<const float4*>(dtok + (int64_t)b * E + c);
  float4 v = *reinterpret_cast<float4*>(g + i);
  v.x = fmaf(d.x, s, v.x); v.y = fmaf(d.y, s, v.y); v.z = fmaf(d.z, s, v.z); v.w = fmaf(d.w, s, v.w);
  *reinterpret_cast<float4*>(g + i) = v;
}
__global__ void __launch_bounds__(256) split_scaled_kernel(const float* __restrict__ src, int64_t rows, int cols8,
                                                           const float* __restrict__ factor, __half* __restrict__ dst,
                                                           int64_t ldd, int64_t lo_off) {
  const int64_t total = rows * cols8;
  const float f = factor ? __ldg(factor) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols8;
    const int c = (int)(i % cols8) * 8;
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src + (r * cols8) * 8 + c));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(src + (r * cols8) * 8 + c + 4));
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    alignas(16) __half hi[8];
    alignas(16) __half lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) hl_split(v[u] * f, hi[u], lo[u]);
    *reinterpret_cast<uint4*>(dst + r * ldd + c) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + r * ldd + c + lo_off) = *reinterpret_cast<const uint4*>(lo);
  }
}

// AFNO2D w[2,nb,bs,bs] ("bio": in,out), b[2,nb,bs] -> the real block form of pack_afno_kernel (norm_misc.cu) stored
// directly as split fp16: Wc16[kap][n][hi 2bs | lo 2bs], bc[nb, 2bs] fp32
__global__ void pack_afno16_kernel(const float* __restrict__ w, const float* __restrict__ b, int nb, int bs,
                                   __half* __restrict__ Wc16, float* __restrict__ bc) {
  const int64_t total = (int64_t)nb * 4 * bs * bs;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    const int k = (int)(i % (2 * bs));
    const int64_t row = i / (2 * bs);                       // kap * 2bs + nn
    const int nn = (int)(row % (2 * bs));
    const int kap = (int)(row / (2 * bs));
    const int ki = k % bs, no = nn % bs;
    const float wr = w[(((int64_t)0 * nb + kap) * bs + ki) * bs + no];
    const float wi = w[(((int64_t)1 * nb + kap) * bs + ki) * bs + no];
    float v;
    if (nn < bs) v = (k < bs) ? wr : -wi;
    else         v = (k < bs) ? wi : wr;
    __half hi, lo;
    hl_split(v, hi, lo);
    Wc16[row * 4 * bs + k] = hi;
    Wc16[row * 4 * bs + 2 * bs + k] = lo;
  }
  if (i < (int64_t)nb * 2 * bs) {
    const int nn = (int)(i % (2 * bs));
    const int kap = (int)(i / (2 * bs));
    bc[i] = b[(((int64_t)(nn / bs)) * nb + kap) * bs + nn % bs];
  }
}
// pack_patch (norm_misc.cu) with the coordinate sums separated: one block per conv0 output channel m
//   rowbias0[(p,q,t), m] = b0[m] + sum_u gx[pP+u] sx[u] + sum_v gy[qP+v] sy[v] + gt[t] st,  sx[u] = sum_v wx[u,v], ...
__global__ void __launch_bounds__(256) pack_patch_fast_kernel(const float* __restrict__ w0, const float* __restrict__ b0,
                                                              const float* __restrict__ gx, const float* __restrict__ gy,
                                                              const float* __restrict__ gt, int mid, int C, int P, int h, int w,
                                                              int T, float* __restrict__ W0p, float* __restrict__ rowbias0) {
  __shared__ double sx[32], sy[32], st_;
  __shared__ double ax[64], ay[64];                      // per latent row p / column q sums (h, w <= 64)
  const int m = blockIdx.x, tid = threadIdx.x, PP = P * P, K = PP * C;
  for (int k = tid; k < K; k += blockDim.x) {
    const int c = k % C, uv = k / C;
    W0p[(int64_t)m * K + k] = w0[((int64_t)m * (C + 3) + c) * PP + uv];
  }
  const float* wx = w0 + ((int64_t)m * (C + 3) + C) * PP;
  const float* wy = wx + PP;
  const float* wt = wy + PP;
  if (tid < P) {
    double a = 0.0, bsum = 0.0;
    for (int v = 0; v < P; ++v) { a += (double)wx[tid * P + v]; bsum += (double)wy[v * P + tid]; }
    sx[tid] = a; sy[tid] = bsum;
  }
  if (tid == 0) {
    double a = 0.0;
    for (int i = 0; i < PP; ++i) a += (double)wt[i];
    st_ = a;
  }
  __syncthreads();
  if (tid < h) { double a = 0.0; for (int u = 0; u < P; ++u) a += (double)gx[tid * P + u] * sx[u]; ax[tid] = a; }
  if (tid >= 64 && tid - 64 < w) { const int q = tid - 64; double a = 0.0; for (int v = 0; v < P; ++v) a += (double)gy[q * P + v] * sy[v]; ay[q] = a; }
  __syncthreads();
  const double bias = (double)b0[m];
  for (int i = tid; i < h * w * T; i += blockDim.x) {
    const int t = i % T; const int pq = i / T; const int q = pq % w, p = pq / w;
    // same summation order per term as the reference fold is not required: every term is accumulated in double
    rowbias0[(int64_t)i * mid + m] = (float)(bias + ax[p] + ay[q] + (double)gt[t] * st_);
  }
}

// generic output-head tail (out_layer_dim != 32): dL/d(out field) -> per-pixel rows with the nout channels zero-padded to
// NPAD = 8 (the contraction engine's minimum K), scaled by S and stored split: g3[tok, hi (uv, 8) | lo (uv, 8)]
__global__ void unshuffle_pad_split_kernel(const float* __restrict__ dout, const float* __restrict__ scale, int B, int h, int w, int P,
                                           int nout, __half* __restrict__ g3) {
  const int PP = P * P;
  const int64_t total = (int64_t)B * h * w * PP * 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = (int)(i & 7);
  const int64_t pix = i >> 3;
  const int64_t tok = pix / PP; const int uv = (int)(pix % PP);
  const int u_ = uv / P, v_ = uv % P;
  const int q = (int)(tok % w); const int64_t r_ = tok / w;
  const int p = (int)(r_ % h); const int64_t b = r_ / h;
  float v = 0.f;
  if (j < nout) v = dout[(((b * h * P + p * P + u_) * (int64_t)(w * P)) + q * P + v_) * nout + j] * (scale ? __ldg(scale) : 1.f);
  __half hi, lo;
  hl_split(v, hi, lo);
  g3[tok * (2 * (int64_t)PP * 8) + uv * 8 + j] = hi;
  g3[tok * (2 * (int64_t)PP * 8) + (int64_t)PP * 8 + uv * 8 + j] = lo;
}
// dst[i] = inv * sum_b src[b * n + i], i < keep  (column sums taken per (u, v) problem of a batched contraction)
__global__ void sum_batches_kernel(const double* __restrict__ src, int nbat, int n, int keep, const float* __restrict__ inv_scale,
                                   float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= keep) return;
  double a = 0.0;
  for (int b = 0; b < nbat; ++b) a += src[(int64_t)b * n + i];
  dst[i] = (float)(a * (double)inv_of(inv_scale));
}

inline unsigned blocks_for(int64_t total, int per) { return (unsigned)ceil_div(total, per); }

}  // namespace

// ================================================================================================ host launchers
int tk_grad_scale(const float* dy, int64_t n, unsigned* bits, float* scale, cudaStream_t st) {
  DPOT_CUDA(cudaMemsetAsync(bits, 0, sizeof(unsigned), st));
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
  absmax_kernel<<<grid, 256, 0, st>>>(dy, n, bits);
  DPOT_LAUNCH_CHECK("absmax_kernel");
  scale_from_max_kernel<<<1, 1, 0, st>>>(bits, scale);
  DPOT_LAUNCH_CHECK("scale_from_max_kernel");
  return 0;
}

int tk_colsum(const void* X, bool f16, int64_t ld, int64_t lo_off, int M, int N, double* out, cudaStream_t st) {
  if (M <= 0) return 0;
  const int rpb = M >= 4096 ? 64 : (M >= 256 ? 32 : M);
  dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div(M, rpb));
  if (f16) colsum2_kernel<true><<<grid, 128, 0, st>>>(X, ld, lo_off, M, N, rpb, out);
  else colsum2_kernel<false><<<grid, 128, 0, st>>>(X, ld, lo_off, M, N, rpb, out);
  DPOT_LAUNCH_CHECK("colsum2_kernel");
  return 0;
}

int tk_gn_bwd(const float* dy, const float* x, const double* stats, const float* gamma, const float* add, int B, int n, int E,
              int groups, float eps, const float* inv_scale, float* scratch, float* dx, __half* dx16, float* dgamma,
              float* dbeta, double* colsum, cudaStream_t st) {
  DPOT_REQUIRE(E % groups == 0 && (E / groups) % 4 == 0 && E % 2 == 0, DPOT_E_BADARG, "gn_bwd: group size must be a multiple of 4");
  float* A1 = scratch; float* A2 = A1 + 2 * (int64_t)B * E; float* coefA = A2 + 2 * (int64_t)B * E; float* coefBC = coefA + (int64_t)B * E;
  DPOT_REQUIRE(n % 2 == 0 && B * groups * 2 * sizeof(float) <= 40 * 1024, DPOT_E_BADARG, "gn_bwd: odd n or too many (sample, group) pairs");
  gn_bwd_reduce2_kernel<<<dim3((unsigned)ceil_div(E, 64), (unsigned)B, 2), 256, 0, st>>>(dy, x, n, E, B, A1, A2);
  DPOT_LAUNCH_CHECK("gn_bwd_reduce2_kernel");
  const int gblocks = (int)ceil_div(B * groups, 4);
  gn_bwd_group2_kernel<<<(unsigned)(gblocks + ceil_div(E, 128)), 128, sizeof(float) * 2 * B * groups, st>>>(
      A1, A2, stats, gamma, B, n, E, groups, eps, gblocks, inv_scale, coefA, coefBC, dgamma, dbeta);
  DPOT_LAUNCH_CHECK("gn_bwd_group2_kernel");
  DPOT_REQUIRE(n % GNA_ROWS == 0, DPOT_E_BADARG, "gn_bwd: latent cells per sample must be a multiple of %d", GNA_ROWS);
  const int64_t rows = (int64_t)B * n;
  dim3 grid((unsigned)ceil_div(E / 4, 128), (unsigned)(rows / GNA_ROWS));
  if (dx16)
    gn_bwd_apply2_kernel<true><<<grid, 128, 0, st>>>(dy, x, coefA, coefBC, add, rows, n, E, E / groups, groups, dx, dx16,
                                                    2 * (int64_t)E, E, colsum);
  else
    gn_bwd_apply2_kernel<false><<<grid, 128, 0, st>>>(dy, x, coefA, coefBC, add, rows, n, E, E / groups, groups, dx, nullptr, 0, 0,
                                                     colsum);
  DPOT_LAUNCH_CHECK("gn_bwd_apply2_kernel");
  return 0;
}

int tk_slab_reduce(const float* slabs, int nslab, int64_t stride, int64_t count, const float* inv_scale, float* dst, cudaStream_t st) {
  if (count % 4 == 0 && stride % 4 == 0 && (reinterpret_cast<uintptr_t>(slabs) | reinterpret_cast<uintptr_t>(dst)) % 16 == 0)
    slab_reduce_kernel<<<blocks_for(count / 4, 256), 256, 0, st>>>(slabs, nslab, stride, count / 4, inv_scale, dst);
  else
    slab_reduce1_kernel<<<blocks_for(count, 256), 256, 0, st>>>(slabs, nslab, stride, count, inv_scale, dst);
  DPOT_LAUNCH_CHECK("slab_reduce_kernel");
  return 0;
}

int tk_finish_double(const double* src, int64_t count, const float* inv_scale, float* dst, cudaStream_t st) {
  finish_double_kernel<<<blocks_for(count, 256), 256, 0, st>>>(src, count, inv_scale, dst);
  DPOT_LAUNCH_CHECK("finish_double_kernel");
  return 0;
}

int tk_scale_copy(const float* src, int64_t count, const float* inv_scale, float* dst, cudaStream_t st) {
  scale_copy_kernel<<<blocks_for(count, 256), 256, 0, st>>>(src, count, inv_scale, dst);
  DPOT_LAUNCH_CHECK("scale_copy_kernel");
  return 0;
}

int tk_unpack_afno_grad(const float* dWc, int nslab, int64_t stride, const double* dbc, int nb, int bs, const float* inv_scale,
                        float* dw, float* db, cudaStream_t st) {
  unpack_afno_grad2_kernel<<<dim3((unsigned)ceil_div(bs, 32), (unsigned)ceil_div(bs, 32), (unsigned)nb), dim3(32, 16), 0, st>>>(
      dWc, nslab, stride, dbc, nb, bs, inv_scale, dw, db);
  DPOT_LAUNCH_CHECK("unpack_afno_grad2_kernel");
  return 0;
}

int tk_unpack_out_grad(const float* dWtT, int nslab, int64_t stride, const double* dbias_t, int E, int old, int P,
                       const float* inv_scale, float* dwt, float* db, cudaStream_t st) {
  const int PP = P * P;
  DPOT_REQUIRE(old <= 65535, DPOT_E_BADARG, "unpack_out_grad: out_layer_dim too large");
  unpack_out_grad_kernel<<<dim3((unsigned)ceil_div(E, 32), (unsigned)ceil_div(PP, 32), (unsigned)old), dim3(32, 8), 0, st>>>(
      dWtT, nslab, stride, E, old, PP, inv_scale, dwt);
  DPOT_LAUNCH_CHECK("unpack_out_grad_kernel");
  if (dbias_t) {
    out_bias_grad_kernel<<<(unsigned)ceil_div(old, 64), 64, 0, st>>>(dbias_t, old, PP, inv_scale, db);
    DPOT_LAUNCH_CHECK("out_bias_grad_kernel");
  }
  return 0;
}

int tk_tail_bwd_supported(int old, int nout) { return old == TB_OLD && nout >= 1 && nout <= TB_NOUT_MAX; }
int tk_tail_bwd_blocks(int B, int h, int w, int P) {
  const int64_t ntiles = ceil_div((int64_t)B * h * w * P * P, TB_PIX);
  return (int)std::min<int64_t>(ntiles, (int64_t)sm_count_cur() * 3);
}
int64_t tk_tail_bwd_part_floats(int nblk) { return (int64_t)nblk * TB_ACC; }
int tk_tail_bwd(const float* Y1pre, const float* dout, const float* scale, const float* w2, const float* b2, const float* w4,
                int B, int h, int w, int P, int nout, int act, __half* g1, float* part, const float* inv_scale, float* dW2,
                float* db2, float* dW4, float* db4, float* db0, cudaStream_t st) {
  DPOT_REQUIRE(tk_tail_bwd_supported(TB_OLD, nout), DPOT_E_UNSUPPORTED, "tail_bwd: nout %d unsupported", nout);
  const int64_t npix = (int64_t)B * h * w * P * P;
  TailBwdArgs a;
  a.Y1pre = Y1pre; a.dout = dout; a.scale = scale; a.w2 = w2; a.b2 = b2; a.w4 = w4; a.g1 = g1; a.part = part;
  a.B = B; a.h = h; a.w = w; a.P = P; a.nout = nout; a.act = act; a.ntiles = ceil_div(npix, TB_PIX);
  const size_t smem = sizeof(float) * (2 * 32 * TB_LD + TB_PIX * 32 + TB_PIX * 8 + 2 * 32 * 32 + TB_NOUT_MAX * 32 + 32);
  static DevOnce attr;
  if (attr.need()) {
    DPOT_CUDA(cudaFuncSetAttribute(tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr.done();
  }
  const int nblk = tk_tail_bwd_blocks(B, h, w, P);
  tail_bwd_kernel<<<(unsigned)nblk, TB_PIX, smem, st>>>(a);
  DPOT_LAUNCH_CHECK("tail_bwd_kernel");
  tail_finish_kernel<<<(unsigned)ceil_div(TB_ACC, 128), 128, 0, st>>>(part, nblk, nout, inv_scale, dW2, db2, dW4, db4, db0);
  DPOT_LAUNCH_CHECK("tail_finish_kernel");
  return 0;
}

int tk_patch_bwd_supported(int mid, int T, int K0) { return mid <= 4 * PB_MID && T <= PB_T && K0 <= 8 * PB_NT; }
int tk_patch_bwd_slabs(int B, int X, int Y, int P) {
  const int64_t npatch = (int64_t)B * (X / P) * (Y / P);
  return (int)std::min<int64_t>(npatch, (int64_t)sm_count_cur() * 2);
}
int tk_patch_bwd(const float* gz, const float* x, const float* W0p, int B, int X, int Y, int T, int C, int P, int mid, int Kp,
                 const float* inv_scale, float* dW0p, float* dx, cudaStream_t st) {
  DPOT_REQUIRE(tk_patch_bwd_supported(mid, T, P * P * C), DPOT_E_UNSUPPORTED, "patch_bwd: geometry unsupported");
  PatchBwdArgs a;
  a.gz = gz; a.x = x; a.W0p = W0p; a.inv_scale = inv_scale; a.dW0p = dW0p; a.dx = dx;
  a.B = B; a.X = X; a.Y = Y; a.T = T; a.C = C; a.P = P; a.mid = mid; a.Kp = Kp; a.K0 = P * P * C; a.h = X / P; a.w = Y / P;
  const size_t smem = sizeof(float) * (PB_T * PB_MID + (size_t)P * P * T * C);
  DPOT_REQUIRE(smem <= 48 * 1024, DPOT_E_UNSUPPORTED, "patch_bwd: patch tile too large");
  const unsigned grid = (unsigned)tk_patch_bwd_slabs(B, X, Y, P);
  if (mid == 35 && T == 10 && a.K0 <= PB_NT) {          // DPOT-Ti/S/M/H at patch 8: one launch, compile-time loops
    a.koff = 0; a.moff = 0; a.mcnt = mid; a.dx_accum = 0;
    if (dx) patch_bwd_kernel<35, 10, true><<<grid, PB_NT, smem, st>>>(a);
    else patch_bwd_kernel<35, 10, false><<<grid, PB_NT, smem, st>>>(a);
    DPOT_LAUNCH_CHECK("patch_bwd_kernel");
    return 0;
  }
  // other geometries (DPOT-L: patch 16 -> 1024 columns, 67 channels): chunks of 256 columns x <= 40 channels per launch
  for (int koff = 0; koff < a.K0; koff += PB_NT)
    for (int moff = 0; moff < mid; moff += PB_MID) {
      a.koff = koff; a.moff = moff; a.mcnt = std::min(PB_MID, mid - moff); a.dx_accum = moff > 0 ? 1 : 0;
      if (dx) patch_bwd_kernel<0, 0, true><<<grid, PB_NT, smem, st>>>(a);
      else patch_bwd_kernel<0, 0, false><<<grid, PB_NT, smem, st>>>(a);
      DPOT_LAUNCH_CHECK("patch_bwd_kernel");
    }
  return 0;
}

int tk_unpack_patch_grad(const float* dW0p, int nslab, const double* drb, const float* gx, const float* gy, const float* gt, int mid, int C,
                         int P, int h, int w, int T, int Kp, const float* inv_scale, float* dw0, float* db0, cudaStream_t st) {
  DPOT_REQUIRE(P <= 16, DPOT_E_UNSUPPORTED, "unpack_patch_grad: patch_size > 16");
  patch_w_reduce_kernel<<<blocks_for((int64_t)mid * P * P * C, 256), 256, 0, st>>>(dW0p, nslab, mid, C, P * P, inv_scale, dw0);
  DPOT_LAUNCH_CHECK("patch_w_reduce_kernel");
  patch_coord_grad_kernel<<<dim3((unsigned)mid, (unsigned)(2 * P + 2)), 256, 0, st>>>(drb, gx, gy, gt, mid, C, P, h, w, T, Kp, inv_scale,
                                                                                    dw0, db0);
  DPOT_LAUNCH_CHECK("patch_coord_grad_kernel");
  return 0;
}

int tk_unshuffle_pad_split(const float* dout, const float* scale, int B, int h, int w, int P, int nout, __half* g3, cudaStream_t st) {
  DPOT_REQUIRE(nout >= 1 && nout <= 8, DPOT_E_UNSUPPORTED, "unshuffle_pad_split: nout %d > 8", nout);
  const int64_t total = (int64_t)B * h * w * P * P * 8;
  unshuffle_pad_split_kernel<<<blocks_for(total, 256), 256, 0, st>>>(dout, scale, B, h, w, P, nout, g3);
  DPOT_LAUNCH_CHECK("unshuffle_pad_split_kernel");
  return 0;
}
int tk_sum_batches(const double* src, int nbat, int n, int keep, const float* inv_scale, float* dst, cudaStream_t st) {
  sum_batches_kernel<<<(unsigned)ceil_div(keep, 128), 128, 0, st>>>(src, nbat, n, keep, inv_scale, dst);
  DPOT_LAUNCH_CHECK("sum_batches_kernel");
  return 0;
}
int tk_pack_afno16(const float* w, const float* b, int nb, int bs, __half* Wc16, float* bc, cudaStream_t st) {
  const int64_t total = (int64_t)nb * 4 * bs * bs;
  pack_afno16_kernel<<<blocks_for(total, 256), 256, 0, st>>>(w, b, nb, bs, Wc16, bc);
  DPOT_LAUNCH_CHECK("pack_afno16_kernel");
  return 0;
}
int tk_pack_patch(const float* w0, const float* b0, const float* gx, const float* gy, const float* gt, int mid, int C, int P, int h,
                  int w, int T, float* W0p, float* rowbias0, cudaStream_t st) {
  if (P > 32 || h > 64 || w > 64) return dpot_pack_patch(w0, b0, gx, gy, gt, mid, C, P, h, w, T, W0p, rowbias0, st);
  pack_patch_fast_kernel<<<(unsigned)mid, 256, 0, st>>>(w0, b0, gx, gy, gt, mid, C, P, h, w, T, W0p, rowbias0);
  DPOT_LAUNCH_CHECK("pack_patch_fast_kernel");
  return 0;
}
int tk_tagg_scale16(const float* w, const float* temb, int T, int E, __half* wts16, __half* Wsum16, cudaStream_t st) {
  DPOT_REQUIRE(E % 8 == 0, DPOT_E_BADARG, "tagg_scale16: E %% 8");
  const int64_t rows = (int64_t)T * E, total = rows * (E / 8);
  tagg_scale16_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 16), 256, 0, st>>>(w, temb, rows, E / 8, wts16);
  DPOT_LAUNCH_CHECK("tagg_scale16_kernel");
  tagg_wsum16_kernel<<<blocks_for((int64_t)E * E, 256), 256, 0, st>>>(w, temb, T, E, Wsum16);
  DPOT_LAUNCH_CHECK("tagg_wsum16_kernel");
  return 0;
}
int tk_tagg_bp16(const float* b2, const float* pos, int E, int n, __half* bp16, cudaStream_t st) {
  tagg_bp16_kernel<<<blocks_for((int64_t)E * n, 256), 256, 0, st>>>(b2, pos, E, n, bp16);
  DPOT_LAUNCH_CHECK("tagg_bp16_kernel");
  return 0;
}
int tk_pad_split(const float* src, int64_t lds, int64_t rows, int cols, int colsp, __half* dst, cudaStream_t st) {
  pad_split_kernel<<<blocks_for(rows * colsp, 256), 256, 0, st>>>(src, lds, rows, cols, colsp, dst);
  DPOT_LAUNCH_CHECK("pad_split_kernel");
  return 0;
}
int tk_tagg_pad_g(const float* dWeffT, int E, int Kp, int T, int mid, int midp, __half* dst, cudaStream_t st) {
  tagg_pad_g_kernel<<<blocks_for((int64_t)T * E * midp, 256), 256, 0, st>>>(dWeffT, E, Kp, T, mid, midp, dst);
  DPOT_LAUNCH_CHECK("tagg_pad_g_kernel");
  return 0;
}
int tk_tagg_dw2_finish(const float* slabs, int T, int E, int mid, int midp, const float* inv_scale, float* dst, cudaStream_t st) {
  tagg_dw2_finish_kernel<<<blocks_for((int64_t)E * mid, 256), 256, 0, st>>>(slabs, T, E, mid, midp, inv_scale, dst);
  DPOT_LAUNCH_CHECK("tagg_dw2_finish_kernel");
  return 0;
}
int tk_tagg_finish(const float* dwt, const float* w, const float* temb, int T, int E, const float* inv_scale, float* dw,
                   float* dtemb, cudaStream_t st) {
  const int64_t rows = (int64_t)T * E;
  tagg_finish_kernel<<<blocks_for(rows, 8), 256, 0, st>>>(dwt, w, temb, rows, E, inv_scale, dw, dtemb);
  DPOT_LAUNCH_CHECK("tagg_finish_kernel");
  return 0;
}
int tk_tagg_gamma_grad(const float* dtemb, const float* gamma, int T, int E, const float* inv_scale, float* dgamma, cudaStream_t st) {
  tagg_gamma_grad_kernel<<<(unsigned)ceil_div(E, 128), 128, 0, st>>>(dtemb, gamma, T, E, inv_scale, dgamma);
  DPOT_LAUNCH_CHECK("tagg_gamma_grad_kernel");
  return 0;
}
int tk_rowsum_scale(const float* X, int rows, int cols, const float* inv_scale, float* rowsum, float* scaled, cudaStream_t st) {
  rowsum_scale_kernel<<<blocks_for(rows, 8), 256, 0, st>>>(X, rows, cols, inv_scale, rowsum, scaled);
  DPOT_LAUNCH_CHECK("rowsum_scale_kernel");
  return 0;
}
int tk_transpose(const float* src, int64_t lds, float* dst, int64_t ldd, int R, int Cc, cudaStream_t st) {
  transpose2_kernel<<<dim3((unsigned)ceil_div(Cc, 32), (unsigned)ceil_div(R, 32)), dim3(32, 8), 0, st>>>(src, lds, dst, ldd, R, Cc);
  DPOT_LAUNCH_CHECK("transpose2_kernel");
  return 0;
}
int tk_add_mean_grad(const float* dtok, int B, int n, int E, float* g, cudaStream_t st) {
  DPOT_REQUIRE(E % 4 == 0, DPOT_E_BADARG, "add_mean_grad: E %% 4");
  const int64_t total4 = (int64_t)B * n * E / 4;
  add_mean_grad_kernel<<<blocks_for(total4, 256), 256, 0, st>>>(dtok, total4, n, E, g);
  DPOT_LAUNCH_CHECK("add_mean_grad_kernel");
  return 0;
}
int tk_split_scaled(const float* src, int64_t rows, int cols, const float* factor, __half* dst, int64_t ldd, int64_t lo_off,
                    cudaStream_t st) {
  DPOT_REQUIRE(cols % 8 == 0, DPOT_E_BADARG, "split_scaled: cols %% 8");
  const int64_t total = rows * (cols / 8);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 16);
  split_scaled_kernel<<<grid, 256, 0, st>>>(src, rows, cols / 8, factor, dst, ldd, lo_off);
  DPOT_LAUNCH_CHECK("split_scaled_kernel");
  return 0;
}

}  // namespace dpot
