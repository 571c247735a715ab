// GroupNorm statistics, weight packing / folding, output-head tail, rollout window advance.
// All of these are HBM/L2-bound streaming kernels: coalesced along the channel axis.
#include "common.cuh"
#include "gemm_common.cuh"
#include <cuda_fp16.h>

namespace dpot {
namespace {

// ---------------------------------------------------------------------------- GroupNorm stats
constexpr int GN_NT = 256, GN_ROWS = 32, GN_KMAX = 8, GN_GMAX = 64;

__global__ void __launch_bounds__(GN_NT) gn_stats_kernel(const float* __restrict__ x, int n, int E, int groups,
                                                         double* __restrict__ stats) {
  __shared__ double sg[2 * GN_GMAX];
  const int tid = threadIdx.x, b = blockIdx.y;
  const int r0 = blockIdx.x * GN_ROWS, r1 = min(n, r0 + GN_ROWS);
  const int gs = E / groups;
  for (int i = tid; i < 2 * groups; i += GN_NT) sg[i] = 0.0;
  __syncthreads();
  const float* xp = x + (int64_t)b * n * E;
  const bool warp_uniform = (gs % 32 == 0);
  for (int c0 = 0; c0 < E; c0 += GN_NT * GN_KMAX) {
    double s1[GN_KMAX], s2[GN_KMAX];
#pragma unroll
    for (int k = 0; k < GN_KMAX; ++k) s1[k] = s2[k] = 0.0;
    for (int r = r0; r < r1; ++r) {
#pragma unroll
      for (int k = 0; k < GN_KMAX; ++k) {
        const int ch = c0 + k * GN_NT + tid;
        if (ch < E) {
          const float v = xp[(int64_t)r * E + ch];
          s1[k] += (double)v;
          s2[k] += (double)v * (double)v;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < GN_KMAX; ++k) {
      const int ch = c0 + k * GN_NT + tid;
      if (warp_uniform) {
        // E is a multiple of 32 here, so a warp is either fully live or fully dead
        if (c0 + k * GN_NT + (tid & ~31) < E) {
          const double t1 = warp_sum(s1[k]), t2 = warp_sum(s2[k]);
          if ((tid & 31) == 0) {
            atomicAdd(&sg[2 * (ch / gs)], t1);
            atomicAdd(&sg[2 * (ch / gs) + 1], t2);
          }
        }
      } else if (ch < E) {
        atomicAdd(&sg[2 * (ch / gs)], s1[k]);
        atomicAdd(&sg[2 * (ch / gs) + 1], s2[k]);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * groups; i += GN_NT) atomicAdd(&stats[(int64_t)b * groups * 2 + i], sg[i]);
}

__global__ void gn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int B, int n, int E, int groups, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * E) return;
  const int b = (int)(i / E), ch = (int)(i % E);
  const int gs = E / groups;
  const double cnt = (double)gs * (double)n;
  const double* s = stats + ((int64_t)b * groups + ch / gs) * 2;
  const double mean = s[0] / cnt;
  double var = s[1] / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const double g = (double)gamma[ch];
  scale[i] = (float)(rstd * g);
  shift[i] = (float)((double)beta[ch] - mean * rstd * g);
}

// ---------------------------------------------------------------------------- weight packing
__global__ void pack_afno_kernel(const float* __restrict__ w, const float* __restrict__ b, int nb, int bs,
                                 float* __restrict__ Wc, float* __restrict__ bc) {
  // Wc[kap][n][k], n,k in [0,2bs): out = [re | im], in = [re | im]
  const int64_t total = (int64_t)nb * 4 * bs * bs;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    const int k = (int)(i % (2 * bs));
    const int nn = (int)((i / (2 * bs)) % (2 * bs));
    const int kap = (int)(i / ((int64_t)4 * bs * bs));
    const int ki = k % bs, no = nn % bs;
    const float wr = w[(((int64_t)0 * nb + kap) * bs + ki) * bs + no];
    const float wi = w[(((int64_t)1 * nb + kap) * bs + ki) * bs + no];
    float v;
    if (nn < bs) v = (k < bs) ? wr : -wi;   // real out = Fr*Wr - Fi*Wi
    else         v = (k < bs) ? wi : wr;    // imag out = Fr*Wi + Fi*Wr
    Wc[i] = v;
  }
  if (i < (int64_t)nb * 2 * bs) {
    const int nn = (int)(i % (2 * bs));
    const int kap = (int)(i / (2 * bs));
    bc[i] = b[(((int64_t)(nn / bs)) * nb + kap) * bs + nn % bs];
  }
}

__global__ void pack_patch_kernel(const float* __restrict__ w0, const float* __restrict__ b0,
                                  const float* __restrict__ gx, const float* __restrict__ gy,
                                  const float* __restrict__ gt, int mid, int C, int P, int h, int w, int T,
                                  float* __restrict__ W0p, float* __restrict__ rowbias0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int K = P * P * C;
  if (i < (int64_t)mid * K) {
    const int k = (int)(i % K), m = (int)(i / K);
    const int c = k % C, uv = k / C;
    W0p[i] = w0[((int64_t)m * (C + 3) + c) * P * P + uv];
  }
  const int64_t nrb = (int64_t)h * w * T * mid;
  if (i < nrb) {
    const int m = (int)(i % mid);
    int64_t r = i / mid;
    const int t = (int)(r % T); r /= T;
    const int q = (int)(r % w); const int p = (int)(r / w);
    double acc = (double)b0[m];
    const float* wx = w0 + ((int64_t)m * (C + 3) + C) * P * P;
    const float* wy = wx + P * P;
    const float* wt = wy + P * P;
    for (int u = 0; u < P; ++u)
      for (int v = 0; v < P; ++v) {
        acc += (double)wx[u * P + v] * (double)gx[p * P + u];
        acc += (double)wy[u * P + v] * (double)gy[q * P + v];
        acc += (double)wt[u * P + v] * (double)gt[t];
      }
    rowbias0[i] = (float)acc;
  }
}

// WeffT[j, t*mid+m] = sum_i W2[i,m] * temb[t,i] * w[t,i,j]
__global__ void fold_weff_kernel(const float* __restrict__ W2, const float* __restrict__ w,
                                 const float* __restrict__ temb, int T, int E, int mid, int Kp,
                                 float* __restrict__ WeffT) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int col = blockIdx.y;  // in [0, Kp)
  if (j >= E) return;
  if (col >= T * mid) { WeffT[(int64_t)j * Kp + col] = 0.f; return; }
  const int t = col / mid, m = col % mid;
  const float* wt = w + (int64_t)t * E * E + j;
  const float* te = temb + (int64_t)t * E;
  double acc = 0.0;
  for (int i = 0; i < E; ++i) acc += (double)(W2[(int64_t)i * mid + m] * te[i]) * (double)wt[(int64_t)i * E];
  WeffT[(int64_t)j * Kp + col] = (float)acc;
}
// Wsum[i,j] = sum_t temb[t,i] * w[t,i,j]
__global__ void fold_wsum_kernel(const float* __restrict__ w, const float* __restrict__ temb, int T, int E,
                                 float* __restrict__ Wsum) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)E * E) return;
  const int i = (int)(idx / E);
  double acc = 0.0;
  for (int t = 0; t < T; ++t) acc += (double)temb[(int64_t)t * E + i] * (double)w[(int64_t)t * E * E + idx];
  Wsum[idx] = (float)acc;
}
// bias_eff[pq, j] = sum_i (b2[i] + pos[i,pq]) * Wsum[i,j]
__global__ void fold_bias_kernel(const float* __restrict__ b2, const float* __restrict__ pos,
                                 const float* __restrict__ Wsum, int E, int n, float* __restrict__ bias_eff) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int pq = blockIdx.y;
  if (j >= E) return;
  double acc = 0.0;
  for (int i = 0; i < E; ++i) acc += (double)(b2[i] + pos[(int64_t)i * n + pq]) * (double)Wsum[(int64_t)i * E + j];
  bias_eff[(int64_t)pq * E + j] = (float)acc;
}

__global__ void pack_out_kernel(const float* __restrict__ wt, const float* __restrict__ bt, int E, int old, int P,
                                float* __restrict__ WtT, float* __restrict__ bias_t) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int N = P * P * old;
  if (i < (int64_t)N * E) {
    const int e = (int)(i % E), nn = (int)(i / E);
    const int o = nn % old, uv = nn / old;
    WtT[i] = wt[((int64_t)e * old + o) * P * P + uv];
  }
  if (i < N) bias_t[i] = bt[i % old];
}

// ---------------------------------------------------------------------------- output tail
constexpr int TAIL_NOUT_MAX = 16;

template <int OLD, bool L2>
__global__ void __launch_bounds__(128) out_tail_kernel(const float* __restrict__ Y1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, const float* __restrict__ w4,
                                                       const float* __restrict__ b4, int B, int h, int w, int P,
                                                       int old, int nout, int act, const float* __restrict__ mu,
                                                       const float* __restrict__ sigma, int Co,
                                                       float* __restrict__ out) {
  extern __shared__ float sw[];  // w2[old*old] (if L2), b2[old], w4[nout*old], b4[nout]
  float* s_w2 = sw;
  float* s_b2 = s_w2 + (L2 ? old * old : 0);
  float* s_w4 = s_b2 + (L2 ? old : 0);
  float* s_b4 = s_w4 + nout * old;
  for (int i = threadIdx.x; i < (L2 ? old * old : 0); i += blockDim.x) s_w2[i] = w2[i];
  for (int i = threadIdx.x; i < (L2 ? old : 0); i += blockDim.x) s_b2[i] = b2[i];
  for (int i = threadIdx.x; i < nout * old; i += blockDim.x) s_w4[i] = w4[i];
  for (int i = threadIdx.x; i < nout; i += blockDim.x) s_b4[i] = b4[i];
  __syncthreads();
  const int64_t npix = (int64_t)B * h * w * P * P;
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  // pix = ((b*h + p)*w + q)*P*P + u*P + v  -- the row-major order of Y1's (row, uv) pairs
  const int uv = (int)(pix % (P * P));
  int64_t r = pix / (P * P);
  const int q = (int)(r % w); r /= w;
  const int p = (int)(r % h); const int b = (int)(r / h);
  const int u = uv / P, v = uv % P;
  const float* src = Y1 + pix * old;
  float y3[TAIL_NOUT_MAX];
#pragma unroll
  for (int c = 0; c < TAIL_NOUT_MAX; ++c) y3[c] = (c < nout) ? s_b4[c] : 0.f;
  if (L2) {
    float y1[OLD];
#pragma unroll
    for (int o = 0; o < OLD; o += 4) {
      const float4 t = *reinterpret_cast<const float4*>(src + o);
      y1[o] = t.x; y1[o + 1] = t.y; y1[o + 2] = t.z; y1[o + 3] = t.w;
    }
    for (int jn = 0; jn < OLD; ++jn) {
      float acc = s_b2[jn];
#pragma unroll
      for (int o = 0; o < OLD; ++o) acc = fmaf(s_w2[jn * OLD + o], y1[o], acc);
      acc = act_apply(acc, act);
#pragma unroll
      for (int c = 0; c < TAIL_NOUT_MAX; ++c)
        if (c < nout) y3[c] = fmaf(s_w4[c * OLD + jn], acc, y3[c]);
    }
  } else {
    for (int o = 0; o < old; ++o) {
      const float t = src[o];
#pragma unroll
      for (int c = 0; c < TAIL_NOUT_MAX; ++c)
        if (c < nout) y3[c] = fmaf(s_w4[c * old + o], t, y3[c]);
    }
  }
  const int X = h * P, Y = w * P;
  float* dst = out + (((int64_t)b * X + p * P + u) * Y + q * P + v) * nout;
#pragma unroll
  for (int c = 0; c < TAIL_NOUT_MAX; ++c) {
    if (c < nout) {
      float val = y3[c];
      if (mu) {  // x * sigma + mu, channel = c % Co   (models/dpot.py:401)
        const int cc = c % Co;
        val = fmaf(val, sigma[(int64_t)b * Co + cc], mu[(int64_t)b * Co + cc]);
      }
      dst[c] = val;
    }
  }
}

__global__ void spatial_mean_kernel(const float* __restrict__ a, int n, int E, float* __restrict__ tok) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (e >= E) return;
  const float* p = a + (int64_t)b * n * E + e;
  double acc = 0.0;
  for (int r = 0; r < n; ++r) acc += (double)p[(int64_t)r * E];
  tok[(int64_t)b * E + e] = (float)(acc / (double)n);
}

// split-fp16 rows [hi E | lo E]: 64 channels x 4 row slices per CTA, fp32 partial sums (<= n/4 terms) combined in double
__global__ void __launch_bounds__(256) spatial_mean16_kernel(const __half* __restrict__ a, int n, int E, float* __restrict__ tok,
                                                             __half* __restrict__ tok16) {
  __shared__ float red[2][4][64];
  const int c = threadIdx.x & 63, sl = threadIdx.x >> 6;
  const int e = blockIdx.x * 64 + c, b = blockIdx.y;
  float s_hi = 0.f, s_lo = 0.f;
  if (e < E) {
    const __half* p = a + (int64_t)b * n * 2 * E + e;
    for (int r = sl; r < n; r += 4) {
      s_hi += __half2float(p[(int64_t)r * 2 * E]);
      s_lo += __half2float(p[(int64_t)r * 2 * E + E]);
    }
  }
  red[0][sl][c] = s_hi; red[1][sl][c] = s_lo;
  __syncthreads();
  if (sl == 0 && e < E) {
    double hi = 0.0, lo = 0.0;
    for (int k = 0; k < 4; ++k) { hi += (double)red[0][k][c]; lo += (double)red[1][k][c]; }
    const float m = (float)((hi + lo * (1.0 / 2048.0)) / (double)n);
    if (tok) tok[(int64_t)b * E + e] = m;
    if (tok16) {                      // the cls head's first contraction reads the token split (rows [hi E | lo E])
      __half h, l;
      hl_split(m, h, l);
      tok16[(int64_t)b * 2 * E + e] = h;
      tok16[(int64_t)b * 2 * E + E + e] = l;
    }
  }
}

// The same with 16-byte loads (E % 8 == 0): a CTA owns 128 channels of one sample, 16 lanes-of-8-channels x 16 row
// slices; every thread keeps 8 loads of 16 B in flight (the 2-byte loads of the kernel above ran at 2 TB/s on an
// L2-resident latent: 16.9 us for 33.5 MB at DPOT-S, B = 32).
__global__ void __launch_bounds__(256) spatial_mean16v_kernel(const __half* __restrict__ a, int n, int E, float* __restrict__ tok,
                                                              __half* __restrict__ tok16) {
  __shared__ float red[2][16][128 + 4];
  const int cg = threadIdx.x & 15, sl = threadIdx.x >> 4;
  const int e0 = blockIdx.x * 128 + cg * 8, b = blockIdx.y;
  float sh[8], slo[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) sh[k] = slo[k] = 0.f;
  if (e0 < E) {
    const __half* p = a + (int64_t)b * n * 2 * E + e0;
    auto acc = [&](const uint4& v, float (&d)[8]) {
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(h[k]); d[2 * k] += f.x; d[2 * k + 1] += f.y; }
    };
    int r = sl;
    for (; r + 48 < n; r += 64) {
      uint4 vh[4], vl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        vh[q] = *reinterpret_cast<const uint4*>(p + (int64_t)(r + 16 * q) * 2 * E);
        vl[q] = *reinterpret_cast<const uint4*>(p + (int64_t)(r + 16 * q) * 2 * E + E);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) { acc(vh[q], sh); acc(vl[q], slo); }
    }
    for (; r < n; r += 16) {
      acc(*reinterpret_cast<const uint4*>(p + (int64_t)r * 2 * E), sh);
      acc(*reinterpret_cast<const uint4*>(p + (int64_t)r * 2 * E + E), slo);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { red[0][sl][cg * 8 + k] = sh[k]; red[1][sl][cg * 8 + k] = slo[k]; }
  __syncthreads();
  const int c = threadIdx.x, e = blockIdx.x * 128 + c;
  if (c < 128 && e < E) {
    double hi = 0.0, lo = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { hi += (double)red[0][k][c]; lo += (double)red[1][k][c]; }
    const float m = (float)((hi + lo * (1.0 / 2048.0)) / (double)n);
    if (tok) tok[(int64_t)b * E + e] = m;
    if (tok16) {
      __half h, l;
      hl_split(m, h, l);
      tok16[(int64_t)b * 2 * E + e] = h;
      tok16[(int64_t)b * 2 * E + E + e] = l;
    }
  }
}

// ---------------------------------------------------------------------------- input statistics
constexpr int IS_CMAX = 16;
__global__ void __launch_bounds__(1024) input_stats_kernel(const float* __restrict__ x, int64_t per_sample, int C,
                                                           int PP, float* __restrict__ musig,
                                                           float* __restrict__ a_scale, float* __restrict__ a_shift) {
  __shared__ double red[2 * IS_CMAX][32];
  __shared__ double fin[2 * IS_CMAX];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* xp = x + (int64_t)b * per_sample;
  double s1[IS_CMAX], s2[IS_CMAX];
#pragma unroll
  for (int c = 0; c < IS_CMAX; ++c) s1[c] = s2[c] = 0.0;
  const int64_t nvec = per_sample / C;
  for (int64_t i = tid; i < nvec; i += blockDim.x) {
#pragma unroll
    for (int c = 0; c < IS_CMAX; ++c)
      if (c < C) {
        const double v = (double)xp[i * C + c];
        s1[c] += v; s2[c] += v * v;
      }
  }
#pragma unroll
  for (int c = 0; c < IS_CMAX; ++c) {
    if (c < C) {
      const double t1 = warp_sum(s1[c]), t2 = warp_sum(s2[c]);
      if ((tid & 31) == 0) { red[2 * c][tid >> 5] = t1; red[2 * c + 1][tid >> 5] = t2; }
    }
  }
  __syncthreads();
  if (tid < 2 * C) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[tid][k];
    fin[tid] = t;
  }
  __syncthreads();
  if (tid < C) {
    const double cnt = (double)nvec;
    const double mean = fin[2 * tid] / cnt;
    double var = (fin[2 * tid + 1] - cnt * mean * mean) / (cnt - 1.0);  // torch.std: unbiased
    if (var < 0.0) var = 0.0;
    const float sg = (float)sqrt(var) + 1e-6f;
    musig[(int64_t)b * 2 * C + tid] = (float)mean;
    musig[(int64_t)b * 2 * C + C + tid] = sg;
  }
  __syncthreads();
  for (int i = tid; i < PP * C; i += blockDim.x) {
    const int c = i % C;
    const float mean = musig[(int64_t)b * 2 * C + c], sg = musig[(int64_t)b * 2 * C + C + c];
    a_scale[(int64_t)b * PP * C + i] = 1.0f / sg;
    a_shift[(int64_t)b * PP * C + i] = -mean / sg;
  }
}

// ---------------------------------------------------------------------------- rollout window
__global__ void window_advance_kernel(const float* __restrict__ xx, const float* __restrict__ im,
                                      float* __restrict__ xx_next, float* __restrict__ pred, int64_t npix, int T,
                                      int Tb, int C, int Ttot, int step) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = npix * T * C;
  if (i < total) {
    const int c = (int)(i % C);
    const int t = (int)((i / C) % T);
    const int64_t pix = i / ((int64_t)C * T);
    float v;
    if (t < T - Tb) v = xx[(pix * T + t + Tb) * C + c];
    else {
      v = im[(pix * Tb + (t - (T - Tb))) * C + c];
      if (pred) pred[(pix * Ttot + (int64_t)step * Tb + (t - (T - Tb))) * C + c] = v;
    }
    xx_next[i] = v;
  }
}

// ring form of the window advance: the new frames overwrite the oldest slots, nothing else moves
__global__ void ring_insert_kernel(const float* __restrict__ im, float* __restrict__ ring, float* __restrict__ pred,
                                   int64_t npix, int T, int Tb, int C, int Ttot, int slot0, int step) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = npix * Tb * C;
  if (i < total) {
    const int c = (int)(i % C);
    const int j = (int)((i / C) % Tb);
    const int64_t pix = i / ((int64_t)C * Tb);
    const float v = im[i];
    int slot = slot0 + j; if (slot >= T) slot -= T;
    ring[(pix * T + slot) * C + c] = v;
    if (pred) pred[(pix * Ttot + (int64_t)step * Tb + j) * C + c] = v;
  }
}

}  // namespace
}  // namespace dpot

using namespace dpot;

extern "C" int dpot_ring_insert(const float* im, float* ring, float* pred, int64_t npix, int32_t T, int32_t Tb,
                                int32_t C, int32_t Ttot, int32_t slot0, int32_t step, void* stream) {
  DPOT_REQUIRE(im && ring, DPOT_E_BADARG, "dpot_ring_insert: null pointer");
  DPOT_REQUIRE(Tb >= 1 && Tb <= T && slot0 >= 0 && slot0 < T, DPOT_E_BADARG, "dpot_ring_insert: bad T_bundle / slot");
  const int64_t total = npix * Tb * C;
  ring_insert_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(im, ring, pred, npix, T, Tb, C, Ttot,
                                                                                    slot0, step);
  DPOT_LAUNCH_CHECK("ring_insert_kernel");
  return 0;
}

extern "C" int dpot_gn_stats(const float* x, int32_t B, int32_t n, int32_t E, int32_t groups, double* stats, void* stream) {
  DPOT_REQUIRE(x && stats, DPOT_E_BADARG, "dpot_gn_stats: null pointer");
  DPOT_REQUIRE(B > 0 && n > 0 && E > 0 && groups > 0 && E % groups == 0 && groups <= GN_GMAX, DPOT_E_BADARG,
               "dpot_gn_stats: bad shape B=%d n=%d E=%d groups=%d", B, n, E, groups);
  cudaStream_t st = as_stream(stream);
  DPOT_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * groups, st));
  dim3 grid((unsigned)ceil_div(n, GN_ROWS), (unsigned)B);
  gn_stats_kernel<<<grid, GN_NT, 0, st>>>(x, n, E, groups, stats);
  DPOT_LAUNCH_CHECK("gn_stats_kernel");
  return 0;
}

extern "C" int dpot_gn_finalize(const double* stats, const float* gamma, const float* beta, int32_t B, int32_t n,
                                int32_t E, int32_t groups, float eps, float* scale, float* shift, void* stream) {
  DPOT_REQUIRE(stats && gamma && beta && scale && shift, DPOT_E_BADARG, "dpot_gn_finalize: null pointer");
  DPOT_REQUIRE(B > 0 && n > 0 && E > 0 && groups > 0 && E % groups == 0, DPOT_E_BADARG, "dpot_gn_finalize: bad shape");
  const int64_t total = (int64_t)B * E;
  gn_finalize_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(stats, gamma, beta, B, n, E, groups,
                                                                                    eps, scale, shift);
  DPOT_LAUNCH_CHECK("gn_finalize_kernel");
  return 0;
}

extern "C" int dpot_pack_afno(const float* w, const float* b, int32_t nb, int32_t bs, float* Wc, float* bc, void* stream) {
  DPOT_REQUIRE(w && b && Wc && bc && nb > 0 && bs > 0, DPOT_E_BADARG, "dpot_pack_afno: bad args");
  const int64_t total = (int64_t)nb * 4 * bs * bs;
  pack_afno_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w, b, nb, bs, Wc, bc);
  DPOT_LAUNCH_CHECK("pack_afno_kernel");
  return 0;
}

extern "C" int dpot_pack_patch(const float* w0, const float* b0, const float* gx, const float* gy, const float* gt,
                               int32_t mid, int32_t C, int32_t P, int32_t h, int32_t w, int32_t T, float* W0p,
                               float* rowbias0, void* stream) {
  DPOT_REQUIRE(w0 && b0 && gx && gy && gt && W0p && rowbias0, DPOT_E_BADARG, "dpot_pack_patch: null pointer");
  const int64_t total = std::max((int64_t)mid * P * P * C, (int64_t)h * w * T * mid);
  pack_patch_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(w0, b0, gx, gy, gt, mid, C, P, h, w, T,
                                                                                   W0p, rowbias0);
  DPOT_LAUNCH_CHECK("pack_patch_kernel");
  return 0;
}

extern "C" int dpot_fold_timeagg(const float* W2, const float* b2, const float* pos, const float* w, const float* temb,
                                 int32_t T, int32_t E, int32_t mid, int32_t n, int32_t Kp, float* WeffT,
                                 float* bias_eff, void* stream) {
  DPOT_REQUIRE(W2 && b2 && pos && w && temb && WeffT && bias_eff, DPOT_E_BADARG, "dpot_fold_timeagg: null pointer");
  DPOT_REQUIRE(Kp >= T * mid && Kp <= 65535 && n <= 65535, DPOT_E_BADARG, "dpot_fold_timeagg: bad Kp/n");
  cudaStream_t st = as_stream(stream);
  fold_weff_kernel<<<dim3((unsigned)ceil_div(E, 128), (unsigned)Kp), 128, 0, st>>>(W2, w, temb, T, E, mid, Kp, WeffT);
  DPOT_LAUNCH_CHECK("fold_weff_kernel");
  // Wsum scratch lives in the tail of bias_eff's arena slot: the caller provides n*E + E*E floats
  float* Wsum = bias_eff + (int64_t)n * E;
  fold_wsum_kernel<<<(unsigned)ceil_div((int64_t)E * E, 256), 256, 0, st>>>(w, temb, T, E, Wsum);
  DPOT_LAUNCH_CHECK("fold_wsum_kernel");
  fold_bias_kernel<<<dim3((unsigned)ceil_div(E, 128), (unsigned)n), 128, 0, st>>>(b2, pos, Wsum, E, n, bias_eff);
  DPOT_LAUNCH_CHECK("fold_bias_kernel");
  return 0;
}

extern "C" int dpot_pack_out(const float* wt, const float* bt, int32_t E, int32_t old, int32_t P, float* WtT,
                             float* bias_t, void* stream) {
  DPOT_REQUIRE(wt && bt && WtT && bias_t, DPOT_E_BADARG, "dpot_pack_out: null pointer");
  const int64_t total = (int64_t)P * P * old * E;
  pack_out_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(wt, bt, E, old, P, WtT, bias_t);
  DPOT_LAUNCH_CHECK("pack_out_kernel");
  return 0;
}

namespace dpot {
int g_tail_engine = 0;   // 0 auto, 1 CUDA cores only (dpot_out_tail_set_engine; tests)
int out_tail_mma_launch(const float* Y1, const float* w2, const float* b2, const float* w4, const float* b4, int B, int h,
                        int w, int P, int old, int nout, int act, const float* mu, const float* sigma, int Co, float* out,
                        cudaStream_t st, bool* served, float* ring, float* pred, int T, int slot0, int Ttot, int step);
}
namespace dpot { extern int g_tail_tc; }
// 0 = auto (tcgen05 tail where the forward can feed it, else warp-MMA, else CUDA cores), 1 = CUDA cores only, 2 = no tcgen05 tail
extern "C" void dpot_out_tail_set_engine(int32_t engine) { dpot::g_tail_engine = engine == 2 ? 0 : engine; dpot::g_tail_tc = engine == 0 ? 1 : 0; }

extern "C" int dpot_out_tail(const float* Y1, const float* w2, const float* b2, const float* w4, const float* b4,
                             int32_t B, int32_t h, int32_t w, int32_t P, int32_t old, int32_t nout, int32_t act,
                             const float* mu, const float* sigma, int32_t Co, float* out, void* stream) {
  DPOT_REQUIRE(Y1 && w4 && b4 && out, DPOT_E_BADARG, "dpot_out_tail: null pointer");
  DPOT_REQUIRE(nout >= 1 && nout <= TAIL_NOUT_MAX, DPOT_E_UNSUPPORTED, "dpot_out_tail: out_channels*out_timesteps=%d > %d", nout, TAIL_NOUT_MAX);
  DPOT_REQUIRE((mu == nullptr) == (sigma == nullptr), DPOT_E_BADARG, "dpot_out_tail: mu/sigma must come together");
  const int64_t npix = (int64_t)B * h * w * P * P;
  const unsigned grid = (unsigned)ceil_div(npix, 128);
  cudaStream_t st = as_stream(stream);
  const bool l2 = (w2 != nullptr);
  if (l2 && g_tail_engine != 1) {   // warp-MMA kernel (out_tail_mma.cu) when the geometry allows
    DPOT_REQUIRE(b2 != nullptr, DPOT_E_BADARG, "dpot_out_tail: b2 missing");
    bool served = false;
    DPOT_CALL(out_tail_mma_launch(Y1, w2, b2, w4, b4, B, h, w, P, old, nout, act, mu, sigma, Co, out, st, &served, nullptr,
                                  nullptr, 0, 0, 0, 0));
    if (served) return 0;
  }
  const size_t smem = sizeof(float) * ((l2 ? (size_t)old * old + old : 0) + (size_t)nout * old + nout);
#define TAIL_CASE(O)                                                                                              \
  case O:                                                                                                         \
    out_tail_kernel<O, true><<<grid, 128, smem, st>>>(Y1, w2, b2, w4, b4, B, h, w, P, old, nout, act, mu, sigma, Co, out); \
    break;
  if (l2) {
    DPOT_REQUIRE(b2 != nullptr, DPOT_E_BADARG, "dpot_out_tail: b2 missing");
    switch (old) {
      TAIL_CASE(4) TAIL_CASE(8) TAIL_CASE(16) TAIL_CASE(32)
      default:
        set_error("dpot_out_tail: fused layer-2 path takes out_layer_dim in {4,8,16,32}, got %d (run layer 2 through dpot_gemm)", old);
        return DPOT_E_UNSUPPORTED;
    }
  } else {
    out_tail_kernel<4, false><<<grid, 128, smem, st>>>(Y1, w2, b2, w4, b4, B, h, w, P, old, nout, act, mu, sigma, Co, out);
  }
#undef TAIL_CASE
  DPOT_LAUNCH_CHECK("out_tail_kernel");
  return 0;
}

// Output tail writing the new frames straight into the autoregressive ring window and the prediction tensor (see
// TailRing, out_tail_mma.cu); geometries the fused kernel does not take go through y + dpot_ring_insert.
extern "C" int dpot_out_tail_ring(const float* Y1, const float* w2, const float* b2, const float* w4, const float* b4,
                                  int32_t B, int32_t h, int32_t w, int32_t P, int32_t old, int32_t nout, int32_t act,
                                  const float* mu, const float* sigma, int32_t Co, float* y_scratch, float* ring, float* pred,
                                  int32_t T, int32_t slot0, int32_t Ttot, int32_t step, void* stream) {
  DPOT_REQUIRE(Y1 && w4 && b4 && ring && y_scratch, DPOT_E_BADARG, "dpot_out_tail_ring: null pointer");
  DPOT_REQUIRE(Co > 0 && nout % Co == 0 && slot0 >= 0 && slot0 < T && nout / Co <= T, DPOT_E_BADARG, "dpot_out_tail_ring: bad geometry");
  cudaStream_t st = as_stream(stream);
  if (w2 != nullptr && g_tail_engine != 1) {
    DPOT_REQUIRE(b2 != nullptr, DPOT_E_BADARG, "dpot_out_tail_ring: b2 missing");
    bool served = false;
    DPOT_CALL(out_tail_mma_launch(Y1, w2, b2, w4, b4, B, h, w, P, old, nout, act, mu, sigma, Co, y_scratch, st, &served, ring,
                                  pred, T, slot0, Ttot, step));
    if (served) return 0;
  }
  DPOT_CALL(dpot_out_tail(Y1, w2, b2, w4, b4, B, h, w, P, old, nout, act, mu, sigma, Co, y_scratch, stream));
  return dpot_ring_insert(y_scratch, ring, pred, (int64_t)B * h * P * w * P, T, nout / Co, Co, Ttot, slot0, step, stream);
}

extern "C" int dpot_spatial_mean(const float* a, int32_t B, int32_t n, int32_t E, float* tok, void* stream) {
  DPOT_REQUIRE(a && tok && B > 0 && n > 0 && E > 0, DPOT_E_BADARG, "dpot_spatial_mean: bad args");
  spatial_mean_kernel<<<dim3((unsigned)ceil_div(E, 128), (unsigned)B), 128, 0, as_stream(stream)>>>(a, n, E, tok);
  DPOT_LAUNCH_CHECK("spatial_mean_kernel");
  return 0;
}

extern "C" int dpot_spatial_mean16(const void* a16, int32_t B, int32_t n, int32_t E, float* tok, void* stream) {
  return dpot_spatial_mean16s(a16, B, n, E, tok, nullptr, stream);
}
extern "C" int dpot_spatial_mean16s(const void* a16, int32_t B, int32_t n, int32_t E, float* tok, void* tok16, void* stream) {
  DPOT_REQUIRE(a16 && (tok || tok16) && B > 0 && n > 0 && E > 0, DPOT_E_BADARG, "dpot_spatial_mean16: bad args");
  if (E % 8 == 0 && reinterpret_cast<uintptr_t>(a16) % 16 == 0) {
    spatial_mean16v_kernel<<<dim3((unsigned)ceil_div(E, 128), (unsigned)B), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const __half*>(a16), n, E, tok, reinterpret_cast<__half*>(tok16));
    DPOT_LAUNCH_CHECK("spatial_mean16v_kernel");
    return 0;
  }
  spatial_mean16_kernel<<<dim3((unsigned)ceil_div(E, 64), (unsigned)B), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const __half*>(a16), n, E, tok, reinterpret_cast<__half*>(tok16));
  DPOT_LAUNCH_CHECK("spatial_mean16_kernel");
  return 0;
}

extern "C" int dpot_input_stats(const float* x, int32_t B, int64_t per_sample, int32_t C, int32_t PP, float* musig,
                                float* a_scale, float* a_shift, void* stream) {
  DPOT_REQUIRE(x && musig && a_scale && a_shift, DPOT_E_BADARG, "dpot_input_stats: null pointer");
  DPOT_REQUIRE(C >= 1 && C <= IS_CMAX && per_sample % C == 0, DPOT_E_UNSUPPORTED, "dpot_input_stats: in_channels=%d > %d", C, IS_CMAX);
  input_stats_kernel<<<(unsigned)B, 1024, 0, as_stream(stream)>>>(x, per_sample, C, PP, musig, a_scale, a_shift);
  DPOT_LAUNCH_CHECK("input_stats_kernel");
  return 0;
}

extern "C" int dpot_window_advance(const float* xx, const float* im, float* xx_next, float* pred, int64_t npix,
                                   int32_t T, int32_t Tb, int32_t C, int32_t Ttot, int32_t step, void* stream) {
  DPOT_REQUIRE(xx && im && xx_next, DPOT_E_BADARG, "dpot_window_advance: null pointer");
  DPOT_REQUIRE(Tb >= 1 && Tb <= T, DPOT_E_BADARG, "dpot_window_advance: T_bundle must be in [1, T]");
  DPOT_REQUIRE(xx != xx_next, DPOT_E_BADARG, "dpot_window_advance: in-place advance is not supported");
  const int64_t total = npix * T * C;
  window_advance_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(xx, im, xx_next, pred, npix, T, Tb, C,
                                                                                       Ttot, step);
  DPOT_LAUNCH_CHECK("window_advance_kernel");
  return 0;
}
