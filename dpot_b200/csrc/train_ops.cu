// Backward-pass building blocks (autograd of models/dpot.py:364-403 through all AR steps,
// train_temporal.py:225-227): weight-gradient contraction, column sums, transposes, GroupNorm
// forward-apply / backward, pixel (un)shuffle of the output head and the AFNO weight-gradient unpack.
#include "common.cuh"
#include "gemm_common.cuh"

namespace dpot {
namespace {

// ------------------------------------------------------------------------------------------------
// dW[n,k] (+)= sum_m X[m,n] * Y'[m,k]      (contraction over the token axis; both operands row-major)
// 64x64 output tile, 256 threads x (4x4), m-slabs of 16.  grid = (K/64, N/64, batch*splits); split-M
// partial sums are combined with fp32 atomics (dW is zeroed by the host wrapper unless accumulating).
constexpr int WG_T = 64, WG_MK = 16, WG_NT = 256;

struct WgradDev {
  const float* X; int64_t ldx;
  float* dW; int64_t ldw;
  int M, N, K;
  int batch, splits;
  int64_t sX, sY, sW;
  GemmDev y;   // Y operand accessor (A/lda/a_mode/patch geometry/a_scale tables), y.M = M, y.K = K
};

__global__ void __launch_bounds__(WG_NT) wgrad_kernel(const WgradDev p) {
  __shared__ __align__(16) float Xs[WG_MK][WG_T + 4];
  __shared__ __align__(16) float Ys[WG_MK][WG_T + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int k0 = blockIdx.x * WG_T, n0 = blockIdx.y * WG_T;
  const int bz = blockIdx.z / p.splits, sp = blockIdx.z % p.splits;
  const int mchunk = (int)(((int64_t)p.M + p.splits - 1) / p.splits);
  const int mbeg = sp * mchunk, mend = min(p.M, mbeg + mchunk);
  const float* __restrict__ X = p.X + bz * p.sX;
  const float* __restrict__ Yb = p.y.A + bz * p.sY;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int m0 = mbeg; m0 < mend; m0 += WG_MK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = tid + j * WG_NT;       // 16 x 64 elements
      const int r = idx / WG_T, c = idx % WG_T;
      const int m = m0 + r;
      Xs[r][c] = (m < mend && n0 + c < p.N) ? X[(int64_t)m * p.ldx + n0 + c] : 0.f;
      Ys[r][c] = (m < mend) ? gemm_load_a(p.y, Yb, m, k0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < WG_MK; ++r) {
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[r][ty * 4]);
      const float4 yv = *reinterpret_cast<const float4*>(&Ys[r][tx * 4]);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ya[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* __restrict__ dW = p.dW + bz * p.sW;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= p.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k >= p.K) continue;
      atomicAdd(&dW[(int64_t)n * p.ldw + k], acc[i][j]);
    }
  }
}

// out[n] (+)= sum_m X[m,n]
__global__ void colsum_kernel(const float* __restrict__ X, int64_t ldx, int M, int N, int rows_per_block,
                              float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc = 0.f;
  for (int m = m0; m < m1; ++m) acc += X[(int64_t)m * ldx + n];
  atomicAdd(&out[n], acc);
}

__global__ void transpose_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int R,
                                 int Cc, int64_t sS, int64_t sD) {
  __shared__ float tile[32][33];
  const float* s = src + blockIdx.z * sS;
  float* d = dst + blockIdx.z * sD;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? s[(int64_t)r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < R) d[(int64_t)c * ldd + r] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                const float* __restrict__ shift, int64_t total, int n, int E, float* __restrict__ out) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= total) return;
  const int64_t row = i / E; const int c = (int)(i % E);
  const int64_t t = (row / n) * E + c;
  const float4 v = *reinterpret_cast<const float4*>(x + i);
  const float4 s = *reinterpret_cast<const float4*>(scale + t);
  const float4 h = *reinterpret_cast<const float4*>(shift + t);
  float4 o;
  o.x = fmaf(v.x, s.x, h.x); o.y = fmaf(v.y, s.y, h.y); o.z = fmaf(v.z, s.z, h.z); o.w = fmaf(v.w, s.w, h.w);
  *reinterpret_cast<float4*>(out + i) = o;
}

// per (sample, channel): A1 = sum_pos dy, A2 = sum_pos dy * xhat
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const double* __restrict__ stats, int n, int E, int groups,
                                                            float eps, float* __restrict__ A1, float* __restrict__ A2) {
  __shared__ float r1[8][32], r2[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y, c = blockIdx.x * 32 + lane;
  float s1 = 0.f, s2 = 0.f;
  if (c < E) {
    const int gs = E / groups;
    const double cnt = (double)gs * n;
    const double* st = stats + ((int64_t)b * groups + c / gs) * 2;
    const double mean = st[0] / cnt;
    double var = st[1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps)), mu = (float)mean;
    const int64_t base = (int64_t)b * n * E + c;
    for (int pos = w; pos < n; pos += 8) {
      const float g = dy[base + (int64_t)pos * E];
      const float xh = (x[base + (int64_t)pos * E] - mu) * rstd;
      s1 += g;
      s2 = fmaf(g, xh, s2);
    }
  }
  r1[w][lane] = s1; r2[w][lane] = s2;
  __syncthreads();
  if (w == 0 && c < E) {
    float t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < 8; ++k) { t1 += r1[k][lane]; t2 += r2[k][lane]; }
    A1[(int64_t)b * E + c] = t1;
    A2[(int64_t)b * E + c] = t2;
  }
}

// per (sample, group): G1 = sum_c gamma*A1, G2 = sum_c gamma*A2 ; per channel: dgamma += sum_b A2, dbeta += sum_b A1
__global__ void gn_bwd_group_kernel(const float* __restrict__ A1, const float* __restrict__ A2,
                                    const float* __restrict__ gamma, int B, int E, int groups, float* __restrict__ G,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int gs = E / groups;
  if (i < B * groups) {
    const int b = i / groups, g = i % groups;
    double g1 = 0.0, g2 = 0.0;
    for (int c = g * gs; c < (g + 1) * gs; ++c) {
      g1 += (double)gamma[c] * A1[(int64_t)b * E + c];
      g2 += (double)gamma[c] * A2[(int64_t)b * E + c];
    }
    G[2 * i] = (float)g1; G[2 * i + 1] = (float)g2;
  }
  if (i < E) {
    double d1 = 0.0, d2 = 0.0;
    for (int b = 0; b < B; ++b) { d1 += A1[(int64_t)b * E + i]; d2 += A2[(int64_t)b * E + i]; }
    dbeta[i] += (float)d1;
    dgamma[i] += (float)d2;
  }
}

// dx = rstd * (gamma*dy - G1/N - xhat*G2/N)  (+ add[i] when given)
__global__ void gn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                    const double* __restrict__ stats, const float* __restrict__ gamma,
                                    const float* __restrict__ G, const float* __restrict__ add, int64_t total, int n,
                                    int E, int groups, float eps, float* __restrict__ dx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t row = i / E; const int c = (int)(i % E);
  const int b = (int)(row / n);
  const int gs = E / groups, g = c / gs;
  const double cnt = (double)gs * n;
  const double* st = stats + ((int64_t)b * groups + g) * 2;
  const double mean = st[0] / cnt;
  double var = st[1] / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float xh = (x[i] - (float)mean) * rstd;
  const float inv = (float)(1.0 / cnt);
  float v = rstd * (gamma[c] * dy[i] - G[2 * (b * groups + g)] * inv - xh * G[2 * (b * groups + g) + 1] * inv);
  if (add) v += add[i];
  dx[i] = v;
}

// out = dy * act'(pre)
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre, int act, int64_t total,
                               float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = dy[i] * act_grad(pre[i], act);
}

// ------------------------------------------------------------------------------------------------
// rows (b,p,q,u,v) x C  <->  field [b, p*P+u, q*P+v, C]
__global__ void pixel_shuffle_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int h, int w, int P,
                                     int Cc, int to_field) {
  const int64_t total = (int64_t)B * h * w * P * P * Cc;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % Cc);
  int64_t pix = i / Cc;
  const int uv = (int)(pix % (P * P)); int64_t r = pix / (P * P);
  const int q = (int)(r % w); r /= w;
  const int p = (int)(r % h); const int b = (int)(r / h);
  const int u = uv / P, v = uv % P;
  const int64_t f = ((((int64_t)b * h * P + p * P + u) * (w * P)) + q * P + v) * Cc + c;
  if (to_field) dst[f] = src[i]; else dst[i] = src[f];
}

// dWc[nb,2bs(out),2bs(in)], dbc[nb,2bs] -> dw[2,nb,bs(in),bs(out)], db[2,nb,bs]   (transpose of pack_afno)
__global__ void unpack_afno_grad_kernel(const float* __restrict__ dWc, const float* __restrict__ dbc, int nb, int bs,
                                        float* __restrict__ dw, float* __restrict__ db) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)nb * bs * bs;
  if (i < per) {
    const int o = (int)(i % bs), ki = (int)((i / bs) % bs), kap = (int)(i / ((int64_t)bs * bs));
    const float* Wk = dWc + (int64_t)kap * 4 * bs * bs;
    const int64_t L = 2 * bs;
    // Wc[n=o][k=ki] = wr, Wc[o][ki+bs] = -wi, Wc[o+bs][ki] = wi, Wc[o+bs][ki+bs] = wr
    dw[i] += Wk[(int64_t)o * L + ki] + Wk[(int64_t)(o + bs) * L + ki + bs];
    dw[per + i] += -Wk[(int64_t)o * L + ki + bs] + Wk[(int64_t)(o + bs) * L + ki];
  }
  if (i < (int64_t)nb * bs) {
    const int o = (int)(i % bs), kap = (int)(i / bs);
    db[i] += dbc[(int64_t)kap * 2 * bs + o];
    db[(int64_t)nb * bs + i] += dbc[(int64_t)kap * 2 * bs + bs + o];
  }
}

}  // namespace
}  // namespace dpot

using namespace dpot;

extern "C" int dpot_wgrad(const dpot_wgrad_args* a, void* stream) {
  DPOT_REQUIRE(a && a->X && a->Y && a->dW, DPOT_E_BADARG, "dpot_wgrad: null pointer");
  DPOT_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0 && a->batch >= 1, DPOT_E_BADARG, "dpot_wgrad: bad shape");
  cudaStream_t st = as_stream(stream);
  WgradDev p;
  memset(&p, 0, sizeof(p));
  p.X = a->X; p.ldx = a->ldx; p.dW = a->dW; p.ldw = a->ldw; p.M = a->M; p.N = a->N; p.K = a->K;
  p.batch = a->batch; p.sX = a->strideX; p.sY = a->strideY; p.sW = a->strideW;
  p.y.A = a->Y; p.y.lda = a->ldy; p.y.M = a->M; p.y.K = a->K;
  p.y.a_scale = a->y_scale; p.y.a_shift = a->y_shift; p.y.a_rps = a->y_rows_per_sample;
  p.y.a_mode = a->y_mode; p.y.pX = a->pX; p.y.pY = a->pY; p.y.pT = a->pT; p.y.pC = a->pC; p.y.pP = a->pP;
  DPOT_REQUIRE((a->y_scale == nullptr) == (a->y_shift == nullptr), DPOT_E_BADARG, "dpot_wgrad: y_scale/y_shift");
  if (a->y_mode == DPOT_A_PATCH) {
    DPOT_REQUIRE(a->pP > 0 && a->pX % a->pP == 0 && a->pY % a->pP == 0 && a->K == a->pP * a->pP * a->pC && a->batch == 1,
                 DPOT_E_BADARG, "dpot_wgrad: patch geometry");
    p.y.ph = a->pX / a->pP; p.y.pw = a->pY / a->pP;
  }
  if (!a->accumulate) {
    for (int b = 0; b < a->batch; ++b)
      DPOT_CUDA(cudaMemset2DAsync(a->dW + b * a->strideW, sizeof(float) * a->ldw, 0, sizeof(float) * a->K, a->N, st));
  }
  if (a->M == 0) return 0;
  const int64_t tiles = ceil_div(a->N, WG_T) * ceil_div(a->K, WG_T) * a->batch;
  int64_t splits = ceil_div(4 * 148, tiles);
  const int64_t max_splits = std::max<int64_t>(1, a->M / 256);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
  dim3 grid((unsigned)ceil_div(a->K, WG_T), (unsigned)ceil_div(a->N, WG_T), (unsigned)(a->batch * splits));
  DPOT_REQUIRE(grid.z <= 65535u, DPOT_E_BADARG, "dpot_wgrad: grid too large");
  wgrad_kernel<<<grid, WG_NT, 0, st>>>(p);
  DPOT_LAUNCH_CHECK("wgrad_kernel");
  return 0;
}

extern "C" int dpot_colsum(const float* X, int64_t ldx, int32_t M, int32_t N, float* out, int32_t accumulate, void* stream) {
  DPOT_REQUIRE(X && out && M >= 0 && N > 0, DPOT_E_BADARG, "dpot_colsum: bad args");
  cudaStream_t st = as_stream(stream);
  if (!accumulate) DPOT_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
  if (M == 0) return 0;
  const int rpb = 256;
  dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div(M, rpb));
  colsum_kernel<<<grid, 128, 0, st>>>(X, ldx, M, N, rpb, out);
  DPOT_LAUNCH_CHECK("colsum_kernel");
  return 0;
}

extern "C" int dpot_transpose(const float* src, int64_t lds, float* dst, int64_t ldd, int32_t R, int32_t Cc, int32_t batch,
                              int64_t stride_src, int64_t stride_dst, void* stream) {
  DPOT_REQUIRE(src && dst && R > 0 && Cc > 0 && batch >= 1, DPOT_E_BADARG, "dpot_transpose: bad args");
  dim3 grid((unsigned)ceil_div(Cc, 32), (unsigned)ceil_div(R, 32), (unsigned)batch);
  transpose_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(src, lds, dst, ldd, R, Cc, stride_src, stride_dst);
  DPOT_LAUNCH_CHECK("transpose_kernel");
  return 0;
}

extern "C" int dpot_gn_apply(const float* x, const float* scale, const float* shift, int32_t B, int32_t n, int32_t E,
                             float* out, void* stream) {
  DPOT_REQUIRE(x && scale && shift && out && E % 4 == 0, DPOT_E_BADARG, "dpot_gn_apply: bad args (E must be a multiple of 4)");
  const int64_t total = (int64_t)B * n * E;
  gn_apply_kernel<<<(unsigned)ceil_div(total / 4, 256), 256, 0, as_stream(stream)>>>(x, scale, shift, total, n, E, out);
  DPOT_LAUNCH_CHECK("gn_apply_kernel");
  return 0;
}

extern "C" int dpot_gn_bwd(const float* dy, const float* x, const double* stats, const float* gamma, const float* add,
                           int32_t B, int32_t n, int32_t E, int32_t groups, float eps, float* scratch, float* dx,
                           float* dgamma, float* dbeta, void* stream) {
  DPOT_REQUIRE(dy && x && stats && gamma && scratch && dx && dgamma && dbeta, DPOT_E_BADARG, "dpot_gn_bwd: null pointer");
  DPOT_REQUIRE(E % groups == 0, DPOT_E_BADARG, "dpot_gn_bwd: E %% groups");
  cudaStream_t st = as_stream(stream);
  float* A1 = scratch; float* A2 = scratch + (int64_t)B * E; float* G = A2 + (int64_t)B * E;   // needs 2*B*E + 2*B*groups floats
  gn_bwd_reduce_kernel<<<dim3((unsigned)ceil_div(E, 32), (unsigned)B), 256, 0, st>>>(dy, x, stats, n, E, groups, eps, A1, A2);
  DPOT_LAUNCH_CHECK("gn_bwd_reduce_kernel");
  const int mx = std::max(B * groups, E);
  gn_bwd_group_kernel<<<(unsigned)ceil_div(mx, 128), 128, 0, st>>>(A1, A2, gamma, B, E, groups, G, dgamma, dbeta);
  DPOT_LAUNCH_CHECK("gn_bwd_group_kernel");
  const int64_t total = (int64_t)B * n * E;
  gn_bwd_apply_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(dy, x, stats, gamma, G, add, total, n, E, groups, eps, dx);
  DPOT_LAUNCH_CHECK("gn_bwd_apply_kernel");
  return 0;
}

extern "C" int dpot_act_bwd(const float* dy, const float* pre, int32_t act, int64_t total, float* out, void* stream) {
  DPOT_REQUIRE(dy && pre && out && total >= 0, DPOT_E_BADARG, "dpot_act_bwd: bad args");
  if (total == 0) return 0;
  act_bwd_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(dy, pre, act, total, out);
  DPOT_LAUNCH_CHECK("act_bwd_kernel");
  return 0;
}

extern "C" int dpot_pixel_shuffle(const float* src, float* dst, int32_t B, int32_t h, int32_t w, int32_t P, int32_t Cc,
                                  int32_t to_field, void* stream) {
  DPOT_REQUIRE(src && dst, DPOT_E_BADARG, "dpot_pixel_shuffle: null pointer");
  const int64_t total = (int64_t)B * h * w * P * P * Cc;
  pixel_shuffle_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(src, dst, B, h, w, P, Cc, to_field);
  DPOT_LAUNCH_CHECK("pixel_shuffle_kernel");
  return 0;
}

extern "C" int dpot_unpack_afno_grad(const float* dWc, const float* dbc, int32_t nb, int32_t bs, float* dw, float* db,
                                     void* stream) {
  DPOT_REQUIRE(dWc && dbc && dw && db, DPOT_E_BADARG, "dpot_unpack_afno_grad: null pointer");
  const int64_t per = (int64_t)nb * bs * bs;
  unpack_afno_grad_kernel<<<(unsigned)ceil_div(per, 256), 256, 0, as_stream(stream)>>>(dWc, dbc, nb, bs, dw, db);
  DPOT_LAUNCH_CHECK("unpack_afno_grad_kernel");
  return 0;
}
