// AFNO2D spectral transforms: GroupNorm-apply + rfft2(ortho) -> kept-mode spectrum, and
// zero-padded irfft2(ortho) + skip (+ GroupNorm-2 statistics).   models/dpot.py:59,62,96-106.
//
// Layout decision: the latent is token-major a[(b,p,q), E], so a warp's 32 lanes map to 32
// adjacent CHANNELS (one coalesced 128 B line per token) and every lane owns whole 1-D
// transforms: an H-point FFT lives entirely in the registers of one thread (radix-2, fully
// unrolled, immediate twiddles).  The row->column turn goes through shared memory laid out
// [pos][channel], which is bank-conflict free for lane==channel.  No shuffles are needed: the
// batch dimension (channels) is the SIMD dimension.  Two real rows are packed into one complex
// transform (z = row_a + i row_b), halving the row-FFT work in both directions.
#include "common.cuh"

namespace dpot {
namespace {

constexpr int NT = 256;

__host__ __device__ constexpr float tw_cos(int j) {  // cos(2 pi j / 32), j in [0,16]
  constexpr float t[17] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                           0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f,
                           0.19509032201612833f, 0.f, -0.19509032201612819f, -0.38268343236508973f,
                           -0.55557023301960196f, -0.70710678118654746f, -0.83146961230254535f,
                           -0.92387953251128674f, -0.98078528040323043f, -1.f};
  return t[j];
}
__host__ __device__ constexpr float tw_sin(int j) {  // sin(2 pi j / 32)
  constexpr float t[17] = {0.f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                           0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f,
                           0.98078528040323043f, 1.f, 0.98078528040323043f, 0.92387953251128674f,
                           0.83146961230254546f, 0.70710678118654757f, 0.55557023301960218f,
                           0.38268343236508989f, 0.19509032201612861f, 0.f};
  return t[j];
}

__host__ __device__ constexpr int bitrev(int i, int n) {
  int r = 0;
  for (int b = 1; b < n; b <<= 1) {
    r = (r << 1) | (i & 1);
    i >>= 1;
  }
  return r;
}

// In-register complex FFT, N a power of two <= 32.  SIGN = -1: forward (e^{-2 pi i jk/N}),
// +1: inverse (unnormalised).  Everything is compile-time indexed after unrolling.
template <int N, int SIGN>
__device__ __forceinline__ void fft_reg(float (&re)[N], float (&im)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int j = bitrev(i, N);
    if (j > i) {
      float t = re[i]; re[i] = re[j]; re[j] = t;
      t = im[i]; im[i] = im[j]; im[j] = t;
    }
  }
#pragma unroll
  for (int len = 2; len <= N; len <<= 1) {
#pragma unroll
    for (int i = 0; i < N; i += len) {
#pragma unroll
      for (int k = 0; k < len / 2; ++k) {
        const float wr = tw_cos(k * (32 / len));
        const float wi = SIGN * tw_sin(k * (32 / len));
        const int a = i + k, b = i + k + len / 2;
        const float xr = re[b] * wr - im[b] * wi;
        const float xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] = re[a] + xr; im[a] = im[a] + xi;
      }
    }
  }
}

template <int H>
struct Cfg {
  static constexpr int CH = (H >= 32) ? 16 : 32;   // channels per CTA
  static constexpr int NTASK = NT / CH;            // concurrent 1-D transforms per channel
  static constexpr int KH = H / 2 + 1;
};

// ------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(NT) afno_fft_fwd_kernel(const float* __restrict__ a, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, int E, int bs, int km1,
                                                          int km2, float* __restrict__ S, float wint) {
  constexpr int CH = Cfg<H>::CH, NTASK = Cfg<H>::NTASK, KH = Cfg<H>::KH, n = H * H;
  extern __shared__ __align__(16) float smem[];
  float* T_s = smem;                                        // [n][CH]
  float2* R_s = reinterpret_cast<float2*>(smem + n * CH);    // [H][KH][CH]

  const int tid = threadIdx.x, c = tid % CH, task0 = tid / CH;
  const int b = blockIdx.y, ch = blockIdx.x * CH + c;
  const bool live = ch < E;
  const float sc = live ? (scale ? scale[(int64_t)b * E + ch] : 1.f) : 0.f;
  const float sh = (live && shift) ? shift[(int64_t)b * E + ch] : 0.f;

  // phase 1: GroupNorm-1 applied on load
  const float* ap = a + (int64_t)b * n * E + ch;
  for (int pos = task0; pos < n; pos += NTASK) T_s[pos * CH + c] = live ? fmaf(ap[(int64_t)pos * E], sc, sh) : 0.f;
  __syncthreads();

  // phase 2: real row transforms, two rows per complex FFT
  for (int pr = task0; pr < H / 2; pr += NTASK) {
    float zr[H], zi[H];
#pragma unroll
    for (int q = 0; q < H; ++q) {
      zr[q] = T_s[((2 * pr) * H + q) * CH + c];
      zi[q] = T_s[((2 * pr + 1) * H + q) * CH + c];
    }
    fft_reg<H, -1>(zr, zi);
#pragma unroll
    for (int k = 0; k < KH; ++k) {
      if (k < km2) {
        const int kn = (H - k) % H;
        const float ar = 0.5f * (zr[k] + zr[kn]), ai = 0.5f * (zi[k] - zi[kn]);
        const float br = 0.5f * (zi[k] + zi[kn]), bi = -0.5f * (zr[k] - zr[kn]);
        R_s[((2 * pr) * KH + k) * CH + c] = make_float2(ar, ai);
        R_s[((2 * pr + 1) * KH + k) * CH + c] = make_float2(br, bi);
      }
    }
  }
  __syncthreads();

  // phase 3: complex column transforms, write the kept modes
  if (!live) return;
  const float norm = 1.0f / (float)H;  // ortho: 1/sqrt(H*W), H == W
  const int kap = ch / bs, j = ch % bs;
  for (int k2 = task0; k2 < km2; k2 += NTASK) {
    float zr[H], zi[H];
#pragma unroll
    for (int p = 0; p < H; ++p) {
      const float2 v = R_s[(p * KH + k2) * CH + c];
      zr[p] = v.x; zi[p] = v.y;
    }
    fft_reg<H, -1>(zr, zi);
#pragma unroll
    for (int k1 = 0; k1 < H; ++k1) {
      if (k1 < km1) {
        float* dst = S + (((int64_t)b * km1 + k1) * km2 + k2) * (2 * E) + (int64_t)kap * 2 * bs + j;
        const float wk = (k2 > 0 && k2 < H / 2) ? norm * wint : norm;
        dst[0] = zr[k1] * wk;
        dst[bs] = zi[k1] * wk;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(NT) afno_fft_inv_kernel(const float* __restrict__ O2, const float* __restrict__ a,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          int E, int bs, int km1, int km2, float* __restrict__ f,
                                                          double* __restrict__ stats, int groups, float wint) {
  constexpr int CH = Cfg<H>::CH, NTASK = Cfg<H>::NTASK, KH = Cfg<H>::KH, n = H * H;
  extern __shared__ __align__(16) float smem[];
  float2* Z_s = reinterpret_cast<float2*>(smem);             // [H][KH][CH]
  double* red = reinterpret_cast<double*>(smem);             // reused for the statistics (after a sync)

  const int tid = threadIdx.x, c = tid % CH, task0 = tid / CH;
  const int b = blockIdx.y, ch = blockIdx.x * CH + c;
  const bool live = ch < E;
  const int kap = live ? ch / bs : 0, j = live ? ch % bs : 0;

  // phase 1: gather the kept modes of this channel chunk (zero elsewhere)
  for (int idx = task0; idx < H * KH; idx += NTASK) {
    const int k1 = idx / KH, k2 = idx % KH;
    float2 v = make_float2(0.f, 0.f);
    if (live && k1 < km1 && k2 < km2) {
      const float* src = O2 + (((int64_t)b * km1 + k1) * km2 + k2) * (2 * E) + (int64_t)kap * 2 * bs + j;
      const float wk = (k2 > 0 && k2 < H / 2) ? wint : 1.f;
      v = make_float2(src[0] * wk, src[bs] * wk);
    }
    Z_s[idx * CH + c] = v;
  }
  __syncthreads();

  // phase 2: inverse complex transform along k1 -> p, in place per column
  for (int k2 = task0; k2 < km2; k2 += NTASK) {
    float zr[H], zi[H];
#pragma unroll
    for (int k1 = 0; k1 < H; ++k1) {
      const float2 v = Z_s[(k1 * KH + k2) * CH + c];
      zr[k1] = v.x; zi[k1] = v.y;
    }
    fft_reg<H, +1>(zr, zi);
#pragma unroll
    for (int p = 0; p < H; ++p) Z_s[(p * KH + k2) * CH + c] = make_float2(zr[p], zi[p]);
  }
  __syncthreads();

  // phase 3: c2r along k2 -> q for two rows at once, + skip of the normalised input
  const float norm = 1.0f / (float)H;
  const float sc = live ? (scale ? scale[(int64_t)b * E + ch] : 1.f) : 0.f;
  const float sh = (live && shift) ? shift[(int64_t)b * E + ch] : 0.f;
  double s1 = 0.0, s2 = 0.0;
  for (int pr = task0; pr < H / 2; pr += NTASK) {
    float zr[H], zi[H];
#pragma unroll
    for (int k = 0; k < KH; ++k) {
      float2 A = make_float2(0.f, 0.f), Bv = make_float2(0.f, 0.f);
      if (k < km2) {
        A = Z_s[((2 * pr) * KH + k) * CH + c];
        Bv = Z_s[((2 * pr + 1) * KH + k) * CH + c];
      }
      if (k == 0 || k == H / 2) {  // c2r ignores Im of the DC and Nyquist columns
        zr[k] = A.x; zi[k] = Bv.x;
      } else {
        zr[k] = A.x - Bv.y; zi[k] = A.y + Bv.x;
        zr[H - k] = A.x + Bv.y; zi[H - k] = -A.y + Bv.x;
      }
    }
    fft_reg<H, +1>(zr, zi);
    if (live) {
      const int64_t base0 = ((int64_t)b * n + (2 * pr) * H) * E + ch;
      const int64_t base1 = base0 + (int64_t)H * E;
#pragma unroll
      for (int q = 0; q < H; ++q) {
        const float v0 = fmaf(zr[q], norm, a ? fmaf(a[base0 + (int64_t)q * E], sc, sh) : 0.f);
        const float v1 = fmaf(zi[q], norm, a ? fmaf(a[base1 + (int64_t)q * E], sc, sh) : 0.f);
        f[base0 + (int64_t)q * E] = v0;
        f[base1 + (int64_t)q * E] = v1;
        s1 += (double)v0 + (double)v1;
        s2 += (double)v0 * v0 + (double)v1 * v1;
      }
    }
  }
  if (stats == nullptr) return;

  // GroupNorm-2 statistics of f: reduce over the CTA per channel, then per group
  __syncthreads();
  red[tid] = s1;
  red[NT + tid] = s2;
  __syncthreads();
  if (tid < CH && live) {
    double t1 = 0.0, t2 = 0.0;
    for (int k = 0; k < NTASK; ++k) {
      t1 += red[k * CH + tid];
      t2 += red[NT + k * CH + tid];
    }
    const int gs = E / groups;
    double* dst = stats + ((int64_t)b * groups + ch / gs) * 2;
    atomicAdd(dst, t1);
    atomicAdd(dst + 1, t2);
  }
}

template <int H>
int launch_fwd(const float* a, const float* scale, const float* shift, int B, int E, int nb, int km1, int km2,
               float* S, float wint, cudaStream_t st) {
  constexpr int CH = Cfg<H>::CH, KH = Cfg<H>::KH;
  const size_t smem = (size_t)H * H * CH * 4 + (size_t)H * KH * CH * 8;
  DPOT_CUDA(cudaFuncSetAttribute(afno_fft_fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(E, CH), (unsigned)B);
  afno_fft_fwd_kernel<H><<<grid, NT, smem, st>>>(a, scale, shift, E, E / nb, km1, km2, S, wint);
  DPOT_LAUNCH_CHECK("afno_fft_fwd_kernel");
  return 0;
}

template <int H>
int launch_inv(const float* O2, const float* a, const float* scale, const float* shift, int B, int E, int nb,
               int km1, int km2, float* f, double* stats, int groups, float wint, cudaStream_t st) {
  constexpr int CH = Cfg<H>::CH, KH = Cfg<H>::KH;
  size_t smem = (size_t)H * KH * CH * 8;
  if (smem < (size_t)2 * NT * 8) smem = (size_t)2 * NT * 8;
  DPOT_CUDA(cudaFuncSetAttribute(afno_fft_inv_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(E, CH), (unsigned)B);
  afno_fft_inv_kernel<H><<<grid, NT, smem, st>>>(O2, a, scale, shift, E, E / nb, km1, km2, f, stats, groups, wint);
  DPOT_LAUNCH_CHECK("afno_fft_inv_kernel");
  return 0;
}

int check_common(int B, int h, int E, int nb, int km1, int km2) {
  DPOT_REQUIRE(B > 0 && E > 0 && nb > 0 && E % nb == 0, DPOT_E_BADARG, "afno_fft: bad B/E/nb (%d,%d,%d)", B, E, nb);
  DPOT_REQUIRE(h == 2 || h == 4 || h == 8 || h == 16 || h == 32, DPOT_E_UNSUPPORTED,
               "afno_fft: latent grid h=%d unsupported (power of two in [2,32] required)", h);
  DPOT_REQUIRE(km1 >= 1 && km1 <= h && km2 >= 1 && km2 <= h / 2 + 1, DPOT_E_BADARG, "afno_fft: bad kept modes (%d,%d)", km1, km2);
  DPOT_REQUIRE(B <= 65535, DPOT_E_BADARG, "afno_fft: B too large");
  return 0;
}

}  // namespace
}  // namespace dpot

using namespace dpot;

extern "C" int dpot_afno_fft_fwd(const float* a, const float* scale, const float* shift, int32_t B, int32_t h,
                                 int32_t E, int32_t nb, int32_t km1, int32_t km2, float* S, float interior_weight,
                                 void* stream) {
  DPOT_REQUIRE(a && S && ((scale == nullptr) == (shift == nullptr)), DPOT_E_BADARG, "dpot_afno_fft_fwd: null pointer");
  DPOT_CALL(check_common(B, h, E, nb, km1, km2));
  cudaStream_t st = as_stream(stream);
  switch (h) {
    case 2: return launch_fwd<2>(a, scale, shift, B, E, nb, km1, km2, S, interior_weight, st);
    case 4: return launch_fwd<4>(a, scale, shift, B, E, nb, km1, km2, S, interior_weight, st);
    case 8: return launch_fwd<8>(a, scale, shift, B, E, nb, km1, km2, S, interior_weight, st);
    case 16: return launch_fwd<16>(a, scale, shift, B, E, nb, km1, km2, S, interior_weight, st);
    default: return launch_fwd<32>(a, scale, shift, B, E, nb, km1, km2, S, interior_weight, st);
  }
}

extern "C" int dpot_afno_fft_inv(const float* O2, const float* a, const float* scale, const float* shift, int32_t B,
                                 int32_t h, int32_t E, int32_t nb, int32_t km1, int32_t km2, float* f,
                                 double* stats_out, int32_t groups, float interior_weight, void* stream) {
  DPOT_REQUIRE(O2 && f && ((scale == nullptr) == (shift == nullptr)), DPOT_E_BADARG, "dpot_afno_fft_inv: null pointer");
  DPOT_CALL(check_common(B, h, E, nb, km1, km2));
  DPOT_REQUIRE(!stats_out || (groups > 0 && E % groups == 0), DPOT_E_BADARG, "dpot_afno_fft_inv: bad groups");
  cudaStream_t st = as_stream(stream);
  switch (h) {
    case 2: return launch_inv<2>(O2, a, scale, shift, B, E, nb, km1, km2, f, stats_out, groups, interior_weight, st);
    case 4: return launch_inv<4>(O2, a, scale, shift, B, E, nb, km1, km2, f, stats_out, groups, interior_weight, st);
    case 8: return launch_inv<8>(O2, a, scale, shift, B, E, nb, km1, km2, f, stats_out, groups, interior_weight, st);
    case 16: return launch_inv<16>(O2, a, scale, shift, B, E, nb, km1, km2, f, stats_out, groups, interior_weight, st);
    default: return launch_inv<32>(O2, a, scale, shift, B, E, nb, km1, km2, f, stats_out, groups, interior_weight, st);
  }
}
