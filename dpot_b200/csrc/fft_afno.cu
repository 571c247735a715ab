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
#include "gemm_common.cuh"
#include "fft_reg.cuh"

namespace dpot {
namespace {

constexpr int NT = 256;

// Compile-time geometry for the hot configurations (SPEC > 0: embed_dim / block size fixed, no mode truncation): the
// 192 stores and 64 loads of a thread then use immediate offsets instead of 64-bit address arithmetic -- ~30 % of the
// instructions of these issue-bound kernels.  SPEC 0 = everything at run time.
template <int SPEC> struct FftSpec { static constexpr int E = 0, bs = 0; };
template <> struct FftSpec<1> { static constexpr int E = 1024, bs = 128; };   // DPOT-S / M
template <> struct FftSpec<2> { static constexpr int E = 2048, bs = 256; };   // DPOT-H
template <> struct FftSpec<3> { static constexpr int E = 1536, bs = 96; };    // DPOT-L
template <> struct FftSpec<4> { static constexpr int E = 512, bs = 128; };    // DPOT-Ti
static inline int fft_spec_of(int E, int bs, int h, int km1, int km2) {
  if (km1 != h || km2 != h / 2 + 1) return 0;
  if (E == 1024 && bs == 128) return 1;
  if (E == 2048 && bs == 256) return 2;
  if (E == 1536 && bs == 96) return 3;
  if (E == 512 && bs == 128) return 4;
  return 0;
}

template <int H>
struct Cfg {
  static constexpr int CH = (H >= 32) ? 16 : 32;   // channels per CTA
  static constexpr int NTASK = NT / CH;            // concurrent 1-D transforms per channel
  static constexpr int KH = H / 2 + 1;
};

// ------------------------------------------------------------------------------------------
// Both kernels pack the DC and Nyquist columns (k2 = 0 and H/2) into ONE complex transform along k1: after
// the real row transform those two columns are real, and before the c2r only the real part of their
// k1-inverse is used (torch.fft.irfft2 ignores Im there), i.e. only the Hermitian part of the column.
// That makes H/2 column transforms and H/2 row-pair transforms per channel: with NT / CH = H/2 concurrent
// tasks (H = 16, 32) every thread owns exactly one transform per phase, and every global access is issued
// as a batch of independent coalesced 128 B lines (lanes = channels) straight from / to registers.
//
// OUT16: the spectrum is stored as split fp16 (DPOT_FMT_HL16; row = [hi 2E | lo 2E] halves), the operand format
// of the f16-split tensor-core engine that consumes it.
template <int H, bool OUT16, bool GN, int SPEC = 0>
__global__ void __launch_bounds__(NT) afno_fft_fwd_kernel(const float* __restrict__ a, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, int E_rt, int bs_rt, int km1_rt,
                                                          int km2_rt, float* __restrict__ S, float wint, const GnRef gn,
                                                          double* __restrict__ csum = nullptr) {
  constexpr int CH = Cfg<H>::CH, NTASK = Cfg<H>::NTASK, HC = (H >= 2) ? H / 2 : 1, n = H * H;
  const int E = SPEC ? FftSpec<SPEC>::E : E_rt, bs = SPEC ? FftSpec<SPEC>::bs : bs_rt;
  const int km1 = SPEC ? H : km1_rt, km2 = SPEC ? H / 2 + 1 : km2_rt;
  extern __shared__ __align__(16) float smem[];
  float2* R_s = reinterpret_cast<float2*>(smem);             // [H][HC][CH]; column 0 = (DC, Nyquist) real pair
  pdl_launch_dependents();
  pdl_wait();

  const int tid = threadIdx.x, c = tid % CH, task0 = tid / CH;
  const int b = blockIdx.y, ch = blockIdx.x * CH + c;
  const bool live = ch < E;
  float sc = live ? (scale ? scale[(int64_t)b * E + ch] : 1.f) : 0.f;
  float sh = (live && shift) ? shift[(int64_t)b * E + ch] : 0.f;
  if (GN && live) gn_affine_ref(gn, b, ch, sc, sh);      // GroupNorm-1 by reference (no finalize launch)

  // phase A: GroupNorm-1 applied on load; real row transforms, two rows per complex FFT
  for (int pr = task0; pr < H / 2; pr += NTASK) {
    float zr[H], zi[H];
    const float* ap = a + ((int64_t)b * n + (2 * pr) * H) * E + ch;
#pragma unroll
    for (int q = 0; q < H; ++q) {
      zr[q] = live ? __ldg(ap + (int64_t)q * E) : 0.f;
      zi[q] = live ? __ldg(ap + (int64_t)(H + q) * E) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < H; ++q) {
      zr[q] = fmaf(zr[q], sc, sh);
      zi[q] = fmaf(zi[q], sc, sh);
    }
    fft_reg<H, -1>(zr, zi);
    R_s[((2 * pr) * HC + 0) * CH + c] = make_float2(zr[0], zr[H / 2]);
    R_s[((2 * pr + 1) * HC + 0) * CH + c] = make_float2(zi[0], zi[H / 2]);
#pragma unroll
    for (int k = 1; k < H / 2; ++k) {
      const int kn = H - k;
      const float ar = 0.5f * (zr[k] + zr[kn]), ai = 0.5f * (zi[k] - zi[kn]);
      const float br = 0.5f * (zi[k] + zi[kn]), bi = -0.5f * (zr[k] - zr[kn]);
      R_s[((2 * pr) * HC + k) * CH + c] = make_float2(ar, ai);
      R_s[((2 * pr + 1) * HC + k) * CH + c] = make_float2(br, bi);
    }
  }
  __syncthreads();

  // phase B: complex column transforms, write the kept modes
  if (!live) return;
  const float norm = 1.0f / (float)H;  // ortho: 1/sqrt(H*W), H == W
  const int kap = ch / bs, j = ch % bs;
  float sum_re = 0.f, sum_im = 0.f;      // column sums of this channel's stored modes (csum != NULL: bias gradient)
  auto store = [&](int k1, int k2, float re, float im) {
    sum_re += re; sum_im += im;
    if (OUT16) {
      __half* dst = reinterpret_cast<__half*>(S) + (((int64_t)b * km1 + k1) * km2 + k2) * (4 * E) + (int64_t)kap * 2 * bs + j;
      __half hi, lo;
      hl_split(re, hi, lo);
      dst[0] = hi; dst[2 * E] = lo;
      hl_split(im, hi, lo);
      dst[bs] = hi; dst[2 * E + bs] = lo;
    } else {
      float* dst = S + (((int64_t)b * km1 + k1) * km2 + k2) * (2 * E) + (int64_t)kap * 2 * bs + j;
      dst[0] = re;
      dst[bs] = im;
    }
  };
  for (int kc = task0; kc < HC; kc += NTASK) {
    if (kc > 0 && kc >= km2) continue;
    float zr[H], zi[H];
#pragma unroll
    for (int p = 0; p < H; ++p) {
      const float2 v = R_s[(p * HC + kc) * CH + c];
      zr[p] = v.x; zi[p] = v.y;
    }
    fft_reg<H, -1>(zr, zi);
    if (kc > 0) {
      const float wk = norm * wint;
#pragma unroll
      for (int k1 = 0; k1 < H; ++k1)
        if (k1 < km1) store(k1, kc, zr[k1] * wk, zi[k1] * wk);
    } else {
      // F = FFT(col_0 + i col_{H/2}) -> X0[k1] = (F[k1] + conj F[-k1]) / 2, X_{H/2}[k1] = (F[k1] - conj F[-k1]) / 2i
      const bool nyq = (H / 2) < km2;
#pragma unroll
      for (int k1 = 0; k1 < H; ++k1) {
        if (k1 < km1) {
          const int kn = (H - k1) % H;
          store(k1, 0, 0.5f * (zr[k1] + zr[kn]) * norm, 0.5f * (zi[k1] - zi[kn]) * norm);
          if (nyq) store(k1, H / 2, 0.5f * (zi[k1] + zi[kn]) * norm, -0.5f * (zr[k1] - zr[kn]) * norm);
        }
      }
    }
  }
  if (csum) {
    atomicAdd(csum + (int64_t)kap * 2 * bs + j, (double)sum_re);
    atomicAdd(csum + (int64_t)kap * 2 * bs + bs + j, (double)sum_im);
  }
}

// ------------------------------------------------------------------------------------------
template <int H, bool GN, int SPEC = 0>
__global__ void __launch_bounds__(NT, H <= 16 ? 3 : 1) afno_fft_inv_kernel(const float* __restrict__ O2, const float* __restrict__ a,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          int E_rt, int bs_rt, int km1_rt, int km2_rt, float* __restrict__ f,
                                                          double* __restrict__ stats, int groups, float wint,
                                                          const GnRef gn) {
  constexpr int CH = Cfg<H>::CH, NTASK = Cfg<H>::NTASK, HC = (H >= 2) ? H / 2 : 1, n = H * H;
  const int E = SPEC ? FftSpec<SPEC>::E : E_rt, bs = SPEC ? FftSpec<SPEC>::bs : bs_rt;
  const int km1 = SPEC ? H : km1_rt, km2 = SPEC ? H / 2 + 1 : km2_rt;
  extern __shared__ __align__(16) float smem[];
  float2* Z_s = reinterpret_cast<float2*>(smem);             // [H][HC][CH]; column 0 = (x_0[p], x_{H/2}[p]) real pair
  double* red = reinterpret_cast<double*>(smem);             // reused for the statistics (after a sync)
  pdl_launch_dependents();
  pdl_wait();

  const int tid = threadIdx.x, c = tid % CH, task0 = tid / CH;
  const int b = blockIdx.y, ch = blockIdx.x * CH + c;
  const bool live = ch < E;
  const int kap = live ? ch / bs : 0, j = live ? ch % bs : 0;

  // phase A: kept modes (zero elsewhere) -> inverse complex transform along k1, one column per thread
  for (int kc = task0; kc < HC; kc += NTASK) {
    float zr[H], zi[H];
    const float* src = O2 + ((int64_t)b * km1 * km2 + kc) * (2 * E) + (int64_t)kap * 2 * bs + j;   // (k1 = 0, k2 = kc)
    const bool kin = live && kc < km2;
#pragma unroll
    for (int k1 = 0; k1 < H; ++k1) {
      const bool ok = kin && k1 < km1;
      zr[k1] = ok ? __ldg(src + (int64_t)k1 * km2 * (2 * E)) : 0.f;
      zi[k1] = ok ? __ldg(src + (int64_t)k1 * km2 * (2 * E) + bs) : 0.f;
    }
    if (kc > 0) {
#pragma unroll
      for (int k1 = 0; k1 < H; ++k1) { zr[k1] *= wint; zi[k1] *= wint; }
    } else {
      // Nyquist column (k2 = H/2), then pack the Hermitian parts: W = Herm(col_0) + i Herm(col_{H/2})
      float yr[H], yi[H];
      const bool nin = live && (H / 2) < km2;
      const float* srcn = src + (int64_t)(H / 2) * (2 * E);
#pragma unroll
      for (int k1 = 0; k1 < H; ++k1) {
        const bool ok = nin && k1 < km1;
        yr[k1] = ok ? __ldg(srcn + (int64_t)k1 * km2 * (2 * E)) : 0.f;
        yi[k1] = ok ? __ldg(srcn + (int64_t)k1 * km2 * (2 * E) + bs) : 0.f;
      }
      float wr[H], wi[H];
#pragma unroll
      for (int k1 = 0; k1 < H; ++k1) {
        const int kn = (H - k1) % H;
        const float s0r = 0.5f * (zr[k1] + zr[kn]), s0i = 0.5f * (zi[k1] - zi[kn]);
        const float s8r = 0.5f * (yr[k1] + yr[kn]), s8i = 0.5f * (yi[k1] - yi[kn]);
        wr[k1] = s0r - s8i;
        wi[k1] = s0i + s8r;
      }
#pragma unroll
      for (int k1 = 0; k1 < H; ++k1) { zr[k1] = wr[k1]; zi[k1] = wi[k1]; }
    }
    fft_reg<H, +1>(zr, zi);
#pragma unroll
    for (int p = 0; p < H; ++p) Z_s[(p * HC + kc) * CH + c] = make_float2(zr[p], zi[p]);
  }

  // skip term: the normalised block input, fetched before the barrier so that its latency overlaps phase A/C
  const float norm = 1.0f / (float)H;
  float sc = live ? (scale ? scale[(int64_t)b * E + ch] : 1.f) : 0.f;
  float sh = (live && shift) ? shift[(int64_t)b * E + ch] : 0.f;
  if (GN && live) gn_affine_ref(gn, b, ch, sc, sh);      // GroupNorm-1 (skip term) by reference
  float s1 = 0.f, s2 = 0.f;      // fp32 over this thread's 2H values; across threads / CTAs in double
  float k0[H], k1v[H];
  auto load_skip = [&](int pr) {
    const float* ap = a + ((int64_t)b * n + (2 * pr) * H) * E + ch;
#pragma unroll
    for (int q = 0; q < H; ++q) {
      k0[q] = (a && live) ? __ldg(ap + (int64_t)q * E) : 0.f;
      k1v[q] = (a && live) ? __ldg(ap + (int64_t)(H + q) * E) : 0.f;
    }
  };
  if (task0 < H / 2) load_skip(task0);
  __syncthreads();                         // Z_s complete
  for (int pr = task0; pr < H / 2; pr += NTASK) {
    const int64_t base0 = ((int64_t)b * n + (2 * pr) * H) * E + ch;
    if (pr != task0) load_skip(pr);
    // phase C: c2r along k2 -> q for two rows at once
    float zr[H], zi[H];
    {
      const float2 A = Z_s[((2 * pr) * HC + 0) * CH + c], Bv = Z_s[((2 * pr + 1) * HC + 0) * CH + c];
      zr[0] = A.x; zi[0] = Bv.x;                 // c2r ignores Im of the DC and Nyquist columns
      zr[H / 2] = A.y; zi[H / 2] = Bv.y;
    }
#pragma unroll
    for (int k = 1; k < H / 2; ++k) {
      const float2 A = Z_s[((2 * pr) * HC + k) * CH + c], Bv = Z_s[((2 * pr + 1) * HC + k) * CH + c];
      zr[k] = A.x - Bv.y; zi[k] = A.y + Bv.x;
      zr[H - k] = A.x + Bv.y; zi[H - k] = -A.y + Bv.x;
    }
    fft_reg<H, +1>(zr, zi);
    if (live) {
#pragma unroll
      for (int q = 0; q < H; ++q) {
        const float v0 = fmaf(zr[q], norm, a ? fmaf(k0[q], sc, sh) : 0.f);
        const float v1 = fmaf(zi[q], norm, a ? fmaf(k1v[q], sc, sh) : 0.f);
        f[base0 + (int64_t)q * E] = v0;
        f[base0 + (int64_t)(H + q) * E] = v1;
        s1 += v0 + v1;
        s2 = fmaf(v0, v0, fmaf(v1, v1, s2));
      }
    }
  }
  if (stats == nullptr) return;

  // GroupNorm-2 statistics of f: reduce over the CTA per channel, then per group
  __syncthreads();
  red[tid] = (double)s1;
  red[NT + tid] = (double)s2;
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

template <int H, bool OUT16 = false, bool GN = false, int SPEC = 0>
int launch_fwd(const float* a, const float* scale, const float* shift, int B, int E, int nb, int km1, int km2,
               float* S, float wint, cudaStream_t st, const GnRef gn = GnRef(), double* csum = nullptr) {
  constexpr int CH = Cfg<H>::CH, KH = Cfg<H>::KH;
  const size_t smem = (size_t)H * (H >= 2 ? H / 2 : 1) * CH * 8;
  (void)KH;
  DPOT_CUDA(cudaFuncSetAttribute(afno_fft_fwd_kernel<H, OUT16, GN, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(E, CH), (unsigned)B);
  DPOT_CUDA(launch_pdl(afno_fft_fwd_kernel<H, OUT16, GN, SPEC>, grid, dim3(NT), smem, st, a, scale, shift, E, E / nb, km1, km2, S, wint, gn, csum));
  DPOT_LAUNCH_CHECK("afno_fft_fwd_kernel");
  return 0;
}

template <int H, bool GN = false, int SPEC = 0>
int launch_inv(const float* O2, const float* a, const float* scale, const float* shift, int B, int E, int nb,
               int km1, int km2, float* f, double* stats, int groups, float wint, cudaStream_t st,
               const GnRef gn = GnRef()) {
  constexpr int CH = Cfg<H>::CH, KH = Cfg<H>::KH;
  size_t smem = (size_t)H * (H >= 2 ? H / 2 : 1) * CH * 8;
  (void)KH;
  if (smem < (size_t)2 * NT * 8) smem = (size_t)2 * NT * 8;
  DPOT_CUDA(cudaFuncSetAttribute(afno_fft_inv_kernel<H, GN, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(E, CH), (unsigned)B);
  DPOT_CUDA(launch_pdl(afno_fft_inv_kernel<H, GN, SPEC>, grid, dim3(NT), smem, st, O2, a, scale, shift, E, E / nb, km1, km2, f, stats, groups, wint, gn));
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

extern "C" int dpot_afno_fft_fwd16(const float* a, const float* scale, const float* shift, int32_t B, int32_t h,
                                   int32_t E, int32_t nb, int32_t km1, int32_t km2, void* S16, void* stream) {
  DPOT_REQUIRE(a && S16 && ((scale == nullptr) == (shift == nullptr)), DPOT_E_BADARG, "dpot_afno_fft_fwd16: null pointer");
  DPOT_CALL(check_common(B, h, E, nb, km1, km2));
  cudaStream_t st = as_stream(stream);
  float* S = reinterpret_cast<float*>(S16);
  switch (h) {
    case 2: return launch_fwd<2, true>(a, scale, shift, B, E, nb, km1, km2, S, 1.0f, st);
    case 4: return launch_fwd<4, true>(a, scale, shift, B, E, nb, km1, km2, S, 1.0f, st);
    case 8: return launch_fwd<8, true>(a, scale, shift, B, E, nb, km1, km2, S, 1.0f, st);
    case 16: return launch_fwd<16, true>(a, scale, shift, B, E, nb, km1, km2, S, 1.0f, st);
    default: return launch_fwd<32, true>(a, scale, shift, B, E, nb, km1, km2, S, 1.0f, st);
  }
}

// fwd16 without GroupNorm and with the interior-column weight: the adjoint of the inverse transform (weight 2) that the
// backward pass feeds straight into the f16-split contractions
extern "C" int dpot_afno_fft_fwd16w(const float* a, int32_t B, int32_t h, int32_t E, int32_t nb, int32_t km1, int32_t km2,
                                    void* S16, float interior_weight, double* colsum, void* stream) {
  DPOT_REQUIRE(a && S16, DPOT_E_BADARG, "dpot_afno_fft_fwd16w: null pointer");
  DPOT_CALL(check_common(B, h, E, nb, km1, km2));
  cudaStream_t st = as_stream(stream);
  float* S = reinterpret_cast<float*>(S16);
  switch (h) {
    case 2: return launch_fwd<2, true>(a, nullptr, nullptr, B, E, nb, km1, km2, S, interior_weight, st, GnRef(), colsum);
    case 4: return launch_fwd<4, true>(a, nullptr, nullptr, B, E, nb, km1, km2, S, interior_weight, st, GnRef(), colsum);
    case 8: return launch_fwd<8, true>(a, nullptr, nullptr, B, E, nb, km1, km2, S, interior_weight, st, GnRef(), colsum);
    case 16: return launch_fwd<16, true>(a, nullptr, nullptr, B, E, nb, km1, km2, S, interior_weight, st, GnRef(), colsum);
    default: return launch_fwd<32, true>(a, nullptr, nullptr, B, E, nb, km1, km2, S, interior_weight, st, GnRef(), colsum);
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

// ---- GroupNorm-by-reference variants (the f16-split inference pipeline, forward.cu): GroupNorm-1 comes as raw
// statistics [B, groups, 2] (double) + gamma/beta instead of finalised scale/shift tables
extern "C" int dpot_afno_fft_fwd16_gn(const float* a, const double* stats1, const float* gamma1, const float* beta1,
                                      int32_t groups, float eps, int32_t B, int32_t h, int32_t E, int32_t nb, int32_t km1,
                                      int32_t km2, void* S16, void* stream) {
  DPOT_REQUIRE(a && S16 && stats1 && gamma1 && beta1, DPOT_E_BADARG, "dpot_afno_fft_fwd16_gn: null pointer");
  DPOT_CALL(check_common(B, h, E, nb, km1, km2));
  DPOT_REQUIRE(groups > 0 && E % groups == 0, DPOT_E_BADARG, "dpot_afno_fft_fwd16_gn: bad groups");
  cudaStream_t st = as_stream(stream);
  float* S = reinterpret_cast<float*>(S16);
  const GnRef gn = make_gn_ref(stats1, gamma1, beta1, groups, eps, E, (int64_t)h * h);
  switch (h) {
    case 2: return launch_fwd<2, true, true>(a, nullptr, nullptr, B, E, nb, km1, km2, S, 1.0f, st, gn);
    case 4: return launch_fwd<4, true, true>(a, nullptr, nullptr, B, E, nb, km1, km2, S, 1.0f, st, gn);
    case 8: return launch_fwd<8, true, true>(a, nullptr, nullptr, B, E, nb, km1, km2, S, 1.0f, st, gn);
    case 16:
      switch (fft_spec_of(E, E / nb, h, km1, km2)) {
        case 1: return launch_fwd<16, true, true, 1>(a, nullptr, nullptr, B, E, nb, km1, km2, S, 1.0f, st, gn);
        case 2: return launch_fwd<16, true, true, 2>(a, nullptr, nullptr, B, E, nb, km1, km2, S, 1.0f, st, gn);
        case 3: return launch_fwd<16, true, true, 3>(a, nullptr, nullptr, B, E, nb, km1, km2, S, 1.0f, st, gn);
        case 4: return launch_fwd<16, true, true, 4>(a, nullptr, nullptr, B, E, nb, km1, km2, S, 1.0f, st, gn);
        default: return launch_fwd<16, true, true>(a, nullptr, nullptr, B, E, nb, km1, km2, S, 1.0f, st, gn);
      }
    default: return launch_fwd<32, true, true>(a, nullptr, nullptr, B, E, nb, km1, km2, S, 1.0f, st, gn);
  }
}

extern "C" int dpot_afno_fft_inv_gn(const float* O2, const float* a, const double* stats1, const float* gamma1,
                                    const float* beta1, int32_t groups, float eps, int32_t B, int32_t h, int32_t E,
                                    int32_t nb, int32_t km1, int32_t km2, float* f, double* stats_out, void* stream) {
  DPOT_REQUIRE(O2 && a && f && stats1 && gamma1 && beta1, DPOT_E_BADARG, "dpot_afno_fft_inv_gn: null pointer");
  DPOT_CALL(check_common(B, h, E, nb, km1, km2));
  DPOT_REQUIRE(groups > 0 && E % groups == 0, DPOT_E_BADARG, "dpot_afno_fft_inv_gn: bad groups");
  cudaStream_t st = as_stream(stream);
  const GnRef gn = make_gn_ref(stats1, gamma1, beta1, groups, eps, E, (int64_t)h * h);
  switch (h) {
    case 2: return launch_inv<2, true>(O2, a, nullptr, nullptr, B, E, nb, km1, km2, f, stats_out, groups, 1.0f, st, gn);
    case 4: return launch_inv<4, true>(O2, a, nullptr, nullptr, B, E, nb, km1, km2, f, stats_out, groups, 1.0f, st, gn);
    case 8: return launch_inv<8, true>(O2, a, nullptr, nullptr, B, E, nb, km1, km2, f, stats_out, groups, 1.0f, st, gn);
    case 16:
      switch (fft_spec_of(E, E / nb, h, km1, km2)) {
        case 1: return launch_inv<16, true, 1>(O2, a, nullptr, nullptr, B, E, nb, km1, km2, f, stats_out, groups, 1.0f, st, gn);
        case 2: return launch_inv<16, true, 2>(O2, a, nullptr, nullptr, B, E, nb, km1, km2, f, stats_out, groups, 1.0f, st, gn);
        case 3: return launch_inv<16, true, 3>(O2, a, nullptr, nullptr, B, E, nb, km1, km2, f, stats_out, groups, 1.0f, st, gn);
        case 4: return launch_inv<16, true, 4>(O2, a, nullptr, nullptr, B, E, nb, km1, km2, f, stats_out, groups, 1.0f, st, gn);
        default: return launch_inv<16, true>(O2, a, nullptr, nullptr, B, E, nb, km1, km2, f, stats_out, groups, 1.0f, st, gn);
      }
    default: return launch_inv<32, true>(O2, a, nullptr, nullptr, B, E, nb, km1, km2, f, stats_out, groups, 1.0f, st, gn);
  }
}
