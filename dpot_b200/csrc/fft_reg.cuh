// In-register FFT building blocks shared by the AFNO spectral kernels (fft_afno.cu, afno_fused.cu): every lane owns
// whole 1-D transforms (lanes = channels), radix-2 generic network for N <= 32 and a radix-4 x radix-4 16-point network.
#pragma once
#include <cuda_runtime.h>

namespace dpot {
namespace {

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

// ---- 16-point transform as radix-4 x radix-4 (n = 4 n1 + n2, k = k1 + 4 k2):
//   X[k1 + 4 k2] = sum_n2 W4^{n2 k2} * ( W16^{n2 k1} * sum_n1 x[4 n1 + n2] W4^{n1 k1} )
// 8 four-point DFTs (adds only) + 9 non-trivial twiddles = 160 instructions instead of ~320 for the generic radix-2
// network (whose multiplications by 0 / 1 the compiler may not fold).  Forward sign; the inverse is the same network on
// swapped (re, im) arrays.
__device__ __forceinline__ void dft4_fwd(float& ar, float& ai, float& br, float& bi, float& cr, float& ci, float& dr, float& di) {
  // in: a = x0, b = x1, c = x2, d = x3 ; out: a = X0, b = X1, c = X2, d = X3 (W4 = -i)
  const float t0r = ar + cr, t0i = ai + ci, t1r = ar - cr, t1i = ai - ci;
  const float t2r = br + dr, t2i = bi + di, t3r = br - dr, t3i = bi - di;
  ar = t0r + t2r; ai = t0i + t2i;
  cr = t0r - t2r; ci = t0i - t2i;
  br = t1r + t3i; bi = t1i - t3r;      // t1 - i t3
  dr = t1r - t3i; di = t1i + t3r;      // t1 + i t3
}
__device__ __forceinline__ void fft16_fwd(float (&re)[16], float (&im)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508978f, R2 = 0.70710678118654757f;
  // step 1: DFT4 over n1 for each n2 (elements n2, n2+4, n2+8, n2+12); result y[n2][k1] stored at index n2 + 4 k1
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2)
    dft4_fwd(re[n2], im[n2], re[n2 + 4], im[n2 + 4], re[n2 + 8], im[n2 + 8], re[n2 + 12], im[n2 + 12]);
  // step 2: twiddles W16^{n2 k1} = cos(2 pi m / 16) - i sin(2 pi m / 16), m = n2 k1 in {1,2,3, 2,4,6, 3,6,9}
  auto tw = [&](int idx, float c, float sn) {   // (x + i y)(c - i sn)
    const float x = re[idx], y = im[idx];
    re[idx] = fmaf(y, sn, x * c);
    im[idx] = fmaf(-x, sn, y * c);
  };
  auto tw45 = [&](int idx) {                    // m = 2: (1 - i)/sqrt 2
    const float x = re[idx], y = im[idx];
    re[idx] = (x + y) * R2; im[idx] = (y - x) * R2;
  };
  auto tw135 = [&](int idx) {                   // m = 6: (-1 - i)/sqrt 2
    const float x = re[idx], y = im[idx];
    re[idx] = (y - x) * R2; im[idx] = -(x + y) * R2;
  };
  auto tw90 = [&](int idx) {                    // m = 4: -i
    const float x = re[idx], y = im[idx];
    re[idx] = y; im[idx] = -x;
  };
  tw(1 + 4, C1, S1); tw45(1 + 8); tw(1 + 12, S1, C1);         // n2 = 1: m = 1, 2, 3
  tw45(2 + 4); tw90(2 + 8); tw135(2 + 12);                     // n2 = 2: m = 2, 4, 6
  tw(3 + 4, S1, C1); tw135(3 + 8); tw(3 + 12, -C1, -S1);       // n2 = 3: m = 3, 6, 9  (W16^9 = -cos(pi/8) + i sin(pi/8))
  // step 3: DFT4 over n2 for each k1 (elements 4 k1 + {0,1,2,3}); output X[k1 + 4 k2] lands at index 4 k1 + k2
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
    dft4_fwd(re[4 * k1], im[4 * k1], re[4 * k1 + 1], im[4 * k1 + 1], re[4 * k1 + 2], im[4 * k1 + 2], re[4 * k1 + 3], im[4 * k1 + 3]);
  // un-permute: X[k1 + 4 k2] is at index 4 k1 + k2 -> transpose of the 4 x 4 index grid (compile-time register renaming)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i + 1; j < 4; ++j) {
      float t = re[4 * i + j]; re[4 * i + j] = re[4 * j + i]; re[4 * j + i] = t;
      t = im[4 * i + j]; im[4 * i + j] = im[4 * j + i]; im[4 * j + i] = t;
    }
}
template <>
__device__ __forceinline__ void fft_reg<16, -1>(float (&re)[16], float (&im)[16]) { fft16_fwd(re, im); }
template <>
__device__ __forceinline__ void fft_reg<16, +1>(float (&re)[16], float (&im)[16]) { fft16_fwd(im, re); }   // conj-swap identity

}  // namespace
}  // namespace dpot
