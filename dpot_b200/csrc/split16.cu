// fp32 -> split fp16 (DPOT_FMT_HL16, include/dpot_b200.h) with an optional per-(sample, column) affine:
// weight packing for the f16-split tensor-core engine, and GroupNorm-apply fused with the split of
// the channel-MLP input (models/dpot.py:175-176).
#include "common.cuh"
#include "gemm_common.cuh"

namespace dpot {
namespace {

template <bool GN>
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int cols8,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        int rps, __half* __restrict__ dst, int64_t ldd, int64_t lo_off,
                                                        const GnRef gn) {
  const int64_t total = rows * cols8;
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols8;
    const int c = (int)(i % cols8) * 8;
    const float4 a = *reinterpret_cast<const float4*>(src + r * lds + c);
    const float4 b = *reinterpret_cast<const float4*>(src + r * lds + c + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (GN) {     // GroupNorm by reference: the 8 columns share one group (host-checked: group size % 8 == 0)
      const int bsm = (int)(r / rps);
      const double* sp = gn.stats + ((int64_t)bsm * gn.groups + c / gn.gs) * 2;
      const double mean = sp[0] * gn.inv_cnt;
      const double var = fma(-mean, mean, sp[1] * gn.inv_cnt);
      const float vv = fmaxf((float)var, 0.f) + gn.eps;
      float rstd = rsqrtf(vv);
      rstd = rstd * fmaf(-0.5f * vv * rstd, rstd, 1.5f);
      const float nm = -(float)mean;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gn.gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gn.gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(gn.beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(gn.beta + c + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float scu = rstd * gg[u];
        v[u] = fmaf(v[u], scu, fmaf(nm, scu, bb[u]));
      }
    } else if (scale) {
      const int64_t o = (r / rps) * (int64_t)(cols8 * 8) + c;
      const float4 s0 = *reinterpret_cast<const float4*>(scale + o), s1 = *reinterpret_cast<const float4*>(scale + o + 4);
      const float4 h0 = *reinterpret_cast<const float4*>(shift + o), h1 = *reinterpret_cast<const float4*>(shift + o + 4);
      const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = fmaf(v[u], s[u], h[u]);
    }
    alignas(16) __half hi[8];
    alignas(16) __half lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) hl_split(v[u], hi[u], lo[u]);
    *reinterpret_cast<uint4*>(dst + r * ldd + c) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(dst + r * ldd + c + lo_off) = *reinterpret_cast<const uint4*>(lo);
  }
}

// GroupNorm-by-reference split, tuned form: one thread owns ONE chunk of 8 columns for a run of consecutive rows of one
// sample, so the statistics, gamma and beta are fetched once and the row loop keeps 4 rows (8 x 16 B) in flight.
constexpr int SPLIT_ROWS = 8;
__global__ void __launch_bounds__(128) split_f16_gn_rows_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int cols8,
                                                                int rps, __half* __restrict__ dst, int64_t ldd, int64_t lo_off,
                                                                const GnRef gn) {
  pdl_launch_dependents();
  pdl_wait();
  const int cc = blockIdx.x * blockDim.x + threadIdx.x;       // column chunk
  if (cc >= cols8) return;
  const int c = cc * 8;
  const int64_t r0 = (int64_t)blockIdx.y * SPLIT_ROWS;         // rps % SPLIT_ROWS == 0: the run stays inside one sample
  const int bsm = (int)(r0 / rps);
  const double* sp = gn.stats + ((int64_t)bsm * gn.groups + c / gn.gs) * 2;
  const double mean = sp[0] * gn.inv_cnt;
  const double var = fma(-mean, mean, sp[1] * gn.inv_cnt);
  const float vv = fmaxf((float)var, 0.f) + gn.eps;
  float rstd = rsqrtf(vv);
  rstd = rstd * fmaf(-0.5f * vv * rstd, rstd, 1.5f);
  const float nm = -(float)mean;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gn.gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gn.gamma + c + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(gn.beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(gn.beta + c + 4));
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float sc[8], sh[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) { sc[u] = rstd * gg[u]; sh[u] = fmaf(nm, sc[u], bb[u]); }
  const int64_t rend = r0 + SPLIT_ROWS < rows ? r0 + SPLIT_ROWS : rows;
  for (int64_t r = r0; r < rend; r += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (r + j < rend) {
        a[j] = __ldcs(reinterpret_cast<const float4*>(src + (r + j) * lds + c));
        b[j] = __ldcs(reinterpret_cast<const float4*>(src + (r + j) * lds + c + 4));
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (r + j < rend) {
        const float v[8] = {a[j].x, a[j].y, a[j].z, a[j].w, b[j].x, b[j].y, b[j].z, b[j].w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float x0 = fmaf(v[2 * u], sc[2 * u], sh[2 * u]), x1 = fmaf(v[2 * u + 1], sc[2 * u + 1], sh[2 * u + 1]);
          const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
          const __half l0 = __float2half_rn((x0 - __half2float(h0)) * HL_SCALE), l1 = __float2half_rn((x1 - __half2float(h1)) * HL_SCALE);
          hi[u] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          lo[u] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        *reinterpret_cast<uint4*>(dst + (r + j) * ldd + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + (r + j) * ldd + c + lo_off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
  }
}

}  // namespace
}  // namespace dpot

using namespace dpot;

extern "C" int dpot_split_f16(const float* src, int64_t lds, int64_t rows, int32_t cols, const float* scale,
                              const float* shift, int32_t rows_per_sample, void* dst, int64_t ldd, int64_t lo_off,
                              void* stream) {
  DPOT_REQUIRE(src && dst && rows >= 0 && cols > 0, DPOT_E_BADARG, "dpot_split_f16: null pointer / bad shape");
  DPOT_REQUIRE(cols % 8 == 0 && lds % 4 == 0 && ldd % 8 == 0 && lo_off % 8 == 0, DPOT_E_ALIGN,
               "dpot_split_f16: cols, ldd, lo_off must be multiples of 8 (lds of 4)");
  DPOT_REQUIRE((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 16 == 0, DPOT_E_ALIGN,
               "dpot_split_f16: pointers must be 16-byte aligned");
  DPOT_REQUIRE((scale == nullptr) == (shift == nullptr) && (!scale || rows_per_sample > 0), DPOT_E_BADARG,
               "dpot_split_f16: scale/shift come together with rows_per_sample");
  if (rows == 0) return 0;
  const int64_t total = rows * (cols / 8);
  const unsigned grid = (unsigned)(ceil_div(total, 256) < 148 * 16 ? ceil_div(total, 256) : 148 * 16);
  DPOT_CUDA(launch_pdl(split_f16_kernel<false>, dim3(grid), dim3(256), 0, as_stream(stream), src, lds, rows, cols / 8, scale, shift,
                       rows_per_sample, reinterpret_cast<__half*>(dst), ldd, lo_off, GnRef()));
  DPOT_LAUNCH_CHECK("split_f16_kernel");
  return 0;
}

// GroupNorm-apply fused with the split, the normalisation given by reference (raw statistics + gamma/beta)
extern "C" int dpot_split_f16_gn(const float* src, int64_t lds, int64_t rows, int32_t cols, const double* stats,
                                 const float* gamma, const float* beta, int32_t groups, float eps, int32_t rows_per_sample,
                                 void* dst, int64_t ldd, int64_t lo_off, void* stream) {
  DPOT_REQUIRE(src && dst && stats && gamma && beta && rows >= 0 && cols > 0, DPOT_E_BADARG, "dpot_split_f16_gn: null pointer / bad shape");
  DPOT_REQUIRE(cols % 8 == 0 && lds % 4 == 0 && ldd % 8 == 0 && lo_off % 8 == 0, DPOT_E_ALIGN,
               "dpot_split_f16_gn: cols, ldd, lo_off must be multiples of 8 (lds of 4)");
  DPOT_REQUIRE((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(gamma) |
                reinterpret_cast<uintptr_t>(beta)) % 16 == 0, DPOT_E_ALIGN, "dpot_split_f16_gn: pointers must be 16-byte aligned");
  DPOT_REQUIRE(groups > 0 && cols % groups == 0 && (cols / groups) % 8 == 0 && rows_per_sample > 0, DPOT_E_BADARG,
               "dpot_split_f16_gn: group size must be a multiple of 8");
  if (rows == 0) return 0;
  const int64_t total = rows * (cols / 8);
  const unsigned grid = (unsigned)(ceil_div(total, 256) < 148 * 16 ? ceil_div(total, 256) : 148 * 16);
  const GnRef gn = make_gn_ref(stats, gamma, beta, groups, eps, cols, rows_per_sample);
  if (rows_per_sample % SPLIT_ROWS == 0 && rows % SPLIT_ROWS == 0 && rows / SPLIT_ROWS <= 65535) {
    dim3 g2((unsigned)ceil_div(cols / 8, 128), (unsigned)(rows / SPLIT_ROWS));
    DPOT_CUDA(launch_pdl(split_f16_gn_rows_kernel, g2, dim3(128), 0, as_stream(stream), src, lds, rows, cols / 8, rows_per_sample,
                         reinterpret_cast<__half*>(dst), ldd, lo_off, gn));
    DPOT_LAUNCH_CHECK("split_f16_gn_rows_kernel");
    return 0;
  }
  DPOT_CUDA(launch_pdl(split_f16_kernel<true>, dim3(grid), dim3(256), 0, as_stream(stream), src, lds, rows, cols / 8,
                       (const float*)nullptr, (const float*)nullptr, rows_per_sample, reinterpret_cast<__half*>(dst), ldd, lo_off, gn));
  DPOT_LAUNCH_CHECK("split_f16_kernel");
  return 0;
}
