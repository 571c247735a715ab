// Warp-level tensor-core helpers (ldmatrix, mma.sync m16n8k16 fp16 -> fp32) for the small-N kernels where a
// tcgen05 tile (UMMA M >= 128, operands staged by TMA) cannot be filled: PatchEmbed conv0 (N = 35) and the
// per-pixel output tail (32 -> 32 -> 4).  fp32-faithful through the split x = hi + lo/2048 (gemm_common.cuh).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace dpot {
namespace {

__device__ __forceinline__ uint32_t smem_u32_generic(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
  r[2] = r[3] = 0u;
}
// D[16x8] += A[16x16] (row) * B[16x8] (col), fp16 operands, fp32 accumulate
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two fp32 -> packed (hi, hi) and (lo, lo) half pairs of the split representation
__device__ __forceinline__ void hl_split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
  const __half l0 = __float2half_rn((x0 - __half2float(h0)) * 2048.0f), l1 = __float2half_rn((x1 - __half2float(h1)) * 2048.0f);
  hi = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
  lo = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
}

}  // namespace
}  // namespace dpot
