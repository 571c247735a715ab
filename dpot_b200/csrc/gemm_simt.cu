// fp32 CUDA-core GEMM engine (DPOT_GEMM_SIMT): the exact-fp32 path of dpot_gemm.
// It serves (a) every contraction on shapes the tcgen05 engine does not take (tiny test
// configs, K or N not tile friendly) and (b) as the numerics cross-check of the tensor-core
// engine.  Register-tiled 128x64x16, 256 threads, 8x4 outputs per thread, register prefetch
// of the next K-slab.  See include/dpot_b200.h for the epilogue contract.
#include "common.cuh"
#include "gemm_common.cuh"

namespace dpot {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4, NT = 256;
constexpr int AS_LD = BM + 4, WS_LD = BN + 4;

template <bool VEC>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const GemmDev p) {
  __shared__ __align__(16) float As[BK][AS_LD];
  __shared__ __align__(16) float Ws[BK][WS_LD];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int bz = blockIdx.z;
  const float* __restrict__ A = p.A + bz * p.sA;
  const float* __restrict__ W = p.W + bz * p.sW;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // register staging for the next slab
  float ra[8], rw[4];
  int64_t prow[8];     // im2col: field offset of (row, k=0) for the 8 rows this thread stages; -1 = out of range
  if (!VEC && p.a_mode == DPOT_A_PATCH) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = m0 + tid / BK + j * (NT / BK);
      prow[j] = (m < p.M) ? patch_offset(p, m, 0) : -1;
    }
  }

  auto load_slab = [&](int k0) {
    if (VEC) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = tid + j * NT;
        const int row = idx >> 2, kq = idx & 3;
        const int m = m0 + row, k = k0 + kq * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < p.M && k < p.K) {
          v = *reinterpret_cast<const float4*>(A + (int64_t)m * p.lda + k);
          if (p.a_scale) {
            const int64_t o = (int64_t)(m / p.a_rps) * p.K + k;
            const float4 s = *reinterpret_cast<const float4*>(p.a_scale + o);
            const float4 h = *reinterpret_cast<const float4*>(p.a_shift + o);
            v.x = fmaf(v.x, s.x, h.x); v.y = fmaf(v.y, s.y, h.y);
            v.z = fmaf(v.z, s.z, h.z); v.w = fmaf(v.w, s.w, h.w);
          }
        }
        ra[j * 4 + 0] = v.x; ra[j * 4 + 1] = v.y; ra[j * 4 + 2] = v.z; ra[j * 4 + 3] = v.w;
      }
      {
        const int row = tid >> 2, kq = tid & 3;
        const int n = n0 + row, k = k0 + kq * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < p.N && k < p.K) v = *reinterpret_cast<const float4*>(W + (int64_t)n * p.ldw + k);
        rw[0] = v.x; rw[1] = v.y; rw[2] = v.z; rw[3] = v.w;
      }
    } else {
      if (p.a_mode == DPOT_A_PATCH) {
        // im2col: the 8 rows of this thread are fixed (row bases precomputed), only k moves with the slab
        const int k = k0 + (tid % BK);
        const bool kin = k < p.K;
        const int c = k % p.pC, uv = k / p.pC;
        const int64_t koff = (((int64_t)(uv / p.pP) * p.pY + (uv % p.pP)) * p.pT) * p.pC + c;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = 0.f;
          if (kin && prow[j] >= 0) {
            v = A[prow[j] + koff];
            if (p.a_scale) {
              const int64_t o = (int64_t)((m0 + tid / BK + j * (NT / BK)) / p.a_rps) * p.K + k;
              v = fmaf(v, p.a_scale[o], p.a_shift[o]);
            }
          }
          ra[j] = v;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = tid + j * NT;
          const int row = idx / BK, kk = idx % BK;
          ra[j] = gemm_load_a(p, A, m0 + row, k0 + kk);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = tid + j * NT;
        const int row = idx / BK, kk = idx % BK;
        const int n = n0 + row, k = k0 + kk;
        rw[j] = (n < p.N && k < p.K) ? W[(int64_t)n * p.ldw + k] : 0.f;
      }
    }
  };
  auto store_slab = [&]() {
    if (VEC) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = tid + j * NT;
        const int row = idx >> 2, kq = idx & 3;
#pragma unroll
        for (int e = 0; e < 4; ++e) As[kq * 4 + e][row] = ra[j * 4 + e];
      }
      const int row = tid >> 2, kq = tid & 3;
#pragma unroll
      for (int e = 0; e < 4; ++e) Ws[kq * 4 + e][row] = rw[e];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = tid + j * NT;
        As[idx % BK][idx / BK] = ra[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = tid + j * NT;
        Ws[idx % BK][idx / BK] = rw[j];
      }
    }
  };

  const int nk = (p.K + BK - 1) / BK;
  load_slab(0);
  for (int kb = 0; kb < nk; ++kb) {
    store_slab();
    __syncthreads();
    if (kb + 1 < nk) load_slab((kb + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * TN]);
      const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float w[TN] = {w0.x, w0.y, w0.z, w0.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ---------------------------------------------------------------------------
  float* __restrict__ C = p.C + bz * p.sC;
  const float* bias = p.bias ? p.bias + bz * p.sBias : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= p.N) continue;
      gemm_epilogue_store(p, C, bias, m, n, acc[i][j], bz * p.sC);
    }
  }
}


// ---- skinny problems (M <= 64 rows: the classification head, models/dpot.py:394-395, runs on M = batch rows; AdaIN
// tables).  The problem is a weight STREAM (N x K fp32, e.g. 4 MB for the 1024 x 1024 head layers, usually from HBM)
// against a tiny A, so the kernel is built for bytes in flight: one warp per output column, the warp's lanes walk K in
// float4s and request a whole SK_KC-wide slice of the weight row (4 independent 512 B warp loads) -- and the NEXT
// slice -- before touching the current one; the CTA stages the matching [32 rows x SK_KC] slice of A in shared memory
// (prologue affine applied once), 32 register accumulators per lane, one shuffle reduction per row at the end.
// History: v1 re-read A from global memory in every warp (376 us for M = 16, N = K = 1024); v2 staged A but kept two
// scalar weight loads in flight per warp on 64 CTAs (~100 us per head at B = 32: 21 % of the whole rollout step).
constexpr int SK_MAXM = 64, SK_WARPS = 8, SK_KC = 512, SK_V = SK_KC / 128;   // SK_V float4 per lane and slice
template <int MT>   // rows handled per pass (32): M is covered in ceil(M/32) passes
__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_kernel(const GemmDev p) {
  extern __shared__ __align__(16) float A_s[];          // [MT][SK_KC]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n0 = blockIdx.x * SK_WARPS + warp;
  const bool live = n0 < p.N;
  const float* __restrict__ wrow = p.W + (int64_t)(live ? n0 : 0) * p.ldw;
  const bool vec = (p.ldw % 4 == 0) && (reinterpret_cast<uintptr_t>(p.W) % 16 == 0);
  const bool fastA = !p.a_scale && p.a_mode == DPOT_A_PLAIN && p.a_fmt == DPOT_FMT_F32 && (p.lda % 4 == 0) &&
                     (reinterpret_cast<uintptr_t>(p.A) % 16 == 0);
  auto load_w = [&](int kc, float4 (&w)[SK_V]) {
#pragma unroll
    for (int v = 0; v < SK_V; ++v) {
      const int k = kc + v * 128 + lane * 4;
      if (!live || k >= p.K) { w[v] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
      if (vec && k + 3 < p.K) w[v] = __ldg(reinterpret_cast<const float4*>(wrow + k));
      else {
        w[v].x = __ldg(wrow + k);
        w[v].y = k + 1 < p.K ? __ldg(wrow + k + 1) : 0.f;
        w[v].z = k + 2 < p.K ? __ldg(wrow + k + 2) : 0.f;
        w[v].w = k + 3 < p.K ? __ldg(wrow + k + 3) : 0.f;
      }
    }
  };
  for (int mb = 0; mb < p.M; mb += MT) {
    float acc[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) acc[i] = 0.f;
    float4 wc[SK_V], wn[SK_V];
    load_w(0, wc);
    for (int kc = 0; kc < p.K; kc += SK_KC) {
      if (kc + SK_KC < p.K) load_w(kc + SK_KC, wn);          // next slice in flight while this one is consumed
      __syncthreads();
      if (fastA && kc + SK_KC <= p.K) {       // plain contiguous rows: 16-byte copies (the staging was most of this kernel's time)
        for (int e = threadIdx.x; e < MT * (SK_KC / 4); e += SK_WARPS * 32) {
          const int i = e / (SK_KC / 4), k4 = e - i * (SK_KC / 4);
          const float4 v = (mb + i < p.M) ? __ldg(reinterpret_cast<const float4*>(p.A + (int64_t)(mb + i) * p.lda + kc) + k4)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
          reinterpret_cast<float4*>(A_s)[e] = v;
        }
      } else {
        for (int e = threadIdx.x; e < MT * SK_KC; e += SK_WARPS * 32) {
          const int i = e / SK_KC, k = e - i * SK_KC;
          A_s[e] = (mb + i < p.M && kc + k < p.K) ? gemm_load_a(p, p.A, mb + i, kc + k) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const float4* ar = reinterpret_cast<const float4*>(A_s + i * SK_KC) + lane;
        float a = acc[i];
#pragma unroll
        for (int v = 0; v < SK_V; ++v) {
          const float4 x = ar[v * 32];
          a = fmaf(x.x, wc[v].x, fmaf(x.y, wc[v].y, fmaf(x.z, wc[v].z, fmaf(x.w, wc[v].w, a))));
        }
        acc[i] = a;
      }
#pragma unroll
      for (int v = 0; v < SK_V; ++v) wc[v] = wn[v];
    }
#pragma unroll
    for (int i = 0; i < MT; ++i) acc[i] = warp_sum(acc[i]);
    // lane i finalises row mb+i (MT == 32: one row per lane)
    float mine = 0.f;
#pragma unroll
    for (int i = 0; i < MT; ++i) mine = (lane == i) ? acc[i] : mine;
    if (live && mb + lane < p.M) gemm_epilogue_store(p, p.C, p.bias, mb + lane, n0, mine);
  }
}

// M <= 32 rows and K <= 1024 (the classification head at every published width up to 1024): the WHOLE A panel
// (<= 128 KB) arrives in shared memory by bulk async copies -- one elected thread, one mbarrier, no per-thread staging
// loop (the slice-by-slice staging of the kernel above exposed the L2 latency twice per launch: 18-21 us per head layer)
// -- while every warp already has its entire weight row (K / 128 float4 per lane) in flight.
constexpr int SW_MAXK = 1024, SW_V = SW_MAXK / 128;
__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_whole_kernel(const GemmDev p) {
  extern __shared__ __align__(128) float A_s[];          // [32][K], then the mbarrier
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n0 = blockIdx.x * SK_WARPS + warp, K = p.K;
  const bool live = n0 < p.N;
  const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(A_s), bar = a_s + 32u * (uint32_t)K * 4u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t row_bytes = (uint32_t)K * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * (uint32_t)p.M) : "memory");
    for (int i = 0; i < p.M; ++i)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(a_s + (uint32_t)i * row_bytes), "l"(p.A + (int64_t)i * p.lda), "r"(row_bytes), "r"(bar) : "memory");
  }
  for (int e = p.M * (K / 4) + threadIdx.x; e < 32 * (K / 4); e += SK_WARPS * 32)      // rows M .. 31: zeros
    reinterpret_cast<float4*>(A_s)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* __restrict__ wrow = p.W + (int64_t)(live ? n0 : 0) * p.ldw;
  float4 w[SW_V];
#pragma unroll
  for (int v = 0; v < SW_V; ++v) {
    const int k = v * 128 + lane * 4;
    w[v] = (live && k < K) ? __ldg(reinterpret_cast<const float4*>(wrow + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar), "r"(0u) : "memory");
  }
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float4* ar = reinterpret_cast<const float4*>(A_s + i * K) + lane;
    float a = 0.f;
#pragma unroll
    for (int v = 0; v < SW_V; ++v)
      if (v * 128 < K) {
        const float4 x = (v * 128 + lane * 4 < K) ? ar[v * 32] : make_float4(0.f, 0.f, 0.f, 0.f);
        a = fmaf(x.x, w[v].x, fmaf(x.y, w[v].y, fmaf(x.z, w[v].z, fmaf(x.w, w[v].w, a))));
      }
    acc[i] = a;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = warp_sum(acc[i]);
  float mine = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) mine = (lane == i) ? acc[i] : mine;
  if (live && lane < p.M) gemm_epilogue_store(p, p.C, p.bias, lane, n0, mine);
}

}  // namespace

int gemm_simt_launch(const GemmDev& p, int batch, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(p.N, BN), (unsigned)ceil_div(p.M, BM), (unsigned)batch);
  DPOT_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, DPOT_E_BADARG, "dpot_gemm: grid too large (M=%d batch=%d)", p.M, batch);
  const bool aligned = (reinterpret_cast<uintptr_t>(p.A) % 16 == 0) && (reinterpret_cast<uintptr_t>(p.W) % 16 == 0) &&
                       (p.lda % 4 == 0) && (p.ldw % 4 == 0) && (p.K % 4 == 0) && (p.sA % 4 == 0) && (p.sW % 4 == 0) &&
                       (!p.a_scale || (reinterpret_cast<uintptr_t>(p.a_scale) % 16 == 0 &&
                                       reinterpret_cast<uintptr_t>(p.a_shift) % 16 == 0));
  if (p.M <= 32 && batch == 1 && p.a_mode == DPOT_A_PLAIN && p.a_fmt == DPOT_FMT_F32 && !p.a_scale && p.K <= SW_MAXK && p.K % 4 == 0 &&
      p.lda % 4 == 0 && p.ldw % 4 == 0 && reinterpret_cast<uintptr_t>(p.A) % 16 == 0 && reinterpret_cast<uintptr_t>(p.W) % 16 == 0) {
    const int smem = 32 * p.K * 4 + 16;
    static DevOnce attr;
    if (attr.need()) {
      DPOT_CUDA(cudaFuncSetAttribute(gemm_skinny_whole_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * SW_MAXK * 4 + 16));
      attr.done();
    }
    gemm_skinny_whole_kernel<<<(unsigned)ceil_div(p.N, SK_WARPS), SK_WARPS * 32, smem, st>>>(p);
    DPOT_LAUNCH_CHECK("gemm_skinny_whole_kernel");
    return 0;
  }
  if (p.M <= SK_MAXM && batch == 1 && p.a_mode == DPOT_A_PLAIN) {
    constexpr int smem = 32 * SK_KC * 4;
    static DevOnce attr;
    if (attr.need()) {
      DPOT_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr.done();
    }
    gemm_skinny_kernel<32><<<(unsigned)ceil_div(p.N, SK_WARPS), SK_WARPS * 32, smem, st>>>(p);
    DPOT_LAUNCH_CHECK("gemm_skinny_kernel");
    return 0;
  }
  if (p.a_mode == DPOT_A_PLAIN && aligned)
    gemm_simt_kernel<true><<<grid, NT, 0, st>>>(p);
  else
    gemm_simt_kernel<false><<<grid, NT, 0, st>>>(p);
  DPOT_LAUNCH_CHECK("gemm_simt_kernel");
  return 0;
}

}  // namespace dpot
