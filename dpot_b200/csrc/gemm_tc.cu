// tcgen05 / TMEM / TMA GEMM engine with fp32-faithful numerics (3xTF32 split), DPOT_GEMM_TC.
//
//   C[m,n] = epilogue( sum_k A'[m,k] W[n,k] )            (contract: include/dpot_b200.h)
//
// Formulation.  The weight tile is the UMMA "A" operand (M = 128 output channels n) and the
// activation tile the UMMA "B" operand (N = BA <= 256 tokens m), both K-major in shared memory,
// so the accumulator lives TRANSPOSED in TMEM: lane = n, column = m.  An epilogue warp then
// holds 32 consecutive output channels of one token per register index, i.e. every global
// store / residual / row-bias access is one coalesced 128 B line.
//
// fp32 parity on tensor cores, two ingredients (both measured on B200):
//  (1) 3xTF32: x = hi + lo with hi = rna_tf32(x), lo = x - hi (exact);
//        D += W_hi*X_hi + W_lo*X_hi + W_hi*X_lo      (the lo*lo term is < 2^-22)
//      Raw fp32 tiles arrive by TMA (128B-swizzled); four converter warps split them IN PLACE
//      (hi overwrites the raw tile, lo goes to a twin buffer at the same swizzled offsets) and apply
//      the optional per-(sample,k) affine of the A operand (GroupNorm-apply fused into the load).
//  (2) short accumulation chains: the tensor core truncates (round-toward-zero) on every
//      accumulate into TMEM, a bias of ~2e-8 per MMA that grows linearly with the chain (7e-6 at
//      K=1024).  The MMA issuer therefore accumulates only `flush` k-blocks in TMEM, ping-ponging
//      two 256-column buffers, and the epilogue warps add each partial sum into fp32 REGISTER
//      accumulators with round-to-nearest.
//
// Warp roles (512 threads, 1 CTA / SM, persistent over tiles; setmaxnreg moves registers to the
// epilogue warpgroups):
//   warp 0      TMA producer            warp 1      MMA issuer (one elected lane)
//   warp 2      TMEM allocator          warp 3      spare
//   warps 4-11  epilogue: TMEM partial sums -> register accumulators -> fused epilogue -> global
//   warps 12-15 converters (hi/lo split + affine)
#include "common.cuh"
#include "gemm_common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>

namespace dpot {
namespace {

constexpr int TN = 128;            // weight rows (output channels) per tile  = UMMA M
constexpr int TM_MAX = 256;        // activation rows (tokens) per tile       = UMMA N
constexpr int BK = 32;             // fp32 elements per k-block = one 128 B swizzle row
constexpr int STAGES = 2;
constexpr int NTHREADS = 512;
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8, CONV_WARP0 = 12, CONV_THREADS = 128;

constexpr uint32_t P_BYTES = TN * 128;          // 16 KB
constexpr uint32_t Q_BYTES = TM_MAX * 128;      // 32 KB
constexpr uint32_t OFF_P_HI = 0, OFF_P_LO = P_BYTES, OFF_Q_HI = 2 * P_BYTES, OFF_Q_LO = 2 * P_BYTES + Q_BYTES;
constexpr uint32_t STAGE_BYTES = 2 * P_BYTES + 2 * Q_BYTES;   // 96 KB
constexpr uint32_t BAR_OFF = STAGES * STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = BAR_OFF + 256 + 1024;        // + barriers + alignment slack

// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=n
__host__ __device__ constexpr uint32_t make_idesc(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// round-to-nearest (ties away) to TF32 = cvt.rna.tf32.f32 for finite inputs, in two integer ops
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

struct TcParams {
  GemmDev g;
  int BA;            // activation rows per tile (UMMA N), multiple of 16, <= 256
  int n_tiles, m_tiles, total_tiles, kblocks;
  int flush;         // k-blocks accumulated inside TMEM before the partial sum is flushed to registers
  long long* trace;  // optional clock64() event log of CTA 0 (debug / profiling), NULL = off
  int trunc;         // 1: hi = hardware truncation of the raw tile (no hi store), 0: explicit round-to-nearest hi
};

constexpr int TRACE_SLOTS = 64;    // events per role
__device__ __forceinline__ void trace_ev(const TcParams& P, int role, int& idx) {
  if (P.trace != nullptr && blockIdx.x == 0 && idx < TRACE_SLOTS) P.trace[role * TRACE_SLOTS + idx++] = clock64();
}

// Split one 128 B operand row (8 swizzled 16 B chunks) into hi (in place) and lo (twin tile).
// All 8 loads are issued before the first use; chunk order is rotated by lane so that a
// quarter-warp touches 8 distinct 16 B bank groups (row pitch is 128 B).
// TRUNC: the UMMA ignores the low 13 mantissa bits of an fp32 container read as TF32, so the raw tile already
// IS the hi operand (hi = trunc(x)); only lo = x - trunc(x) is written (halves the converter's smem stores).
template <bool AFFINE, bool TRUNC>
__device__ __forceinline__ void split_row(uint32_t row_hi, uint32_t lo_delta, int lane, uint32_t rsw,
                                          const float* __restrict__ sc, const float* __restrict__ sh) {
  float4 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = lds128(row_hi + (((uint32_t)(j + lane) & 7u) << 4));
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t cp = (uint32_t)(j + lane) & 7u;
    float4 x = v[j];
    if (AFFINE) {
      const uint32_t cl = cp ^ rsw;                 // logical k-chunk stored at physical chunk cp
      const float4 s = __ldg(reinterpret_cast<const float4*>(sc) + cl);
      const float4 h = __ldg(reinterpret_cast<const float4*>(sh) + cl);
      x.x = fmaf(x.x, s.x, h.x); x.y = fmaf(x.y, s.y, h.y); x.z = fmaf(x.z, s.z, h.z); x.w = fmaf(x.w, s.w, h.w);
    }
    float4 hi, lo;
    if (TRUNC) {
      hi.x = tf32_trunc(x.x); hi.y = tf32_trunc(x.y); hi.z = tf32_trunc(x.z); hi.w = tf32_trunc(x.w);
    } else {
      hi.x = tf32_rna(x.x); hi.y = tf32_rna(x.y); hi.z = tf32_rna(x.z); hi.w = tf32_rna(x.w);
    }
    lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
    if (!TRUNC) sts128(row_hi + (cp << 4), hi);
    else if (AFFINE) sts128(row_hi + (cp << 4), x);      // the transformed value must replace the raw one
    sts128(row_hi + lo_delta + (cp << 4), lo);
  }
}

// ACT_MODE: 0 = no activation, 1 = GELU (erf), 2 = runtime switch over the other activations
template <int ACT_MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapA, const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;     // 1024 B aligned (128B swizzle atom)
  // barrier map: [0,2) full  [2,4) converted  [4,6) empty  [6,8) tmem_full  [8,10) tmem_empty ; then tmem base slot
  const uint32_t bar0 = smem0 + BAR_OFF;
  auto BAR = [&](int which, int idx) -> uint32_t { return bar0 + 8u * (which * 2 + idx); };
  const uint32_t tmem_slot = bar0 + 8u * 10;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmDev& g = P.g;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapA);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(BAR(0, s), 1);
      mbar_init(BAR(1, s), CONV_THREADS);
      mbar_init(BAR(2, s), 1);
      mbar_init(BAR(3, s), 1);
      mbar_init(BAR(4, s), EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int KB = P.kblocks;
  const int F = P.flush;
  const uint32_t stage_tx = P_BYTES + (uint32_t)P.BA * 128u;

  if (warp < 4) {
    reg_dec<56>();
    if (warp == 0) {
      // ================================ TMA producer ================================
      if (elect_one()) {
        int s = 0; uint32_t ph = 0; int tr = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
          const int mt = tile % P.m_tiles; const int rest = tile / P.m_tiles;
          const int nt = rest % P.n_tiles; const int bz = rest / P.n_tiles;
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(BAR(2, s), ph ^ 1);
            trace_ev(P, 0, tr);
            mbar_expect_tx(BAR(0, s), stage_tx);
            const uint32_t sb = smem0 + (uint32_t)s * STAGE_BYTES;
            tma_load_3d(sb + OFF_P_HI, &mapW, BAR(0, s), kb * BK, nt * TN, bz);       // dims (k, n, batch)
            tma_load_3d(sb + OFF_Q_HI, &mapA, BAR(0, s), kb * BK, bz, mt * P.BA);     // dims (k, batch, m)
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ==================================
      if (elect_one()) {
        const uint32_t idesc = make_idesc((uint32_t)P.BA);
        int s = 0; uint32_t ph = 0; uint32_t gc = 0;       // gc: global chunk counter -> TMEM buffer ring
        int tr = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
          for (int kb0 = 0; kb0 < KB; kb0 += F, ++gc) {
            const uint32_t buf = gc & 1u, bph = (gc >> 1) & 1u;
            mbar_wait(BAR(4, buf), bph ^ 1);        // epilogue has drained this TMEM buffer
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * TM_MAX;
            const int kb1 = min(KB, kb0 + F);
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(BAR(1, s), ph);             // TMA landed AND converters finished the split
              tc_fence_after();
              trace_ev(P, 1, tr);
              const uint32_t sb = smem0 + (uint32_t)s * STAGE_BYTES;
#pragma unroll
              for (int k4 = 0; k4 < BK / 8; ++k4) {
                const uint64_t p_hi = make_smem_desc(sb + OFF_P_HI + k4 * 32);
                const uint64_t p_lo = make_smem_desc(sb + OFF_P_LO + k4 * 32);
                const uint64_t q_hi = make_smem_desc(sb + OFF_Q_HI + k4 * 32);
                const uint64_t q_lo = make_smem_desc(sb + OFF_Q_LO + k4 * 32);
                umma_tf32(d_tmem, p_hi, q_hi, idesc, (kb > kb0 || k4 > 0) ? 1u : 0u);
                umma_tf32(d_tmem, p_lo, q_hi, idesc, 1u);
                umma_tf32(d_tmem, p_hi, q_lo, idesc, 1u);
              }
              umma_commit(BAR(2, s));               // stage reusable once these MMAs retire
              if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit(BAR(3, buf));               // partial accumulator complete -> epilogue warps
          }
        }
      }
    }
  } else if (warp >= CONV_WARP0) {
    // ================================ converters ==================================
    reg_dec<72>();
    const int ct = threadIdx.x - CONV_WARP0 * 32;     // 0..127
    const bool affine = g.a_scale != nullptr;
    int s = 0; uint32_t ph = 0; int tr = 0, tr2 = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const int m0 = (tile % P.m_tiles) * P.BA;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(BAR(0, s), ph);
        if (ct == 0) trace_ev(P, 2, tr);
        const uint32_t sb = smem0 + (uint32_t)s * STAGE_BYTES;
        if (P.trunc) split_row<false, true>(sb + OFF_P_HI + (uint32_t)ct * 128u, P_BYTES, lane, 0, nullptr, nullptr);
        else split_row<false, false>(sb + OFF_P_HI + (uint32_t)ct * 128u, P_BYTES, lane, 0, nullptr, nullptr);   // weight row ct
#pragma unroll
        for (int half = 0; half < 2; ++half) {                                                        // token rows
          const int r = ct + half * 128;
          if (r < P.BA) {
            const uint32_t row = sb + OFF_Q_HI + (uint32_t)r * 128u;
            const int m = m0 + r;
            if (affine && m < g.M) {
              const int64_t tbl = (int64_t)(m / g.a_rps) * g.K + (int64_t)kb * BK;
              if (P.trunc) split_row<true, true>(row, Q_BYTES, lane, (uint32_t)r & 7u, g.a_scale + tbl, g.a_shift + tbl);
              else split_row<true, false>(row, Q_BYTES, lane, (uint32_t)r & 7u, g.a_scale + tbl, g.a_shift + tbl);
            } else {
              if (P.trunc) split_row<false, true>(row, Q_BYTES, lane, 0, nullptr, nullptr);
              else split_row<false, false>(row, Q_BYTES, lane, 0, nullptr, nullptr);
            }
          }
        }
        fence_proxy_async();                      // generic-proxy writes -> visible to the UMMA (async proxy)
        mbar_arrive(BAR(1, s));
        if (ct == 0) trace_ev(P, 3, tr2);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ================================ epilogue ====================================
    reg_inc<192>();
    const int quarter = warp & 3;                     // TMEM lane quarter this warp may read
    const int cpar = (warp - EPI_WARP0) >> 2;         // this warp owns the 32-column chunks 2*i + cpar
    const int nchunks = (P.BA + 31) / 32;
    uint32_t gc = 0;
    int tr = 0, tr2 = 0, tr3 = 0;
    const bool tracer = (warp == EPI_WARP0 && lane == 0);
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const int mt = tile % P.m_tiles; const int rest = tile / P.m_tiles;
      const int nt = rest % P.n_tiles; const int bz = rest / P.n_tiles;
      float acc[4][32];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[i][j] = 0.f;
      for (int kb0 = 0; kb0 < KB; kb0 += F, ++gc) {
        const uint32_t buf = gc & 1u, bph = (gc >> 1) & 1u;
        mbar_wait(BAR(3, buf), bph);
        tc_fence_after();
        if (tracer) trace_ev(P, 4, tr);
        const uint32_t t_row = tmem_base + buf * TM_MAX + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ci = 2 * i + cpar;
          if (ci < nchunks) {
            uint32_t r[32];
            tmem_ld32(t_row + (uint32_t)(ci * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[i][j] += __uint_as_float(r[j]);
          }
        }
        tc_fence_before();
        mbar_arrive(BAR(4, buf));
      }
      if (tracer) trace_ev(P, 5, tr2);
      // ---- final epilogue.  The accumulators are register arrays (static indices only).  A fully
      // unrolled 128-element epilogue is ~100 KB of straight-line code (instruction-cache thrash) and
      // parking the values in shared memory queues behind the converters' saturated LSU traffic, so the
      // 16 groups of 8 registers are selected by a uniform switch and consumed by ONE shared body that
      // keeps 8 independent elements in flight.  lanes = 32 consecutive output channels n -> every global
      // access is one coalesced 128 B line.
      const int n = nt * TN + quarter * 32 + lane;
      if (n < g.N) {
        float* __restrict__ C = g.C + (int64_t)bz * g.sC;
        const float bias_n = g.bias ? g.bias[(int64_t)bz * g.sBias + n] : 0.f;
        float st1 = 0.f, st2 = 0.f;        // GroupNorm statistics of what this thread stores
#pragma unroll 1
        for (int b = 0; b < 16; ++b) {
          const int c0 = (2 * (b >> 2) + cpar) * 32 + (b & 3) * 8;     // first tile column of this group
          const int m0 = mt * P.BA + c0;
          const int cnt = min(8, min(P.BA - c0, g.M - m0));
          if (cnt <= 0) continue;
          float t[8];
#define DPOT_GRP(B, I, J0) case B: _Pragma("unroll") for (int u = 0; u < 8; ++u) t[u] = acc[I][J0 + u]; break;
          switch (b) {
            DPOT_GRP(0, 0, 0) DPOT_GRP(1, 0, 8) DPOT_GRP(2, 0, 16) DPOT_GRP(3, 0, 24)
            DPOT_GRP(4, 1, 0) DPOT_GRP(5, 1, 8) DPOT_GRP(6, 1, 16) DPOT_GRP(7, 1, 24)
            DPOT_GRP(8, 2, 0) DPOT_GRP(9, 2, 8) DPOT_GRP(10, 2, 16) DPOT_GRP(11, 2, 24)
            DPOT_GRP(12, 3, 0) DPOT_GRP(13, 3, 8) DPOT_GRP(14, 3, 16)
            default: _Pragma("unroll") for (int u = 0; u < 8; ++u) t[u] = acc[3][24 + u]; break;
          }
#undef DPOT_GRP
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] += bias_n;
          if (g.rowbias) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (u < cnt) t[u] += g.rowbias[(int64_t)((m0 + u) % g.rb_period) * g.ldrb + n];
          }
          if (g.C_pre) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (u < cnt) g.C_pre[(int64_t)bz * g.sC + gemm_c_offset(g, m0 + u) + n] = t[u];
          }
          if (ACT_MODE == 1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = gelu_select(t[u]);
          } else if (ACT_MODE == 2) {
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = act_apply(t[u], g.act);
          }
          if (g.dact_src) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (u < cnt) t[u] *= act_grad(g.dact_src[(int64_t)bz * g.sC + (int64_t)(m0 + u) * g.ldc + n], g.dact);
          }
          if (g.c_scale) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (u < cnt) {
                const int64_t o = (int64_t)((m0 + u) / g.c_rps) * g.N + n;
                t[u] = fmaf(t[u], g.c_scale[o], g.c_shift[o]);
              }
          }
          if (g.residual) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (u < cnt) t[u] += g.residual[(int64_t)(m0 + u) * g.ldr + n];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (u < cnt) {
              C[gemm_c_offset(g, m0 + u) + n] = t[u];
              st1 += t[u];
              st2 = fmaf(t[u], t[u], st2);
            }
        }
        if (g.out_stats) {   // host guarantees: the tile lies in one sample, the warp's 32 channels in one group
          double d1 = (double)st1, d2 = (double)st2;
          const unsigned msk = __activemask();
          for (int o = 16; o > 0; o >>= 1) {
            d1 += __shfl_xor_sync(msk, d1, o);
            d2 += __shfl_xor_sync(msk, d2, o);
          }
          if (lane == 0) {
            const int smp = (mt * P.BA) / g.st_rps, grp = n / (g.N / g.st_groups);
            double* dst = g.out_stats + ((int64_t)smp * g.st_groups + grp) * 2;
            atomicAdd(dst, d1);
            atomicAdd(dst + 1, d2);
          }
        }
      }
      if (tracer) trace_ev(P, 6, tr3);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 3-D fp32 tensor map, 128B swizzle.  dims/strides fastest-first; strides in bytes for dims 1,2.
int encode_map_t(CUtensorMap* out, CUtensorMapDataType dt, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t s1, uint64_t s2, uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn enc = get_encode();
  DPOT_REQUIRE(enc != nullptr, DPOT_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1, s2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DPOT_REQUIRE(r == CUDA_SUCCESS, DPOT_E_BADARG,
               "cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u,%u)", (int)r,
               (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1,
               (unsigned long long)s2, b0, b1, b2);
  return 0;
}

int encode_map(CUtensorMap* out, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
               uint32_t b0, uint32_t b1, uint32_t b2) {
  return encode_map_t(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, d0, d1, d2, s1, s2, b0, b1, b2);
}

bool device_ok() {
  static int ok = -1;
  if (ok < 0) ok = dpot_device_supported();
  return ok == 1;
}

int g_flush = 2;
int g_trunc = 0;
long long* g_trace = nullptr;   // k-blocks (of 32) per in-TMEM accumulation chain; tunable for experiments

}  // namespace

// shared with the f16-split engine (gemm_tc16.cu)
bool tc_device_ok() { return device_ok() && get_encode() != nullptr; }
int tc_encode_map_f16(void* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                      uint32_t b0, uint32_t b1, uint32_t b2) {
  return encode_map_t(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, d0, d1, d2, s1, s2, b0, b1, b2);
}

bool gemm_tc_supports(const GemmDev& p, int batch) {
  if (!device_ok() || get_encode() == nullptr) return false;
  if (p.a_mode != DPOT_A_PLAIN || p.c_mode != DPOT_A_PLAIN) return false;
  if (p.K < BK || p.K % BK != 0) return false;
  if (p.M < 64 || p.N < 32) return false;                       // tiny problems stay on the SIMT engine
  if (p.lda % 4 || p.ldw % 4) return false;
  if ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.W)) % 16) return false;
  if (p.a_scale && ((reinterpret_cast<uintptr_t>(p.a_scale) | reinterpret_cast<uintptr_t>(p.a_shift)) % 16)) return false;
  if (batch > 1) {
    if (p.sA <= 0 || p.sW <= 0 || p.sA % 4 || p.sW % 4) return false;
    if (p.lda % p.sA != 0) return false;           // tensor-map strides ascending multiples: (k, batch, m)
    if (p.sW % p.ldw != 0 || p.sW < (int64_t)p.N * p.ldw) return false;   // (k, n, batch)
  }
  return true;
}

bool gemm_tc_fuses_stats(const GemmDev& p) {
  const int BA = p.M >= TM_MAX ? TM_MAX : (int)round_up(p.M, 16);
  return p.st_groups > 0 && p.st_rps > 0 && p.st_rps % BA == 0 && p.N % 32 == 0 && (p.N / p.st_groups) % 32 == 0;
}

int gemm_tc_launch(const GemmDev& p, int batch, cudaStream_t st) {
  TcParams P;
  P.g = p;
  P.BA = p.M >= TM_MAX ? TM_MAX : (int)round_up(p.M, 16);
  P.n_tiles = (int)ceil_div(p.N, TN);
  P.m_tiles = (int)ceil_div(p.M, P.BA);
  P.total_tiles = P.n_tiles * P.m_tiles * batch;
  P.kblocks = p.K / BK;
  P.flush = g_flush < 1 ? 1 : g_flush;
  P.trace = g_trace;
  P.trunc = g_trunc;

  alignas(64) CUtensorMap mapW, mapA;
  const uint64_t sWb = batch > 1 ? (uint64_t)p.sW * 4 : (uint64_t)p.ldw * 4 * (uint64_t)p.N;
  DPOT_CALL(encode_map(&mapW, p.W, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)batch, (uint64_t)p.ldw * 4, sWb, BK, TN, 1));
  const uint64_t sAb = batch > 1 ? (uint64_t)p.sA * 4 : (uint64_t)p.lda * 4;
  DPOT_CALL(encode_map(&mapA, p.A, (uint64_t)p.K, (uint64_t)batch, (uint64_t)p.M, sAb, (uint64_t)p.lda * 4, BK, 1,
                       (uint32_t)P.BA));

  const int sm_count = sm_count_cur();
  static DevOnce attr_set;
  if (attr_set.need()) {
    DPOT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    DPOT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    DPOT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set.done();
  }
  const int grid = P.total_tiles < sm_count ? P.total_tiles : sm_count;
  if (p.act == DPOT_ACT_NONE)
    gemm_tc_kernel<0><<<grid, NTHREADS, SMEM_BYTES, st>>>(mapW, mapA, P);
  else if (p.act == DPOT_ACT_GELU)
    gemm_tc_kernel<1><<<grid, NTHREADS, SMEM_BYTES, st>>>(mapW, mapA, P);
  else
    gemm_tc_kernel<2><<<grid, NTHREADS, SMEM_BYTES, st>>>(mapW, mapA, P);
  DPOT_LAUNCH_CHECK("gemm_tc_kernel");
  return 0;
}

}  // namespace dpot

extern "C" int dpot_tc_available(void) { return (dpot::device_ok() && dpot::get_encode() != nullptr) ? 1 : 0; }
// experiment knob: k-blocks (32 fp32 each) accumulated in TMEM between register flushes (default 4)
// debug: device buffer of 7*64 int64 receiving clock64() events of CTA 0 (NULL disables)
extern "C" void dpot_tc_set_trace(long long* dev_buf) { dpot::g_trace = dev_buf; }
// experiment knob: 1 = rely on the hardware's TF32 truncation for the hi operand (see split_row)
extern "C" int dpot_tc_set_trunc(int on) {
  const int old = dpot::g_trunc;
  if (on >= 0) dpot::g_trunc = on ? 1 : 0;
  return old;
}
extern "C" int dpot_tc_set_flush(int kblocks) {
  const int old = dpot::g_flush;
  if (kblocks >= 1) dpot::g_flush = kblocks;
  return old;
}
