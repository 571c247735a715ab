// tcgen05 / TMEM / TMA GEMM engine on split-fp16 operands (DPOT_GEMM_TC16): fp32-faithful results at
// twice the TF32 MMA rate and half the shared-memory operand traffic of the 3xTF32 engine (gemm_tc.cu).
//
//   C[m,n] = epilogue( sum_k A[m,k] W[n,k] ),   A, W stored as DPOT_FMT_HL16 (include/dpot_b200.h):
//   x = hi + lo/2048, hi = fp16_rn(x), lo = fp16_rn((x - hi) * 2048).
//
//   sum_k A W = sum_k Ahi Whi  +  2^-11 * sum_k (Ahi Wlo + Alo Whi)  +  O(2^-22)
//               `---- D1 ----'          `----------- D2 -----------'
// Products of two fp16 are exact in the fp32 accumulator; D1 and D2 are separate TMEM accumulators
// (D2's accumulation error is scaled by 2^-11, so only D1's chain length matters: K/16 MMAs, measured
// whole-model effect 2e-6 rel-L2 at K = 1024).  The lo halves are pre-scaled by 2^11 so
// they keep a full fp16 significand at every magnitude (no subnormal loss for small activations).
//
// Operands arrive already split from their producers (FFT kernel, previous GEMM epilogue, weight
// packing): no converter warps, no extra bytes -- a split element is 4 bytes like the fp32 it replaces.
//
// Tile: weights are the UMMA A operand (M = 128 output channels n), activations the B operand
// (BA <= 128 tokens m), so the accumulator is TRANSPOSED in TMEM (lane = n, column = m) and a warp's
// 32 lanes always address 32 adjacent channels of one token (coalesced epilogue, no smem staging).
// Per 16-deep k-step two MMAs:   D[0,2BA)  += Whi * [Ahi ; Alo]^T      (N = 2*BA: D1 | D2 side by side)
//                                D[BA,2BA) += Wlo * Ahi^T              (N = BA)
// TMEM: 2 x 256 columns (double-buffered tiles), so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (384 threads, 1 CTA / SM, persistent): warp 0 TMA producer, warp 1 MMA issuer,
// warp 2 TMEM allocator, warps 4-11 epilogue (two per TMEM lane quarter, interleaved 8-column groups).
#include "common.cuh"
#include "gemm_common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

namespace dpot {
namespace {

constexpr int TN = 128;            // weight rows (output channels) per tile = UMMA M
constexpr int TA = 128;            // max activation rows (tokens) per tile
constexpr int BKH = 64;            // halves per k-block = one 128 B swizzle row
constexpr int STAGES = 3;
constexpr int NTHREADS = 384;
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8;

constexpr uint32_t W_BYTES = TN * 128;          // 16 KB per plane
constexpr uint32_t A_BYTES = TA * 128;          // 16 KB per plane (max)
constexpr uint32_t OFF_W_HI = 0, OFF_W_LO = W_BYTES, OFF_A = 2 * W_BYTES;   // A_hi at OFF_A, A_lo at OFF_A + BA*128
constexpr uint32_t STAGE_BYTES = 2 * W_BYTES + 2 * A_BYTES;   // 64 KB
constexpr uint32_t BAR_OFF = STAGES * STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = BAR_OFF + 256 + 1024;

// instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=n
__host__ __device__ constexpr uint32_t make_idesc16(uint32_t n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

struct Tc16Params {
  GemmDev g;
  int BA;            // activation rows per tile, multiple of 16, <= 128
  int n_tiles, m_tiles, total_tiles, kblocks;
};

// ACT_MODE: 0 = none, 1 = GELU (erf), 2 = runtime switch.  OUT16: store the result as DPOT_FMT_HL16.
// SIDE: the epilogue has side inputs (row-periodic bias, residual, per-sample affine); without them that code
// (and its registers) is compiled out -- the AFNO GEMMs (K = 2*bs) are epilogue-bound.
template <int ACT_MODE, bool OUT16, bool SIDE>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc16_kernel(const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWl,
                 const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                 const Tc16Params P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // barriers: [0,S) full  [S,2S) empty  [2S,2S+2) tmem_full  [2S+2,2S+4) tmem_empty ; then the TMEM base slot
  const uint32_t bar0 = smem0 + BAR_OFF;
  auto FULL = [&](int s) -> uint32_t { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) -> uint32_t { return bar0 + 8u * (STAGES + s); };
  auto TFULL = [&](int b) -> uint32_t { return bar0 + 8u * (2 * STAGES + b); };
  auto TEMPTY = [&](int b) -> uint32_t { return bar0 + 8u * (2 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmDev& g = P.g;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&mapWh); tma_prefetch_desc(&mapWl); tma_prefetch_desc(&mapAh); tma_prefetch_desc(&mapAl);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(TFULL(b), 1); mbar_init(TEMPTY(b), EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int KB = P.kblocks, BA = P.BA;
  const uint32_t a_plane = (uint32_t)BA * 128u;
  const uint32_t stage_tx = 2u * W_BYTES + 2u * a_plane;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const int mt = tile % P.m_tiles; const int rest = tile / P.m_tiles;
        const int nt = rest % P.n_tiles; const int bz = rest / P.n_tiles;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(EMPTY(s), ph ^ 1);
          mbar_expect_tx(FULL(s), stage_tx);
          const uint32_t sb = smem0 + (uint32_t)s * STAGE_BYTES;
          tma_load_3d(sb + OFF_W_HI, &mapWh, FULL(s), kb * BKH, nt * TN, bz);        // dims (k, n, batch)
          tma_load_3d(sb + OFF_W_LO, &mapWl, FULL(s), kb * BKH, nt * TN, bz);
          tma_load_3d(sb + OFF_A, &mapAh, FULL(s), kb * BKH, bz, mt * BA);           // dims (k, batch, m)
          tma_load_3d(sb + OFF_A + a_plane, &mapAl, FULL(s), kb * BKH, bz, mt * BA);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      const uint32_t idesc_a = make_idesc16((uint32_t)(2 * BA));
      const uint32_t idesc_b = make_idesc16((uint32_t)BA);
      int s = 0; uint32_t ph = 0; uint32_t tc = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++tc) {
        const uint32_t buf = tc & 1u, bph = (tc >> 1) & 1u;
        mbar_wait(TEMPTY(buf), bph ^ 1);          // epilogue has drained this TMEM buffer
        tc_fence_after();
        const uint32_t d1 = tmem_base + buf * 256u;
        const uint32_t d2 = d1 + (uint32_t)BA;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(FULL(s), ph);
          tc_fence_after();
          const uint32_t sb = smem0 + (uint32_t)s * STAGE_BYTES;
#pragma unroll
          for (int k4 = 0; k4 < BKH / 16; ++k4) {
            const uint64_t w_hi = make_smem_desc(sb + OFF_W_HI + k4 * 32);
            const uint64_t w_lo = make_smem_desc(sb + OFF_W_LO + k4 * 32);
            const uint64_t a_hl = make_smem_desc(sb + OFF_A + k4 * 32);      // rows [0,BA) = hi, [BA,2BA) = lo
            umma_f16(d1, w_hi, a_hl, idesc_a, (kb > 0 || k4 > 0) ? 1u : 0u);
            umma_f16(d2, w_lo, a_hl, idesc_b, 1u);
          }
          umma_commit(EMPTY(s));
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(TFULL(buf));
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ================================ epilogue ====================================
    const int quarter = warp & 3;                     // TMEM lane quarter this warp may read
    const int half = (warp - EPI_WARP0) >> 2;         // interleaved 8-column groups: g = half, half+2, ...
    const int ngroups = BA / 8;
    __half* const Ch_base = reinterpret_cast<__half*>(g.C);
    uint32_t tc = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++tc) {
      const int mt = tile % P.m_tiles; const int rest = tile / P.m_tiles;
      const int nt = rest % P.n_tiles; const int bz = rest / P.n_tiles;
      const uint32_t buf = tc & 1u, bph = (tc >> 1) & 1u;
      const int n = nt * TN + quarter * 32 + lane;
      const bool nok = n < g.N;
      const float bias_n = (g.bias && nok) ? g.bias[(int64_t)bz * g.sBias + n] : 0.f;
      float st1 = 0.f, st2 = 0.f;
      // side inputs (row-periodic bias, residual) of a group are fetched one group ahead: their global-load
      // latency hides behind the arithmetic of the previous group instead of serialising the epilogue
      float nrb[8], nrs[8];
      auto load_side = [&](int gi) {
        if (!SIDE) return;
        const int m0 = mt * BA + gi * 8;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool ok = nok && (m0 + u) < g.M;
          nrb[u] = (g.rowbias && ok) ? __ldg(g.rowbias + (int64_t)((m0 + u) % g.rb_period) * g.ldrb + n) : 0.f;
          nrs[u] = (g.residual && ok) ? __ldg(g.residual + (int64_t)(m0 + u) * g.ldr + n) : 0.f;
        }
      };
      if (half < ngroups) load_side(half);
      mbar_wait(TFULL(buf), bph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + buf * 256u + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int gi = half; gi < ngroups; gi += 2) {
        const int c0 = gi * 8;
        uint32_t r1[8], r2[8];
        tmem_ld8(t_row + (uint32_t)c0, r1);
        tmem_ld8(t_row + (uint32_t)(BA + c0), r2);
        float rb[8], rs[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { rb[u] = SIDE ? nrb[u] : 0.f; rs[u] = SIDE ? nrs[u] : 0.f; }
        if (gi + 2 < ngroups) load_side(gi + 2);
        tmem_ld_wait();
        const int m0 = mt * BA + c0;
        const int cnt = min(8, g.M - m0);
        if (cnt <= 0 || !nok) continue;
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = fmaf(__uint_as_float(r2[u]), HL_INV, __uint_as_float(r1[u])) + bias_n + rb[u];
        if (ACT_MODE == 1) {
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] = gelu_select(t[u]);
        } else if (ACT_MODE == 2) {
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] = act_apply(t[u], g.act);
        }
        if (SIDE && g.c_scale) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (u < cnt) {
              const int64_t o = (int64_t)((m0 + u) / g.c_rps) * g.N + n;
              t[u] = fmaf(t[u], g.c_scale[o], g.c_shift[o]);
            }
        }
        if (SIDE) {
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] += rs[u];
        }
        if (OUT16) {
          __half* __restrict__ cp = Ch_base + (int64_t)bz * g.sC + (int64_t)m0 * g.ldc + n;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (u < cnt) {
              __half hi, lo;
              hl_split(t[u], hi, lo);
              cp[(int64_t)u * g.ldc] = hi;
              cp[(int64_t)u * g.ldc + g.c_lo] = lo;
            }
        } else {
          float* __restrict__ cp = g.C + (int64_t)bz * g.sC + (int64_t)m0 * g.ldc + n;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (u < cnt) {
              cp[(int64_t)u * g.ldc] = t[u];
              st1 += t[u];
              st2 = fmaf(t[u], t[u], st2);
            }
        }
      }
      tc_fence_before();
      mbar_arrive(TEMPTY(buf));
      if (!OUT16 && g.out_stats && nok) {   // host guarantees: tile within one sample, the warp's 32 channels in one group
        double d1 = (double)st1, d2 = (double)st2;
        const unsigned msk = __activemask();
        for (int o = 16; o > 0; o >>= 1) {
          d1 += __shfl_xor_sync(msk, d1, o);
          d2 += __shfl_xor_sync(msk, d2, o);
        }
        if (lane == 0) {
          const int smp = (mt * BA) / g.st_rps, grp = n / (g.N / g.st_groups);
          double* dst = g.out_stats + ((int64_t)smp * g.st_groups + grp) * 2;
          atomicAdd(dst, d1);
          atomicAdd(dst + 1, d2);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_sm_count = 0;
// Tokens per tile.  The persistent grid runs ceil(tiles / #SM) rounds of one tile per SM; a slightly smaller
// tile (e.g. 112 instead of 128 rows: 592 = 4 x 148 tiles for M = 8192, N = 1024) can fill the last round.
int pick_ba(int M, int other_tiles, int divides = 0) {   // divides > 0: BA must divide it (fused GroupNorm statistics)
  if (M <= TA) return (int)round_up(M, 16);
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  int best = TA; int64_t best_cost = -1;
  for (int ba = TA; ba >= 64; ba -= 16) {
    if (divides > 0 && divides % ba != 0) continue;
    const int64_t tiles = ceil_div(M, ba) * other_tiles;
    const int64_t cost = ceil_div(tiles, sms) * (ba + 12);      // + fixed per-tile overhead (pipeline fill, epilogue tail)
    if (best_cost < 0 || cost < best_cost) { best = ba; best_cost = cost; }
  }
  return best;
}

}  // namespace

bool gemm_tc16_supports(const GemmDev& p, int batch) {
  if (!tc_device_ok()) return false;
  if (p.a_fmt != DPOT_FMT_HL16 || p.w_fmt != DPOT_FMT_HL16) return false;
  if (p.a_mode != DPOT_A_PLAIN || p.c_mode != DPOT_A_PLAIN) return false;
  if (p.a_scale || p.C_pre || p.dact_src || p.c_group) return false;
  if (p.K < 8 || p.K % 8 != 0 || p.M < 1 || p.N < 1) return false;
  if (p.lda % 8 || p.ldw % 8 || p.a_lo % 8 || p.w_lo % 8) return false;
  if ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.W)) % 16) return false;
  if (batch > 1) {
    if (p.sA <= 0 || p.sW <= 0 || p.sA % 8 || p.sW % 8) return false;
  }
  return true;
}

bool gemm_tc16_fuses_stats(const GemmDev& p) {
  const int BA = pick_ba(p.M, (int)ceil_div(p.N, TN), p.st_rps);
  return p.c_fmt == DPOT_FMT_F32 && p.st_groups > 0 && p.st_rps > 0 && p.st_rps % BA == 0 && p.N % 32 == 0 &&
         (p.N / p.st_groups) % 32 == 0;
}

int gemm_tc16_launch(const GemmDev& p, int batch, cudaStream_t st) {
  Tc16Params P;
  P.g = p;
  if (g_sm_count == 0) {
    int dev = 0;
    DPOT_CUDA(cudaGetDevice(&dev));
    DPOT_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  P.n_tiles = (int)ceil_div(p.N, TN);
  P.BA = pick_ba(p.M, P.n_tiles * batch, p.out_stats ? p.st_rps : 0);
  P.m_tiles = (int)ceil_div(p.M, P.BA);
  P.total_tiles = P.n_tiles * P.m_tiles * batch;
  P.kblocks = (int)ceil_div(p.K, BKH);

  const __half* Ah = reinterpret_cast<const __half*>(p.A);
  const __half* Wh = reinterpret_cast<const __half*>(p.W);
  alignas(64) CUtensorMap mWh, mWl, mAh, mAl;
  const uint64_t sWb = batch > 1 ? (uint64_t)p.sW * 2 : (uint64_t)p.ldw * 2 * (uint64_t)p.N;
  DPOT_CALL(tc_encode_map_f16(&mWh, Wh, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)batch, (uint64_t)p.ldw * 2, sWb, BKH, TN, 1));
  DPOT_CALL(tc_encode_map_f16(&mWl, Wh + p.w_lo, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)batch, (uint64_t)p.ldw * 2, sWb, BKH, TN, 1));
  const uint64_t sAb = batch > 1 ? (uint64_t)p.sA * 2 : (uint64_t)p.lda * 2;
  DPOT_CALL(tc_encode_map_f16(&mAh, Ah, (uint64_t)p.K, (uint64_t)batch, (uint64_t)p.M, sAb, (uint64_t)p.lda * 2, BKH, 1, (uint32_t)P.BA));
  DPOT_CALL(tc_encode_map_f16(&mAl, Ah + p.a_lo, (uint64_t)p.K, (uint64_t)batch, (uint64_t)p.M, sAb, (uint64_t)p.lda * 2, BKH, 1, (uint32_t)P.BA));

  const int sm_count = g_sm_count;
  const int grid = P.total_tiles < sm_count ? P.total_tiles : sm_count;
  const int am = p.act == DPOT_ACT_NONE ? 0 : (p.act == DPOT_ACT_GELU ? 1 : 2);
  const bool o16 = p.c_fmt == DPOT_FMT_HL16;
  const bool side = p.rowbias || p.residual || p.c_scale;
#define DPOT_TC16_LAUNCH(AM, O16, SD)                                                                                  \
  do {                                                                                                                 \
    static bool attr = false;                                                                                          \
    if (!attr) {                                                                                                       \
      DPOT_CUDA(cudaFuncSetAttribute(gemm_tc16_kernel<AM, O16, SD>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                     (int)SMEM_BYTES));                                                                \
      attr = true;                                                                                                     \
    }                                                                                                                  \
    gemm_tc16_kernel<AM, O16, SD><<<grid, NTHREADS, SMEM_BYTES, st>>>(mWh, mWl, mAh, mAl, P);                          \
  } while (0)
#define DPOT_TC16_SD(AM, O16) do { if (side) DPOT_TC16_LAUNCH(AM, O16, true); else DPOT_TC16_LAUNCH(AM, O16, false); } while (0)
#define DPOT_TC16_O(AM) do { if (o16) DPOT_TC16_SD(AM, true); else DPOT_TC16_SD(AM, false); } while (0)
  if (am == 0) DPOT_TC16_O(0);
  else if (am == 1) DPOT_TC16_O(1);
  else DPOT_TC16_O(2);
#undef DPOT_TC16_O
#undef DPOT_TC16_SD
#undef DPOT_TC16_LAUNCH
  DPOT_LAUNCH_CHECK("gemm_tc16_kernel");
  return 0;
}

}  // namespace dpot

extern "C" int dpot_tc16_available(void) { return dpot::tc_device_ok() ? 1 : 0; }
