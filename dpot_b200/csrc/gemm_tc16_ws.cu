// Weight-stationary variant of the f16-split tcgen05 GEMM (gemm_tc16.cu) for SHORT-K batched problems -- the AFNO
// block-diagonal complex MLP (models/dpot.py:72-94): K = N = 2*bs = 256, M = B*kept modes, batch = n_blocks.
//
// With K = 256 a 128 x 128 tile is only 4 k-blocks of MMA work (1.6 us); the generic kernel re-loads the 128 KB weight
// tile for every token tile (half of its L2->SM traffic) and its two TMEM buffers cannot hide the MMA -> epilogue ->
// MMA hand-off (measured: MMA-only 14 us vs 6.4 us of tensor time).  Here a CTA owns ONE (batch, 128-channel) weight
// tile for its whole life: hi/lo planes of all k-blocks stay in shared memory (128 KB), only token tiles of BA = 64
// rows stream through a 6-stage ring (16 KB per k-block), and the 512 TMEM columns form FOUR accumulator buffers
// (D1 | D2 = 2 x 64 columns each), so three tiles of epilogue latency overlap the MMAs.
// Same numerics, operand format and epilogue arithmetic as gemm_tc16.cu (D1 += Whi Ahi, D2 += Whi Alo + Wlo Ahi).
#include "common.cuh"
#include "gemm_common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

namespace dpot {
namespace {

constexpr int WS_TN = 128, WS_BA = 64, WS_BKH = 64, WS_KB_MAX = 4, WS_STAGES = 6, WS_NBUF = 4;
constexpr int WS_EPI_WARP0 = 4, WS_EPI_WARPS = 16, WS_EPI_PARTS = WS_EPI_WARPS / 4;
constexpr int WS_NTHREADS = 32 * (WS_EPI_WARP0 + WS_EPI_WARPS);
constexpr uint32_t WS_W_PLANE = WS_TN * 128;                  // 16 KB: 128 rows x one k-block
constexpr uint32_t WS_W_BYTES = WS_KB_MAX * 2 * WS_W_PLANE;   // 128 KB: [kb][hi | lo]
constexpr uint32_t WS_A_PLANE = WS_BA * 128;                  // 8 KB
constexpr uint32_t WS_STAGE = 2 * WS_A_PLANE;                 // 16 KB: [hi | lo] token rows of one k-block
constexpr uint32_t WS_BAR_OFF = WS_W_BYTES + WS_STAGES * WS_STAGE;
constexpr uint32_t WS_SMEM = WS_BAR_OFF + 256 + 1024;

__host__ __device__ constexpr uint32_t ws_idesc(uint32_t n) {   // D=f32, A=B=f16, K-major, M=128, N=n
  return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

struct WsParams {
  GemmDev g;
  int n_tiles, m_tiles, kblocks, cpg;     // cpg = CTAs per (batch, n-tile) group
};

template <int ACT_MODE, bool OUT16>
__global__ void __launch_bounds__(WS_NTHREADS, 1)
gemm_tc16_ws_kernel(const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWl,
                    const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                    const WsParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a0 = smem0 + WS_W_BYTES;
  const uint32_t bar0 = smem0 + WS_BAR_OFF;
  // barriers: [0,S) full  [S,2S) empty  [2S,2S+4) tmem_full  [2S+4,2S+8) tmem_empty  [2S+8] weights ; then the TMEM slot
  auto FULL = [&](int s) -> uint32_t { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) -> uint32_t { return bar0 + 8u * (WS_STAGES + s); };
  auto TFULL = [&](int b) -> uint32_t { return bar0 + 8u * (2 * WS_STAGES + b); };
  auto TEMPTY = [&](int b) -> uint32_t { return bar0 + 8u * (2 * WS_STAGES + WS_NBUF + b); };
  const uint32_t WFULL = bar0 + 8u * (2 * WS_STAGES + 2 * WS_NBUF);
  const uint32_t tmem_slot = WFULL + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmDev& g = P.g;
  pdl_launch_dependents();
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&mapWh); tma_prefetch_desc(&mapWl); tma_prefetch_desc(&mapAh); tma_prefetch_desc(&mapAl);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < WS_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int b = 0; b < WS_NBUF; ++b) { mbar_init(TFULL(b), 1); mbar_init(TEMPTY(b), WS_EPI_WARPS); }
    mbar_init(WFULL, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // PDL: the weight tile is a parameter (packed long before the predecessor kernel), so the producer requests it
  // BEFORE pdl_wait(): on a ~20 us kernel the prologue + 128 KB weight load (~3 us) overlap the predecessor's tail.
  // Everything that touches activations is behind the wait: the producer's token loads directly, the MMAs and the
  // epilogue's stores through the barrier chain that starts at those loads.

  const int KB = P.kblocks;
  const int group = blockIdx.x / P.cpg, member = blockIdx.x % P.cpg;
  const int nt = group % P.n_tiles, bz = group / P.n_tiles;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      mbar_expect_tx(WFULL, (uint32_t)KB * 2u * WS_W_PLANE);          // the weight tile, once
      for (int kb = 0; kb < KB; ++kb) {
        tma_load_3d(smem0 + (uint32_t)kb * 2u * WS_W_PLANE, &mapWh, WFULL, kb * WS_BKH, nt * WS_TN, bz);           // (k, n, batch)
        tma_load_3d(smem0 + (uint32_t)kb * 2u * WS_W_PLANE + WS_W_PLANE, &mapWl, WFULL, kb * WS_BKH, nt * WS_TN, bz);
      }
      pdl_wait();
      int s = 0; uint32_t ph = 0;
      for (int mt = member; mt < P.m_tiles; mt += P.cpg) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(EMPTY(s), ph ^ 1);
          mbar_expect_tx(FULL(s), WS_STAGE);
          const uint32_t sb = a0 + (uint32_t)s * WS_STAGE;
          tma_load_3d(sb, &mapAh, FULL(s), kb * WS_BKH, bz, mt * WS_BA);             // (k, batch, m)
          tma_load_3d(sb + WS_A_PLANE, &mapAl, FULL(s), kb * WS_BKH, bz, mt * WS_BA);
          if (++s == WS_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      constexpr uint32_t idesc_2ba = ws_idesc(2 * WS_BA), idesc_ba = ws_idesc(WS_BA);
      mbar_wait(WFULL, 0);
      tc_fence_after();
      int s = 0; uint32_t ph = 0; uint32_t tc = 0;
      for (int mt = member; mt < P.m_tiles; mt += P.cpg, ++tc) {
        const uint32_t buf = tc % WS_NBUF, bph = (tc / WS_NBUF) & 1u;
        mbar_wait(TEMPTY(buf), bph ^ 1);
        tc_fence_after();
        const uint32_t d1 = tmem_base + buf * (2u * WS_BA);
        const uint32_t d2 = d1 + (uint32_t)WS_BA;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(FULL(s), ph);
          tc_fence_after();
          const uint32_t sb = a0 + (uint32_t)s * WS_STAGE;
          const uint32_t wb = smem0 + (uint32_t)kb * 2u * WS_W_PLANE;
#pragma unroll
          for (int k4 = 0; k4 < WS_BKH / 16; ++k4) {
            const uint64_t w_hi = make_smem_desc(wb + k4 * 32);
            const uint64_t w_lo = make_smem_desc(wb + WS_W_PLANE + k4 * 32);
            const uint64_t a_hl = make_smem_desc(sb + k4 * 32);              // rows [0,BA) = hi, [BA,2BA) = lo
            umma_f16(d1, w_hi, a_hl, idesc_2ba, (kb > 0 || k4 > 0) ? 1u : 0u);
            umma_f16(d2, w_lo, a_hl, idesc_ba, 1u);
          }
          umma_commit(EMPTY(s));
          if (++s == WS_STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(TFULL(buf));
      }
    }
  } else if (warp >= WS_EPI_WARP0) {
    // ================================ epilogue ====================================
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may read
    const int part = (warp - WS_EPI_WARP0) >> 2;         // 8-column groups part, part + 4
    __half* const Ch_base = reinterpret_cast<__half*>(g.C);
    const int n = nt * WS_TN + quarter * 32 + lane;
    const bool nok = n < g.N;
    const float bias_n = (g.bias && nok) ? g.bias[(int64_t)bz * g.sBias + n] : 0.f;
    pdl_wait();
    uint32_t tc = 0;
    for (int mt = member; mt < P.m_tiles; mt += P.cpg, ++tc) {
      const uint32_t buf = tc % WS_NBUF, bph = (tc / WS_NBUF) & 1u;
      mbar_wait(TFULL(buf), bph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + buf * (2u * WS_BA) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
      for (int it = 0; it < WS_BA / 8 / WS_EPI_PARTS; ++it) {
        const int c0 = (part + it * WS_EPI_PARTS) * 8;
        uint32_t r1[8], r2[8];
        tmem_ld8(t_row + (uint32_t)c0, r1);
        tmem_ld8(t_row + (uint32_t)(WS_BA + c0), r2);
        tmem_ld_wait();
        const int m0 = mt * WS_BA + c0;
        const int cnt = min(8, g.M - m0);
        if (cnt <= 0 || !nok) continue;
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = fmaf(__uint_as_float(r2[u]), HL_INV, __uint_as_float(r1[u])) + bias_n;
        if (ACT_MODE == 1) {
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] = gelu_fast(t[u]);
        } else if (ACT_MODE == 2) {
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] = act_apply(t[u], g.act);
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
            if (u < cnt) cp[(int64_t)u * g.ldc] = t[u];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(TEMPTY(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_ws_mode = -1;    // -1 auto, 0 never, 2 auto without the early (PDL) start (dpot_tc16_set_ws; tests / experiments)
int g_ws_pdl = 1;      // launch with programmatic stream serialization so that prologue + weight load start early

}  // namespace

// The weight-stationary plan serves batched / multi-n-tile problems whose whole K fits the resident weight buffer and
// that have no epilogue side inputs; every (batch, n-tile) group gets the same number of CTAs.
bool gemm_tc16_ws_takes(const GemmDev& p, int batch, int sms) {
  if (g_ws_mode == 0) return false;
  if (p.c_fmt == DPOT_FMT_HL16G32) return false;
  if (p.K > WS_KB_MAX * WS_BKH || p.rowbias || p.residual || p.c_scale || p.out_stats) return false;
  const int groups = (int)ceil_div(p.N, WS_TN) * batch;
  if (groups < 2 || groups > sms) return false;
  const int cpg = sms / groups, m_tiles = (int)ceil_div(p.M, WS_BA);
  return m_tiles >= 2 * cpg;          // at least two token tiles per CTA, else the weight load is not amortised
}

int gemm_tc16_ws_launch(const GemmDev& p, int batch, int sms, cudaStream_t st) {
  WsParams P;
  P.g = p;
  P.n_tiles = (int)ceil_div(p.N, WS_TN);
  P.m_tiles = (int)ceil_div(p.M, WS_BA);
  P.kblocks = (int)ceil_div(p.K, WS_BKH);
  const int groups = P.n_tiles * batch;
  P.cpg = sms / groups;
  if (P.cpg > P.m_tiles) P.cpg = P.m_tiles;
  const __half* Ah = reinterpret_cast<const __half*>(p.A);
  const __half* Wh = reinterpret_cast<const __half*>(p.W);
  alignas(64) CUtensorMap mWh, mWl, mAh, mAl;
  const uint64_t sWb = batch > 1 ? (uint64_t)p.sW * 2 : (uint64_t)p.ldw * 2 * (uint64_t)p.N;
  DPOT_CALL(tc_encode_map_f16(&mWh, Wh, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)batch, (uint64_t)p.ldw * 2, sWb, WS_BKH, WS_TN, 1));
  DPOT_CALL(tc_encode_map_f16(&mWl, Wh + p.w_lo, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)batch, (uint64_t)p.ldw * 2, sWb, WS_BKH, WS_TN, 1));
  const uint64_t sAb = batch > 1 ? (uint64_t)p.sA * 2 : (uint64_t)p.lda * 2;
  DPOT_CALL(tc_encode_map_f16(&mAh, Ah, (uint64_t)p.K, (uint64_t)batch, (uint64_t)p.M, sAb, (uint64_t)p.lda * 2, WS_BKH, 1, WS_BA));
  DPOT_CALL(tc_encode_map_f16(&mAl, Ah + p.a_lo, (uint64_t)p.K, (uint64_t)batch, (uint64_t)p.M, sAb, (uint64_t)p.lda * 2, WS_BKH, 1, WS_BA));
  const int grid = groups * P.cpg;
  const int am = p.act == DPOT_ACT_NONE ? 0 : (p.act == DPOT_ACT_GELU ? 1 : 2);
  const bool o16 = p.c_fmt == DPOT_FMT_HL16;
#define DPOT_WS_LAUNCH(AM, O16)                                                                                       \
  do {                                                                                                                \
    static DevOnce attr;                                                                                         \
    if (attr.need()) {                                                                                                      \
      DPOT_CUDA(cudaFuncSetAttribute(gemm_tc16_ws_kernel<AM, O16>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                     (int)WS_SMEM));                                                                  \
      attr.done();                                                                                                    \
    }                                                                                                                 \
    DPOT_CUDA(launch_pdl_if(g_pdl != 0 || g_ws_pdl != 0, gemm_tc16_ws_kernel<AM, O16>, dim3(grid), dim3(WS_NTHREADS), WS_SMEM, st, mWh, mWl, mAh, mAl, P)); \
  } while (0)
#define DPOT_WS_O(AM) do { if (o16) DPOT_WS_LAUNCH(AM, true); else DPOT_WS_LAUNCH(AM, false); } while (0)
  if (am == 0) DPOT_WS_O(0); else if (am == 1) DPOT_WS_O(1); else DPOT_WS_O(2);
#undef DPOT_WS_O
#undef DPOT_WS_LAUNCH
  DPOT_LAUNCH_CHECK("gemm_tc16_ws_kernel");
  return 0;
}

}  // namespace dpot

extern "C" void dpot_tc16_set_ws(int32_t mode) { dpot::g_ws_mode = mode == 2 ? -1 : mode; dpot::g_ws_pdl = mode == 2 ? 0 : 1; }
