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
// Warp roles (640 threads, 1 CTA / SM, persistent): warp 0 TMA producer, warp 1 MMA issuer,
// warp 2 TMEM allocator, warps 4-19 epilogue (four per TMEM lane quarter, interleaved 8-column groups).
#include "common.cuh"
#include "gemm_common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

namespace dpot {
namespace {

constexpr int TN = 128;            // weight rows (output channels) per CTA tile = TMEM lanes
constexpr int TA = 128;            // max activation rows (tokens) per tile
constexpr int BKH = 64;            // halves per k-block = one 128 B swizzle row
constexpr int EPI_WARP0 = 4, EPI_WARPS = 16, EPI_PARTS = EPI_WARPS / 4;
constexpr int NTHREADS = 32 * (EPI_WARP0 + EPI_WARPS);

constexpr uint32_t W_BYTES = TN * 128;          // 16 KB per plane
constexpr uint32_t OFF_W_HI = 0, OFF_W_LO = W_BYTES, OFF_A = 2 * W_BYTES;   // A_hi at OFF_A, A_lo one plane later

// CG = 1: one CTA per tile (128 channels x BA tokens).  CG = 2: a CTA PAIR (tcgen05 cta_group::2) owns a
// 256-channel x BA-token tile: each CTA stages its own 128 weight rows and HALF of the tokens, the pair's MMA
// (UMMA M = 256) reads the token operand from both shared memories -- 48 KB instead of 64 KB of L2->SM traffic
// per k-block and CTA.  The L2->SM fabric (~6300 B/clk chip-wide, measured) is what bounds this engine.
template <int CG>
struct Geo {
  static constexpr int STAGES = CG == 2 ? 4 : 3;
  static constexpr uint32_t A_BYTES = (TA / CG) * 128;                // per plane (max)
  static constexpr uint32_t STAGE_BYTES = 2 * W_BYTES + 2 * A_BYTES;  // 64 KB / 48 KB
  static constexpr uint32_t BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr uint32_t SMEM_BYTES = BAR_OFF + 256 + 1024;
};

// instruction descriptor: D=f32, A=B=f16, both K-major, M=m, N=n
// bit 15 / 16: the UMMA A (weight role) / B (token role) operand tile is MN-major (transposed storage)
__host__ __device__ constexpr uint32_t make_idesc16(uint32_t n, uint32_t m = 128u, uint32_t a_mn = 0u, uint32_t b_mn = 0u) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// MN-major, 128B-swizzled operand tile as two-dimensional TMA boxes of (64 MN elements = 128 B) x (64 k rows) leave it:
// the 8 x 128 B swizzle atoms of one box follow each other along k (SBO = 1024 B), the next 64 MN elements are the
// next box (LBO = 8192 B).  One MMA (k = 16) covers two atoms: the k-step inside a stage is 2048 B.
constexpr uint32_t MN_BOX_BYTES = 64 * 128;
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(MN_BOX_BYTES >> 4) << 16;          // leading byte offset: between 64-element MN chunks
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: between 8-row k groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct Tc16Params {
  GemmDev g;
  int BA;            // activation rows per tile, multiple of 16, <= 128
  int n_tiles, m_tiles, total_tiles, kblocks;
  int dbg;           // experiment mask (dpot_tc16_set_debug): 1 = no TMA loads, 2 = no MMAs, 4 = no epilogue work
  int batch;         // independent problems; total_tiles also spans the g.ksplit contraction chunks (bz = b + batch * chunk)
  int w_bmul, a_bmul; // 0: the operand is shared by every problem of the batch (batch stride 0), else 1
  int single;        // 1: half-precision operand mode -- only the hi planes are loaded and multiplied (1 MMA per product)
};

// ACT_MODE: 0 = none, 1 = GELU (erf), 2 = runtime switch.  OUT16: store the result as DPOT_FMT_HL16.
// SIDE: side inputs of the epilogue.  0 = none (that code and its registers are compiled out -- the AFNO GEMMs,
// K = 2*bs, are epilogue-bound); 1 = row-periodic bias only, 2 = residual only: ALL of the warp's side values of a tile
// are requested before the accumulator is even waited for (their latency is multiple microseconds while the TMA
// stream saturates the L2 -> SM fabric; a one-group-ahead prefetch left the epilogue latency-bound: +14 us on the
// K = 352 time-aggregation GEMM); 3 = any combination incl. the per-sample affine, fetched one group ahead.
template <int CG, int ACT_MODE, bool OUT16, int SIDE>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc16_kernel(const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWl,
                 const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                 const Tc16Params P) {
  constexpr int STAGES = Geo<CG>::STAGES;
  constexpr uint32_t STAGE_BYTES = Geo<CG>::STAGE_BYTES, BAR_OFF = Geo<CG>::BAR_OFF;
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
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;     // 0 = leader of the pair (issues the MMAs)
  pdl_launch_dependents();

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&mapWh); tma_prefetch_desc(&mapWl); tma_prefetch_desc(&mapAh); tma_prefetch_desc(&mapAl);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    // tmem_empty lives in the leader and counts one arrival per epilogue warp of every CTA of the group
    for (int b = 0; b < 2; ++b) { mbar_init(TFULL(b), 1); mbar_init(TEMPTY(b), CG * EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_2cta(tmem_slot, 512); else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();      // the prologue above overlapped the predecessor's tail; operands / side inputs are read below

  const int KB = P.kblocks, BA = P.BA;
  const int rows_a = BA / CG;                                  // token rows this CTA stages (a multiple of 8)
  const uint32_t a_bytes = (uint32_t)rows_a * 128u;
  const uint32_t a_plane = CG == 2 ? ((a_bytes + 1023u) & ~1023u) : a_bytes;   // lo plane on a swizzle-atom boundary
  const uint32_t stage_tx = P.single ? W_BYTES + a_bytes : 2u * W_BYTES + 2u * a_bytes;       // bytes this CTA loads per stage
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      const uint32_t full_leader = CG == 2 ? mapa_rank(FULL(0), 0) : FULL(0);   // completion barrier of the group
      int s = 0; uint32_t ph = 0;
      for (int tile = tile0; tile < P.total_tiles; tile += tile_step) {
        const int mt = tile % P.m_tiles; const int rest = tile / P.m_tiles;
        const int nt = rest % P.n_tiles; const int bz = rest / P.n_tiles;
        const int n0 = (nt * CG + (int)rank) * TN, m0 = mt * BA + (int)rank * rows_a;
        const int bb = bz % P.batch, k00 = (bz / P.batch) * (int)g.kchunk;     // problem of the batch, contraction chunk
        const int bw = bb * P.w_bmul, ba = bb * P.a_bmul;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(EMPTY(s), ph ^ 1);
          const uint32_t sb = smem0 + (uint32_t)s * STAGE_BYTES;
          const int k0 = k00 + kb * BKH;
          if (P.dbg & 1) {                      // experiment: the MMA pipeline alone (operands = whatever is in smem)
            if (rank == 0) mbar_arrive(FULL(s));
          } else {
            uint32_t fb = FULL(s);
            if (CG == 2) {
              if (rank == 0) mbar_expect_tx(FULL(s), 2u * stage_tx);   // the leader's barrier collects both CTAs' bytes
              fb = full_leader + 8u * s;
            } else {
              mbar_expect_tx(FULL(s), stage_tx);
            }
            auto ld = [&](uint32_t dst, const CUtensorMap* mp, int c0, int c1, int c2) {
              if (CG == 2) tma_load_3d_2cta(dst, mp, fb, c0, c1, c2); else tma_load_3d(dst, mp, fb, c0, c1, c2);
            };
            if (g.w_tr) {                       // stored [K, N]: dims (n, k, batch), two boxes of 64 channels per plane
#pragma unroll
              for (int j = 0; j < TN / 64; ++j) {
                ld(sb + OFF_W_HI + j * MN_BOX_BYTES, &mapWh, n0 + 64 * j, k0, bw);
                if (!P.single) ld(sb + OFF_W_LO + j * MN_BOX_BYTES, &mapWl, n0 + 64 * j, k0, bw);
              }
            } else {
              ld(sb + OFF_W_HI, &mapWh, k0, n0, bw);                   // dims (k, n, batch)
              if (!P.single) ld(sb + OFF_W_LO, &mapWl, k0, n0, bw);
            }
            if (g.a_tr) {                       // stored [K, M]: dims (m, k, batch), rows_a / 64 boxes per plane
              for (int j = 0; j < rows_a / 64; ++j) {
                ld(sb + OFF_A + j * MN_BOX_BYTES, &mapAh, m0 + 64 * j, k0, ba);
                if (!P.single) ld(sb + OFF_A + a_plane + j * MN_BOX_BYTES, &mapAl, m0 + 64 * j, k0, ba);
              }
            } else {
              ld(sb + OFF_A, &mapAh, k0, ba, m0);                      // dims (k, batch, m)
              if (!P.single) ld(sb + OFF_A + a_plane, &mapAl, k0, ba, m0);
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (rank == 0 && elect_one()) {
      const uint32_t wmn = g.w_tr ? 1u : 0u, amn = g.a_tr ? 1u : 0u;
      const uint32_t idesc_2ba = make_idesc16((uint32_t)(2 * BA), 128u, wmn, amn);   // CG 1: [hi ; lo] token rows in one MMA
      const uint32_t idesc_ba = make_idesc16((uint32_t)BA, 128u * CG, wmn, amn);
      const uint32_t w_kstep = g.w_tr ? 2048u : 32u, a_kstep = g.a_tr ? 2048u : 32u;
      // descriptor = constant high part (layout, LBO / SBO) | start address: one OR per operand inside the loop
      const uint64_t w_hi_bits = (g.w_tr ? make_smem_desc_mn(0u) : make_smem_desc(0u));
      const uint64_t a_hi_bits = (g.a_tr ? make_smem_desc_mn(0u) : make_smem_desc(0u));
      auto DW = [&](uint32_t addr) -> uint64_t { return w_hi_bits | (uint64_t)((addr & 0x3FFFFu) >> 4); };
      auto DA = [&](uint32_t addr) -> uint64_t { return a_hi_bits | (uint64_t)((addr & 0x3FFFFu) >> 4); };
      const bool single = P.single != 0;
      int s = 0; uint32_t ph = 0; uint32_t tc = 0;
      for (int tile = tile0; tile < P.total_tiles; tile += tile_step, ++tc) {
        const uint32_t buf = tc & 1u, bph = (tc >> 1) & 1u;
        mbar_wait(TEMPTY(buf), bph ^ 1);          // every epilogue warp (of both CTAs) has drained this TMEM buffer
        tc_fence_after();
        const uint32_t d1 = tmem_base + buf * 256u;
        const uint32_t d2 = d1 + (uint32_t)BA;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(FULL(s), ph);
          tc_fence_after();
          const uint32_t sb = smem0 + (uint32_t)s * STAGE_BYTES;
          if (P.dbg & 2) {                      // experiment: the load pipeline alone
          } else if (single) {                  // half-precision operand mode: hi x hi only
#pragma unroll
            for (int k4 = 0; k4 < BKH / 16; ++k4) {
              const uint32_t acc = (kb > 0 || k4 > 0) ? 1u : 0u;
              if (CG == 2) umma_f16_2cta(d1, DW(sb + OFF_W_HI + k4 * w_kstep), DA(sb + OFF_A + k4 * a_kstep), idesc_ba, acc);
              else umma_f16(d1, DW(sb + OFF_W_HI + k4 * w_kstep), DA(sb + OFF_A + k4 * a_kstep), idesc_ba, acc);
            }
          } else {
#pragma unroll
            for (int k4 = 0; k4 < BKH / 16; ++k4) {
              const uint64_t w_hi = DW(sb + OFF_W_HI + k4 * w_kstep);
              const uint64_t w_lo = DW(sb + OFF_W_LO + k4 * w_kstep);
              const uint64_t a_hi = DA(sb + OFF_A + k4 * a_kstep);      // CG 1: rows [0,BA) = hi, [BA,2BA) = lo
              const uint32_t acc = (kb > 0 || k4 > 0) ? 1u : 0u;
              if (CG == 2) {
                // the pair's token operand: first BA/2 rows from the leader's shared memory, the rest from the peer's
                const uint64_t a_lo = DA(sb + OFF_A + a_plane + k4 * a_kstep);
                umma_f16_2cta(d1, w_hi, a_hi, idesc_ba, acc);
                umma_f16_2cta(d2, w_hi, a_lo, idesc_ba, acc);
                umma_f16_2cta(d2, w_lo, a_hi, idesc_ba, 1u);
              } else {
                umma_f16(d1, w_hi, a_hi, idesc_2ba, acc);
                umma_f16(d2, w_lo, a_hi, idesc_ba, 1u);
              }
            }
          }
          if (CG == 2) umma_commit_2cta(EMPTY(s)); else umma_commit(EMPTY(s));
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (CG == 2) umma_commit_2cta(TFULL(buf)); else umma_commit(TFULL(buf));
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ================================ epilogue ====================================
    const int quarter = warp & 3;                     // TMEM lane quarter this warp may read
    const int half = (warp - EPI_WARP0) >> 2;         // interleaved 8-column groups: g = half, half + EPI_PARTS, ...
    const int ngroups = BA / 8;
    __half* const Ch_base = reinterpret_cast<__half*>(g.C);
    const uint32_t tempty_leader = CG == 2 ? mapa_rank(TEMPTY(0), 0) : TEMPTY(0);
    const bool epi_single = P.single != 0;
    uint32_t tc = 0;
    for (int tile = tile0; tile < P.total_tiles; tile += tile_step, ++tc) {
      const int mt = tile % P.m_tiles; const int rest = tile / P.m_tiles;
      const int nt = rest % P.n_tiles; const int bz = rest / P.n_tiles;
      const uint32_t buf = tc & 1u, bph = (tc >> 1) & 1u;
      const int n = (nt * CG + (int)rank) * TN + quarter * 32 + lane;
      const bool nok = n < g.N;
      const int bb = bz % P.batch;
      const int64_t c_boff = (int64_t)bb * g.sC + (int64_t)(bz / P.batch) * g.sC2;   // problem of the batch + contraction chunk
      const bool dact_on = SIDE == 2 && g.dact_src != nullptr;   // side value = pre-activation whose act' multiplies the result
      const float bias_n = (g.bias && nok) ? g.bias[(int64_t)bb * g.sBias + n] : 0.f;
      // fused GroupNorm statistics: a tile of <= 128 tokens touches at most two samples (slot 0 / slot 1)
      float st1a = 0.f, st2a = 0.f, st1b = 0.f, st2b = 0.f;
      float csum = 0.f;                                 // column sum of this lane's channel over the warp's token groups
      const int smp0 = g.out_stats ? (mt * BA) / g.st_rps : 0;
      const int m_next = (smp0 + 1) * g.st_rps;       // first token of the next sample
      // side inputs (row-periodic bias, residual) of a group are fetched one group ahead: their global-load
      // latency hides behind the arithmetic of the previous group instead of serialising the epilogue
      constexpr bool DEEP = (SIDE == 1 || SIDE == 2);
      constexpr int EPI_ITERS = (TA / 8 + EPI_PARTS - 1) / EPI_PARTS;      // groups per warp and tile (max)
      float sd[DEEP ? EPI_ITERS : 1][8];                                   // DEEP: every side value of this warp's tile
      float nrb[8], nrs[8];                                                // SIDE 3: one group ahead
      auto side_ptr = [&](int m) -> const float* {
        return SIDE == 1 ? g.rowbias + (int64_t)(m % g.rb_period) * g.ldrb + n
                         : (dact_on ? g.dact_src + (int64_t)bb * g.sDact + (int64_t)m * g.lddact + n : g.residual + (int64_t)m * g.ldr + n);
      };
      auto load_side = [&](int gi) {
        if (SIDE != 3) return;
        const int m0 = mt * BA + gi * 8;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool ok = nok && (m0 + u) < g.M;
          nrb[u] = (g.rowbias && ok) ? __ldg(g.rowbias + (int64_t)((m0 + u) % g.rb_period) * g.ldrb + n) : 0.f;
          nrs[u] = (g.residual && ok) ? __ldg(g.residual + (int64_t)(m0 + u) * g.ldr + n) : 0.f;
        }
      };
      if (DEEP) {
#pragma unroll
        for (int it = 0; it < EPI_ITERS; ++it) {
          const int m0 = mt * BA + (half + it * EPI_PARTS) * 8;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            sd[it][u] = (nok && half + it * EPI_PARTS < ngroups && m0 + u < g.M) ? __ldg(side_ptr(m0 + u)) : 0.f;
        }
      } else if (half < ngroups) load_side(half);
      mbar_wait(TFULL(buf), bph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + buf * 256u + ((uint32_t)(quarter * 32) << 16);
      auto do_group = [&](int gi, const float (&sv)[8]) {      // sv: this group's deep-prefetched side values (DEEP only)
        const int c0 = gi * 8;
        uint32_t r1[8], r2[8];
        tmem_ld8(t_row + (uint32_t)c0, r1);
        tmem_ld8(t_row + (uint32_t)(BA + c0), r2);          // (half-precision operand mode: stale columns, ignored below)
        float rb[8], rs[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          rb[u] = SIDE == 3 ? nrb[u] : (SIDE == 1 ? sv[u] : 0.f);
          rs[u] = SIDE == 3 ? nrs[u] : (SIDE == 2 ? sv[u] : 0.f);
        }
        if (gi + EPI_PARTS < ngroups) load_side(gi + EPI_PARTS);
        tmem_ld_wait();
        const int m0 = mt * BA + c0;
        const int cnt = min(8, g.M - m0);
        if (cnt <= 0 || !nok) return;
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          t[u] = (epi_single ? __uint_as_float(r1[u]) : fmaf(__uint_as_float(r2[u]), HL_INV, __uint_as_float(r1[u]))) + bias_n + rb[u];
        if (g.C_pre) {                                   // training: keep the pre-activation (fp32) for the backward pass
          float* __restrict__ pp = g.C_pre + (int64_t)bb * g.sPre + (int64_t)m0 * g.ldpre + n;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (u < cnt) pp[(int64_t)u * g.ldpre] = t[u];
        }
        if (ACT_MODE == 1) {
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] = gelu_select(t[u]);
        } else if (ACT_MODE == 2) {
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] = act_apply(t[u], g.act);
        }
        if (SIDE == 3 && g.c_scale) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (u < cnt) {
              const int64_t o = (int64_t)((m0 + u) / g.c_rps) * g.N + n;
              t[u] = fmaf(t[u], g.c_scale[o], g.c_shift[o]);
            }
        }
        if (SIDE == 2 && dact_on) {
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] *= act_grad_fast(rs[u], g.dact);
        } else if (SIDE >= 2) {
#pragma unroll
          for (int u = 0; u < 8; ++u) t[u] += rs[u];
        }
        if (g.out_colsum) {
#pragma unroll
          for (int u = 0; u < 8; ++u) csum += u < cnt ? t[u] : 0.f;
        }
        if (OUT16) {
          const bool grp = g.c_fmt == DPOT_FMT_HL16G32;            // [hi 32 | lo 32] record per 32-column group
          const int64_t ncol = grp ? (int64_t)((n >> 5) << 6) + (n & 31) : n, lo_off = grp ? 32 : g.c_lo;
          __half* __restrict__ cp = Ch_base + c_boff + (int64_t)m0 * g.ldc + ncol;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (u < cnt) {
              __half hi, lo;
              hl_split(t[u], hi, lo);
              cp[(int64_t)u * g.ldc] = hi;
              cp[(int64_t)u * g.ldc + lo_off] = lo;
            }
        } else {
          float* __restrict__ cp = g.C + c_boff + (int64_t)m0 * g.ldc + n;
          float a1 = 0.f, a2 = 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (u < cnt) {
              cp[(int64_t)u * g.ldc] = t[u];
              a1 += t[u];
              a2 = fmaf(t[u], t[u], a2);
            }
          const bool nxt = m0 >= m_next;             // st_rps % 8 == 0: a group of 8 tokens never straddles samples
          st1a += nxt ? 0.f : a1; st2a += nxt ? 0.f : a2;
          st1b += nxt ? a1 : 0.f; st2b += nxt ? a2 : 0.f;
        }
      };
      if (!(P.dbg & 4)) {                      // (dbg 4: experiment without epilogue work)
        if (DEEP) {
#pragma unroll
          for (int it = 0; it < EPI_ITERS; ++it)
            if (half + it * EPI_PARTS < ngroups) do_group(half + it * EPI_PARTS, sd[it]);
        } else {
#pragma unroll 1
          for (int gi = half; gi < ngroups; gi += EPI_PARTS) do_group(gi, sd[0]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty_leader + 8u * buf); else mbar_arrive(TEMPTY(buf));
      }
      if (g.out_colsum && nok) atomicAdd(g.out_colsum + (int64_t)bb * g.N + n, (double)csum);
      if (!OUT16 && g.out_stats) {   // host guarantees: the warp's 32 channels lie in one group
#pragma unroll
        for (int slot = 0; slot < 2; ++slot) {
          const int smp = smp0 + slot;
          if (slot == 1 && (mt * BA + BA <= m_next || m_next >= g.M)) break;   // uniform: the tile ends inside sample smp0
          // fp32 warp reduction (32 partial sums of <= 64 elements each; the fp64 pipe is slow and these shuffles sit on
          // the epilogue's critical path); the cross-tile accumulation below stays in double
          float f1 = nok ? (slot ? st1b : st1a) : 0.f, f2 = nok ? (slot ? st2b : st2a) : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            f1 += __shfl_xor_sync(0xffffffffu, f1, o);
            f2 += __shfl_xor_sync(0xffffffffu, f2, o);
          }
          const double d1 = (double)f1, d2 = (double)f2;
          if (lane == 0 && nok) {
            const int grp = n / (g.N / g.st_groups);
            double* dst = g.out_stats + ((int64_t)smp * g.st_groups + grp) * 2;
            atomicAdd(dst, d1);
            atomicAdd(dst + 1, d2);
          }
        }
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: no CTA may exit while its peer can still reach its smem / TMEM
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

int g_sm_count = 0;
int g_dbg = 0;
int g_single = 0;        // dpot_tc16_set_precision: 1 = half-precision operand mode (hi planes only)
int g_pair_mode = -1;    // -1: auto (cost model below), 0: never, 1: pairs whenever legal (dpot_tc16_set_pair)

bool stats_any_ba(const GemmDev& p) { return p.st_rps % 8 == 0 && p.st_rps >= TA; }

// Tile plan.  The persistent grid runs ceil(tiles / #units) rounds of one tile per unit (unit = SM, or SM pair); a
// slightly smaller tile (e.g. 112 instead of 128 tokens: 592 = 4 x 148 tiles for M = 8192, N = 1024) can fill the
// last round.  Cost of a plan = rounds x (BA x k-blocks + fixed per-tile overhead), in token-k-block units; a CTA
// pair moves 25 % fewer operand bytes per MMA (measured: ~7 % faster per tile at K = 1024) but pays a larger
// per-tile overhead (cluster-wide handshakes), so short-K problems stay on single-CTA tiles.
// divides > 0: BA must divide it (fused GroupNorm statistics when a tile may not straddle samples).
struct Plan { int cg, BA, n_tiles, m_tiles; int64_t cost; };
Plan plan_for(const GemmDev& p, int batch, int cg, int sms, int divides) {
  Plan pl;
  pl.cg = cg;
  pl.n_tiles = (int)ceil_div(p.N, TN * cg);
  const int other = pl.n_tiles * batch, units = sms / cg, step = 16;   // pairs: BA/2 rows per CTA, a multiple of 8
  const int64_t kblocks = ceil_div(p.ksplit > 1 ? p.kchunk : p.K, BKH), ovh = cg == 2 ? 320 : 192;
  pl.BA = 0; pl.cost = -1;
  if (p.a_tr) {   // transposed token operand: whole 64-row TMA boxes (a pair stages BA / 2 rows per CTA)
    for (int ba = TA; ba >= (cg == 2 ? 128 : 64); ba -= 64) {
      const int64_t tiles = ceil_div(p.M, ba) * other;
      const int64_t cost = ceil_div(tiles, units) * (ba * kblocks + ovh);
      if (pl.cost < 0 || cost < pl.cost) { pl.BA = ba; pl.cost = cost; }
    }
  } else if (p.M <= TA) {
    pl.BA = (int)round_up(p.M, step);
    pl.cost = ceil_div(other, units) * (pl.BA * kblocks + ovh);
  } else {
    for (int ba = TA; ba >= (cg == 2 ? 96 : 64); ba -= step) {
      if (divides > 0 && divides % ba != 0) continue;
      const int64_t tiles = ceil_div(p.M, ba) * other;
      const int64_t cost = ceil_div(tiles, units) * (ba * kblocks + ovh);
      if (pl.cost < 0 || cost < pl.cost) { pl.BA = ba; pl.cost = cost; }
    }
  }
  if (cg == 2 && pl.cost >= 0) pl.cost = pl.cost * 93 / 100;
  pl.m_tiles = pl.BA > 0 ? (int)ceil_div(p.M, pl.BA) : 0;
  return pl;
}
Plan make_plan(const GemmDev& p, int batch) {
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  const int divides = (p.st_groups > 0 && p.st_rps > 0 && !stats_any_ba(p)) ? p.st_rps : 0;
  Plan one = plan_for(p, batch, 1, sms, divides);
  // pairs need >= 256 output channels to fill both CTAs and enough tokens to split
  if (g_pair_mode == 0 || p.N <= TN || p.M < 32 || sms % 2 != 0 || (p.a_tr && p.M < 128)) return one;
  if ((g_dbg & 3) == 3) return one;   // experiment "no TMA + no MMA": a pair commit with nothing in flight never reaches the peer (measured: trap)
  Plan two = plan_for(p, batch, 2, sms, divides);
  if (two.cost < 0) return one;
  if (one.cost < 0 || g_pair_mode == 1) return two;
  return two.cost < one.cost ? two : one;
}

}  // namespace

bool gemm_tc16_supports(const GemmDev& p, int batch) {
  if (!tc_device_ok()) return false;
  if (p.a_fmt != DPOT_FMT_HL16 || p.w_fmt != DPOT_FMT_HL16) return false;
  if (p.a_mode != DPOT_A_PLAIN || p.c_mode != DPOT_A_PLAIN) return false;
  if (p.a_scale || p.c_group) return false;
  if (p.K < 8 || p.M < 1 || p.N < 1) return false;
  if (!p.a_tr && !p.w_tr && p.K % 8 != 0) return false;      // K-major rows must keep TMA's 16-byte pitch
  if (p.a_tr && p.M % 8 != 0) return false;                  // transposed storage: the row is the M / N axis
  if (p.w_tr && p.N % 8 != 0) return false;
  if (p.lda % 8 || p.ldw % 8 || p.a_lo % 8 || p.w_lo % 8) return false;
  if ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.W)) % 16) return false;
  if (batch > 1) {
    if (p.sA < 0 || p.sW < 0 || p.sA % 8 || p.sW % 8) return false;      // stride 0: the operand is shared by the batch
  }
  return true;
}

bool gemm_tc16_fuses_stats(const GemmDev& p) {
  if (!(p.c_fmt == DPOT_FMT_F32 && p.st_groups > 0 && p.st_rps > 0 && p.N % 32 == 0 && (p.N / p.st_groups) % 32 == 0))
    return false;
  if (stats_any_ba(p)) return true;
  const Plan pl = make_plan(p, 1);
  return p.st_rps % pl.BA == 0;
}

int gemm_tc16_launch(const GemmDev& p, int batch, cudaStream_t st) {
  Tc16Params P;
  P.g = p;
  g_sm_count = sm_count_cur();
  GemmDev q = p;
  if (!p.out_stats) { q.st_groups = 0; q.st_rps = 0; }   // the tile plan only honours the statistics geometry when they are fused
  const Plan pl = make_plan(q, batch * p.ksplit);
  const int CGn = pl.cg;
  P.n_tiles = pl.n_tiles;
  P.BA = pl.BA;
  P.m_tiles = pl.m_tiles;
  P.total_tiles = P.n_tiles * P.m_tiles * batch * p.ksplit;
  P.kblocks = (int)ceil_div(p.ksplit > 1 ? p.kchunk : p.K, BKH);
  P.dbg = g_dbg;
  P.batch = batch;
  P.single = g_single;
  P.w_bmul = (batch > 1 && p.sW == 0) ? 0 : 1;
  P.a_bmul = (batch > 1 && p.sA == 0) ? 0 : 1;
  const int batchW = P.w_bmul ? batch : 1, batchA = P.a_bmul ? batch : 1;
  if (p.ksplit <= 1) { P.g.kchunk = 0; P.g.sC2 = 0; P.g.ksplit = 1; }

  const __half* Ah = reinterpret_cast<const __half*>(p.A);
  const __half* Wh = reinterpret_cast<const __half*>(p.W);
  alignas(64) CUtensorMap mWh, mWl, mAh, mAl;
  const uint64_t sWh = (batch > 1 && p.sW > 0) ? (uint64_t)p.sW * 2 : 0;      // batch stride in bytes, 0 = derive a dummy
  const uint64_t sAh = (batch > 1 && p.sA > 0) ? (uint64_t)p.sA * 2 : 0;
  if (p.w_tr) {   // stored [K, N] (row stride ldw): dims (n, k, batch), boxes of 64 channels x 64 k rows
    const uint64_t sWt = sWh ? sWh : (uint64_t)p.ldw * 2 * (uint64_t)p.K;
    DPOT_CALL(tc_encode_map_f16(&mWh, Wh, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)batchW, (uint64_t)p.ldw * 2, sWt, 64, BKH, 1));
    DPOT_CALL(tc_encode_map_f16(&mWl, Wh + p.w_lo, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)batchW, (uint64_t)p.ldw * 2, sWt, 64, BKH, 1));
  } else {
    const uint64_t sWk = sWh ? sWh : (uint64_t)p.ldw * 2 * (uint64_t)p.N;
    DPOT_CALL(tc_encode_map_f16(&mWh, Wh, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)batchW, (uint64_t)p.ldw * 2, sWk, BKH, TN, 1));
    DPOT_CALL(tc_encode_map_f16(&mWl, Wh + p.w_lo, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)batchW, (uint64_t)p.ldw * 2, sWk, BKH, TN, 1));
  }
  const uint32_t box_m = (uint32_t)(P.BA / CGn);
  if (p.a_tr) {   // stored [K, M] (row stride lda): dims (m, k, batch)
    const uint64_t sAt = sAh ? sAh : (uint64_t)p.lda * 2 * (uint64_t)p.K;
    DPOT_CALL(tc_encode_map_f16(&mAh, Ah, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)batchA, (uint64_t)p.lda * 2, sAt, 64, BKH, 1));
    DPOT_CALL(tc_encode_map_f16(&mAl, Ah + p.a_lo, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)batchA, (uint64_t)p.lda * 2, sAt, 64, BKH, 1));
  } else {
    const uint64_t sAb = sAh ? sAh : (uint64_t)p.lda * 2;
    DPOT_CALL(tc_encode_map_f16(&mAh, Ah, (uint64_t)p.K, (uint64_t)batchA, (uint64_t)p.M, sAb, (uint64_t)p.lda * 2, BKH, 1, box_m));
    DPOT_CALL(tc_encode_map_f16(&mAl, Ah + p.a_lo, (uint64_t)p.K, (uint64_t)batchA, (uint64_t)p.M, sAb, (uint64_t)p.lda * 2, BKH, 1, box_m));
  }

  const int units = g_sm_count / CGn;
  const int grid = CGn * (P.total_tiles < units ? P.total_tiles : units);
  const int am = p.act == DPOT_ACT_NONE ? 0 : (p.act == DPOT_ACT_GELU ? 1 : 2);
  const bool o16 = p.c_fmt != DPOT_FMT_F32;
  const int side = p.c_scale || (p.rowbias && p.residual) ? 3 : (p.rowbias ? 1 : ((p.residual || p.dact_src) ? 2 : 0));
  if (P.g.C_pre && P.g.ldpre <= 0) { P.g.ldpre = p.ldc; P.g.sPre = p.sC; }          // fp32 result: C's geometry
  if (P.g.dact_src && P.g.lddact <= 0) { P.g.lddact = p.ldc; P.g.sDact = p.sC; }
#define DPOT_TC16_LAUNCH(CGV, AM, O16, SD)                                                                             \
  do {                                                                                                                 \
    static DevOnce attr;                                                                                          \
    if (attr.need()) {                                                                                                       \
      DPOT_CUDA(cudaFuncSetAttribute(gemm_tc16_kernel<CGV, AM, O16, SD>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                     (int)Geo<CGV>::SMEM_BYTES));                                                      \
      attr.done();                                                                                                     \
    }                                                                                                                  \
    cudaLaunchConfig_t lc = {};                                                                                        \
    lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(NTHREADS); lc.dynamicSmemBytes = Geo<CGV>::SMEM_BYTES;       \
    lc.stream = st;                                                                                                    \
    cudaLaunchAttribute at[2];                                                                                         \
    at[0].id = cudaLaunchAttributeClusterDimension;                                                                    \
    at[0].val.clusterDim.x = CGV; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;                              \
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                     \
    at[1].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;                                                  \
    lc.attrs = at; lc.numAttrs = 2;                                                                                    \
    DPOT_CUDA(cudaLaunchKernelEx(&lc, gemm_tc16_kernel<CGV, AM, O16, SD>, mWh, mWl, mAh, mAl, P));                     \
  } while (0)
#define DPOT_TC16_SD(CGV, AM, O16)                                                                                     \
  do {                                                                                                                 \
    if (side == 0) DPOT_TC16_LAUNCH(CGV, AM, O16, 0);                                                                  \
    else if (side == 1) DPOT_TC16_LAUNCH(CGV, AM, O16, 1);                                                             \
    else if (side == 2) DPOT_TC16_LAUNCH(CGV, AM, O16, 2);                                                             \
    else DPOT_TC16_LAUNCH(CGV, AM, O16, 3);                                                                            \
  } while (0)
#define DPOT_TC16_O(CGV, AM) do { if (o16) DPOT_TC16_SD(CGV, AM, true); else DPOT_TC16_SD(CGV, AM, false); } while (0)
#define DPOT_TC16_A(CGV) do { if (am == 0) DPOT_TC16_O(CGV, 0); else if (am == 1) DPOT_TC16_O(CGV, 1); else DPOT_TC16_O(CGV, 2); } while (0)
  if (CGn == 2) DPOT_TC16_A(2); else DPOT_TC16_A(1);
#undef DPOT_TC16_A
#undef DPOT_TC16_O
#undef DPOT_TC16_SD
#undef DPOT_TC16_LAUNCH
  DPOT_LAUNCH_CHECK("gemm_tc16_kernel");
  return 0;
}

}  // namespace dpot

extern "C" int dpot_tc16_available(void) { return dpot::tc_device_ok() ? 1 : 0; }
// experiment / test knob: -1 auto, 0 single-CTA tiles only, 1 CTA pairs whenever legal
extern "C" void dpot_tc16_set_pair(int32_t mode) { dpot::g_pair_mode = mode; }
extern "C" void dpot_tc16_set_debug(int32_t mask) { dpot::g_dbg = mask; }
extern "C" int dpot_tc16_set_precision(int32_t mode) { const int prev = dpot::g_single; if (mode >= 0) dpot::g_single = mode ? 1 : 0; return prev; }
