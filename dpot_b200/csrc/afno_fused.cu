// The fused AFNO2D mixer (models/dpot.py:51-110 + the GroupNorm around it, :165-176): ONE kernel per block.
//
//   f = irfft2( W2 * act( W1 * rfft2( GN1(a) ) + b1 ) + b2 ) + GN1(a),   plus the GroupNorm-2 statistics of f
//
// Work unit = (sample b, channel block kappa): 16 x 16 positions x 128 channels.  The spectrum (144 kept modes x 128
// complex channels) and the hidden layer never leave the SM:
//
//   phase A   GroupNorm-1 on load -> rfft2 on the CUDA cores (one thread owns a whole 16-point transform in registers,
//             lanes = channels, two real rows per complex transform, DC / Nyquist columns packed; the row -> column
//             turn goes through a 64 KB shared-memory scratch) -> the spectrum is written as split fp16 (hi | lo*2048)
//             straight into the 128B-swizzled K-major UMMA operand tile X[144 modes][256 = re 128 | im 128]
//   layer 1   tcgen05.mma, kind::f16, accumulators in TMEM.  Weights are the UMMA A operand (M = 128 outputs), the
//             modes the B operand (N = 144): the accumulator is TRANSPOSED, lane = output channel, column = mode.
//             Complex structure: D_re = Wr X_re - Wi X_im, D_im = Wr X_im + Wi X_re -- the SAME two 128 x 128 weight
//             tiles feed both accumulators, the minus sign is the instruction descriptor's negate-A bit.
//   E1        TMEM -> registers -> *1/s + b1 -> act -> split fp16 -> operand tile O1 (over X, which is dead by then)
//   layer 2   the same contraction with W2 on O1
//   E2        a thread owns ONE channel (TMEM lane) and reads 16 consecutive columns = the 16 k1 modes of one k2:
//             the inverse transform along k1 runs straight out of TMEM, then the c2r rows through shared memory
//             (the dead operand tile), + skip (GN1(a) again), store f, GroupNorm-2 statistics.
//
// fp32 parity on fp16 tensor cores: activations are split  x = hi + lo/2048  (DPOT_FMT_HL16 semantics), weights are
// stored per layer as THREE fp16 planes of s*w (s = a power of two that puts max|w| at 2^13..2^14, so even the
// reference's default init of ~1e-5 keeps full precision):  P1 = fp16(s w),  P2 = P1 / 2048 (exact),  P3 = fp16(s w - P1).
// Then  s w x = P1 hi + P2 lo + P3 hi + O(2^-22): three MMAs per product into ONE accumulator (288 TMEM columns per
// layer; two accumulators per output, as the generic engine keeps them, would need 576 > 512 columns).
//
// Warp roles (576 threads, 1 CTA / SM, persistent over units): warps 0-15 compute (FFT / epilogues), warp 16 = weight
// producer (cp.async.bulk of pre-swizzled 16 KB weight chunks through a 3-stage ring; stages 1-2 double as the second
// half of the FFT scratch while phase A runs), warp 17 = TMEM allocator + MMA issuer.
#include "common.cuh"
#include "gemm_common.cuh"
#include "tc_ptx.cuh"
#include "fft_reg.cuh"

#include <cuda_fp16.h>

namespace dpot {
namespace {

constexpr int AF_H = 16, AF_BS = 128, AF_KM2 = 9, AF_MODES = AF_H * AF_KM2;      // 144 kept modes, row m = k2*16 + k1
constexpr int AF_CWARPS = 16, AF_CTHREADS = AF_CWARPS * 32, AF_THREADS = AF_CTHREADS + 64;
constexpr int AF_CHUNKS = 12;                                    // weight chunks per layer: (Wr|Wi) x (k-block 0|1) x 3 planes
constexpr uint32_t AF_PLANE = AF_MODES * 128;                    // one operand plane of one k-block: 144 rows x 128 B
constexpr uint32_t AF_OPER = 8 * AF_PLANE;                       // 4 k-blocks x (hi, lo) = 147456 B
constexpr uint32_t AF_WCHUNK = 128 * 128;                        // 128 output rows x 64 halves = 16 KB
constexpr int AF_NST = 3;
constexpr uint32_t AF_OFF_W = AF_OPER;
constexpr uint32_t AF_OFF_SCR = AF_OFF_W + AF_WCHUNK;            // 64 KB scratch = ring stages 1, 2 + 32 KB of its own
constexpr uint32_t AF_OFF_BAR = AF_OFF_W + AF_NST * AF_WCHUNK + 32768;
constexpr uint32_t AF_SMEM = AF_OFF_BAR + 128 + 1024;
static_assert(AF_SMEM <= 232448, "shared memory budget");
static_assert(AF_OFF_SCR + 65536 == AF_OFF_BAR, "scratch = stages 1-2 + 32 KB");
static_assert(16 * 8 * 128 * 8 <= AF_OPER, "inverse-transform scratch lives in the dead operand tile");

// packed-weight arena of one block (floats): [2 layers][nb][12 chunks][4096] | bias [2][nb][2][128] | inv_s[2] (+ amax scratch)
__host__ __device__ inline int64_t af_chunk_floats() { return AF_WCHUNK / 4; }
__host__ __device__ inline int64_t af_bias_off(int nb) { return (int64_t)2 * nb * AF_CHUNKS * (AF_WCHUNK / 4); }
__host__ __device__ inline int64_t af_scale_off(int nb) { return af_bias_off(nb) + (int64_t)2 * nb * 2 * AF_BS; }
__host__ __device__ inline int64_t af_total_floats(int nb) { return af_scale_off(nb) + 8; }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, __half v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(__half_as_ushort(v)) : "memory");
}
__device__ __forceinline__ void sts_f2(uint32_t addr, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// instruction descriptor: D = f32, A = B = f16, K-major, M = 128, N = 144 (+ negate A)
__host__ __device__ constexpr uint32_t af_idesc(bool neg_a) {
  return (1u << 4) | ((uint32_t)(AF_MODES >> 3) << 17) | ((128u >> 4) << 24) | (neg_a ? (1u << 13) : 0u);
}

struct AfArgs {
  const float* lat;        // [B*256, E] block input (residual stream)
  float* f;                // [B*256, E] output
  const float* packed;     // af_total_floats(nb) floats (dpot_afno_fused_pack)
  double* stats2;          // [B, groups, 2] GroupNorm-2 statistics (accumulated; zeroed by the caller) or nullptr
  GnRef gn;                // GroupNorm-1 by reference
  int B, E, nb, act, groups;
  float* dbg;              // test hook: units x (X operand | O1 operand) images, 2 * AF_OPER bytes per unit, or nullptr
  long long* trace;        // profiling hook: clock64() of CTA 0 / thread 0 at the phase boundaries, 8 per unit, or nullptr
  // GroupNorm-2 applied in the kernel (the unit owns whole groups): n2 = split fp16 of GN2(f), rows [hi E | lo E]; or nullptr
  __half* n2; const float* gamma2; const float* beta2; float eps2;
};

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// TMEM columns: layer 1 accumulates D_re at 0 and D_im at 144; layer 2 accumulates D_re at 0 (layer 1's D_re has been
// consumed by then) and D_im at 288, so that its first half can run while E1 still reads layer 1's D_im.
constexpr uint32_t AF_T_RE = 0, AF_T_IM1 = AF_MODES, AF_T_IM2 = 2 * AF_MODES;

// ESPEC: embed_dim as a compile-time constant (0 = run time): every global load / store of a transform then uses an
// immediate offset instead of 64-bit address arithmetic.
template <int ACT_MODE, int ESPEC>
__global__ void __launch_bounds__(AF_THREADS, 1) afno_fused_kernel(const AfArgs P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const S = smem_raw + (smem0 - smem_u32(smem_raw));        // generic view of the aligned base (shared space)
  const uint32_t bar0 = smem0 + AF_OFF_BAR;
  auto FULL = [&](int s) -> uint32_t { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) -> uint32_t { return bar0 + 8u * (AF_NST + s); };
  const uint32_t XRDY = bar0 + 8u * (2 * AF_NST), M1DONE = XRDY + 8u, O1RE = XRDY + 16u, O1IM = XRDY + 24u, M2DONE = XRDY + 32u;
  const uint32_t tmem_slot = XRDY + 40u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == AF_CWARPS && elect_one()) {
    for (int s = 0; s < AF_NST; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    mbar_init(XRDY, AF_CWARPS); mbar_init(M1DONE, 1); mbar_init(O1RE, AF_CWARPS); mbar_init(O1IM, AF_CWARPS); mbar_init(M2DONE, 1);
    fence_barrier_init();
  }
  if (warp == AF_CWARPS + 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  const int E = ESPEC ? ESPEC : P.E;
  const int nb = ESPEC ? ESPEC / AF_BS : P.nb;
  const int units = P.B * nb;
  const float* const bias_all = P.packed + af_bias_off(nb);
  const float inv_s1 = P.packed[af_scale_off(nb)], inv_s2 = P.packed[af_scale_off(nb) + 1];

  if (warp == AF_CWARPS) {
    // ======================================= weight producer =======================================
    // per unit and layer: 12 chunks of 16 KB, each feeding both accumulators.  (Measured: streaming layer 2's chunks
    // twice so that its Re(hidden) products could overlap E1 made things slower -- the MMA phases already run at the
    // ~30 B/clk/SM the L2 -> SM fabric delivers, 576 tensor-clocks per chunk.)
    if (elect_one()) {
      int s = 0; uint32_t ph = 0; uint32_t it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const int kap = u % nb;
        for (int pass = 0; pass < 2; ++pass) {
          const int layer = pass;
          const float* src = P.packed + ((int64_t)(layer * nb + kap) * AF_CHUNKS) * (AF_WCHUNK / 4);
          for (int ci = 0; ci < AF_CHUNKS; ++ci) {
            // ring stages 1 and 2 are FFT scratch until this unit's spectrum is complete
            if (pass == 0 && ci == 1) mbar_wait(XRDY, it & 1u);
            mbar_wait(EMPTY(s), ph ^ 1u);
            mbar_expect_tx(FULL(s), AF_WCHUNK);
            bulk_g2s(smem0 + AF_OFF_W + (uint32_t)s * AF_WCHUNK, src + (int64_t)ci * (AF_WCHUNK / 4), AF_WCHUNK, FULL(s));
            if (++s == AF_NST) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == AF_CWARPS + 1) {
    // ========================================= MMA issuer ==========================================
    if (elect_one()) {
      int s = 0; uint32_t ph = 0; uint32_t it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        // layer 1 accumulates into (D_re @ 0, D_im @ 144), layer 2 into (D_re @ 0, D_im @ 288)
        for (int layer = 0; layer < 2; ++layer) {
          mbar_wait(layer == 0 ? XRDY : O1IM, it & 1u);
          tc_fence_after();
          const uint32_t d_re = tmem_base + AF_T_RE, d_im = tmem_base + (layer == 0 ? AF_T_IM1 : AF_T_IM2);
          for (int ci = 0; ci < AF_CHUNKS; ++ci) {
            const int T = ci / 6, kb = (ci / 3) & 1, plane = ci % 3;
            mbar_wait(FULL(s), ph);
            tc_fence_after();
            const uint32_t wb = smem0 + AF_OFF_W + (uint32_t)s * AF_WCHUNK;
            const uint32_t bpl = (plane == 1) ? AF_PLANE : 0u;                       // P2 multiplies the lo plane
            const uint32_t x_re = smem0 + (uint32_t)(kb * 2) * AF_PLANE + bpl;        // operand k-block kb      (re part)
            const uint32_t x_im = smem0 + (uint32_t)((2 + kb) * 2) * AF_PLANE + bpl;  // operand k-block 2 + kb  (im part)
            const uint32_t b0 = T == 0 ? x_re : x_im, b1 = T == 0 ? x_im : x_re;
            const uint32_t id0 = af_idesc(T == 1), id1 = af_idesc(false);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t acc = (ci > 0 || ks > 0) ? 1u : 0u;
              const uint64_t ad = make_smem_desc(wb + ks * 32);
              umma_f16(d_re, ad, make_smem_desc(b0 + ks * 32), id0, acc);
              umma_f16(d_im, ad, make_smem_desc(b1 + ks * 32), id1, acc);
            }
            umma_commit(EMPTY(s));
            if (++s == AF_NST) { s = 0; ph ^= 1u; }
          }
          umma_commit(layer == 0 ? M1DONE : M2DONE);
        }
      }
    }
  } else {
    // ======================================= compute warps =========================================
    const int q4 = warp & 3, sl = warp >> 2;          // TMEM lane quarter / slot among the four warps of a quarter
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const int b = u / nb, kap = u % nb;
      const uint32_t par = it & 1u;
      const float* const lat_u = P.lat + (int64_t)b * 256 * E + (int64_t)kap * AF_BS;
      float* const f_u = P.f ? P.f + (int64_t)b * 256 * E + (int64_t)kap * AF_BS : nullptr;   // nullptr: f stays on chip (GroupNorm-2 in the kernel)
      const bool tr = P.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
      long long* const trow_t = P.trace + (int64_t)it * 8;
      if (tr) trow_t[0] = clock64();

      // ---------------- phase A: GroupNorm-1 + rfft2 -> operand tile X (two passes of 64 channels) ----------------
      {
        const int c = (warp & 1) * 32 + lane, task = warp >> 1;                       // channel in pass, task 0..7
        float2* const R = reinterpret_cast<float2*>(S + AF_OFF_SCR);                  // [16 p][8 kc][64 c]
        // second pass of this unit: pull its 64 KB towards L2 while the first pass computes (one 128 B line per thread)
        prefetch_l2(lat_u + (int64_t)(threadIdx.x >> 1) * E + 64 + (threadIdx.x & 1) * 32);
        // 128B-swizzle byte offset of this thread's channel inside a row whose index is j modulo 8, plus that row
        uint32_t swr[8];
#pragma unroll
        for (int jx = 0; jx < 8; ++jx)
          swr[jx] = ((((uint32_t)c >> 3) ^ (uint32_t)jx) << 4) + ((uint32_t)c & 7u) * 2u + (uint32_t)jx * 128u + (uint32_t)task * 2048u;
#pragma unroll 1
        for (int cp = 0; cp < 2; ++cp) {
          const int chb = cp * 64 + c;
          float sc = 1.f, sh = 0.f;
          gn_affine_ref(P.gn, b, kap * AF_BS + chb, sc, sh);
          {   // A1: row pair `task`; the halving of the two-for-one untangling is folded into A2's normalisation
            float zr[16], zi[16];
            const float* ap = lat_u + (int64_t)(2 * task) * 16 * E + chb;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              zr[q] = __ldg(ap + (int64_t)q * E);
              zi[q] = __ldg(ap + (int64_t)(16 + q) * E);
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) { zr[q] = fmaf(zr[q], sc, sh); zi[q] = fmaf(zi[q], sc, sh); }
            fft_reg<16, -1>(zr, zi);
            float2* r0 = R + ((2 * task) * 8) * 64 + c;
            float2* r1 = r0 + 8 * 64;
            r0[0] = make_float2(zr[0], zr[8]);
            r1[0] = make_float2(zi[0], zi[8]);
#pragma unroll
            for (int k = 1; k < 8; ++k) {
              const int kn = 16 - k;
              r0[k * 64] = make_float2(zr[k] + zr[kn], zi[k] - zi[kn]);          // 2 x (spectrum of row 2*task)
              r1[k * 64] = make_float2(zi[k] + zi[kn], zr[kn] - zr[k]);          // 2 x (spectrum of row 2*task + 1)
            }
          }
          named_bar_sync(1, AF_CTHREADS);
          {   // A2: column `task` (0 = DC and Nyquist packed), write split fp16 into the swizzled operand tile
            float zr[16], zi[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const float2 v = R[(p * 8 + task) * 64 + c];
              zr[p] = v.x; zi[p] = v.y;
            }
            fft_reg<16, -1>(zr, zi);
            uint8_t* const re_hi = S + (uint32_t)(cp * 2) * AF_PLANE;                  // k-block cp (re); im: k-block 2 + cp
            constexpr uint32_t IM = 4u * AF_PLANE, LO = AF_PLANE;
            if (task > 0) {
              const float norm = 0.5f / 16.0f;
#pragma unroll
              for (int k1 = 0; k1 < 16; ++k1) {
                uint8_t* const o = re_hi + swr[k1 & 7] + (k1 >> 3) * 1024;
                __half hi, lo;
                hl_split(zr[k1] * norm, hi, lo);
                *reinterpret_cast<__half*>(o) = hi; *reinterpret_cast<__half*>(o + LO) = lo;
                hl_split(zi[k1] * norm, hi, lo);
                *reinterpret_cast<__half*>(o + IM) = hi; *reinterpret_cast<__half*>(o + IM + LO) = lo;
              }
            } else {
              // F = FFT(col_0 + i col_8) -> X_0[k1] = (F[k1] + conj F[-k1]) / 2,  X_8[k1] = (F[k1] - conj F[-k1]) / 2i
              const float norm = 0.5f / 16.0f;
#pragma unroll
              for (int k1 = 0; k1 < 16; ++k1) {
                const int kn = (16 - k1) & 15;
                uint8_t* const o = re_hi + swr[k1 & 7] + (k1 >> 3) * 1024;           // task == 0: row k1 (k2 = 0)
                __half hi, lo;
                hl_split((zr[k1] + zr[kn]) * norm, hi, lo);
                *reinterpret_cast<__half*>(o) = hi; *reinterpret_cast<__half*>(o + LO) = lo;
                hl_split((zi[k1] - zi[kn]) * norm, hi, lo);
                *reinterpret_cast<__half*>(o + IM) = hi; *reinterpret_cast<__half*>(o + IM + LO) = lo;
                uint8_t* const o8 = o + 8 * 2048;                                    // row 128 + k1 (k2 = 8)
                hl_split((zi[k1] + zi[kn]) * norm, hi, lo);
                *reinterpret_cast<__half*>(o8) = hi; *reinterpret_cast<__half*>(o8 + LO) = lo;
                hl_split((zr[kn] - zr[k1]) * norm, hi, lo);
                *reinterpret_cast<__half*>(o8 + IM) = hi; *reinterpret_cast<__half*>(o8 + IM + LO) = lo;
              }
            }
          }
          if (cp == 0) named_bar_sync(1, AF_CTHREADS);       // the scratch is rewritten by the second pass
        }
      }
      tc_fence_before();            // this warp's TMEM reads of the previous unit precede the new MMAs
      fence_proxy_async();          // generic-proxy operand writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(XRDY);
      if (tr) trow_t[1] = clock64();

      if (P.dbg) {                  // test hook: image of the X operand tile (E1 overwrites it: barrier on both sides)
        named_bar_sync(1, AF_CTHREADS);
        float* dst = P.dbg + (int64_t)u * (2 * AF_OPER / 4);
        for (uint32_t i = threadIdx.x; i < AF_OPER / 4; i += AF_CTHREADS) dst[i] = reinterpret_cast<const float*>(S)[i];
        named_bar_sync(1, AF_CTHREADS);
      }
      // next unit of this CTA: pull its first 64 channels towards L2 while the tensor core works (one line per thread)
      if (u + (int)gridDim.x < units) {
        const int un = u + (int)gridDim.x;
        prefetch_l2(P.lat + (int64_t)(un / nb) * 256 * E + (int64_t)(un % nb) * AF_BS + (int64_t)(threadIdx.x >> 1) * E + (threadIdx.x & 1) * 32);
      }

      // ---------------- E1: layer-1 accumulators -> + b1 -> act -> split fp16 -> operand tile O1 ----------------
      // All 16 warps convert the real half, then the imaginary half; slot sl of a lane quarter owns 36 of the 144 modes.
      {
        const int j = q4 * 32 + lane;
        const uint32_t chunk = (uint32_t)((q4 & 1) * 4 + (lane >> 3));
        const int mb = sl * 36;                                                       // first mode; mb & 7 = 4 * (sl & 1)
        uint32_t swr[8];
#pragma unroll
        for (int jx = 0; jx < 8; ++jx) {
          const uint32_t rj = (uint32_t)((jx + 4 * (sl & 1)) & 7);                    // row index modulo 8 of local row jx
          swr[jx] = ((chunk ^ rj) << 4) + ((uint32_t)lane & 7u) * 2u + (uint32_t)(mb + jx) * 128u;
        }
        mbar_wait(M1DONE, par);
        tc_fence_after();
        if (tr) trow_t[2] = clock64();
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
          const float bias = __ldg(bias_all + ((int64_t)(0 * nb + kap) * 2 + part) * AF_BS + j);
          uint8_t* const o_hi = S + (uint32_t)((part * 2 + (q4 >> 1)) * 2) * AF_PLANE;
          const uint32_t tcol = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(part * AF_MODES + mb);
          uint32_t v[36];
          {
            uint32_t (&v0)[16] = *reinterpret_cast<uint32_t (*)[16]>(&v[0]);
            uint32_t (&v1)[16] = *reinterpret_cast<uint32_t (*)[16]>(&v[16]);
            uint32_t (&v2)[4] = *reinterpret_cast<uint32_t (*)[4]>(&v[32]);
            tmem_ld16(tcol, v0);
            tmem_ld16(tcol + 16u, v1);
            tmem_ld4(tcol + 32u, v2);
            tmem_ld_wait();
          }
#pragma unroll
          for (int i = 0; i < 36; ++i) {
            float x = fmaf(__uint_as_float(v[i]), inv_s1, bias);
            x = ACT_MODE == 1 ? gelu_select(x) : act_apply(x, P.act);
            __half hi, lo;
            hl_split(x, hi, lo);
            uint8_t* const o = o_hi + swr[i & 7] + (i >> 3) * 1024;
            *reinterpret_cast<__half*>(o) = hi;
            *reinterpret_cast<__half*>(o + AF_PLANE) = lo;
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(part == 0 ? O1RE : O1IM);
        }
      }
      if (tr) trow_t[3] = clock64();

      if (P.dbg) {                  // test hook: image of the O1 operand tile (read while layer 2 reads it, before E2 reuses it)
        named_bar_sync(1, AF_CTHREADS);
        float* dst = P.dbg + (int64_t)u * (2 * AF_OPER / 4) + AF_OPER / 4;
        for (uint32_t i = threadIdx.x; i < AF_OPER / 4; i += AF_CTHREADS) dst[i] = reinterpret_cast<const float*>(S)[i];
        named_bar_sync(1, AF_CTHREADS);
      }

      // ---------------- E2: layer-2 accumulators -> inverse transforms -> + skip -> f, GroupNorm-2 statistics -----
      const int j = q4 * 32 + lane;                                                   // channel in block = TMEM lane
      float2* const Z = reinterpret_cast<float2*>(S);                                 // [16 p][8 kc][128 j] over the dead operand
      // the skip rows of this thread's first row pair: requested before the wait, consumed in E2b
      mbar_wait(M2DONE, par);
      tc_fence_after();
      if (tr) trow_t[4] = clock64();
      {
        const float b2r = __ldg(bias_all + ((int64_t)(1 * nb + kap) * 2 + 0) * AF_BS + j);
        const float b2i = __ldg(bias_all + ((int64_t)(1 * nb + kap) * 2 + 1) * AF_BS + j);
        const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16);
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int kc = sl + 4 * t;
          float zr[16], zi[16];
          {
            uint32_t rr[16], ri[16];
            tmem_ld16(trow + AF_T_RE + (uint32_t)(kc * 16), rr);
            tmem_ld16(trow + AF_T_IM2 + (uint32_t)(kc * 16), ri);
            tmem_ld_wait();
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
              zr[k1] = fmaf(__uint_as_float(rr[k1]), inv_s2, b2r);
              zi[k1] = fmaf(__uint_as_float(ri[k1]), inv_s2, b2i);
            }
          }
          if (kc == 0) {
            // Nyquist column (k2 = 8), then pack the Hermitian parts: W = Herm(col_0) + i Herm(col_8)
            // (torch.fft.irfft2 ignores the imaginary part of the k1-inverse of these two columns)
            uint32_t rr[16], ri[16];
            tmem_ld16(trow + AF_T_RE + 128u, rr);
            tmem_ld16(trow + AF_T_IM2 + 128u, ri);
            tmem_ld_wait();
            float wr[16], wi[16];
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
              const int kn = (16 - k1) & 15;
              const float yr = fmaf(__uint_as_float(rr[k1]), inv_s2, b2r), yrn = fmaf(__uint_as_float(rr[kn]), inv_s2, b2r);
              const float yi = fmaf(__uint_as_float(ri[k1]), inv_s2, b2i), yin = fmaf(__uint_as_float(ri[kn]), inv_s2, b2i);
              const float s0r = 0.5f * (zr[k1] + zr[kn]), s0i = 0.5f * (zi[k1] - zi[kn]);
              const float s8r = 0.5f * (yr + yrn), s8i = 0.5f * (yi - yin);
              wr[k1] = s0r - s8i;
              wi[k1] = s0i + s8r;
            }
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) { zr[k1] = wr[k1]; zi[k1] = wi[k1]; }
          }
          fft_reg<16, +1>(zr, zi);
          float2* zo = Z + kc * 128 + j;
#pragma unroll
          for (int p = 0; p < 16; ++p) zo[p * 8 * 128] = make_float2(zr[p], zi[p]);
        }
      }
      tc_fence_before();
      named_bar_sync(1, AF_CTHREADS);                      // Z complete
      if (tr) trow_t[5] = clock64();
      {
        float sc = 1.f, sh = 0.f;
        gn_affine_ref(P.gn, b, kap * AF_BS + j, sc, sh);
        float s1 = 0.f, s2 = 0.f;
        const float norm = 1.0f / 16.0f;
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int pr = sl + 4 * t;
          const float* ap = lat_u + (int64_t)(2 * pr) * 16 * E + j;
          float k0[16], k1v[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            k0[q] = __ldg(ap + (int64_t)q * E);
            k1v[q] = __ldg(ap + (int64_t)(16 + q) * E);
          }
          float zr[16], zi[16];
          const float2* za = Z + ((2 * pr) * 8) * 128 + j;
          const float2* zb = za + 8 * 128;
          {
            const float2 A = za[0], Bv = zb[0];
            zr[0] = A.x; zi[0] = Bv.x;                 // c2r ignores Im of the DC and Nyquist columns
            zr[8] = A.y; zi[8] = Bv.y;
          }
#pragma unroll
          for (int k = 1; k < 8; ++k) {
            const float2 A = za[k * 128], Bv = zb[k * 128];
            zr[k] = A.x - Bv.y; zi[k] = A.y + Bv.x;
            zr[16 - k] = A.x + Bv.y; zi[16 - k] = -A.y + Bv.x;
          }
          fft_reg<16, +1>(zr, zi);
          float* fp = f_u + (int64_t)(2 * pr) * 16 * E + j;
          // GroupNorm-2 in the kernel: the two output rows go back into this thread's OWN 32 slots of Z (the only ones it
          // has just read; no other thread touches them) and are normalised from there once the unit's statistics are known
          float2* const zs0 = Z + ((2 * pr) * 8) * 128 + j;
          float2* const zs1 = zs0 + 8 * 128;
#pragma unroll
          for (int q = 0; q < 16; q += 2) {
            const float v0a = fmaf(zr[q], norm, fmaf(k0[q], sc, sh)), v0b = fmaf(zr[q + 1], norm, fmaf(k0[q + 1], sc, sh));
            const float v1a = fmaf(zi[q], norm, fmaf(k1v[q], sc, sh)), v1b = fmaf(zi[q + 1], norm, fmaf(k1v[q + 1], sc, sh));
            if (P.f) {
              fp[(int64_t)q * E] = v0a; fp[(int64_t)(q + 1) * E] = v0b;
              fp[(int64_t)(16 + q) * E] = v1a; fp[(int64_t)(17 + q) * E] = v1b;
            }
            if (P.n2) { zs0[(q >> 1) * 128] = make_float2(v0a, v0b); zs1[(q >> 1) * 128] = make_float2(v1a, v1b); }
            s1 += (v0a + v0b) + (v1a + v1b);
            s2 = fmaf(v0a, v0a, fmaf(v0b, v0b, fmaf(v1a, v1a, fmaf(v1b, v1b, s2))));
          }
        }
        if (P.stats2 || P.n2) { s1 = warp_sum(s1); s2 = warp_sum(s2); }
        if (P.stats2) {        // the warp's 32 channels lie in one group (host guarantees (E / groups) % 32 == 0)
          if (lane == 0) {
            double* dst = P.stats2 + ((int64_t)b * P.groups + (kap * AF_BS + q4 * 32) / (E / P.groups)) * 2;
            atomicAdd(dst, (double)s1);
            atomicAdd(dst + 1, (double)s2);
          }
        }
        if (P.n2) {
          // GroupNorm-2 + fp16 split of the channel-MLP input in the same launch: a unit covers whole groups (group size
          // 32 / 64 / 128 channels of its 128), so its statistics are complete once the 16 warps have met; every thread
          // then takes its 64 values back from shared memory (stashed above) and writes them normalised and split --
          // f itself need not go to global memory at all (P.f == nullptr: inference, nobody else reads it).
          float* red = reinterpret_cast<float*>(S + 131072);       // the 16 KB of the operand tile beyond Z
          if (lane == 0) { red[2 * warp] = s1; red[2 * warp + 1] = s2; }
          named_bar_sync(1, AF_CTHREADS);
          const int gs = E / P.groups, qpg = gs >> 5, q0 = (q4 / qpg) * qpg;
          double t1 = 0.0, t2 = 0.0;
          for (int qq = q0; qq < q0 + qpg; ++qq)
#pragma unroll
            for (int ss = 0; ss < 4; ++ss) { t1 += (double)red[2 * (ss * 4 + qq)]; t2 += (double)red[2 * (ss * 4 + qq) + 1]; }
          const double inv_cnt = 1.0 / ((double)gs * 256.0);
          const double mean = t1 * inv_cnt;
          const double var = fma(-mean, mean, t2 * inv_cnt);
          const float vv = fmaxf((float)var, 0.f) + P.eps2;
          float rstd = rsqrtf(vv);
          rstd = rstd * fmaf(-0.5f * vv * rstd, rstd, 1.5f);
          const int ch = kap * AF_BS + j;
          const float sc2 = rstd * __ldg(P.gamma2 + ch), sh2 = fmaf(-(float)mean, sc2, __ldg(P.beta2 + ch));
#pragma unroll 1
          for (int t = 0; t < 2; ++t) {
            const int pr = sl + 4 * t;
            const float2* const zs0 = Z + ((2 * pr) * 8) * 128 + j;
            const float2* const zs1 = zs0 + 8 * 128;
            __half* np_ = P.n2 + ((int64_t)b * 256 + (2 * pr) * 16) * (2 * (int64_t)E) + ch;
            float v0[16], v1[16];
#pragma unroll
            for (int q = 0; q < 16; q += 2) {
              const float2 a0 = zs0[(q >> 1) * 128], a1 = zs1[(q >> 1) * 128];
              v0[q] = a0.x; v0[q + 1] = a0.y; v1[q] = a1.x; v1[q + 1] = a1.y;
            }
            // adjacent lanes own adjacent channels: the even lane stores (hi_j, hi_j+1) into the hi plane, the odd lane
            // (lo_j-1, lo_j) into the lo plane -- one 4-byte store per value instead of two 2-byte stores
            uint32_t* const np32 = reinterpret_cast<uint32_t*>((lane & 1) ? np_ + E - 1 : np_);
            auto put = [&](float v, int row) {
              __half hi, lo;
              hl_split(fmaf(v, sc2, sh2), hi, lo);
              const uint32_t w = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
              const uint32_t pw = __shfl_xor_sync(0xffffffffu, w, 1);
              np32[(int64_t)row * E] = (lane & 1) ? __byte_perm(pw, w, 0x7632) : __byte_perm(w, pw, 0x5410);   // row stride 2E halves = E words
            };
#pragma unroll
            for (int q = 0; q < 16; ++q) { put(v0[q], q); put(v1[q], 16 + q); }
          }
        }
      }
      named_bar_sync(1, AF_CTHREADS);                      // Z (the operand tile) is free for the next unit's spectrum
      if (tr) trow_t[6] = clock64();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AF_CWARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- weight packing --------------------------------------------------------------------------------------------
__global__ void af_amax_kernel(const float* __restrict__ w, int64_t n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));      // non-negative floats order like their bit patterns
}

// one thread per (layer, kappa, T, kb, output row o, 16-byte chunk of 8 k): writes the three planes
__global__ void af_pack_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                               const float* __restrict__ b2, int nb, float* __restrict__ packed) {
  const unsigned int* amax = reinterpret_cast<const unsigned int*>(packed + af_scale_off(nb) + 2);
  const int64_t total = (int64_t)2 * nb * 4 * 128 * 8;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx < 2) {
    const float a = __uint_as_float(amax[idx]);
    int e = 0;
    if (a > 0.f && isfinite(a)) frexpf(a, &e);           // a = m * 2^e, m in [0.5, 1)  ->  a * 2^(14 - e) in [2^13, 2^14)
    else e = 14;
    packed[af_scale_off(nb) + idx] = ldexpf(1.0f, e - 14);       // 1 / s
  }
  if (idx < (int64_t)2 * nb * 2 * AF_BS) {                       // biases [layer][kappa][part][j] <- b[part][kappa][j]
    const int jj = (int)(idx % AF_BS), part = (int)((idx / AF_BS) % 2), kap = (int)((idx / (2 * AF_BS)) % nb), layer = (int)(idx / (2 * AF_BS * nb));
    const float* bsrc = layer == 0 ? b1 : b2;
    packed[af_bias_off(nb) + idx] = bsrc[((int64_t)part * nb + kap) * AF_BS + jj];
  }
  if (idx >= total) return;
  const int ck = (int)(idx % 8), o = (int)((idx / 8) % 128), tk = (int)((idx / 1024) % 4), kap = (int)((idx / 4096) % nb),
            layer = (int)(idx / (4096 * (int64_t)nb));
  const int T = tk >> 1, kb = tk & 1;
  const float* w = layer == 0 ? w1 : w2;                         // [2][nb][bs in][bs out]
  const float a = __uint_as_float(amax[layer]);
  int e = 14;
  if (a > 0.f && isfinite(a)) frexpf(a, &e);
  const float s = ldexpf(1.0f, 14 - e);
  __half p1[8], p2[8], p3[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int i = kb * 64 + ck * 8 + t;
    const float v = s * w[(((int64_t)T * nb + kap) * AF_BS + i) * AF_BS + o];
    p1[t] = __float2half_rn(v);
    const float h = __half2float(p1[t]);
    p2[t] = __float2half_rn(h * (1.0f / 2048.0f));
    p3[t] = __float2half_rn(v - h);
  }
  const int64_t chunk0 = ((int64_t)(layer * nb + kap) * AF_CHUNKS + (T * 2 + kb) * 3) * (AF_WCHUNK / 2);    // halves
  const uint32_t off = (uint32_t)o * 64u + ((((uint32_t)ck) ^ ((uint32_t)o & 7u)) << 3);                     // halves, 128B swizzle
  __half* dst = reinterpret_cast<__half*>(packed);
  *reinterpret_cast<uint4*>(dst + chunk0 + off) = *reinterpret_cast<const uint4*>(p1);
  *reinterpret_cast<uint4*>(dst + chunk0 + (AF_WCHUNK / 2) + off) = *reinterpret_cast<const uint4*>(p2);
  *reinterpret_cast<uint4*>(dst + chunk0 + 2 * (AF_WCHUNK / 2) + off) = *reinterpret_cast<const uint4*>(p3);
}

int g_fused_mode = -1;     // -1 auto (fused when supported), 0 never
long long* g_fused_trace = nullptr;

}  // namespace
}  // namespace dpot

using namespace dpot;

extern "C" void dpot_afno_set_fused(int32_t mode) { dpot::g_fused_mode = mode; }
extern "C" void dpot_afno_fused_set_trace(long long* dev_buf) { dpot::g_fused_trace = dev_buf; }

extern "C" int dpot_afno_fused_supported(int32_t h, int32_t E, int32_t nb, int32_t km1, int32_t km2, int32_t groups) {
  if (dpot::g_fused_mode == 0 || !tc_device_ok()) return 0;
  if (h != AF_H || nb <= 0 || E != nb * AF_BS || km1 != AF_H || km2 != AF_KM2) return 0;
  if (groups <= 0 || E % groups != 0 || (E / groups) % 32 != 0) return 0;
  return 1;
}

extern "C" int64_t dpot_afno_fused_packed_floats(int32_t nb) { return nb > 0 ? af_total_floats(nb) : -1; }

extern "C" int dpot_afno_fused_pack(const float* w1, const float* b1, const float* w2, const float* b2, int32_t nb,
                                    int32_t bs, float* packed, void* stream) {
  DPOT_REQUIRE(w1 && b1 && w2 && b2 && packed && nb > 0, DPOT_E_BADARG, "dpot_afno_fused_pack: null pointer / bad nb");
  DPOT_REQUIRE(bs == AF_BS, DPOT_E_UNSUPPORTED, "dpot_afno_fused_pack: block size %d (needs %d)", bs, AF_BS);
  DPOT_REQUIRE(reinterpret_cast<uintptr_t>(packed) % 16 == 0, DPOT_E_ALIGN, "dpot_afno_fused_pack: packed must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  unsigned int* amax = reinterpret_cast<unsigned int*>(packed + af_scale_off(nb) + 2);
  DPOT_CUDA(cudaMemsetAsync(amax, 0, 2 * sizeof(unsigned int), st));
  const int64_t nw = (int64_t)2 * nb * AF_BS * AF_BS;
  af_amax_kernel<<<64, 256, 0, st>>>(w1, nw, amax);
  DPOT_LAUNCH_CHECK("af_amax_kernel");
  af_amax_kernel<<<64, 256, 0, st>>>(w2, nw, amax + 1);
  DPOT_LAUNCH_CHECK("af_amax_kernel");
  const int64_t total = (int64_t)2 * nb * 4 * 128 * 8;
  af_pack_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(w1, b1, w2, b2, nb, packed);
  DPOT_LAUNCH_CHECK("af_pack_kernel");
  return 0;
}

extern "C" int dpot_afno_fused_gn2(const float* lat, const double* stats1, const float* gamma1, const float* beta1,
                                   int32_t groups, float eps, int32_t B, int32_t h, int32_t E, int32_t nb, const float* packed,
                                   int32_t act, float* f, double* stats2, float* dbg, void* n2_16, const float* gamma2,
                                   const float* beta2, float eps2, void* stream);
extern "C" int dpot_afno_fused(const float* lat, const double* stats1, const float* gamma1, const float* beta1,
                               int32_t groups, float eps, int32_t B, int32_t h, int32_t E, int32_t nb, const float* packed,
                               int32_t act, float* f, double* stats2, float* dbg, void* stream) {
  return dpot_afno_fused_gn2(lat, stats1, gamma1, beta1, groups, eps, B, h, E, nb, packed, act, f, stats2, dbg, nullptr, nullptr,
                             nullptr, 0.f, stream);
}
extern "C" int dpot_afno_fused_gn2(const float* lat, const double* stats1, const float* gamma1, const float* beta1,
                                   int32_t groups, float eps, int32_t B, int32_t h, int32_t E, int32_t nb, const float* packed,
                                   int32_t act, float* f, double* stats2, float* dbg, void* n2_16, const float* gamma2,
                                   const float* beta2, float eps2, void* stream) {
  DPOT_REQUIRE(lat && stats1 && gamma1 && beta1 && packed && (f || n2_16), DPOT_E_BADARG, "dpot_afno_fused: null pointer");
  DPOT_REQUIRE(lat != f, DPOT_E_BADARG, "dpot_afno_fused: in-place operation is not supported (the skip term re-reads the input)");
  DPOT_REQUIRE(B > 0 && dpot_afno_fused_supported(h, E, nb, h, h / 2 + 1, groups), DPOT_E_UNSUPPORTED,
               "dpot_afno_fused: geometry h=%d E=%d nb=%d groups=%d is not served by the fused mixer", h, E, nb, groups);
  DPOT_REQUIRE(reinterpret_cast<uintptr_t>(packed) % 16 == 0, DPOT_E_ALIGN, "dpot_afno_fused: packed must be 16-byte aligned");
  DPOT_REQUIRE(!n2_16 || (gamma2 && beta2 && AF_BS % (E / groups) == 0 && (E / groups) % 32 == 0), DPOT_E_BADARG,
               "dpot_afno_fused_gn2: GroupNorm-2 in the kernel needs gamma2 / beta2 and a group size of 32, 64 or 128 channels");
  AfArgs P;
  P.lat = lat; P.f = f; P.packed = packed; P.stats2 = stats2; P.dbg = dbg; P.trace = g_fused_trace;
  P.n2 = reinterpret_cast<__half*>(n2_16); P.gamma2 = gamma2; P.beta2 = beta2; P.eps2 = eps2;
  P.gn = make_gn_ref(stats1, gamma1, beta1, groups, eps, E, (int64_t)h * h);
  P.B = B; P.E = E; P.nb = nb; P.act = act; P.groups = groups;
  const int units = B * nb, sms = sm_count_cur();
  const int grid = units < sms ? units : sms;
  cudaStream_t st = as_stream(stream);
#define DPOT_AF_LAUNCH(AM, ES)                                                                                         \
  do {                                                                                                                 \
    static DevOnce attr;                                                                                               \
    if (attr.need()) {                                                                                                 \
      DPOT_CUDA(cudaFuncSetAttribute(afno_fused_kernel<AM, ES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AF_SMEM)); \
      attr.done();                                                                                                     \
    }                                                                                                                  \
    DPOT_CUDA(launch_pdl(afno_fused_kernel<AM, ES>, dim3(grid), dim3(AF_THREADS), AF_SMEM, st, P));                    \
  } while (0)
#define DPOT_AF_E(AM) do { if (E == 1024) DPOT_AF_LAUNCH(AM, 1024); else if (E == 512) DPOT_AF_LAUNCH(AM, 512); else DPOT_AF_LAUNCH(AM, 0); } while (0)
  if (act == DPOT_ACT_GELU) DPOT_AF_E(1); else DPOT_AF_E(2);
#undef DPOT_AF_E
#undef DPOT_AF_LAUNCH
  DPOT_LAUNCH_CHECK("afno_fused_kernel");
  return 0;
}
