// Output tail on tcgen05 (models/dpot.py:317-321,397-401):  y2 = act(W2 y1 + b2),  y3 = W4 y2 + b4,  pixel shuffle.
//
// The warp-MMA version (out_tail_mma.cu) is bound by the legacy mma.sync issue rate (~1/8 of the tcgen05 MAC rate on
// B200).  Here the per-pixel 32 -> 32 layer is a UMMA with PIXELS as the M dimension (128 TMEM lanes =
// 128 pixels, so an epilogue thread owns all 32 hidden channels of its pixel) and the 32 x 32 weight as the N operand:
//   * the ConvTranspose GEMM writes y1 as DPOT_FMT_HL16G32: one 128-byte record [hi 32 | lo 32] per pixel, i.e. a dense
//     [pixels x 64 halves] matrix -- a plain K-major SWIZZLE_128B operand with K' = 64 that TMA loads directly;
//   * with A' = [ahi | alo] the two accumulators are D1 = A'[:, :32] Whi^T (2 k-steps) and D2 = A' [Wlo | Whi]^T
//     (4 k-steps): both lo-order products in ONE pass; y = D1 + D2/2048 as everywhere else (gemm_tc16.cu);
//   * 64 TMEM columns per 128-pixel tile -> EIGHT accumulator buffers, an 8-stage TMA ring of 16 KB tiles;
//   * the epilogue thread adds b2, applies the activation to its 32 values and does the 32 -> nout projection with
//     exact fp32 FMAs against W4 from shared memory (128 FMAs: not worth a second MMA + split), then stores the pixel
//     (out[B,X,Y,nout], or straight into the autoregressive ring window / prediction tensor).
#include "common.cuh"
#include "gemm_common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

namespace dpot {
int g_tail_tc = 1;     // 0 = never use the tcgen05 tail (dpot_out_tail_set_engine)
namespace {

constexpr int OT_STAGES = 8, OT_NBUF = 8, OT_ROWS = 128;
constexpr int OT_EPI_WARP0 = 4, OT_EPI_WARPS = 12, OT_EPI_CLASSES = OT_EPI_WARPS / 4;
constexpr int OT_NTHREADS = 32 * (OT_EPI_WARP0 + OT_EPI_WARPS);
constexpr uint32_t OT_TILE = OT_ROWS * 128;                 // 16 KB: 128 pixels x [hi 32 | lo 32] halves
constexpr uint32_t OT_W_OFF = OT_STAGES * OT_TILE;          // two 4 KB weight operands: [Whi | 0], [Wlo | Whi]
constexpr uint32_t OT_P_OFF = OT_W_OFF + 2 * 4096;          // fp32 parameters: b2[32], W4^T[32][8], b4[8]
constexpr uint32_t OT_BAR_OFF = OT_P_OFF + (32 + 256 + 8) * 4;
constexpr uint32_t OT_SMEM = OT_BAR_OFF + 512 + 1024;

struct OtRing { float* ring; float* pred; int T, slot0, Ttot, step; };

struct OtParams {
  const float* w2; const float* b2; const float* w4; const float* b4;
  const float* mu; const float* sigma;
  float* out; OtRing rg;
  int B, h, w, P, nout, act, Co;
  int64_t npix; int ntiles;
};

template <int ACT_MODE, bool RING>
__global__ void __launch_bounds__(OT_NTHREADS, 1)
out_tail_tc_kernel(const __grid_constant__ CUtensorMap mapY, const OtParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const sm = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t bar0 = smem0 + OT_BAR_OFF;
  auto FULL = [&](int s) -> uint32_t { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) -> uint32_t { return bar0 + 8u * (OT_STAGES + s); };
  auto TFULL = [&](int b) -> uint32_t { return bar0 + 8u * (2 * OT_STAGES + b); };
  auto TEMPTY = [&](int b) -> uint32_t { return bar0 + 8u * (2 * OT_STAGES + OT_NBUF + b); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * OT_STAGES + 2 * OT_NBUF);
  float* const prm = reinterpret_cast<float*>(sm + OT_P_OFF);     // b2[32] | W4T[32][8] | b4[8]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  pdl_launch_dependents();
  if (warp == 0 && elect_one()) tma_prefetch_desc(&mapY);
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < OT_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int b = 0; b < OT_NBUF; ++b) { mbar_init(TFULL(b), 1); mbar_init(TEMPTY(b), 4); }   // 4 warps read one buffer
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  // ---- weights -> the two K' = 64 operands in the 128B-swizzled K-major layout (row n = 128 B, 16-byte chunk c at
  // c ^ (n & 7)); parameters only, so this runs before pdl_wait()
  for (int e = tid; e < 32 * 32; e += OT_NTHREADS) {
    const int n = e >> 5, k = e & 31;
    __half hi, lo;
    hl_split(P.w2[n * 32 + k], hi, lo);
    auto put = [&](int op, int kk, __half v) {
      const uint32_t off = (uint32_t)n * 128u + ((((uint32_t)kk >> 3) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)kk & 7u) * 2u;
      *reinterpret_cast<__half*>(sm + OT_W_OFF + op * 4096 + off) = v;
    };
    put(0, k, hi); put(0, 32 + k, __float2half_rn(0.f));     // [Whi | 0]
    put(1, k, lo); put(1, 32 + k, hi);                        // [Wlo | Whi]
  }
  for (int e = tid; e < 32; e += OT_NTHREADS) prm[e] = P.b2[e];
  for (int e = tid; e < 32 * 8; e += OT_NTHREADS) { const int k = e >> 3, c = e & 7; prm[32 + e] = c < P.nout ? P.w4[c * 32 + k] : 0.f; }
  for (int e = tid; e < 8; e += OT_NTHREADS) prm[32 + 256 + e] = e < P.nout ? P.b4[e] : 0.f;
  fence_proxy_async();            // generic-proxy writes of the weight operands -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        mbar_wait(EMPTY(s), ph ^ 1);
        mbar_expect_tx(FULL(s), OT_TILE);
        tma_load_3d(smem0 + (uint32_t)s * OT_TILE, &mapY, FULL(s), 0, tile * OT_ROWS, 0);     // dims (k' = 64, pixel, 1)
        if (++s == OT_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16 -> f32, M = 128, N = 32
      int s = 0; uint32_t ph = 0; uint32_t tc = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++tc) {
        const uint32_t buf = tc % OT_NBUF, bph = (tc / OT_NBUF) & 1u;
        mbar_wait(TEMPTY(buf), bph ^ 1);
        mbar_wait(FULL(s), ph);
        tc_fence_after();
        const uint32_t ab = smem0 + (uint32_t)s * OT_TILE, wa = smem0 + OT_W_OFF, wb = wa + 4096;
        const uint32_t d1 = tmem_base + buf * 64u, d2 = d1 + 32u;
#pragma unroll
        for (int k4 = 0; k4 < 2; ++k4)      // D1 = ahi Whi^T : the hi half of the record only
          umma_f16(d1, make_smem_desc(ab + k4 * 32), make_smem_desc(wa + k4 * 32), idesc, k4 > 0 ? 1u : 0u);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)      // D2 = [ahi | alo] [Wlo | Whi]^T
          umma_f16(d2, make_smem_desc(ab + k4 * 32), make_smem_desc(wb + k4 * 32), idesc, k4 > 0 ? 1u : 0u);
        umma_commit(EMPTY(s));
        umma_commit(TFULL(buf));
        if (++s == OT_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= OT_EPI_WARP0) {
    // ================================ epilogue ====================================
    const int quarter = warp & 3;                         // TMEM lane quarter = pixels 32q .. 32q+31 of the tile
    const int par = (warp - OT_EPI_WARP0) >> 2;           // this warp takes every OT_EPI_CLASSES-th tile of the CTA's sequence
    const int X = P.h * P.P, Y = P.w * P.P, PP = P.P * P.P;
    const float4* w4t = reinterpret_cast<const float4*>(prm + 32);
    uint32_t tc = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++tc) {
      if ((int)(tc % OT_EPI_CLASSES) != par) continue;
      const uint32_t buf = tc % OT_NBUF, bph = (tc / OT_NBUF) & 1u;
      mbar_wait(TFULL(buf), bph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + buf * 64u + ((uint32_t)(quarter * 32) << 16);
      uint32_t r1[32], r2[32];
      tmem_ld32(t_row, r1);
      tmem_ld32(t_row + 32u, r2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(TEMPTY(buf));           // accumulators are in registers: the buffer is free again
      float y3[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) y3[c] = prm[32 + 256 + c];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float v = fmaf(__uint_as_float(r2[k]), HL_INV, __uint_as_float(r1[k])) + prm[k];
        v = ACT_MODE == 1 ? gelu_fast(v) : act_apply(v, P.act);
        const float4 wlo = w4t[2 * k], whi = w4t[2 * k + 1];
        y3[0] = fmaf(v, wlo.x, y3[0]); y3[1] = fmaf(v, wlo.y, y3[1]); y3[2] = fmaf(v, wlo.z, y3[2]); y3[3] = fmaf(v, wlo.w, y3[3]);
        if (P.nout > 4) {
          y3[4] = fmaf(v, whi.x, y3[4]); y3[5] = fmaf(v, whi.y, y3[5]); y3[6] = fmaf(v, whi.z, y3[6]); y3[7] = fmaf(v, whi.w, y3[7]);
        }
      }
      // ---- pixel shuffle store: pix = ((b*h + p)*w + q)*P*P + u*P + v
      const int64_t pix = (int64_t)tile * OT_ROWS + quarter * 32 + lane;
      if (pix >= P.npix) continue;
      const uint32_t patch = (uint32_t)(pix / PP), uv = (uint32_t)(pix - (int64_t)patch * PP);
      const uint32_t q = patch % (uint32_t)P.w, bp = patch / (uint32_t)P.w;
      const uint32_t p = bp % (uint32_t)P.h, b = bp / (uint32_t)P.h;
      const uint32_t u = uv / (uint32_t)P.P, v = uv - u * (uint32_t)P.P;
      if (P.mu) {   // x * sigma + mu, channel = c % Co   (models/dpot.py:401)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < P.nout) y3[c] = fmaf(y3[c], P.sigma[(int64_t)b * P.Co + c % P.Co], P.mu[(int64_t)b * P.Co + c % P.Co]);
      }
      const int64_t pl = ((int64_t)b * X + p * P.P + u) * Y + q * P.P + v;       // linear pixel index
      if (RING) {
        const int To = P.nout / P.Co;
        for (int to = 0; to < To; ++to) {
          int slot = P.rg.slot0 + to; if (slot >= P.rg.T) slot -= P.rg.T;
          float* dr = P.rg.ring + (pl * P.rg.T + slot) * P.Co;
          float* dp = P.rg.pred ? P.rg.pred + (pl * P.rg.Ttot + (int64_t)P.rg.step * To + to) * P.Co : nullptr;
          if (P.Co == 4) {
            const float4 val = to == 0 ? make_float4(y3[0], y3[1], y3[2], y3[3]) : make_float4(y3[4], y3[5], y3[6], y3[7]);
            *reinterpret_cast<float4*>(dr) = val;
            if (dp) *reinterpret_cast<float4*>(dp) = val;
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c >= to * P.Co && c < (to + 1) * P.Co) { dr[c - to * P.Co] = y3[c]; if (dp) dp[c - to * P.Co] = y3[c]; }
          }
        }
      } else {
        float* dst = P.out + pl * P.nout;
        if (P.nout == 4) *reinterpret_cast<float4*>(dst) = make_float4(y3[0], y3[1], y3[2], y3[3]);
        else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < P.nout) dst[c] = y3[c];
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

}  // namespace

bool out_tail_tc_takes(int old, int nout, int Co) {
  return g_tail_tc != 0 && tc_device_ok() && old == 32 && nout >= 1 && nout <= 8 && Co >= 1 && nout % Co == 0;
}

// y1g: the ConvTranspose GEMM's result in DPOT_FMT_HL16G32 = dense [npix, 64] halves.  *served = false: geometry not taken.
int out_tail_tc_launch(const void* y1g, const float* w2, const float* b2, const float* w4, const float* b4, int B, int h, int w,
                       int P, int old, int nout, int act, const float* mu, const float* sigma, int Co, float* out, float* ring,
                       float* pred, int T, int slot0, int Ttot, int step, cudaStream_t st, bool* served) {
  *served = false;
  if (!out_tail_tc_takes(old, nout, Co)) return 0;
  if ((reinterpret_cast<uintptr_t>(y1g) % 16) != 0) return 0;
  if (ring ? (Co == 4 && ((reinterpret_cast<uintptr_t>(ring) % 16) != 0 || (pred && (reinterpret_cast<uintptr_t>(pred) % 16) != 0)))
           : (nout == 4 && (reinterpret_cast<uintptr_t>(out) % 16) != 0))
    return 0;
  OtParams Pm;
  Pm.w2 = w2; Pm.b2 = b2; Pm.w4 = w4; Pm.b4 = b4; Pm.mu = mu; Pm.sigma = sigma; Pm.out = out;
  Pm.rg.ring = ring; Pm.rg.pred = pred; Pm.rg.T = T; Pm.rg.slot0 = slot0; Pm.rg.Ttot = Ttot; Pm.rg.step = step;
  Pm.B = B; Pm.h = h; Pm.w = w; Pm.P = P; Pm.nout = nout; Pm.act = act; Pm.Co = Co;
  Pm.npix = (int64_t)B * h * w * P * P;
  if (Pm.npix >= (1ll << 31) - OT_ROWS) return 0;
  Pm.ntiles = (int)ceil_div(Pm.npix, OT_ROWS);
  alignas(64) CUtensorMap mY;
  DPOT_CALL(tc_encode_map_f16(&mY, y1g, 64, (uint64_t)Pm.npix, 1, 128, (uint64_t)Pm.npix * 128, 64, OT_ROWS, 1));
  int dev = 0, sms = 148;
  DPOT_CUDA(cudaGetDevice(&dev));
  DPOT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = Pm.ntiles < sms ? Pm.ntiles : sms;
#define DPOT_OTT(AM, RG)                                                                                               \
  do {                                                                                                                 \
    static DevOnce attr;                                                                                          \
    if (attr.need()) {                                                                                                       \
      DPOT_CUDA(cudaFuncSetAttribute(out_tail_tc_kernel<AM, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OT_SMEM)); \
      attr.done();                                                                                                     \
    }                                                                                                                  \
    DPOT_CUDA(launch_pdl(out_tail_tc_kernel<AM, RG>, dim3(grid), dim3(OT_NTHREADS), OT_SMEM, st, mY, Pm));             \
  } while (0)
  if (act == DPOT_ACT_GELU) { if (ring) DPOT_OTT(1, true); else DPOT_OTT(1, false); }
  else { if (ring) DPOT_OTT(2, true); else DPOT_OTT(2, false); }
#undef DPOT_OTT
  *served = true;
  DPOT_LAUNCH_CHECK("out_tail_tc_kernel");
  return 0;
}

}  // namespace dpot

using namespace dpot;

extern "C" int dpot_out_tail_tc(const void* Y1g, const float* w2, const float* b2, const float* w4, const float* b4, int32_t B,
                                int32_t h, int32_t w, int32_t P, int32_t old, int32_t nout, int32_t act, const float* mu,
                                const float* sigma, int32_t Co, float* out, float* ring, float* pred, int32_t T, int32_t slot0,
                                int32_t Ttot, int32_t step, void* stream) {
  DPOT_REQUIRE(Y1g && w2 && b2 && w4 && b4 && (out || ring), DPOT_E_BADARG, "dpot_out_tail_tc: null pointer");
  DPOT_REQUIRE((mu == nullptr) == (sigma == nullptr), DPOT_E_BADARG, "dpot_out_tail_tc: mu/sigma must come together");
  DPOT_REQUIRE(!ring || (slot0 >= 0 && slot0 < T && Co > 0 && nout / Co <= T), DPOT_E_BADARG, "dpot_out_tail_tc: bad ring geometry");
  bool served = false;
  DPOT_CALL(out_tail_tc_launch(Y1g, w2, b2, w4, b4, B, h, w, P, old, nout, act, mu, sigma, Co, out, ring, pred, T, slot0, Ttot, step,
                               as_stream(stream), &served));
  DPOT_REQUIRE(served, DPOT_E_UNSUPPORTED, "dpot_out_tail_tc: geometry not served (out_layer_dim must be 32, nout <= 8, 16-byte aligned buffers)");
  return 0;
}
extern "C" int dpot_out_tail_tc_supported(int32_t old, int32_t nout, int32_t Co) { return out_tail_tc_takes(old, nout, Co) ? 1 : 0; }
