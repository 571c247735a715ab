// PatchEmbed conv0 + activation on tcgen05 (models/dpot.py:199-200, 375; see patch_embed.cu for the math and the
// ring-in-time field layout).  The warp-MMA kernel is bound by the legacy mma.sync issue rate; here the
// im2col rows of an item (QPB patches x T frames <= 128 rows) are the M dimension of a UMMA and the mid (<= 48)
// output channels its N dimension.  Per image row u of the patch (a 32-deep k-slab when P*C = 32):
//   producer warp   : cp.async.bulk of the contiguous field run of the item -> raw fp32 ring (as in patch_embed.cu)
//   converter warps : raw slab -> one 128-byte record [hi 32 | lo 32] per im2col row, written in the 128B-swizzled
//                     K-major operand layout (k' = 64 per slab), optional input-normalisation affine applied on the way
//   MMA thread      : D1 += A'[:, :32] Whi(u)^T (2 k-steps), D2 += A' [Wlo(u) | Whi(u)]^T (4 k-steps); the 2 x 8 weight
//                     operands (96 KB) are resident in shared memory; 4 TMEM accumulator buffers
//   epilogue warps  : thread = im2col row: D1 + D2/2048 + row bias (conv bias + coordinate channels), activation, split
//                     fp16 (or fp32) store of its mid values into z1[(b,p,q), t*mid + m]
#include "common.cuh"
#include "gemm_common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

namespace dpot {
int g_patch_tc = 1;     // 0 = never use the tcgen05 PatchEmbed (dpot_patch_embed_set_engine)
namespace {

constexpr int PT_RAW = 4, PT_AST = 3, PT_NBUF = 4, PT_NMAX = 48, PT_ROWS = 128;
constexpr int PT_CONV_WARP0 = 4, PT_CONV_WARPS = 8, PT_EPI_WARP0 = 12, PT_EPI_WARPS = 8, PT_EPI_CLASSES = PT_EPI_WARPS / 4;
constexpr int PT_NTHREADS = 32 * (PT_EPI_WARP0 + PT_EPI_WARPS);
constexpr uint32_t PT_A_BYTES = PT_ROWS * 128;              // 16 KB operand tile per slab
constexpr uint32_t PT_W_OP = PT_NMAX * 128;                 // 6 KB per weight operand
constexpr int PT_SLABS_MAX = 8;

struct PtArgs {
  const float* x; const float* W0p; const float* rowbias0; const float* a_scale; const float* a_shift;
  float* z1;
  int B, X, Y, T, C, P, mid, act, t0, Kp, h, w, QPB, QSPLIT, K0, PC, nitems, out16;
  uint32_t slab_bytes;    // full-size raw slab
};

template <int ACT_MODE>
__global__ void __launch_bounds__(PT_NTHREADS, 1) patch_embed_tc_kernel(const PtArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const sm = smem_raw + (smem0 - smem_u32(smem_raw));
  // layout: A' ring | weight operands [slab][a|b] | raw ring | barriers
  const uint32_t OFF_W = PT_AST * PT_A_BYTES;
  const uint32_t OFF_RAW = OFF_W + (uint32_t)a.P * 2u * PT_W_OP;
  const uint32_t OFF_BAR = OFF_RAW + PT_RAW * ((a.slab_bytes + 127u) & ~127u);
  const uint32_t raw_stride = (a.slab_bytes + 127u) & ~127u;
  const uint32_t bar0 = smem0 + OFF_BAR;
  auto RFULL = [&](int s) -> uint32_t { return bar0 + 8u * s; };
  auto REMPTY = [&](int s) -> uint32_t { return bar0 + 8u * (PT_RAW + s); };
  auto AFULL = [&](int s) -> uint32_t { return bar0 + 8u * (2 * PT_RAW + s); };
  auto AEMPTY = [&](int s) -> uint32_t { return bar0 + 8u * (2 * PT_RAW + PT_AST + s); };
  auto TFULL = [&](int b) -> uint32_t { return bar0 + 8u * (2 * PT_RAW + 2 * PT_AST + b); };
  auto TEMPTY = [&](int b) -> uint32_t { return bar0 + 8u * (2 * PT_RAW + 2 * PT_AST + PT_NBUF + b); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * PT_RAW + 2 * PT_AST + 2 * PT_NBUF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int TC = a.T * a.C;
  pdl_launch_dependents();
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < PT_RAW; ++s) { mbar_init(RFULL(s), 1); mbar_init(REMPTY(s), PT_CONV_WARPS); }
    for (int s = 0; s < PT_AST; ++s) { mbar_init(AFULL(s), PT_CONV_WARPS); mbar_init(AEMPTY(s), 1); }
    for (int b = 0; b < PT_NBUF; ++b) { mbar_init(TFULL(b), 1); mbar_init(TEMPTY(b), 4); }      // one warp per lane quarter reads a buffer
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  // ---- weight operands (parameters only: before pdl_wait).  Slab u, operand a = [Whi(u) | 0], b = [Wlo(u) | Whi(u)]:
  // row n = output channel (zero beyond mid), 128 bytes, 16-byte chunk c stored at c ^ (n & 7)
  for (int e = tid; e < PT_NMAX * a.K0; e += PT_NTHREADS) {
    const int n = e / a.K0, kk = e - n * a.K0, u = kk / a.PC, k = kk - u * a.PC;      // PC == 32
    __half hi = __float2half_rn(0.f), lo = hi;
    if (n < a.mid) hl_split(__ldg(a.W0p + (int64_t)n * a.K0 + kk), hi, lo);
    auto put = [&](int op, int kq, __half v) {
      const uint32_t off = (uint32_t)n * 128u + ((((uint32_t)kq >> 3) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)kq & 7u) * 2u;
      *reinterpret_cast<__half*>(sm + OFF_W + (uint32_t)(u * 2 + op) * PT_W_OP + off) = v;
    };
    put(0, k, hi); put(0, 32 + k, __float2half_rn(0.f));
    put(1, k, lo); put(1, 32 + k, hi);
  }
  for (uint32_t e = tid; e < PT_AST * PT_A_BYTES / 16; e += PT_NTHREADS)       // rows no item ever owns stay zero
    reinterpret_cast<uint4*>(sm)[e] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  const int64_t run_stride = (int64_t)a.Y * TC;
  const int rows_full = a.QPB * a.T;

  if (warp == 0) {
    // ================================ producer: raw slabs ================================
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
        int it = item;
        const int qs = it % a.QSPLIT; it /= a.QSPLIT;
        const int p = it % a.h, b = it / a.h;
        const int q0 = qs * a.QPB;
        const uint32_t bytes = (uint32_t)(min(a.QPB, a.w - q0) * a.P * TC * 4);
        const float* run0 = a.x + (((int64_t)b * a.X + (int64_t)p * a.P) * a.Y + (int64_t)q0 * a.P) * TC;
        for (int u = 0; u < a.P; ++u) {
          mbar_wait(REMPTY(s), ph ^ 1);
          mbar_expect_tx(RFULL(s), bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem0 + OFF_RAW + (uint32_t)s * raw_stride), "l"(run0 + (int64_t)u * run_stride), "r"(bytes), "r"(RFULL(s))
                       : "memory");
          if (++s == PT_RAW) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(PT_NMAX >> 3) << 17) | ((128u >> 4) << 24);   // M = 128, N = 48
      int as = 0; uint32_t aph = 0; uint32_t tc = 0;
      for (int item = blockIdx.x; item < a.nitems; item += gridDim.x, ++tc) {
        const uint32_t buf = tc % PT_NBUF, bph = (tc / PT_NBUF) & 1u;
        mbar_wait(TEMPTY(buf), bph ^ 1);
        tc_fence_after();
        const uint32_t d1 = tmem_base + buf * 128u, d2 = d1 + 64u;
        for (int u = 0; u < a.P; ++u) {
          mbar_wait(AFULL(as), aph);
          tc_fence_after();
          const uint32_t ab = smem0 + (uint32_t)as * PT_A_BYTES;
          const uint32_t wa = smem0 + OFF_W + (uint32_t)(u * 2) * PT_W_OP, wb = wa + PT_W_OP;
#pragma unroll
          for (int k4 = 0; k4 < 2; ++k4)
            umma_f16(d1, make_smem_desc(ab + k4 * 32), make_smem_desc(wa + k4 * 32), idesc, (u > 0 || k4 > 0) ? 1u : 0u);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_f16(d2, make_smem_desc(ab + k4 * 32), make_smem_desc(wb + k4 * 32), idesc, (u > 0 || k4 > 0) ? 1u : 0u);
          umma_commit(AEMPTY(as));
          if (++as == PT_AST) { as = 0; aph ^= 1; }
        }
        umma_commit(TFULL(buf));
      }
    }
  } else if (warp >= PT_CONV_WARP0 && warp < PT_EPI_WARP0) {
    // ================================ converters ================================
    const int ct = (warp - PT_CONV_WARP0) * 32 + lane, nct = PT_CONV_WARPS * 32;
    int rs = 0, as = 0; uint32_t rph = 0, aph = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      int it = item;
      const int qs = it % a.QSPLIT; it /= a.QSPLIT;
      const int b = it / a.h;
      const int R = min(a.QPB, a.w - qs * a.QPB) * a.T;          // live rows of this item
      for (int u = 0; u < a.P; ++u) {
        mbar_wait(RFULL(rs), rph);
        mbar_wait(AEMPTY(as), aph ^ 1);
        const float* raw = reinterpret_cast<const float*>(sm + OFF_RAW + (uint32_t)rs * raw_stride);
        uint8_t* A = sm + (uint32_t)as * PT_A_BYTES;
        const float* scl = a.a_scale ? a.a_scale + (int64_t)b * a.K0 + u * a.PC : nullptr;
        const float* shf = a.a_scale ? a.a_shift + (int64_t)b * a.K0 + u * a.PC : nullptr;
        // task = (row r, pixel pair j): two float4 (pixels v = 2j, 2j+1; 4 channels) -> 16 B of hi + 16 B of lo
        for (int task = ct; task < rows_full * 4; task += nct) {
          const int r = task >> 2, j = task & 3;
          const int ql = r / a.T, t = r - ql * a.T;
          int slot = t + a.t0; if (slot >= a.T) slot -= a.T;
          float v[8];
          if (r < R) {
            const float4 p0 = *reinterpret_cast<const float4*>(raw + ((ql * a.P + 2 * j) * a.T + slot) * 4);
            const float4 p1 = *reinterpret_cast<const float4*>(raw + ((ql * a.P + 2 * j + 1) * a.T + slot) * 4);
            v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
            if (scl) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], scl[8 * j + i], shf[8 * j + i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
          }
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __half h0 = __float2half_rn(v[2 * i]), h1 = __float2half_rn(v[2 * i + 1]);
            const __half l0 = __float2half_rn((v[2 * i] - __half2float(h0)) * HL_SCALE);
            const __half l1 = __float2half_rn((v[2 * i + 1] - __half2float(h1)) * HL_SCALE);
            hi[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            lo[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
          }
          uint8_t* row = A + (uint32_t)r * 128u;
          *reinterpret_cast<uint4*>(row + (((uint32_t)j ^ ((uint32_t)r & 7u)) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(row + (((uint32_t)(4 + j) ^ ((uint32_t)r & 7u)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();          // this thread's operand writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) { mbar_arrive(AFULL(as)); mbar_arrive(REMPTY(rs)); }
        if (++rs == PT_RAW) { rs = 0; rph ^= 1; }
        if (++as == PT_AST) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp >= PT_EPI_WARP0) {
    // ================================ epilogue ================================
    const int quarter = warp & 3;
    const int cls = (warp - PT_EPI_WARP0) >> 2;             // this warp takes every PT_EPI_CLASSES-th item of the CTA
    const int r = quarter * 32 + lane;                      // im2col row of the item = TMEM lane
    const int ql = r / a.T, t = r - ql * a.T;
    uint32_t tc = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x, ++tc) {
      if ((int)(tc % PT_EPI_CLASSES) != cls) continue;
      int it = item;
      const int qs = it % a.QSPLIT; it /= a.QSPLIT;
      const int p = it % a.h, b = it / a.h;
      const int q0 = qs * a.QPB;
      const int R = min(a.QPB, a.w - q0) * a.T;
      const uint32_t buf = tc % PT_NBUF, bph = (tc / PT_NBUF) & 1u;
      const bool live = r < R;
      const int q = q0 + ql;
      const int64_t tok = ((int64_t)b * a.h + p) * a.w + q;
      const float* rb = a.rowbias0 + (((int64_t)p * a.w + q) * a.T + t) * a.mid;
      __half* dst16 = reinterpret_cast<__half*>(a.z1) + tok * (2 * (int64_t)a.Kp) + t * a.mid;
      float* dst32 = a.z1 + tok * a.Kp + t * a.mid;
      float bva[PT_NMAX];                                  // the row's bias values: in flight while the MMAs still run
#pragma unroll
      for (int i = 0; i < PT_NMAX; ++i) bva[i] = (live && i < a.mid) ? __ldg(rb + i) : 0.f;
      mbar_wait(TFULL(buf), bph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + buf * 128u + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < PT_NMAX; c0 += 16) {
        uint32_t r1[16], r2[16];
        tmem_ld16(t_row + (uint32_t)c0, r1);
        tmem_ld16(t_row + 64u + (uint32_t)c0, r2);
        float bv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) bv[i] = bva[c0 + i];
        tmem_ld_wait();
        if (c0 + 16 >= PT_NMAX) {                          // last chunk read: the accumulator buffer is free again
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(TEMPTY(buf));
        }
        if (live) {            // (no `continue`: every lane must reach the next chunk's aligned tcgen05.ld together)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (c0 + i < a.mid) {
              float v = fmaf(__uint_as_float(r2[i]), HL_INV, __uint_as_float(r1[i])) + bv[i];
              v = ACT_MODE == 1 ? gelu_fast(v) : act_apply(v, a.act);
              if (a.out16) {
                __half hi, lo;
                hl_split(v, hi, lo);
                dst16[c0 + i] = hi;
                dst16[c0 + i + a.Kp] = lo;
              } else {
                dst32[c0 + i] = v;
              }
            }
          }
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

// *served = false: geometry not taken (the caller falls back to the warp-MMA / CUDA-core kernels of patch_embed.cu)
int patch_embed_tc_launch(const float* x, int t0, const float* W0p, const float* rowbias0, const float* a_scale, const float* a_shift,
                          int B, int X, int Y, int T, int C, int P, int mid, int act, void* z1, int Kp, int out_fmt, cudaStream_t st,
                          bool* served) {
  *served = false;
  if (g_patch_tc == 0 || !tc_device_ok() || C != 4 || P * C != 32 || mid > PT_NMAX || P > PT_SLABS_MAX) return 0;
  if ((reinterpret_cast<uintptr_t>(x) % 16) != 0 || (reinterpret_cast<uintptr_t>(W0p) % 4) != 0) return 0;
  const int w = Y / P, h = X / P;
  int qpb = 0;
  for (int q = w; q >= 1; --q)                     // largest run of patches with <= 128 rows, preferring divisors of w
    if ((int64_t)q * T <= PT_ROWS) { if (qpb == 0) qpb = q; if (w % q == 0) { qpb = q; break; } }
  if (qpb == 0) return 0;
  PtArgs a;
  a.x = x; a.W0p = W0p; a.rowbias0 = rowbias0; a.a_scale = a_scale; a.a_shift = a_shift; a.z1 = reinterpret_cast<float*>(z1);
  a.B = B; a.X = X; a.Y = Y; a.T = T; a.C = C; a.P = P; a.mid = mid; a.act = act; a.t0 = t0; a.Kp = Kp; a.h = h; a.w = w;
  a.QPB = qpb; a.QSPLIT = (int)ceil_div(w, qpb); a.K0 = P * P * C; a.PC = P * C;
  a.out16 = out_fmt == DPOT_FMT_HL16 ? 1 : 0;
  a.slab_bytes = (uint32_t)(qpb * P * T * C * 4);
  const int64_t nitems = (int64_t)B * h * a.QSPLIT;
  if (nitems >= (1ll << 30) || a.slab_bytes % 16 != 0) return 0;
  a.nitems = (int)nitems;
  const size_t smem = (size_t)PT_AST * PT_A_BYTES + (size_t)P * 2 * PT_W_OP + (size_t)PT_RAW * ((a.slab_bytes + 127u) & ~127u) + 512 + 1024;
  if (smem > 220 * 1024) return 0;
  int dev = 0, sms = 148;
  DPOT_CUDA(cudaGetDevice(&dev));
  DPOT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = a.nitems < sms ? a.nitems : sms;
#define DPOT_PT(AM)                                                                                                   \
  do {                                                                                                                \
    DPOT_CUDA(cudaFuncSetAttribute(patch_embed_tc_kernel<AM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    DPOT_CUDA(launch_pdl(patch_embed_tc_kernel<AM>, dim3(grid), dim3(PT_NTHREADS), smem, st, a));                     \
  } while (0)
  if (act == DPOT_ACT_GELU) DPOT_PT(1); else DPOT_PT(2);
#undef DPOT_PT
  *served = true;
  DPOT_LAUNCH_CHECK("patch_embed_tc_kernel");
  return 0;
}

}  // namespace dpot
