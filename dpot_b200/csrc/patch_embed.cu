// PatchEmbed conv0 + activation (models/dpot.py:199-200, 375) straight from the field layout, with the
// coordinate channels of get_grid_3d (:350-360) folded into a row-bias table (dpot_pack_patch).
//
//   z1[(b,p,q), t*mid + m] = act( rowbias0[(p,q,t), m] + sum_{u,v,c} W0p[m,(u,v,c)] * x[b, pP+u, qP+v, slot(t), c] )
//
// The field is a ring in time: logical frame t lives in slot (t + t0) % T, so the autoregressive
// window advance (train_temporal.py:219) is "overwrite the oldest slot" instead of a copy of the window.
//
// Work item = (b, p, run of QPB patches along q), all T frames: R = QPB*T output rows, N = mid (35 for
// DPOT-S), K = P*P*C walked in P slabs of P*C (one image row u per slab).  For fixed (b, p, u) the slab
// is ONE contiguous run of QPB*P*T*C floats of the field -> fully coalesced float4 loads, transposed into
// shared memory as As[k][row]; fp32 FMA register tiles of 8 rows x 4 columns (N is too small for a
// tensor-core tile to pay; the kernel is bound by the fp32 FMA pipe, HBM traffic is the 84 MB field once).
#include "common.cuh"
#include "gemm_common.cuh"

namespace dpot {
namespace {

struct PatchArgs {
  const float* x; const float* W0p; const float* rowbias0; const float* a_scale; const float* a_shift;
  float* z1;
  int B, X, Y, T, C, P, mid, act, t0, Kp;
  int h, w, QPB, QSPLIT, R, Rp, NRG, NCG, Np, PC, K0;
};

template <bool OUT16, bool VEC4>
__global__ void __launch_bounds__(256) patch_embed_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                       // [PC][Rp]
  float* Ws = smem + (size_t)a.PC * a.Rp; // [PC][Np]

  const int tid = threadIdx.x, nthr = blockDim.x;
  int item = blockIdx.x;
  const int qs = item % a.QSPLIT; item /= a.QSPLIT;
  const int p = item % a.h; const int b = item / a.h;
  const int q0 = qs * a.QPB;
  const int nq = min(a.QPB, a.w - q0);
  const int R = nq * a.T;                 // live rows of this item
  const int TC = a.T * a.C;
  const int L = nq * a.P * TC;            // floats per slab run

  const int tx = tid % a.NCG, ty = tid / a.NCG;
  const bool worker = ty < a.NRG;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  constexpr int NPRE = 8;
  int soff[NPRE];                          // (k << 16) | row of the float4 this thread stages in iteration it
  if (VEC4) {
#pragma unroll
    for (int it = 0; it < NPRE; ++it) {
      const int e = (tid + it * nthr) * 4;
      const int c = e % a.C; int r = e / a.C;
      const int slot = r % a.T; r /= a.T;
      const int v = r % a.P; const int ql = r / a.P;
      int t = slot - a.t0; if (t < 0) t += a.T;
      soff[it] = ((v * a.C + c) << 16) | (ql * a.T + t);
    }
  }

  for (int u = 0; u < a.P; ++u) {
    // ---- stage slab u: field run -> As[(v,c)][(q,t)], weights -> Ws[(v,c)][m]
    const float* run = a.x + (((int64_t)b * a.X + (int64_t)p * a.P + u) * a.Y + (int64_t)q0 * a.P) * TC;
    const float* scl = a.a_scale ? a.a_scale + (int64_t)b * a.K0 + u * a.PC : nullptr;
    const float* shf = a.a_scale ? a.a_shift + (int64_t)b * a.K0 + u * a.PC : nullptr;
    if (VEC4) {
#pragma unroll
      for (int it = 0; it < NPRE; ++it) {          // smem offsets precomputed once (same for every slab)
        const int e = (tid + it * nthr) * 4;
        if (e < L) {
          const float4 v4 = __ldg(reinterpret_cast<const float4*>(run + e));
          const int k = soff[it] >> 16, o = (soff[it] >> 16) * a.Rp + (soff[it] & 0xFFFF);
          float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float val = vv[j];
            if (scl) val = fmaf(val, scl[k + j], shf[k + j]);
            As[o + j * a.Rp] = val;
          }
        }
      }
      for (int e = (tid + NPRE * nthr) * 4; e < L; e += nthr * 4) {
        const float4 v4 = __ldg(reinterpret_cast<const float4*>(run + e));
        const int c = e % a.C; int r = e / a.C;
        const int slot = r % a.T; r /= a.T;
        const int v = r % a.P; const int ql = r / a.P;
        int t = slot - a.t0; if (t < 0) t += a.T;
        const int row = ql * a.T + t;
        const int k = v * a.C + c;
        float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float val = vv[j];
          if (scl) val = fmaf(val, scl[k + j], shf[k + j]);
          As[(k + j) * a.Rp + row] = val;
        }
      }
    } else {
      for (int e = tid; e < L; e += nthr) {
        const int c = e % a.C; int r = e / a.C;
        const int slot = r % a.T; r /= a.T;
        const int v = r % a.P; const int ql = r / a.P;
        int t = slot - a.t0; if (t < 0) t += a.T;
        const int k = v * a.C + c;
        float val = __ldg(run + e);
        if (scl) val = fmaf(val, scl[k], shf[k]);
        As[k * a.Rp + ql * a.T + t] = val;
      }
    }
    for (int e = tid; e < a.mid * a.PC; e += nthr) {
      const int k = e % a.PC, m = e / a.PC;
      Ws[k * a.Np + m] = __ldg(a.W0p + (int64_t)m * a.K0 + u * a.PC + k);
    }
    if (u == 0) {   // zero the padding once (rows >= R, columns >= mid are read by the register tiles)
      for (int e = tid; e < a.PC * (a.Rp - R); e += nthr) As[(e / (a.Rp - R)) * a.Rp + R + e % (a.Rp - R)] = 0.f;
      for (int e = tid; e < a.PC * (a.Np - a.mid); e += nthr)
        Ws[(e / (a.Np - a.mid)) * a.Np + a.mid + e % (a.Np - a.mid)] = 0.f;
    }
    __syncthreads();
    if (worker) {
      const float* ap = As + ty * 8;
      const float* wp = Ws + tx * 4;
#pragma unroll 4
      for (int kk = 0; kk < a.PC; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(ap + kk * a.Rp);
        const float4 a1 = *reinterpret_cast<const float4*>(ap + kk * a.Rp + 4);
        const float4 w0 = *reinterpret_cast<const float4*>(wp + kk * a.Np);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float wv[4] = {w0.x, w0.y, w0.z, w0.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }

  if (!worker) return;
  // ---- epilogue: + row bias (conv bias and coordinate channels), activation, store
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = ty * 8 + i;
    if (r >= R) continue;
    const int ql = r / a.T, t = r % a.T;
    const int q = q0 + ql;
    const int64_t tok = ((int64_t)b * a.h + p) * a.w + q;
    const float* rb = a.rowbias0 + (((int64_t)p * a.w + q) * a.T + t) * a.mid;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = tx * 4 + j;
      if (m >= a.mid) continue;
      const float v = act_apply(acc[i][j] + rb[m], a.act);
      if (OUT16) {
        __half* dst = reinterpret_cast<__half*>(a.z1) + tok * (2 * (int64_t)a.Kp) + t * a.mid + m;
        __half hi, lo;
        hl_split(v, hi, lo);
        dst[0] = hi;
        dst[a.Kp] = lo;
      } else {
        a.z1[tok * a.Kp + t * a.mid + m] = v;
      }
    }
  }
}

}  // namespace
}  // namespace dpot

using namespace dpot;

extern "C" int dpot_patch_embed(const float* x, int32_t t0, const float* W0p, const float* rowbias0,
                                const float* a_scale, const float* a_shift, int32_t B, int32_t X, int32_t Y, int32_t T,
                                int32_t C, int32_t P, int32_t mid, int32_t act, void* z1, int32_t Kp, int32_t out_fmt,
                                void* stream) {
  DPOT_REQUIRE(x && W0p && rowbias0 && z1, DPOT_E_BADARG, "dpot_patch_embed: null pointer");
  DPOT_REQUIRE(B > 0 && P > 0 && X % P == 0 && Y % P == 0 && T > 0 && C > 0 && mid > 0 && Kp >= T * mid, DPOT_E_BADARG,
               "dpot_patch_embed: bad geometry");
  DPOT_REQUIRE(t0 >= 0 && t0 < T, DPOT_E_BADARG, "dpot_patch_embed: ring offset t0=%d outside [0,%d)", t0, T);
  DPOT_REQUIRE((a_scale == nullptr) == (a_shift == nullptr), DPOT_E_BADARG, "dpot_patch_embed: a_scale/a_shift come together");
  DPOT_REQUIRE(out_fmt == DPOT_FMT_F32 || out_fmt == DPOT_FMT_HL16, DPOT_E_BADARG, "dpot_patch_embed: bad out_fmt");
  PatchArgs a;
  a.x = x; a.W0p = W0p; a.rowbias0 = rowbias0; a.a_scale = a_scale; a.a_shift = a_shift; a.z1 = reinterpret_cast<float*>(z1);
  a.B = B; a.X = X; a.Y = Y; a.T = T; a.C = C; a.P = P; a.mid = mid; a.act = act; a.t0 = t0; a.Kp = Kp;
  a.h = X / P; a.w = Y / P;
  a.PC = P * C; a.K0 = P * P * C;
  a.NCG = (int)ceil_div(mid, 4); a.Np = a.NCG * 4;
  // rows per item: as many patches along q as keep the CTA at <= 256 threads
  int qpb = a.w;
  while (qpb > 1 && ceil_div((int64_t)qpb * T, 8) * a.NCG > 256) qpb = (qpb + 1) / 2;
  if (qpb == a.w && a.w >= 16) qpb = a.w / 2;      // finer items balance the 148 SMs better
  DPOT_REQUIRE(ceil_div((int64_t)qpb * T, 8) * a.NCG <= 256, DPOT_E_UNSUPPORTED,
               "dpot_patch_embed: T=%d x mid=%d does not fit one CTA", T, mid);
  a.QPB = qpb; a.QSPLIT = (int)ceil_div(a.w, qpb);
  a.R = qpb * T; a.Rp = (int)round_up(a.R, 8); a.NRG = a.Rp / 8;
  const int nthr = (int)round_up((int64_t)a.NRG * a.NCG, 32);
  const size_t smem = sizeof(float) * ((size_t)a.PC * a.Rp + (size_t)a.PC * a.Np);
  DPOT_REQUIRE(smem <= 200 * 1024, DPOT_E_UNSUPPORTED, "dpot_patch_embed: slab of %zu bytes exceeds shared memory", smem);
  const bool vec4 = (C % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0);
  const bool o16 = out_fmt == DPOT_FMT_HL16;
  const unsigned grid = (unsigned)((int64_t)B * a.h * a.QSPLIT);
  cudaStream_t st = as_stream(stream);
#define DPOT_PE_LAUNCH(O16, V4)                                                                                         \
  do {                                                                                                                  \
    if (smem > 48 * 1024)                                                                                               \
      DPOT_CUDA(cudaFuncSetAttribute(patch_embed_kernel<O16, V4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    patch_embed_kernel<O16, V4><<<grid, nthr, smem, st>>>(a);                                                           \
  } while (0)
  if (o16) { if (vec4) DPOT_PE_LAUNCH(true, true); else DPOT_PE_LAUNCH(true, false); }
  else { if (vec4) DPOT_PE_LAUNCH(false, true); else DPOT_PE_LAUNCH(false, false); }
#undef DPOT_PE_LAUNCH
  DPOT_LAUNCH_CHECK("patch_embed_kernel");
  return 0;
}
