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
#include "mma_sync.cuh"

namespace dpot {
namespace {

int g_patch_engine = 0;    // 0 auto, 1 fp32 CUDA cores only, 2 mma only (dpot_patch_embed_set_engine; tests)

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


// ------------------------------------------------------------------------------------------
// Tensor-core version (warp-level mma.sync m16n8k16, fp16 x fp16 -> fp32) with fp32-faithful numerics:
// x = hi + lo/2048 and W = hi + lo/2048 (DPOT_FMT_HL16 arithmetic), D1 += Ahi Whi, D2 += Ahi Wlo + Alo Whi,
// result = D1 + D2/2048 (products of two fp16 are exact in the fp32 accumulator).  N = mid is far too small
// for a tcgen05 tile (UMMA M >= 128 would waste 73 % of the tensor core and need the im2col operand in a
// TMA-describable layout); at N = 35 the kernel is bound by reading the field once from HBM.
//
// CTA = one item (b, p, QPB patches along q) x all T frames: rows r = ql*T + t, one warp per 16-row tile.
// Per image row u the item's slab is ONE contiguous run of the field: float4 loads -> hi/lo split ->
// shared memory [row][k] (double-buffered), ldmatrix fragments, NT8 n-tiles of 8 output channels.
template <int NT8, bool OUT16, int KS>
__global__ void __launch_bounds__(256) patch_embed_mma_kernel(const PatchArgs a, const int nitems) {
  constexpr int NF = KS <= 2 ? 4 : 8;      // float4 per thread per slab (host guarantees L4 <= NF * nthr)
  extern __shared__ __align__(16) uint8_t smem_mma[];
  const int KW = a.K0 + 8;                 // halves per weight row (padded: conflict-free ldmatrix)
  const int KA = a.PC + 8;                 // halves per activation row
  const int rows_pad = (int)(blockDim.x >> 5) * 16;
  __half* Wh = reinterpret_cast<__half*>(smem_mma);            // [NT8*8][KW]
  __half* Wl = Wh + (size_t)NT8 * 8 * KW;
  __half* Ab = Wl + (size_t)NT8 * 8 * KW;                      // [2 buffers][hi, lo][rows_pad][KA]
  const size_t a_plane = (size_t)rows_pad * KA;

  const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
  const int TC = a.T * a.C;
  pdl_launch_dependents();

  // ---- weights -> split fp16 in shared memory, once per (persistent) CTA; rows >= mid are zero.  The packed weights
  // were produced long before the predecessor kernel: this staging runs BEFORE pdl_wait() and overlaps its tail.
  {
    const int nvec = a.mid * a.K0 / 4;                         // K0 % 4 == 0 (C % 4 == 0)
    const float4* W4 = reinterpret_cast<const float4*>(a.W0p);
    for (int e0 = tid; e0 < nvec; e0 += 4 * nthr) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (e0 + j * nthr < nvec) ? __ldg(W4 + e0 + j * nthr) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = (e0 + j * nthr) * 4;
        if (e >= nvec * 4) continue;
        const int n = e / a.K0, k = e % a.K0;
        uint32_t h0, l0, h1, l1;
        hl_split2(v[j].x, v[j].y, h0, l0);
        hl_split2(v[j].z, v[j].w, h1, l1);
        *reinterpret_cast<uint2*>(Wh + (size_t)n * KW + k) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(Wl + (size_t)n * KW + k) = make_uint2(l0, l1);
      }
    }
    for (int e = a.mid * a.K0 + tid; e < NT8 * 8 * a.K0; e += nthr) {
      Wh[(size_t)(e / a.K0) * KW + e % a.K0] = __float2half_rn(0.f);
      Wl[(size_t)(e / a.K0) * KW + e % a.K0] = __float2half_rn(0.f);
    }
  }
  // activation rows beyond a short item (ragged last q-run) must read as zero: clear everything once
  for (int e = tid; e < (int)(4 * a_plane / 2); e += nthr) reinterpret_cast<uint32_t*>(Ab)[e] = 0u;
  __syncthreads();
  pdl_wait();      // the field (written by the previous step's output tail) is read from here on

  // ldmatrix lane addressing: lane -> (matrix j = lane / 8, row i = lane % 8)
  const int lj = lane >> 3, li = lane & 7;
  const uint32_t a_lane = (uint32_t)(((warp * 16 + (lj & 1) * 8 + li) * KA + (lj >> 1) * 8) * 2);       // A: (rows, k) quads
  const uint32_t w_lane = (uint32_t)((((lj >> 1) * 8 + li) * KW + (lj & 1) * 8) * 2);                  // W: (n-tile pair, k)
  const uint32_t Ab_s = smem_u32_generic(Ab), Wh_s = smem_u32_generic(Wh), Wl_s = smem_u32_generic(Wl);
  const int g = lane >> 2, tg = lane & 3;

  // per-thread staging map (identical for every slab and item): float4 index -> destination (row*KA + k)
  int doff[NF];
#pragma unroll
  for (int it = 0; it < NF; ++it) {
    const int e = (tid + it * nthr) * 4;
    const int c = e % a.C; int r = e / a.C;
    const int slot = r % a.T; r /= a.T;
    const int v = r % a.P; const int ql = r / a.P;
    int t = slot - a.t0; if (t < 0) t += a.T;
    doff[it] = (ql * a.T + t) * KA + v * a.C + c;              // ql >= nq of the item is masked at load time
  }

  struct Item { int b, p, q0, L4; const float* run0; };
  auto item_of = [&](int item) {
    Item I;
    const int qs = item % a.QSPLIT; item /= a.QSPLIT;
    I.p = item % a.h; I.b = item / a.h;
    I.q0 = qs * a.QPB;
    const int nq = min(a.QPB, a.w - I.q0);
    I.L4 = nq * a.P * TC / 4;
    I.run0 = a.x + (((int64_t)I.b * a.X + (int64_t)I.p * a.P) * a.Y + (int64_t)I.q0 * a.P) * TC;
    return I;
  };
  const int64_t run_stride = (int64_t)a.Y * TC;      // one image row
  float4 pre[NF];
  auto prefetch = [&](const Item& I, int u) {
    const float4* run = reinterpret_cast<const float4*>(I.run0 + (int64_t)u * run_stride);
#pragma unroll
    for (int it = 0; it < NF; ++it)
      pre[it] = (tid + it * nthr < I.L4) ? __ldg(run + tid + it * nthr) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto convert = [&](const Item& I, int u, int buf) {
    __half* Ah = Ab + (size_t)buf * 2 * a_plane;
    __half* Al = Ah + a_plane;
    const float* scl = a.a_scale ? a.a_scale + (int64_t)I.b * a.K0 + u * a.PC : nullptr;
    const float* shf = a.a_scale ? a.a_shift + (int64_t)I.b * a.K0 + u * a.PC : nullptr;
#pragma unroll
    for (int it = 0; it < NF; ++it) {
      if ((tid + it * nthr) * 4 >= a.QPB * a.P * TC) continue;       // beyond the full-size slab
      float vv[4] = {pre[it].x, pre[it].y, pre[it].z, pre[it].w};
      if (scl && tid + it * nthr < I.L4) {
        const int k = doff[it] % KA;
#pragma unroll
        for (int j = 0; j < 4; ++j) vv[j] = fmaf(vv[j], scl[k + j], shf[k + j]);
      }
      uint32_t h0, l0, h1, l1;
      hl_split2(vv[0], vv[1], h0, l0);
      hl_split2(vv[2], vv[3], h1, l1);
      *reinterpret_cast<uint2*>(Ah + doff[it]) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(Al + doff[it]) = make_uint2(l0, l1);
    }
  };

  int item = blockIdx.x;
  if (item >= nitems) return;
  Item cur = item_of(item);
  prefetch(cur, 0);
  int buf = 0;
  while (true) {
    float d1[NT8][4], d2[NT8][4];
#pragma unroll
    for (int i = 0; i < NT8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { d1[i][j] = 0.f; d2[i][j] = 0.f; }
    const int next_item = item + gridDim.x;
    Item nxt = cur;
    if (next_item < nitems) nxt = item_of(next_item);

    for (int u = 0; u < a.P; ++u, buf ^= 1) {
      convert(cur, u, buf);
      if (u + 1 < a.P) prefetch(cur, u + 1);
      else if (next_item < nitems) prefetch(nxt, 0);           // the next item's first slab flies during this item's tail
      __syncthreads();
      const uint32_t Ah_s = Ab_s + (uint32_t)(buf * 2 * a_plane * 2);
      const uint32_t Al_s = Ah_s + (uint32_t)(a_plane * 2);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[4], al[4];
        ldmatrix_x4(Ah_s + a_lane + ks * 32, ah);
        ldmatrix_x4(Al_s + a_lane + ks * 32, al);
        const uint32_t wk = (uint32_t)((u * a.PC + ks * 16) * 2);
#pragma unroll
        for (int np = 0; np < (NT8 + 1) / 2; ++np) {
          uint32_t wh[4], wl[4];
          const uint32_t wrow = (uint32_t)(np * 16 * KW * 2);
          if (2 * np + 1 < NT8) {
            ldmatrix_x4(Wh_s + w_lane + wrow + wk, wh);
            ldmatrix_x4(Wl_s + w_lane + wrow + wk, wl);
          } else {                                   // odd tail: lanes 16-31 would address rows past the table
            ldmatrix_x2(Wh_s + w_lane + wrow + wk, wh);
            ldmatrix_x2(Wl_s + w_lane + wrow + wk, wl);
          }
          mma_f16(d1[2 * np], ah, wh[0], wh[1]);
          mma_f16(d2[2 * np], ah, wl[0], wl[1]);
          mma_f16(d2[2 * np], al, wh[0], wh[1]);
          if (2 * np + 1 < NT8) {
            mma_f16(d1[2 * np + 1], ah, wh[2], wh[3]);
            mma_f16(d2[2 * np + 1], ah, wl[2], wl[3]);
            mma_f16(d2[2 * np + 1], al, wh[2], wh[3]);
          }
        }
      }
    }

    // ---- epilogue: + row bias (conv bias and coordinate channels), activation, store
    const int R = min(a.QPB, a.w - cur.q0) * a.T;
    const bool is_gelu = a.act == DPOT_ACT_GELU;
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int r = warp * 16 + g + hrow * 8;
      if (r >= R) continue;
      const int ql = r / a.T, t = r - ql * a.T;
      const int q = cur.q0 + ql;
      const int64_t tok = ((int64_t)cur.b * a.h + cur.p) * a.w + q;
      const float* rb = a.rowbias0 + (((int64_t)cur.p * a.w + q) * a.T + t) * a.mid + tg * 2;
      float bv[NT8][2];
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt)        // all bias loads in flight before the first use
#pragma unroll
        for (int j = 0; j < 2; ++j) bv[nt][j] = (nt * 8 + tg * 2 + j < a.mid) ? __ldg(rb + nt * 8 + j) : 0.f;
      __half* dst16 = reinterpret_cast<__half*>(a.z1) + tok * (2 * (int64_t)a.Kp) + t * a.mid + tg * 2;
      float* dst32 = a.z1 + tok * a.Kp + t * a.mid + tg * 2;
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (nt * 8 + tg * 2 + j >= a.mid) continue;
          const float pre_act = fmaf(d2[nt][hrow * 2 + j], HL_INV, d1[nt][hrow * 2 + j]) + bv[nt][j];
          const float v = is_gelu ? gelu_fast(pre_act) : act_apply(pre_act, a.act);
          if (OUT16) {
            __half hi, lo;
            hl_split(v, hi, lo);
            dst16[nt * 8 + j] = hi;
            dst16[nt * 8 + j + a.Kp] = lo;
          } else {
            dst32[nt * 8 + j] = v;
          }
        }
    }
    if (next_item >= nitems) break;
    item = next_item;
    cur = nxt;
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
  cudaStream_t st = as_stream(stream);

  // ---- tensor-core path (mma.sync on split fp16): needs float4-able runs and 16-deep k-steps per image row
  const int nt8 = (int)ceil_div(mid, 8);
  if (g_patch_engine != 1 && vec4 && a.PC % 16 == 0 && a.PC <= 64 && nt8 <= 9) {
    const int ks = a.PC / 16;                   // k-steps per image-row slab: 1, 2 or 4
    const int nf = ks <= 2 ? 4 : 8;
    int q = a.w, best = 0;
    for (; q >= 1; --q) {                      // largest run of patches that fits 8 warps, preferring no padded rows
      const int64_t rows = (int64_t)q * T, nw = ceil_div(rows, 16);
      if (nw > 8 || (int64_t)q * P * T * C / 4 > nf * 32 * nw) continue;
      if (best == 0) best = q;
      if (rows % 16 == 0) { best = q; break; }
      if (q < best / 2) break;
    }
    if (best > 0 && best == a.w && a.w >= 16 && ((int64_t)(a.w / 2) * T) % 16 == 0) best = a.w / 2;
    // instantiated (k-steps, n-tiles) combinations: mid = out_channels*P + 3 with P = 4 / 8 / 16 (models/dpot.py:278)
    const bool inst = (ks == 1 && nt8 <= 3) || (ks == 2 && nt8 >= 2 && nt8 <= 5) || (ks == 4 && (nt8 == 3 || nt8 == 5));
    if (best > 0 && inst) {
      PatchArgs m = a;
      m.QPB = best; m.QSPLIT = (int)ceil_div(a.w, best);
      const int nw = (int)ceil_div((int64_t)best * T, 16);
      const size_t smem_m = 2 * ((size_t)nt8 * 8 * (a.K0 + 8) * 2 + 2 * (size_t)nw * 16 * (a.PC + 8) * 2);
      if (smem_m <= 160 * 1024) {
        const int nitems = (int)((int64_t)B * a.h * m.QSPLIT);
        int dev = 0, sms = 148;
        DPOT_CUDA(cudaGetDevice(&dev));
        DPOT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        int per_sm = (int)((220 * 1024) / (smem_m + 1024));
        per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
        const unsigned grid_m = (unsigned)(nitems < sms * per_sm ? nitems : sms * per_sm);   // persistent: weights staged once per CTA
#define DPOT_PEM_LAUNCH(NT, O16, KSV)                                                                                   \
  do {                                                                                                                  \
    if (smem_m > 48 * 1024)                                                                                             \
      DPOT_CUDA(cudaFuncSetAttribute(patch_embed_mma_kernel<NT, O16, KSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem_m));                                                                     \
    DPOT_CUDA(launch_pdl(patch_embed_mma_kernel<NT, O16, KSV>, dim3(grid_m), dim3(nw * 32), smem_m, st, m, nitems));    \
  } while (0)
#define DPOT_PEM(NT, KSV) do { if (o16) DPOT_PEM_LAUNCH(NT, true, KSV); else DPOT_PEM_LAUNCH(NT, false, KSV); } while (0)
        switch (ks * 16 + nt8) {
          case 16 + 1: DPOT_PEM(1, 1); break;
          case 16 + 2: DPOT_PEM(2, 1); break;
          case 16 + 3: DPOT_PEM(3, 1); break;
          case 32 + 2: DPOT_PEM(2, 2); break;
          case 32 + 3: DPOT_PEM(3, 2); break;
          case 32 + 4: DPOT_PEM(4, 2); break;
          case 32 + 5: DPOT_PEM(5, 2); break;
          case 64 + 3: DPOT_PEM(3, 4); break;
          case 64 + 5: DPOT_PEM(5, 4); break;
          default: break;
        }
#undef DPOT_PEM
#undef DPOT_PEM_LAUNCH
        DPOT_LAUNCH_CHECK("patch_embed_mma_kernel");
        return 0;
      }
    }
  }
  DPOT_REQUIRE(g_patch_engine != 2, DPOT_E_UNSUPPORTED, "dpot_patch_embed: the mma engine does not take this geometry");
  const unsigned grid = (unsigned)((int64_t)B * a.h * a.QSPLIT);
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

extern "C" void dpot_patch_embed_set_engine(int32_t engine) { dpot::g_patch_engine = engine; }
