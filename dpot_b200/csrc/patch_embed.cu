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

int g_patch_engine = 0;    // 0 auto, 1 fp32 CUDA cores only, 2 warp-MMA only, 3 tcgen05 only (dpot_patch_embed_set_engine; tests)
}
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
// Warp-specialised, persistent: warp 0 is a producer that streams the item's image-row slabs (ONE contiguous run of
// the field each) into a ring of shared-memory stages with cp.async.bulk + mbarriers; warps 1..NW own one 16-row
// tile of the item each: they wait for a slab, pull their mma A fragments straight out of the raw fp32 slab (one
// float2 = one pixel's channel pair per register pair), release the stage, split to hi/lo in registers and run the
// MMAs against the weight table (ldmatrix).  No CTA-wide barrier after the one-time weight staging.
constexpr int PE_STAGES = 6;
template <int NT8, bool OUT16, int KS>
__global__ void __launch_bounds__(288) patch_embed_mma_kernel(const PatchArgs a, const int nitems) {
  extern __shared__ __align__(128) uint8_t smem_mma[];
  const int KW = a.K0 + 8;                 // halves per weight row (padded: conflict-free ldmatrix)
  const int NW = (int)(blockDim.x >> 5) - 1;                   // consumer warps = 16-row tiles per item
  const int TC = a.T * a.C;
  const uint32_t slab_bytes = (uint32_t)(a.QPB * a.P * TC * 4);   // full-size slab (a multiple of 16)
  __half* Wh = reinterpret_cast<__half*>(smem_mma);            // [NT8*8][KW]
  __half* Wl = Wh + (size_t)NT8 * 8 * KW;
  const uint32_t w_bytes = (uint32_t)(2 * NT8 * 8 * KW * 2);
  const uint32_t ring_off = (w_bytes + 127u) & ~127u;
  float* ring = reinterpret_cast<float*>(smem_mma + ring_off);                      // [PE_STAGES][slab]
  const uint32_t bar0 = smem_u32_generic(smem_mma + ring_off + PE_STAGES * slab_bytes);   // full[S], empty[S]
  auto FULL = [&](int st) -> uint32_t { return bar0 + 8u * st; };
  auto EMPTY = [&](int st) -> uint32_t { return bar0 + 8u * (PE_STAGES + st); };

  const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
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
  if (tid == 0) {
    for (int st = 0; st < PE_STAGES; ++st) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(FULL(st)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(EMPTY(st)), "r"(NW));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();      // the field (written by the previous step's output tail) is read from here on

  auto mbar_wait_ = [&](uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t it = 0; !ok; ++it) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
      if (it > 20000000u) __trap();          // a protocol bug must fail loudly, never hang the GPU
    }
  };
  const int64_t run_stride = (int64_t)a.Y * TC;      // one image row

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        int it = item;
        const int qs = it % a.QSPLIT; it /= a.QSPLIT;
        const int p = it % a.h, b = it / a.h;
        const int q0 = qs * a.QPB;
        const uint32_t bytes = (uint32_t)(min(a.QPB, a.w - q0) * a.P * TC * 4);
        const float* run0 = a.x + (((int64_t)b * a.X + (int64_t)p * a.P) * a.Y + (int64_t)q0 * a.P) * TC;
        for (int u = 0; u < a.P; ++u) {
          mbar_wait_(EMPTY(st), ph ^ 1);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(FULL(st)), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32_generic(ring) + (uint32_t)st * slab_bytes), "l"(run0 + (int64_t)u * run_stride), "r"(bytes),
                         "r"(FULL(st)) : "memory");
          if (++st == PE_STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
    return;
  }

  // ================================ consumers ================================
  const int cw = warp - 1;                                     // 16-row tile of the item
  const int lj = lane >> 3, li = lane & 7;
  const uint32_t w_lane = (uint32_t)((((lj >> 1) * 8 + li) * KW + (lj & 1) * 8) * 2);                  // W: (n-tile pair, k)
  const uint32_t Wh_s = smem_u32_generic(Wh), Wl_s = smem_u32_generic(Wl);
  const int g = lane >> 2, tg = lane & 3;
  // float offset of this lane's float2 inside a slab, per (k-step, k-half): pixel column v = ks*4 + tg/2 + 2j, channels (tg&1)*2
  int voff[KS][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int j = 0; j < 2; ++j) voff[ks][j] = (ks * 4 + (tg >> 1) + 2 * j) * TC + (tg & 1) * 2;
  // this lane's two rows r = (q_local, t): float offset of (pixel column 0, channel 0) of the row inside a slab
  int roff[2], rq[2], rt[2];
#pragma unroll
  for (int hr = 0; hr < 2; ++hr) {
    const int r = cw * 16 + g + hr * 8;
    rq[hr] = r / a.T; rt[hr] = r - rq[hr] * a.T;
    int slot = rt[hr] + a.t0; if (slot >= a.T) slot -= a.T;
    roff[hr] = (rq[hr] * a.P * a.T + slot) * a.C;
  }
  const bool is_gelu = a.act == DPOT_ACT_GELU;

  int st = 0; uint32_t ph = 0;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    int it = item;
    const int qs = it % a.QSPLIT; it /= a.QSPLIT;
    const int p = it % a.h, b = it / a.h;
    const int q0 = qs * a.QPB;
    const int R = min(a.QPB, a.w - q0) * a.T;        // live rows of this item
    const bool rok[2] = {cw * 16 + g < R, cw * 16 + g + 8 < R};
    float d1[NT8][4], d2[NT8][4];
#pragma unroll
    for (int i = 0; i < NT8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { d1[i][j] = 0.f; d2[i][j] = 0.f; }

    for (int u = 0; u < a.P; ++u) {
      mbar_wait_(FULL(st), ph);
      const float* slab = ring + (size_t)st * (slab_bytes / 4);
      float2 v[KS][4];                               // [k-step][row g klo, row g+8 klo, row g khi, row g+8 khi]
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          v[ks][i] = rok[i & 1] ? *reinterpret_cast<const float2*>(slab + roff[i & 1] + voff[ks][i >> 1]) : make_float2(0.f, 0.f);
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(EMPTY(st)) : "memory");   // slab consumed
      if (++st == PE_STAGES) { st = 0; ph ^= 1; }
      if (a.a_scale) {                               // input normalisation tables (normalize=True), k = (u, v, c)
        const float* scl = a.a_scale + (int64_t)b * a.K0 + u * a.PC;
        const float* shf = a.a_shift + (int64_t)b * a.K0 + u * a.PC;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k = (ks * 4 + (tg >> 1) + 2 * (i >> 1)) * a.C + (tg & 1) * 2;
            if (rok[i & 1]) {
              v[ks][i].x = fmaf(v[ks][i].x, scl[k], shf[k]);
              v[ks][i].y = fmaf(v[ks][i].y, scl[k + 1], shf[k + 1]);
            }
          }
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) hl_split2(v[ks][i].x, v[ks][i].y, ah[i], al[i]);
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
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      if (!rok[hrow]) continue;
      const int q = q0 + rq[hrow], t = rt[hrow];
      const int64_t tok = ((int64_t)b * a.h + p) * a.w + q;
      const float* rb = a.rowbias0 + (((int64_t)p * a.w + q) * a.T + t) * a.mid + tg * 2;
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
          const float vv = is_gelu ? gelu_fast(pre_act) : act_apply(pre_act, a.act);
          if (OUT16) {
            __half hi, lo;
            hl_split(vv, hi, lo);
            dst16[nt * 8 + j] = hi;
            dst16[nt * 8 + j + a.Kp] = lo;
          } else {
            dst32[nt * 8 + j] = vv;
          }
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
  cudaStream_t st = as_stream(stream);

  DPOT_REQUIRE(g_patch_engine != 3, DPOT_E_UNSUPPORTED, "dpot_patch_embed: engine 3 (the tcgen05 PatchEmbed) was removed: measured slower than the warp-MMA kernel (DESIGN 4.3)");
  // ---- tensor-core path (mma.sync on split fp16): needs float4-able runs and 16-deep k-steps per image row
  const int nt8 = (int)ceil_div(mid, 8);
  if (g_patch_engine != 1 && vec4 && C == 4 && a.PC % 16 == 0 && a.PC <= 64 && nt8 <= 9) {
    const int ks = a.PC / 16;                   // k-steps per image-row slab: 1, 2 or 4
    // items = runs of QPB patches along q: at most 8 tiles of 16 rows, preferring runs without padded rows
    int best = 0;
    for (int q = a.w; q >= 1; --q) {
      const int64_t rows = (int64_t)q * T;
      if (ceil_div(rows, 16) > 8) continue;
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
      const size_t slab = (size_t)best * P * T * C * 4;
      const size_t w_b = ((size_t)2 * nt8 * 8 * (a.K0 + 8) * 2 + 127) & ~(size_t)127;
      const size_t smem_m = w_b + PE_STAGES * slab + 2 * PE_STAGES * 8 + 128;
      if (smem_m <= 200 * 1024 && slab % 16 == 0) {
        const int nitems = (int)((int64_t)B * a.h * m.QSPLIT);
        int dev = 0, sms = 148;
        DPOT_CUDA(cudaGetDevice(&dev));
        DPOT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        int per_sm = (int)((224 * 1024) / (smem_m + 1024));
        per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
        const unsigned grid_m = (unsigned)(nitems < sms * per_sm ? nitems : sms * per_sm);   // persistent: weights staged once per CTA
#define DPOT_PEM_LAUNCH(NT, O16, KSV)                                                                                   \
  do {                                                                                                                  \
    if (smem_m > 48 * 1024)                                                                                             \
      DPOT_CUDA(cudaFuncSetAttribute(patch_embed_mma_kernel<NT, O16, KSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem_m));                                                                     \
    DPOT_CUDA(launch_pdl(patch_embed_mma_kernel<NT, O16, KSV>, dim3(grid_m), dim3((nw + 1) * 32), smem_m, st, m, nitems)); \
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
