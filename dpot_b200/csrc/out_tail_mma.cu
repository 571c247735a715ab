// Output tail on warp-level tensor cores: per pixel  y2 = act(W2 y1 + b2) (old -> old),  y3 = W4 y2 + b4 (old -> nout),
// pixel shuffle (+ de-normalisation)   (models/dpot.py:317-321,397-401).
//
// The previous CUDA-core kernel spent one shared-memory broadcast load per FMA (LSU-bound, 107 us for DPOT-S B=32);
// here a warp owns 16 pixels per step: its y1 rows arrive as two fully used 32 B sectors per pixel and k-step, are
// split to fp16 hi/lo in registers, both weight matrices live in REGISTERS as mma B fragments for the whole kernel
// (no shared-memory traffic in the loop), and the accumulator fragment of layer 2 is re-used directly as the A
// fragment of the final projection.  fp32-faithful: D1 += Ahi Whi, D2 += Ahi Wlo + Alo Whi, y = D1 + D2/2048.
#include "common.cuh"
#include "gemm_common.cuh"
#include "mma_sync.cuh"

namespace dpot {
namespace {

// RING: instead of out[B,X,Y,nout] the T_out new frames go straight into the autoregressive window, a ring in time
// (frame j -> slot (slot0 + j) % T of ring[B,X,Y,T,Co]), and into the prediction tensor pred[B,X,Y,Ttot,Co] at
// frame step*T_out + j: the window advance of train_temporal.py:219 / evaluate.py:207 without a copy kernel.
struct TailRing { float* ring; float* pred; int T, slot0, Ttot, step; };

template <int OLD, int ACT_MODE, bool RING>
__global__ void __launch_bounds__(128, 4) out_tail_mma_kernel(const float* __restrict__ Y1, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, const float* __restrict__ w4,
                                                           const float* __restrict__ b4, int B, int h, int w, int P,
                                                           int nout, int act, const float* __restrict__ mu,
                                                           const float* __restrict__ sigma, int Co,
                                                           float* __restrict__ out, const TailRing rg) {
  constexpr int KS = OLD / 16, NT = OLD / 8, KW = OLD + 8;
  __shared__ __align__(16) __half s_w[2][OLD + 8][KW];      // [hi|lo][n: W2 rows, then 8 rows of W4][k]
  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  pdl_launch_dependents();
  // parameters only up to pdl_wait(): the weight fragments are built while the predecessor (the ConvTranspose GEMM) drains
  for (int e = tid; e < (OLD + 8) * OLD; e += blockDim.x) {
    const int n = e / OLD, k = e % OLD;
    float v = 0.f;
    if (n < OLD) v = w2[n * OLD + k];
    else if (n - OLD < nout) v = w4[(n - OLD) * OLD + k];
    __half hi, lo;
    hl_split(v, hi, lo);
    s_w[0][n][k] = hi;
    s_w[1][n][k] = lo;
  }
  __syncthreads();
  // weight fragments -> registers: layer 2 [KS][NT] and the projection [KS] (one n-tile of 8, columns >= nout are zero)
  uint32_t w2h[KS][NT][2], w2l[KS][NT][2], w4h[KS][2], w4l[KS][2];
  {
    const int lj = lane >> 3, li = lane & 7;
    const uint32_t base_h = smem_u32_generic(&s_w[0][0][0]), base_l = smem_u32_generic(&s_w[1][0][0]);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        const uint32_t off = (uint32_t)((((np * 2 + (lj >> 1)) * 8 + li) * KW + ks * 16 + (lj & 1) * 8) * 2);
        uint32_t r[4];
        ldmatrix_x4(base_h + off, r);
        w2h[ks][2 * np][0] = r[0]; w2h[ks][2 * np][1] = r[1]; w2h[ks][2 * np + 1][0] = r[2]; w2h[ks][2 * np + 1][1] = r[3];
        ldmatrix_x4(base_l + off, r);
        w2l[ks][2 * np][0] = r[0]; w2l[ks][2 * np][1] = r[1]; w2l[ks][2 * np + 1][0] = r[2]; w2l[ks][2 * np + 1][1] = r[3];
      }
      const uint32_t off4 = (uint32_t)(((OLD + li) * KW + ks * 16 + (lj & 1) * 8) * 2);
      uint32_t r[4];
      ldmatrix_x2(base_h + off4, r);
      w4h[ks][0] = r[0]; w4h[ks][1] = r[1];
      ldmatrix_x2(base_l + off4, r);
      w4l[ks][0] = r[0]; w4l[ks][1] = r[1];
    }
  }
  float bias2[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { bias2[nt][0] = b2[nt * 8 + tg * 2]; bias2[nt][1] = b2[nt * 8 + tg * 2 + 1]; }
  const int c0 = tg * 2;
  const float bias4_0 = c0 < nout ? b4[c0] : 0.f, bias4_1 = c0 + 1 < nout ? b4[c0 + 1] : 0.f;

  pdl_wait();
  const int64_t npix = (int64_t)B * h * w * P * P;
  const int64_t ntile = (npix + 15) / 16;
  const int X = h * P, Y = w * P;
  const int tiles_per_patch = P * P / 16;
  const int64_t wstep = (int64_t)gridDim.x * (blockDim.x >> 5);
  // y1 rows of a 16-pixel tile (row g and g+8; k = ks*16 + tg*2 (+8)); the NEXT tile's rows are in flight while
  // this one is computed (the kernel is a stream over 67 MB with ~1 us of arithmetic per tile and warp)
  float2 v[KS][4];
  auto load_tile = [&](int64_t tile) {
    const int64_t pa = tile * 16 + g, pb = pa + 8;
    const bool oka = pa < npix, okb = pb < npix;
    const float2* r0 = reinterpret_cast<const float2*>(Y1 + pa * OLD);
    const float2* r1 = reinterpret_cast<const float2*>(Y1 + pb * OLD);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      v[ks][0] = oka ? __ldg(r0 + ks * 8 + tg) : make_float2(0.f, 0.f);
      v[ks][1] = okb ? __ldg(r1 + ks * 8 + tg) : make_float2(0.f, 0.f);
      v[ks][2] = oka ? __ldg(r0 + ks * 8 + 4 + tg) : make_float2(0.f, 0.f);
      v[ks][3] = okb ? __ldg(r1 + ks * 8 + 4 + tg) : make_float2(0.f, 0.f);
    }
  };
  const int64_t tile_first = (int64_t)blockIdx.x * (blockDim.x >> 5) + (tid >> 5);
  if (tile_first < ntile) load_tile(tile_first);
  for (int64_t tile = tile_first; tile < ntile; tile += wstep) {
    const int64_t pix0 = tile * 16 + g, pix1 = pix0 + 8;
    uint32_t ah[KS][4], al[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) hl_split2(v[ks][i].x, v[ks][i].y, ah[ks][i], al[ks][i]);
    if (tile + wstep < ntile) load_tile(tile + wstep);
    // ---- layer 2
    float d1[NT][4], d2[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) { d1[nt][j] = 0.f; d2[nt][j] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        mma_f16(d1[nt], ah[ks], w2h[ks][nt][0], w2h[ks][nt][1]);
        mma_f16(d2[nt], ah[ks], w2l[ks][nt][0], w2l[ks][nt][1]);
        mma_f16(d2[nt], al[ks], w2h[ks][nt][0], w2h[ks][nt][1]);
      }
    // ---- bias + activation; the accumulator fragment of n-tiles (2s, 2s+1) IS the A fragment of k-step s
    uint32_t yh[KS][4], yl[KS][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        t[j] = fmaf(d2[nt][j], HL_INV, d1[nt][j]) + bias2[nt][j & 1];
        t[j] = ACT_MODE == 1 ? gelu_fast(t[j]) : act_apply(t[j], act);
      }
      hl_split2(t[0], t[1], yh[nt >> 1][(nt & 1) * 2], yl[nt >> 1][(nt & 1) * 2]);          // row g
      hl_split2(t[2], t[3], yh[nt >> 1][(nt & 1) * 2 + 1], yl[nt >> 1][(nt & 1) * 2 + 1]);  // row g + 8
    }
    // ---- projection to nout channels
    float e1[4] = {0.f, 0.f, 0.f, 0.f}, e2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      mma_f16(e1, yh[ks], w4h[ks][0], w4h[ks][1]);
      mma_f16(e2, yh[ks], w4l[ks][0], w4l[ks][1]);
      mma_f16(e2, yl[ks], w4h[ks][0], w4h[ks][1]);
    }
    // ---- pixel shuffle store: pix = ((b*h + p)*w + q)*P*P + u*P + v.  P*P % 16 == 0 (host-checked): the 16 pixels
    // of a tile belong to ONE patch, so (b, p, q) is a warp-uniform 32-bit computation
    {
      const uint32_t patch = (uint32_t)(tile / tiles_per_patch);
      const uint32_t uv0 = (uint32_t)(tile - (int64_t)patch * tiles_per_patch) * 16u;
      const uint32_t q = patch % (uint32_t)w, bp = patch / (uint32_t)w;
      const uint32_t p = bp % (uint32_t)h, b = bp / (uint32_t)h;
      float sg0 = 1.f, sg1 = 1.f, mu0 = 0.f, mu1 = 0.f;
      if (mu) {   // x * sigma + mu, channel = c % Co   (models/dpot.py:401)
        const int ca = c0 % Co, cb = (c0 + 1) % Co;
        sg0 = sigma[(int64_t)b * Co + ca]; mu0 = mu[(int64_t)b * Co + ca];
        sg1 = sigma[(int64_t)b * Co + cb]; mu1 = mu[(int64_t)b * Co + cb];
      }
      const int64_t pix_base = ((int64_t)b * X + p * P) * Y + q * P;       // linear pixel index of the patch corner
      float* base = out + pix_base * nout + c0;
      const int to = RING ? c0 / Co : 0, co = RING ? c0 - to * Co : 0;      // Co even: (c0, c0+1) lie in one frame
      int slot = RING ? rg.slot0 + to : 0;
      if (RING && slot >= rg.T) slot -= rg.T;
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        if (c0 >= nout || (hr ? pix1 : pix0) >= npix) continue;
        const uint32_t uv = uv0 + (uint32_t)g + (hr ? 8u : 0u);
        const uint32_t u = uv / (uint32_t)P, v = uv - u * (uint32_t)P;
        const float y0 = fmaf(fmaf(e2[hr * 2], HL_INV, e1[hr * 2]) + bias4_0, sg0, mu0);
        const float y1v = fmaf(fmaf(e2[hr * 2 + 1], HL_INV, e1[hr * 2 + 1]) + bias4_1, sg1, mu1);
        if (RING) {
          const int64_t pl = pix_base + (int64_t)u * Y + v;
          *reinterpret_cast<float2*>(rg.ring + (pl * rg.T + slot) * Co + co) = make_float2(y0, y1v);
          if (rg.pred)
            *reinterpret_cast<float2*>(rg.pred + (pl * rg.Ttot + (int64_t)rg.step * (nout / Co) + to) * Co + co) = make_float2(y0, y1v);
        } else {
          float* dst = base + ((int64_t)u * Y + v) * nout;
          if ((nout & 1) == 0) *reinterpret_cast<float2*>(dst) = make_float2(y0, y1v);
          else { dst[0] = y0; if (c0 + 1 < nout) dst[1] = y1v; }
        }
      }
    }
  }
}

}  // namespace

// *served = false: this geometry is not taken (the caller falls back to the CUDA-core kernel)
int out_tail_mma_launch(const float* Y1, const float* w2, const float* b2, const float* w4, const float* b4, int B, int h,
                        int w, int P, int old, int nout, int act, const float* mu, const float* sigma, int Co, float* out,
                        cudaStream_t st, bool* served, float* ring, float* pred, int T, int slot0, int Ttot, int step) {
  *served = false;
  if (ring && (nout % 2 != 0 || Co % 2 != 0 || nout % Co != 0 || (reinterpret_cast<uintptr_t>(ring) % 8) != 0 ||
               (pred && (reinterpret_cast<uintptr_t>(pred) % 8) != 0)))
    return 0;
  if (!(old == 16 || old == 32) || nout > 8 || (P * P) % 16 != 0 || (reinterpret_cast<uintptr_t>(Y1) % 8) != 0 ||
      (!ring && (nout & 1) == 0 && (reinterpret_cast<uintptr_t>(out) % 8) != 0))
    return 0;
  TailRing rg; rg.ring = ring; rg.pred = pred; rg.T = T; rg.slot0 = slot0; rg.Ttot = Ttot; rg.step = step;
  const int64_t npix = (int64_t)B * h * w * P * P;
  const int64_t ntile = (npix + 15) / 16;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (ntile + 3) / 4;
  const unsigned grid = (unsigned)(want < (int64_t)sms * 4 ? want : (int64_t)sms * 4);   // one resident wave (4 CTAs / SM)
#define DPOT_OTM2(O, AM, RG) DPOT_CUDA(launch_pdl(out_tail_mma_kernel<O, AM, RG>, dim3(grid), dim3(128), 0, st, Y1, w2, b2, w4, b4, B, h, w, P, nout, act, mu, sigma, Co, out, rg))
#define DPOT_OTM(O, AM) do { if (ring) DPOT_OTM2(O, AM, true); else DPOT_OTM2(O, AM, false); } while (0)
  if (old == 32) { if (act == DPOT_ACT_GELU) DPOT_OTM(32, 1); else DPOT_OTM(32, 2); }
  else { if (act == DPOT_ACT_GELU) DPOT_OTM(16, 1); else DPOT_OTM(16, 2); }
#undef DPOT_OTM
#undef DPOT_OTM2
  *served = true;
  DPOT_LAUNCH_CHECK("out_tail_mma_kernel");
  return 0;
}

}  // namespace dpot
