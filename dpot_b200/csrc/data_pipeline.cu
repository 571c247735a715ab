// GPU-side batch assembly of the reference's data path (utils/griddataset.py): pad_data (:88-100: bilinear resize of every
// (t, c) plane to res x res with torch's align_corners=False rule, channels padded with 1.0), the training window
// (:152-157: T_in input frames from a per-sample start, the next T_ar frames as targets) and the masks (:156 ones for
// training, get_target_mask :102-116 for evaluation).  One pass: every output element is written once, raw samples are read
// through the 4 neighbours of the bilinear stencil (L2-resident: a raw sample is 2.9 MB).
#include "common.cuh"

namespace dpot {
namespace {

struct AsmArgs {
  const float* raw; const int* t_start;
  float* xx; float* yy; float* msk;
  int B, H0, W0, T0, C0, res, T_in, T_ar, C, mask_mode, pred_channels;
  float sh, sw;     // H0 / res, W0 / res as torch computes them (float)
};

// torch area_pixel_compute_source_index, align_corners = False, bilinear: max(scale * (dst + 0.5) - 0.5, 0)
__device__ __forceinline__ void src_index(int dst, float scale, int n_in, int& i0, int& i1, float& l0, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) assemble_batch_kernel(const AsmArgs a) {
  const int TT = a.T_in + a.T_ar;
  const int64_t per_pix = (int64_t)TT * a.C;
  const int64_t total = (int64_t)a.B * a.res * a.res * per_pix;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % a.C);
    int64_t r = i / a.C;
    const int t = (int)(r % TT); r /= TT;
    const int y = (int)(r % a.res); r /= a.res;
    const int x = (int)(r % a.res); const int b = (int)(r / a.res);
    float v = 1.0f;                                         // channel padding (:98-99)
    if (c < a.C0) {
      const int ts = a.t_start[b] + t;
      int x0, x1, y0, y1; float lx0, lx1, ly0, ly1;
      src_index(x, a.sh, a.H0, x0, x1, lx0, lx1);
      src_index(y, a.sw, a.W0, y0, y1, ly0, ly1);
      const float* p = a.raw + (int64_t)b * a.H0 * a.W0 * a.T0 * a.C0 + (int64_t)ts * a.C0 + c;
      const int64_t sx = (int64_t)a.W0 * a.T0 * a.C0, sy = (int64_t)a.T0 * a.C0;
      // torch: lx0 * (ly0 * v00 + ly1 * v01) + lx1 * (ly0 * v10 + ly1 * v11)
      v = lx0 * (ly0 * __ldg(p + x0 * sx + y0 * sy) + ly1 * __ldg(p + x0 * sx + y1 * sy)) +
          lx1 * (ly0 * __ldg(p + x1 * sx + y0 * sy) + ly1 * __ldg(p + x1 * sx + y1 * sy));
    }
    const int64_t pix = ((int64_t)b * a.res + x) * a.res + y;
    if (t < a.T_in) a.xx[(pix * a.T_in + t) * a.C + c] = v;
    else a.yy[(pix * a.T_ar + (t - a.T_in)) * a.C + c] = v;
    if (a.msk && t == 0) {
      float m = 1.0f;
      if (a.mask_mode == 1) {                               // get_target_mask: the grid points of the original resolution
        int kx = a.res / a.H0, ky = a.res / a.W0;
        kx = kx == 0 ? 1 : kx; ky = ky == 0 ? 1 : ky;
        m = (x % kx == 0 && y % ky == 0 && c < a.pred_channels) ? 1.0f : 0.0f;
      }
      a.msk[pix * a.C + c] = m;
    }
  }
}

}  // namespace
}  // namespace dpot

using namespace dpot;

extern "C" int dpot_assemble_batch(const float* raw, const int32_t* t_start, int32_t B, int32_t H0, int32_t W0, int32_t T0,
                                   int32_t C0, int32_t res, int32_t T_in, int32_t T_ar, int32_t C, int32_t mask_mode,
                                   int32_t pred_channels, float* xx, float* yy, float* msk, void* stream) {
  DPOT_REQUIRE(raw && t_start && xx && yy, DPOT_E_BADARG, "dpot_assemble_batch: null pointer");
  DPOT_REQUIRE(B > 0 && H0 > 0 && W0 > 0 && T0 > 0 && C0 > 0 && res > 0 && T_in > 0 && T_ar > 0 && C >= C0, DPOT_E_BADARG,
               "dpot_assemble_batch: bad shape (the padded channel count must be >= the sample's)");
  DPOT_REQUIRE(T_in + T_ar <= T0, DPOT_E_BADARG, "dpot_assemble_batch: the window of %d frames does not fit %d", T_in + T_ar, T0);
  DPOT_REQUIRE(mask_mode == 0 || mask_mode == 1, DPOT_E_BADARG, "dpot_assemble_batch: mask_mode must be 0 (ones) or 1 (target mask)");
  AsmArgs a;
  a.raw = raw; a.t_start = t_start; a.xx = xx; a.yy = yy; a.msk = msk;
  a.B = B; a.H0 = H0; a.W0 = W0; a.T0 = T0; a.C0 = C0; a.res = res; a.T_in = T_in; a.T_ar = T_ar; a.C = C;
  a.mask_mode = mask_mode; a.pred_channels = pred_channels > 0 ? pred_channels : C0;
  a.sh = (float)H0 / (float)res; a.sw = (float)W0 / (float)res;
  const int64_t total = (int64_t)B * res * res * (T_in + T_ar) * C;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count_cur() * 16);
  assemble_batch_kernel<<<grid, 256, 0, as_stream(stream)>>>(a);
  DPOT_LAUNCH_CHECK("assemble_batch_kernel");
  return 0;
}
