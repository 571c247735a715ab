// Model geometry, packed-weight arena layout and GEMM-argument helpers shared by the inference forward (forward.cu)
// and the training step (train_step.cu).
#pragma once
#include "common.cuh"
#include <string.h>

namespace dpot {

struct Dims {
  int P, C, Co, T, To, nb, E, old, depth, hid, ncls, h, n, mid, K0, Kp, bs, km1, km2, NP;
};

static inline int make_dims(const dpot_config* c, Dims& d) {
  DPOT_REQUIRE(c != nullptr, DPOT_E_BADARG, "null config");
  DPOT_REQUIRE(c->patch_size > 0 && c->img_size > 0 && c->img_size % c->patch_size == 0, DPOT_E_BADARG,
               "img_size %d must be a multiple of patch_size %d", c->img_size, c->patch_size);
  DPOT_REQUIRE(c->n_blocks > 0 && c->embed_dim % c->n_blocks == 0, DPOT_E_BADARG, "embed_dim %% n_blocks != 0");
  DPOT_REQUIRE(c->embed_dim % 8 == 0, DPOT_E_BADARG, "embed_dim must be divisible by the 8 GroupNorm groups");
  d.P = c->patch_size; d.C = c->in_channels; d.Co = c->out_channels; d.T = c->in_timesteps; d.To = c->out_timesteps;
  d.nb = c->n_blocks; d.E = c->embed_dim; d.old = c->out_layer_dim; d.depth = c->depth; d.hid = c->hidden_dim;
  d.ncls = c->n_cls; d.h = c->img_size / c->patch_size; d.n = d.h * d.h;
  d.mid = d.Co * d.P + 3;                       // models/dpot.py:278
  d.K0 = d.P * d.P * d.C;
  d.Kp = (int)round_up((int64_t)d.T * d.mid, 32);
  d.bs = d.E / d.nb;
  d.km1 = c->modes < d.h ? c->modes : d.h;       // python slicing clamps, models/dpot.py:70-94
  d.km2 = c->modes < d.h / 2 + 1 ? c->modes : d.h / 2 + 1;
  d.NP = d.P * d.P * d.old;
  DPOT_REQUIRE(d.km1 >= 1, DPOT_E_BADARG, "modes must be >= 1");
  DPOT_REQUIRE(d.h == 2 || d.h == 4 || d.h == 8 || d.h == 16 || d.h == 32, DPOT_E_UNSUPPORTED,
               "latent grid img_size/patch_size = %d unsupported (power of two in [2,32])", d.h);
  return 0;
}

static inline int64_t slot(int64_t n) { return round_up(n, 64); }  // 256 B granules

// geometry served by the fused AFNO mixer kernel (afno_fused.cu): latent 16 x 16, block size 128, no mode truncation
static inline bool fused_geometry(const Dims& d) {
  return d.h == 16 && d.bs == 128 && d.km1 == 16 && d.km2 == 9 && (d.E / 8) % 32 == 0;
}

struct Packed {
  int64_t W0p, rowbias0, WeffT, bias_eff, blocks, blk_stride, Wc1, bc1, Wc2, bc2, WtT, bias_t, total;
  // split-fp16 (DPOT_FMT_HL16) copies for the f16-split tensor-core engine; same float counts as the fp32 originals
  int64_t WeffT16, WtT16, Wc1_16, Wc2_16, fc1_16, fc2_16;   // the last four are offsets inside a block slab
  int64_t Wfus;                                             // fused AFNO mixer arena inside a block slab (afno_fused.cu)
  int64_t cls0_16, cls2_16, cls4_16;                        // split-fp16 copies of the cls head weights
};
static inline Packed packed_layout(const Dims& d) {
  Packed L; int64_t o = 0;
  L.W0p = o; o += slot((int64_t)d.mid * d.K0);
  L.rowbias0 = o; o += slot((int64_t)d.n * d.T * d.mid);
  L.WeffT = o; o += slot((int64_t)d.E * d.Kp);
  L.bias_eff = o; o += slot((int64_t)d.n * d.E + (int64_t)d.E * d.E);  // + Wsum scratch
  L.blocks = o;
  int64_t b = 0;
  L.Wc1 = b; b += slot((int64_t)d.nb * 4 * d.bs * d.bs);
  L.bc1 = b; b += slot(2 * d.E);
  L.Wc2 = b; b += slot((int64_t)d.nb * 4 * d.bs * d.bs);
  L.bc2 = b; b += slot(2 * d.E);
  L.Wc1_16 = b; b += slot((int64_t)d.nb * 4 * d.bs * d.bs);
  L.Wc2_16 = b; b += slot((int64_t)d.nb * 4 * d.bs * d.bs);
  L.fc1_16 = b; b += slot((int64_t)d.hid * d.E);
  L.fc2_16 = b; b += slot((int64_t)d.hid * d.E);
  L.Wfus = b; b += fused_geometry(d) ? slot(dpot_afno_fused_packed_floats(d.nb)) : 0;
  L.blk_stride = b; o += b * d.depth;
  L.WtT = o; o += slot((int64_t)d.NP * d.E);
  L.bias_t = o; o += slot(d.NP);
  L.WeffT16 = o; o += slot((int64_t)d.E * d.Kp);
  L.WtT16 = o; o += slot((int64_t)d.NP * d.E);
  L.cls0_16 = o; o += slot((int64_t)d.E * d.E);
  L.cls2_16 = o; o += slot((int64_t)d.E * d.E);
  L.cls4_16 = o; o += slot((int64_t)d.ncls * d.E);
  L.total = o;
  return L;
}

static inline dpot_gemm_args gemm_args(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int M,
                                int N, int K, const float* bias, int act, int engine) {
  dpot_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  g.bias = bias; g.act = act; g.batch = 1; g.engine = engine; g.a_mode = DPOT_A_PLAIN;
  return g;
}

// split-fp16 operands: A16 rows of `lda_f` floats' worth of bytes ([hi lda_f halves | lo lda_f halves]), likewise W16
static inline dpot_gemm_args gemm16_args(const float* A16, int64_t a_row_f, const float* W16, int64_t w_row_f, float* C,
                                  int64_t ldc, int M, int N, int K, const float* bias, int act) {
  dpot_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.A = A16; g.lda = 2 * a_row_f; g.a_lo_off = a_row_f; g.a_fmt = DPOT_FMT_HL16;
  g.W = W16; g.ldw = 2 * w_row_f; g.w_lo_off = w_row_f; g.w_fmt = DPOT_FMT_HL16;
  g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  g.bias = bias; g.act = act; g.batch = 1; g.engine = DPOT_GEMM_TC16; g.a_mode = DPOT_A_PLAIN;
  return g;
}
static inline void out16(dpot_gemm_args& g, int64_t c_row_f) {   // store the result split: rows of c_row_f floats' worth
  g.c_fmt = DPOT_FMT_HL16; g.ldc = 2 * c_row_f; g.c_lo_off = c_row_f;
}

// The f16-split pipeline serves every dense contraction when the device has tcgen05 and the row
// lengths keep the 16-byte alignment TMA needs.
static inline bool use_tc16(const Dims& d, int engine) {
  if (engine != DPOT_GEMM_AUTO && engine != DPOT_GEMM_TC16) return false;
  if (!dpot_tc16_available()) return false;
  return d.E % 8 == 0 && d.Kp % 8 == 0 && (2 * d.bs) % 8 == 0 && d.hid % 8 == 0 && d.NP % 8 == 0 && d.mid > 0;
}


}  // namespace dpot
