// dpot_gemm: argument validation and engine dispatch (see include/dpot_b200.h).
#include "common.cuh"
#include "gemm_common.cuh"

using namespace dpot;

namespace dpot {
int g_chain_max = 2048;   // dpot_tc16_set_chain: longest accumulation chain of one launch of the f16-split engine (0 = unlimited)
}

static int gemm_one(const dpot_gemm_args* a, void* stream);

// A tcgen05 fp32 accumulator truncates (round toward zero) on every accumulation step, so the error of a contraction
// grows LINEARLY with its depth: measured 1.2e-9 * K relative L2 on this engine (tools/numerics_longk.py: K = 8192 ->
// 9.7e-6, cuBLAS fp32 1.6e-6).  The channel MLP's fc2 of DPOT-M / L / H is 4096 / 6144 / 8192 deep and 12 / 24 / 27
// blocks stack it, which put the full-depth forwards of L and H at 1.4e-5 (tests/test_full_depth_gpu.py).  Chained
// form: the contraction runs as ceil(K / k_chain) launches over K-ranges; every launch but the last stores its fp32
// partial sum (bias and row bias included) in `scratch`, the next one reads it back as a per-row bias -- an fp32
// round-to-nearest add in the epilogue -- and the last launch applies the activation / scale / residual / statistics
// / output format of the original call.  The error becomes that of one k_chain-deep chain (2048: 2.5e-6).
static bool chain_form_ok(const dpot_gemm_args* a) {
  return a->a_fmt == DPOT_FMT_HL16 && a->w_fmt == DPOT_FMT_HL16 && a->batch == 1 && !a->a_trans && !a->w_trans && a->k_split <= 1 &&
         a->a_mode == DPOT_A_PLAIN && !a->C_pre && !a->dact_src && !a->a_scale;
}

extern "C" int dpot_gemm_chained(const dpot_gemm_args* a, int32_t k_chain, float* scratch, int64_t ld_scratch, void* stream) {
  DPOT_REQUIRE(a != nullptr, DPOT_E_BADARG, "dpot_gemm_chained: null args");
  if (k_chain <= 0 || a->K <= k_chain) return gemm_one(a, stream);
  DPOT_REQUIRE(chain_form_ok(a), DPOT_E_BADARG,
               "dpot_gemm_chained: needs the plain forward form of the f16-split engine (batch 1, no transposes / k_split / C_pre / dact)");
  DPOT_REQUIRE(scratch && ld_scratch >= a->N && reinterpret_cast<uintptr_t>(scratch) % 16 == 0, DPOT_E_BADARG,
               "dpot_gemm_chained: scratch [M, ld_scratch >= N] fp32, 16-byte aligned");
  const int nch = (a->K + k_chain - 1) / k_chain;
  const int64_t chunk = ((a->K + nch - 1) / nch + 63) / 64 * 64;       // equal chains, whole 64-half k-blocks
  for (int64_t k0 = 0; k0 < a->K; k0 += chunk) {
    dpot_gemm_args c = *a;
    const bool first = k0 == 0, last = k0 + chunk >= a->K;
    // A / W point to halves on this engine: a K-range is a column offset in both planes (lo plane = + *_lo_off)
    c.A = reinterpret_cast<const float*>(reinterpret_cast<const uint16_t*>(a->A) + k0);
    c.W = reinterpret_cast<const float*>(reinterpret_cast<const uint16_t*>(a->W) + k0);
    c.K = (int32_t)(last ? a->K - k0 : chunk);
    if (!first) { c.bias = nullptr; c.rowbias = scratch; c.rowbias_period = a->M; c.ldrb = ld_scratch; }
    if (!last) {
      c.C = scratch; c.ldc = ld_scratch; c.c_fmt = DPOT_FMT_F32; c.c_lo_off = 0; c.c_group = 0; c.c_group_stride = 0; c.c_mode = DPOT_A_PLAIN;
      c.act = DPOT_ACT_NONE; c.residual = nullptr; c.c_scale = c.c_shift = nullptr; c.out_stats = nullptr; c.out_colsum = nullptr;
    }
    DPOT_CALL(gemm_one(&c, stream));
  }
  return 0;
}

extern "C" int dpot_tc16_set_chain(int32_t k_chain) {
  const int prev = dpot::g_chain_max;
  if (k_chain >= 0) dpot::g_chain_max = k_chain;
  return prev;
}

// Entry point: long contractions whose result is plain fp32 chain in place (the partial sums live in C itself).
extern "C" int dpot_gemm(const dpot_gemm_args* a, void* stream) {
  DPOT_REQUIRE(a != nullptr, DPOT_E_BADARG, "dpot_gemm: null args");
  const int kc = dpot::g_chain_max;
  if (kc > 0 && a->K > kc && a->A && a->W && a->C && chain_form_ok(a) && a->c_fmt == DPOT_FMT_F32 && a->c_group == 0 &&
      a->c_mode == DPOT_A_PLAIN && a->ldc >= a->N && reinterpret_cast<uintptr_t>(a->C) % 16 == 0)
    return dpot_gemm_chained(a, kc, a->C, a->ldc, stream);
  return gemm_one(a, stream);
}

static int gemm_one(const dpot_gemm_args* a, void* stream) {
  DPOT_REQUIRE(a->A && a->W && a->C, DPOT_E_BADARG, "dpot_gemm: null A/W/C");
  DPOT_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0, DPOT_E_BADARG, "dpot_gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  DPOT_REQUIRE(a->batch >= 1, DPOT_E_BADARG, "dpot_gemm: batch must be >= 1");
  DPOT_REQUIRE(a->act >= DPOT_ACT_NONE && a->act <= DPOT_ACT_SILU, DPOT_E_BADARG, "dpot_gemm: bad act %d", a->act);
  DPOT_REQUIRE(!a->rowbias || a->rowbias_period > 0, DPOT_E_BADARG, "dpot_gemm: rowbias needs a period");
  DPOT_REQUIRE((a->a_scale == nullptr) == (a->a_shift == nullptr), DPOT_E_BADARG, "dpot_gemm: a_scale/a_shift must come together");
  DPOT_REQUIRE(!a->a_scale || a->a_rows_per_sample > 0, DPOT_E_BADARG, "dpot_gemm: a_rows_per_sample");
  DPOT_REQUIRE((a->c_scale == nullptr) == (a->c_shift == nullptr), DPOT_E_BADARG, "dpot_gemm: c_scale/c_shift must come together");
  DPOT_REQUIRE(!a->c_scale || a->c_rows_per_sample > 0, DPOT_E_BADARG, "dpot_gemm: c_rows_per_sample");
  DPOT_REQUIRE(a->a_mode == DPOT_A_PLAIN || a->a_mode == DPOT_A_PATCH, DPOT_E_BADARG, "dpot_gemm: bad a_mode");
  if (a->M == 0) return 0;

  GemmDev p;
  p.A = a->A; p.lda = a->lda; p.W = a->W; p.ldw = a->ldw; p.C = a->C; p.ldc = a->ldc;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias = a->bias; p.rowbias = a->rowbias; p.rb_period = a->rowbias_period; p.ldrb = a->ldrb;
  p.residual = a->residual; p.ldr = a->ldr; p.act = a->act;
  p.a_scale = a->a_scale; p.a_shift = a->a_shift; p.a_rps = a->a_rows_per_sample;
  p.c_scale = a->c_scale; p.c_shift = a->c_shift; p.c_rps = a->c_rows_per_sample;
  p.c_group = a->c_group; p.c_group_stride = a->c_group_stride;
  p.sA = a->strideA; p.sW = a->strideW; p.sC = a->strideC; p.sBias = a->strideBias;
  p.a_mode = a->a_mode; p.pX = a->pX; p.pY = a->pY; p.pT = a->pT; p.pC = a->pC; p.pP = a->pP;
  p.ph = p.pw = 0;
  if (a->a_mode == DPOT_A_PATCH) { p.ph = a->pX / (a->pP > 0 ? a->pP : 1); p.pw = a->pY / (a->pP > 0 ? a->pP : 1); }
  p.C_pre = a->C_pre; p.dact_src = a->dact_src; p.dact = a->dact; p.c_mode = a->c_mode;
  p.a_fmt = a->a_fmt; p.w_fmt = a->w_fmt; p.c_fmt = a->c_fmt; p.a_lo = a->a_lo_off; p.w_lo = a->w_lo_off; p.c_lo = a->c_lo_off;
  p.a_tr = a->a_trans; p.w_tr = a->w_trans; p.ksplit = a->k_split > 1 ? a->k_split : 1; p.kchunk = a->k_chunk; p.sC2 = a->strideC_split;
  p.ldpre = a->ld_pre; p.sPre = a->stride_pre; p.lddact = a->ld_dact; p.sDact = a->stride_dact;
  p.out_colsum = a->out_colsum;
  DPOT_REQUIRE(!a->out_colsum || (a->a_fmt == DPOT_FMT_HL16 && a->k_split <= 1), DPOT_E_BADARG, "dpot_gemm: out_colsum needs the TC16 engine without k_split");
  const bool bw_form = a->a_trans || a->w_trans || a->k_split > 1;
  DPOT_REQUIRE(!bw_form || a->a_fmt == DPOT_FMT_HL16, DPOT_E_BADARG, "dpot_gemm: a_trans / w_trans / k_split need split-fp16 operands");
  if (a->k_split > 1)
    DPOT_REQUIRE(a->k_chunk > 0 && a->k_chunk % 64 == 0 && (int64_t)a->k_split * a->k_chunk >= a->K && !a->bias && !a->rowbias &&
                 !a->residual && a->act == DPOT_ACT_NONE && !a->c_scale && !a->out_stats && !a->C_pre && !a->dact_src,
                 DPOT_E_BADARG, "dpot_gemm: k_split needs k_chunk %% 64 == 0 covering K and a plain epilogue");
  DPOT_REQUIRE((a->a_fmt == DPOT_FMT_F32 || a->a_fmt == DPOT_FMT_HL16) && a->a_fmt == a->w_fmt, DPOT_E_BADARG,
               "dpot_gemm: a_fmt and w_fmt must both be F32 or both HL16");
  DPOT_REQUIRE(a->c_fmt == DPOT_FMT_F32 || a->c_fmt == DPOT_FMT_HL16 || a->c_fmt == DPOT_FMT_HL16G32, DPOT_E_BADARG, "dpot_gemm: bad c_fmt");
  if (a->c_fmt == DPOT_FMT_HL16G32)
    DPOT_REQUIRE(a->a_fmt == DPOT_FMT_HL16 && a->N % 32 == 0 && a->ldc >= 2 * (int64_t)a->N && a->batch == 1 &&
                 !a->dact_src && a->c_mode == DPOT_A_PLAIN && !a->out_stats, DPOT_E_BADARG,
                 "dpot_gemm: the grouped split-fp16 output needs split-fp16 operands, N %% 32 == 0 and ldc >= 2N halves");
  if (a->c_fmt == DPOT_FMT_HL16)
    DPOT_REQUIRE(((!a->C_pre && !a->dact_src) || a->a_fmt == DPOT_FMT_HL16) && a->c_mode == DPOT_A_PLAIN && !a->out_stats && a->c_lo_off > 0,
                 DPOT_E_BADARG, "dpot_gemm: split-fp16 output excludes patch scatter/out_stats (and C_pre/dact_src off the TC16 engine)");
  if (a->a_fmt == DPOT_FMT_HL16 && (a->C_pre || a->dact_src)) {
    DPOT_REQUIRE(!a->C_pre || a->ld_pre >= a->N, DPOT_E_BADARG, "dpot_gemm: C_pre on the TC16 engine needs ld_pre");
    DPOT_REQUIRE(!a->dact_src || (a->ld_dact >= a->N && !a->residual && !a->rowbias && !a->c_scale), DPOT_E_BADARG,
                 "dpot_gemm: dact_src on the TC16 engine needs ld_dact and excludes residual / rowbias / c_scale");
  }
  if (a->c_mode == DPOT_A_PATCH) {
    DPOT_REQUIRE(a->a_mode == DPOT_A_PLAIN && a->pP > 0 && a->pX % a->pP == 0 && a->pY % a->pP == 0 &&
                 a->N == a->pP * a->pP * a->pC && a->batch == 1 && !a->C_pre && !a->dact_src && !a->residual,
                 DPOT_E_BADARG, "dpot_gemm: bad patch-scatter output geometry");
    p.ph = a->pX / a->pP; p.pw = a->pY / a->pP;
  }
  DPOT_REQUIRE(!a->dact_src || (a->c_group == 0), DPOT_E_BADARG, "dpot_gemm: dact_src needs a plain C layout");
  p.out_stats = nullptr; p.st_groups = a->stats_groups; p.st_rps = a->stats_rows_per_sample;
  if (a->out_stats) {
    DPOT_REQUIRE(a->batch == 1 && a->c_group == 0 && a->ldc == a->N, DPOT_E_BADARG, "dpot_gemm: out_stats needs a plain contiguous C");
    DPOT_REQUIRE(a->stats_groups > 0 && a->N % a->stats_groups == 0 && a->stats_rows_per_sample > 0 &&
                 a->M % a->stats_rows_per_sample == 0, DPOT_E_BADARG, "dpot_gemm: bad out_stats geometry");
  }
  if (a->a_mode == DPOT_A_PATCH) {
    DPOT_REQUIRE(a->pP > 0 && a->pX % a->pP == 0 && a->pY % a->pP == 0, DPOT_E_BADARG, "dpot_gemm: patch geometry");
    p.ph = a->pX / a->pP; p.pw = a->pY / a->pP;
    DPOT_REQUIRE(a->K == a->pP * a->pP * a->pC, DPOT_E_BADARG, "dpot_gemm: patch K must be P*P*C");
    DPOT_REQUIRE(a->M % (p.ph * p.pw * a->pT) == 0, DPOT_E_BADARG, "dpot_gemm: patch M must be B*h*w*T");
    DPOT_REQUIRE(a->batch == 1, DPOT_E_BADARG, "dpot_gemm: patch mode is not batched");
  }
  cudaStream_t st = as_stream(stream);
  int engine = a->engine;
  const int B = a->out_stats ? a->M / a->stats_rows_per_sample : 0;
  if (a->a_fmt == DPOT_FMT_HL16) {   // split-fp16 operands: only the f16 tensor-core engine reads them
    DPOT_REQUIRE(engine == DPOT_GEMM_AUTO || engine == DPOT_GEMM_TC16, DPOT_E_BADARG,
                 "dpot_gemm: split-fp16 operands need the TC16 engine");
    DPOT_REQUIRE(gemm_tc16_supports(p, a->batch), DPOT_E_UNSUPPORTED,
                 "dpot_gemm: f16-split tcgen05 engine does not take this problem (M=%d N=%d K=%d)", a->M, a->N, a->K);
    if (a->out_stats && gemm_tc16_fuses_stats(p)) {
      DPOT_CUDA(cudaMemsetAsync(a->out_stats, 0, sizeof(double) * 2 * (size_t)B * a->stats_groups, st));
      p.out_stats = a->out_stats;
      return gemm_tc16_launch(p, a->batch, st);
    }
    DPOT_CALL(gemm_tc16_launch(p, a->batch, st));
    if (a->out_stats)
      return dpot_gn_stats(a->C, B, a->stats_rows_per_sample, a->N, a->stats_groups, a->out_stats, stream);
    return 0;
  }
  DPOT_REQUIRE(engine != DPOT_GEMM_TC16, DPOT_E_BADARG, "dpot_gemm: the TC16 engine needs split-fp16 operands");
  if (a->c_fmt == DPOT_FMT_HL16) {
    DPOT_REQUIRE(engine != DPOT_GEMM_TC && a->batch == 1, DPOT_E_UNSUPPORTED,
                 "dpot_gemm: split-fp16 output from fp32 operands is served by the SIMT engine, batch 1");
    engine = DPOT_GEMM_SIMT;
  }
  if (engine == DPOT_GEMM_AUTO) engine = gemm_tc_supports(p, a->batch) ? DPOT_GEMM_TC : DPOT_GEMM_SIMT;
  if (engine == DPOT_GEMM_TC) {
    DPOT_REQUIRE(gemm_tc_supports(p, a->batch), DPOT_E_UNSUPPORTED,
                 "dpot_gemm: tcgen05 engine does not take this problem (M=%d N=%d K=%d)", a->M, a->N, a->K);
    if (a->out_stats && gemm_tc_fuses_stats(p)) {
      DPOT_CUDA(cudaMemsetAsync(a->out_stats, 0, sizeof(double) * 2 * (size_t)B * a->stats_groups, st));
      p.out_stats = a->out_stats;
      return gemm_tc_launch(p, a->batch, st);
    }
    DPOT_CALL(gemm_tc_launch(p, a->batch, st));
  } else {
    DPOT_CALL(gemm_simt_launch(p, a->batch, st));
  }
  if (a->out_stats)   // engines without a fused reduction: one streaming pass over the (L2-hot) result
    return dpot_gn_stats(a->C, B, a->stats_rows_per_sample, a->N, a->stats_groups, a->out_stats, stream);
  return 0;
}
