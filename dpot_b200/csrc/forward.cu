// Whole-model inference forward (DPOTNet.forward under no_grad, models/dpot.py:364-403):
// host-side orchestration of the kernels of this library on one stream.  No allocation, no
// synchronisation: packed weights and the activation workspace are caller-provided arenas.
#include "common.cuh"
#include "model_layout.cuh"
#include <string.h>

namespace dpot {

struct Work {
  int64_t z1, lat0, lat1, f, n2, hid, S, O1, sc1, sh1, sc2, sh2, st1, st2, Y1, Y2, tok, c1, c2, musig, asc, ash, smu, ssg, total;
};
static Work work_layout(const Dims& d, const dpot_config* c, int B) {
  Work L; int64_t o = 0;
  const int64_t Mt = (int64_t)B * d.n, Ms = (int64_t)B * d.km1 * d.km2;
  L.z1 = o; o += slot(Mt * d.Kp);
  L.lat0 = o; o += slot(Mt * d.E);
  L.lat1 = o; o += slot(Mt * d.E);
  L.f = o; o += slot(Mt * d.E);
  L.n2 = o; o += slot(Mt * d.E);      // split-fp16 GroupNorm-2 output / last latent (TC16 pipeline)
  L.hid = o; o += slot(Mt * d.hid);
  L.S = o; o += slot(Ms * 2 * d.E);
  L.O1 = o; o += slot(Ms * 2 * d.E);
  L.sc1 = o; o += slot((int64_t)B * d.E);
  L.sh1 = o; o += slot((int64_t)B * d.E);
  L.sc2 = o; o += slot((int64_t)B * d.E);
  L.sh2 = o; o += slot((int64_t)B * d.E);
  L.st1 = o; o += slot((int64_t)B * 8 * 2 * 2);   // doubles
  L.st2 = o; o += slot((int64_t)B * 8 * 2 * 2);
  L.Y1 = o; o += slot(Mt * d.NP);
  const bool fused_tail = (d.old == 4 || d.old == 8 || d.old == 16 || d.old == 32);
  L.Y2 = o; o += fused_tail ? 0 : slot(Mt * d.NP);
  L.tok = o; o += slot((int64_t)B * d.E);
  L.c1 = o; o += slot((int64_t)B * d.E);
  L.c2 = o; o += slot((int64_t)B * d.E);
  L.musig = o; o += c->normalize ? slot((int64_t)B * 2 * d.C) : 0;
  L.asc = o; o += c->normalize ? slot((int64_t)B * d.K0) : 0;
  L.ash = o; o += c->normalize ? slot((int64_t)B * d.K0) : 0;
  L.smu = o; o += c->normalize ? slot((int64_t)B * d.E) : 0;
  L.ssg = o; o += c->normalize ? slot((int64_t)B * d.E) : 0;
  L.total = o;
  return L;
}

// destination of the new frames when the forward is one step of an autoregressive rollout (dpot_rollout_step)
struct RingOut { float* ring; float* pred; int slot0, Ttot, step; };

static int run_tail(const float* Y, const float* w2, const float* b2, const dpot_params* prm, int B, const Dims& d, int nout,
                    int act, const float* mu_c, const float* sg_c, float* y, const RingOut* ro, void* stream) {
  if (ro)
    return dpot_out_tail_ring(Y, w2, b2, prm->out4_w, prm->out4_b, B, d.h, d.h, d.P, d.old, nout, act, mu_c, sg_c, d.Co, y,
                              ro->ring, ro->pred, d.T, ro->slot0, ro->Ttot, ro->step, stream);
  return dpot_out_tail(Y, w2, b2, prm->out4_w, prm->out4_b, B, d.h, d.h, d.P, d.old, nout, act, mu_c, sg_c, d.Co, y, stream);
}

// The classification head (a spatial mean + three M = B skinny contractions, ~60 us of latency-bound work) depends only on
// the last latent, like the output head: it runs on a library-owned side stream between a fork event (after the last
// block) and a join event (end of the forward) -- also under stream capture, where the events become graph edges.
// One side stream per device: concurrent forwards from SEVERAL host threads on one device are not supported with it.
int g_fused_gn2 = 1;     // dpot_afno_set_fused_gn2: GroupNorm-2 + split inside the fused mixer (the unit's f tile stays in shared
                         // memory, f never goes to global memory): the separate 12 us pass per block is gone, the in-kernel
                         // pass costs ~8.7 us per launch -> +1.5 % on the rollout (tools/ab_gn2.py)
constexpr int AF_UNIT_CH = 128;
int g_cls_tc = 1;        // dpot_set_cls_engine: 1 = the cls head on the f16-split engine, 0 = CUDA-core skinny contractions
int g_cls_overlap = 0;   // measured (r02t): the side stream takes SMs from the persistent contraction kernels of the output head: -3 %
struct SideStream { cudaStream_t st = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static SideStream* side_stream() {
  static SideStream tab[64];
  SideStream& s = tab[cur_dev()];
  if (!s.st) {
    if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &s;
}

// ---- the f16-split tensor-core pipeline (DPOT_GEMM_TC16) ---------------------------------------
// Every activation that feeds a dense contraction is written by its producer as split fp16
// (DPOT_FMT_HL16, 4 bytes per element like fp32): conv0 epilogue -> z1, forward FFT -> S, GEMM
// epilogues -> O1 / hidden, GroupNorm-2 apply -> n2.  The residual stream, the inverse-FFT input
// and the head stay fp32.
static int forward_tc16(const dpot_config* cfg, const dpot_params* prm, const float* packed, const float* x, int t0, int B,
                        float* y, float* cls, float* ws, const Dims& d, const Packed& PL, const Work& WL, const RingOut* ro,
                        void* stream) {
  cudaStream_t st = as_stream(stream);
  const int Mt = B * d.n, Ms = B * d.km1 * d.km2, act = cfg->act, R = cfg->img_size;
  const int groups = 8;
  // PatchEmbed conv0 + act on the CUDA cores (N = 4P+3 is tiny), result stored split
  if (d.Kp != d.T * d.mid) DPOT_CUDA(cudaMemsetAsync(ws + WL.z1, 0, sizeof(float) * (size_t)Mt * d.Kp, st));
  DPOT_CALL(dpot_patch_embed(x, t0, packed + PL.W0p, packed + PL.rowbias0, cfg->normalize ? ws + WL.asc : nullptr,
                             cfg->normalize ? ws + WL.ash : nullptr, B, R, R, d.T, d.C, d.P, d.mid, act, ws + WL.z1, d.Kp,
                             DPOT_FMT_HL16, stream));
  // folded (conv 1x1 + pos_embed + TimeAggregator) GEMM (+ AdaIN) with GroupNorm-1 statistics
  float* lat = ws + WL.lat0;
  float* lat_next = ws + WL.lat1;
  double* st1 = reinterpret_cast<double*>(ws + WL.st1);
  double* st2 = reinterpret_cast<double*>(ws + WL.st2);
  {
    dpot_gemm_args g = gemm16_args(ws + WL.z1, d.Kp, packed + PL.WeffT16, d.Kp, lat, d.E, Mt, d.E, d.Kp, nullptr, DPOT_ACT_NONE);
    g.rowbias = packed + PL.bias_eff; g.rowbias_period = d.n; g.ldrb = d.E;
    if (cfg->normalize) { g.c_scale = ws + WL.ssg; g.c_shift = ws + WL.smu; g.c_rows_per_sample = d.n; }
    g.out_stats = st1; g.stats_groups = groups; g.stats_rows_per_sample = d.n;
    DPOT_CALL(dpot_gemm(&g, stream));
  }
  const int64_t kb = 2 * d.bs;
  for (int i = 0; i < d.depth; ++i) {
    const dpot_block_params& bp = prm->blocks[i];
    const float* pk = packed + PL.blocks + (int64_t)i * PL.blk_stride;
    // GroupNorm by reference: the consumers derive their affines from the raw statistics (no finalize launches)
    bool fused_gn2 = false;
    const bool gnref = (d.E / groups) % 8 == 0 && (reinterpret_cast<uintptr_t>(bp.norm2_w) | reinterpret_cast<uintptr_t>(bp.norm2_b)) % 16 == 0;
    if (fused_geometry(d) && dpot_afno_fused_supported(d.h, d.E, d.nb, d.km1, d.km2, groups)) {
      // the whole mixer in one launch: spectrum and hidden layer stay on chip; with whole GroupNorm groups inside a work
      // unit the kernel also applies GroupNorm-2 and writes the channel MLP's split-fp16 operand itself
      fused_gn2 = gnref && g_fused_gn2 && AF_UNIT_CH % (d.E / groups) == 0 && (d.E / groups) % 32 == 0;
      if (fused_gn2) {
        DPOT_CALL(dpot_afno_fused_gn2(lat, st1, bp.norm1_w, bp.norm1_b, groups, 1e-5f, B, d.h, d.E, d.nb, pk + PL.Wfus, act,
                                      nullptr /* f stays on chip */, nullptr, nullptr, ws + WL.n2, bp.norm2_w, bp.norm2_b, 1e-5f, stream));
      } else {
        DPOT_CUDA(cudaMemsetAsync(st2, 0, sizeof(double) * 2 * groups * B, st));
        DPOT_CALL(dpot_afno_fused(lat, st1, bp.norm1_w, bp.norm1_b, groups, 1e-5f, B, d.h, d.E, d.nb, pk + PL.Wfus, act,
                                  ws + WL.f, st2, nullptr, stream));
      }
    } else {
    DPOT_CALL(dpot_afno_fft_fwd16_gn(lat, st1, bp.norm1_w, bp.norm1_b, groups, 1e-5f, B, d.h, d.E, d.nb, d.km1, d.km2,
                                     ws + WL.S, stream));
    {
      // block-diagonal complex MLP as nb real GEMMs of size [Ms, 2bs] x [2bs, 2bs]
      dpot_gemm_args g = gemm16_args(ws + WL.S, 2 * d.E, pk + PL.Wc1_16, kb, ws + WL.O1, 0, Ms, (int)kb, (int)kb, pk + PL.bc1, act);
      g.batch = d.nb; g.strideA = kb; g.strideW = 2 * kb * kb; g.strideC = kb; g.strideBias = kb;
      out16(g, 2 * d.E);
      DPOT_CALL(dpot_gemm(&g, stream));
      g = gemm16_args(ws + WL.O1, 2 * d.E, pk + PL.Wc2_16, kb, ws + WL.S, 2 * d.E, Ms, (int)kb, (int)kb, pk + PL.bc2, DPOT_ACT_NONE);
      g.batch = d.nb; g.strideA = kb; g.strideW = 2 * kb * kb; g.strideC = kb; g.strideBias = kb;
      DPOT_CALL(dpot_gemm(&g, stream));     // O2 overwrites S as fp32 (the inverse FFT reads fp32)
    }
    DPOT_CUDA(cudaMemsetAsync(st2, 0, sizeof(double) * 2 * groups * B, st));
    DPOT_CALL(dpot_afno_fft_inv_gn(ws + WL.S, lat, st1, bp.norm1_w, bp.norm1_b, groups, 1e-5f, B, d.h, d.E, d.nb, d.km1, d.km2,
                                   ws + WL.f, st2, stream));
    }
    // GroupNorm-2 apply fused with the fp16 split of the channel-MLP input
    if (fused_gn2) {
    } else if (gnref) {
      DPOT_CALL(dpot_split_f16_gn(ws + WL.f, d.E, Mt, d.E, st2, bp.norm2_w, bp.norm2_b, groups, 1e-5f, d.n, ws + WL.n2, 2 * d.E, d.E, stream));
    } else {
      DPOT_CALL(dpot_gn_finalize(st2, bp.norm2_w, bp.norm2_b, B, d.n, d.E, groups, 1e-5f, ws + WL.sc2, ws + WL.sh2, stream));
      DPOT_CALL(dpot_split_f16(ws + WL.f, d.E, Mt, d.E, ws + WL.sc2, ws + WL.sh2, d.n, ws + WL.n2, 2 * d.E, d.E, stream));
    }
    {
      dpot_gemm_args g = gemm16_args(ws + WL.n2, d.E, pk + PL.fc1_16, d.E, ws + WL.hid, 0, Mt, d.hid, d.E, bp.fc1_b, act);
      out16(g, d.hid);
      DPOT_CALL(dpot_gemm(&g, stream));
      g = gemm16_args(ws + WL.hid, d.hid, pk + PL.fc2_16, d.hid, lat_next, d.E, Mt, d.E, d.hid, bp.fc2_b, DPOT_ACT_NONE);
      g.residual = lat; g.ldr = d.E;
      if (i + 1 < d.depth) { g.out_stats = st1; g.stats_groups = groups; g.stats_rows_per_sample = d.n; }
      else { g.C = ws + WL.n2; out16(g, d.E); }   // last block: the latent is only read by the output GEMM (and the cls mean) -> split fp16
      // hid > 2048 (DPOT-M / L / H): accumulation chains of <= dpot_tc16_set_chain halves (gemm.cu), partial sums in lat_next
      DPOT_CALL(dpot_gemm_chained(&g, dpot_tc16_set_chain(-1), lat_next, d.E, stream));
    }
    float* tmp = lat; lat = lat_next; lat_next = tmp;
  }

  SideStream* side = (cls && g_cls_overlap && d.depth > 0) ? side_stream() : nullptr;
  if (cls) {   // classification head (M = B rows: skinny CUDA-core GEMMs), on the side stream when available
    void* cs = stream;
    if (side) {
      DPOT_CUDA(cudaEventRecord(side->fork, st));
      DPOT_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
      cs = side->st;
    }
    if (d.depth > 0 && g_cls_tc) {
      // mean -> split token; three contractions on the f16-split engine with split intermediates
      DPOT_CALL(dpot_spatial_mean16s(ws + WL.n2, B, d.n, d.E, nullptr, ws + WL.tok, cs));
      dpot_gemm_args g = gemm16_args(ws + WL.tok, d.E, packed + PL.cls0_16, d.E, ws + WL.c1, 0, B, d.E, d.E, prm->cls0_b, act);
      out16(g, d.E);
      DPOT_CALL(dpot_gemm(&g, cs));
      g = gemm16_args(ws + WL.c1, d.E, packed + PL.cls2_16, d.E, ws + WL.c2, 0, B, d.E, d.E, prm->cls2_b, act);
      out16(g, d.E);
      DPOT_CALL(dpot_gemm(&g, cs));
      g = gemm16_args(ws + WL.c2, d.E, packed + PL.cls4_16, d.E, cls, d.ncls, B, d.ncls, d.E, prm->cls4_b, DPOT_ACT_NONE);
      DPOT_CALL(dpot_gemm(&g, cs));
    } else {
    if (d.depth > 0) DPOT_CALL(dpot_spatial_mean16(ws + WL.n2, B, d.n, d.E, ws + WL.tok, cs));
    else DPOT_CALL(dpot_spatial_mean(lat, B, d.n, d.E, ws + WL.tok, cs));
    dpot_gemm_args g = gemm_args(ws + WL.tok, d.E, prm->cls0_w, d.E, ws + WL.c1, d.E, B, d.E, d.E, prm->cls0_b, act, DPOT_GEMM_AUTO);
    DPOT_CALL(dpot_gemm(&g, cs));
    g = gemm_args(ws + WL.c1, d.E, prm->cls2_w, d.E, ws + WL.c2, d.E, B, d.E, d.E, prm->cls2_b, act, DPOT_GEMM_AUTO);
    DPOT_CALL(dpot_gemm(&g, cs));
    g = gemm_args(ws + WL.c2, d.E, prm->cls4_w, d.E, cls, d.ncls, B, d.ncls, d.E, prm->cls4_b, DPOT_ACT_NONE, DPOT_GEMM_AUTO);
    DPOT_CALL(dpot_gemm(&g, cs));
    }
    if (side) DPOT_CUDA(cudaEventRecord(side->join, side->st));
  }
  auto join_side = [&]() -> int {
    if (side) DPOT_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    return 0;
  };

  // output head: ConvTranspose as a GEMM on the split latent (written split by the last fc2 unless the cls head
  // needed it in fp32), then the fused per-pixel tail
  if (d.depth == 0) DPOT_CALL(dpot_split_f16(lat, d.E, Mt, d.E, nullptr, nullptr, 0, ws + WL.n2, 2 * d.E, d.E, stream));
  {
    dpot_gemm_args g = gemm16_args(ws + WL.n2, d.E, packed + PL.WtT16, d.E, ws + WL.Y1, d.NP, Mt, d.NP, d.E, packed + PL.bias_t, act);
    const int nout = d.Co * d.To;
    // tcgen05 tail: the GEMM writes one [hi 32 | lo 32] record per pixel (DPOT_FMT_HL16G32), the tail's TMA operand
    const bool tc_tail = d.old == 32 && dpot_out_tail_tc_supported(d.old, nout, d.Co) &&
                         (reinterpret_cast<uintptr_t>(y) % 16 == 0) && (!ro || (reinterpret_cast<uintptr_t>(ro->ring) % 16 == 0 &&
                                                                               (!ro->pred || reinterpret_cast<uintptr_t>(ro->pred) % 16 == 0)));
    if (tc_tail) { g.c_fmt = DPOT_FMT_HL16G32; g.ldc = 2 * (int64_t)d.NP; }
    DPOT_CALL(dpot_gemm(&g, stream));
    const bool fused_tail = (d.old == 4 || d.old == 8 || d.old == 16 || d.old == 32);
    if (cfg->normalize) DPOT_REQUIRE(d.C == d.Co, DPOT_E_UNSUPPORTED, "normalize=True needs in_channels == out_channels");
    const float* mu_c = nullptr; const float* sg_c = nullptr;
    if (cfg->normalize) {
      const float* mu = ws + WL.musig;
      float* mc = ws + WL.smu; float* sc = ws + WL.ssg;
      DPOT_CUDA(cudaMemcpy2DAsync(mc, sizeof(float) * d.C, mu, sizeof(float) * 2 * d.C, sizeof(float) * d.C, B, cudaMemcpyDeviceToDevice, st));
      DPOT_CUDA(cudaMemcpy2DAsync(sc, sizeof(float) * d.C, mu + d.C, sizeof(float) * 2 * d.C, sizeof(float) * d.C, B, cudaMemcpyDeviceToDevice, st));
      mu_c = mc; sg_c = sc;
    }
    if (tc_tail) {
      DPOT_CALL(dpot_out_tail_tc(ws + WL.Y1, prm->out2_w, prm->out2_b, prm->out4_w, prm->out4_b, B, d.h, d.h, d.P, d.old, nout, act,
                                 mu_c, sg_c, d.Co, y, ro ? ro->ring : nullptr, ro ? ro->pred : nullptr, d.T, ro ? ro->slot0 : 0,
                                 ro ? ro->Ttot : 0, ro ? ro->step : 0, stream));
      return join_side();
    }
    if (fused_tail) {
      DPOT_CALL(run_tail(ws + WL.Y1, prm->out2_w, prm->out2_b, prm, B, d, nout, act, mu_c, sg_c, y, ro, stream));
    } else {
      g = gemm_args(ws + WL.Y1, d.old, prm->out2_w, d.old, ws + WL.Y2, d.old, Mt * d.P * d.P, d.old, d.old, prm->out2_b, act, DPOT_GEMM_AUTO);
      DPOT_CALL(dpot_gemm(&g, stream));
      DPOT_CALL(run_tail(ws + WL.Y2, nullptr, nullptr, prm, B, d, nout, act, mu_c, sg_c, y, ro, stream));
    }
  }
  return join_side();
}

}  // namespace dpot

using namespace dpot;

extern "C" int64_t dpot_packed_floats(const dpot_config* cfg) {
  Dims d;
  if (make_dims(cfg, d) != 0) return -1;
  return packed_layout(d).total;
}

extern "C" int64_t dpot_workspace_floats(const dpot_config* cfg, int32_t B) {
  Dims d;
  if (make_dims(cfg, d) != 0 || B <= 0) return -1;
  return work_layout(d, cfg, B).total;
}

extern "C" int dpot_pack_weights(const dpot_config* cfg, const dpot_params* prm, float* packed, void* stream) {
  Dims d;
  DPOT_CALL(make_dims(cfg, d));
  DPOT_REQUIRE(prm && packed && prm->blocks, DPOT_E_BADARG, "dpot_pack_weights: null pointer");
  const Packed L = packed_layout(d);
  DPOT_CALL(dpot_pack_patch(prm->pe0_w, prm->pe0_b, prm->grid_x, prm->grid_y, prm->grid_t, d.mid, d.C, d.P, d.h, d.h,
                            d.T, packed + L.W0p, packed + L.rowbias0, stream));
  DPOT_CALL(dpot_fold_timeagg(prm->pe2_w, prm->pe2_b, prm->pos_embed, prm->tagg_w, prm->temb, d.T, d.E, d.mid, d.n,
                              d.Kp, packed + L.WeffT, packed + L.bias_eff, stream));
  for (int i = 0; i < d.depth; ++i) {
    float* base = packed + L.blocks + (int64_t)i * L.blk_stride;
    const dpot_block_params& b = prm->blocks[i];
    DPOT_CALL(dpot_pack_afno(b.w1, b.b1, d.nb, d.bs, base + L.Wc1, base + L.bc1, stream));
    DPOT_CALL(dpot_pack_afno(b.w2, b.b2, d.nb, d.bs, base + L.Wc2, base + L.bc2, stream));
  }
  DPOT_CALL(dpot_pack_out(prm->out0_w, prm->out0_b, d.E, d.old, d.P, packed + L.WtT, packed + L.bias_t, stream));
  if (use_tc16(d, DPOT_GEMM_AUTO)) {   // split-fp16 copies of every GEMM weight
    DPOT_CALL(dpot_split_f16(packed + L.WeffT, d.Kp, d.E, d.Kp, nullptr, nullptr, 0, packed + L.WeffT16, 2 * d.Kp, d.Kp, stream));
    DPOT_CALL(dpot_split_f16(packed + L.WtT, d.E, d.NP, d.E, nullptr, nullptr, 0, packed + L.WtT16, 2 * d.E, d.E, stream));
    if (prm->cls0_w && prm->cls2_w && prm->cls4_w) {   // the cls head on the same engine (M = B rows: 8 tiles, latency ~1/2 of the CUDA-core form)
      DPOT_CALL(dpot_split_f16(prm->cls0_w, d.E, d.E, d.E, nullptr, nullptr, 0, packed + L.cls0_16, 2 * d.E, d.E, stream));
      DPOT_CALL(dpot_split_f16(prm->cls2_w, d.E, d.E, d.E, nullptr, nullptr, 0, packed + L.cls2_16, 2 * d.E, d.E, stream));
      DPOT_CALL(dpot_split_f16(prm->cls4_w, d.E, d.ncls, d.E, nullptr, nullptr, 0, packed + L.cls4_16, 2 * d.E, d.E, stream));
    }
    for (int i = 0; i < d.depth; ++i) {
      float* base = packed + L.blocks + (int64_t)i * L.blk_stride;
      const dpot_block_params& b = prm->blocks[i];
      const int64_t kb = 2 * d.bs;
      DPOT_CALL(dpot_split_f16(base + L.Wc1, kb, (int64_t)d.nb * kb, (int)kb, nullptr, nullptr, 0, base + L.Wc1_16, 2 * kb, kb, stream));
      DPOT_CALL(dpot_split_f16(base + L.Wc2, kb, (int64_t)d.nb * kb, (int)kb, nullptr, nullptr, 0, base + L.Wc2_16, 2 * kb, kb, stream));
      DPOT_CALL(dpot_split_f16(b.fc1_w, d.E, d.hid, d.E, nullptr, nullptr, 0, base + L.fc1_16, 2 * d.E, d.E, stream));
      DPOT_CALL(dpot_split_f16(b.fc2_w, d.hid, d.E, d.hid, nullptr, nullptr, 0, base + L.fc2_16, 2 * d.hid, d.hid, stream));
      if (fused_geometry(d)) DPOT_CALL(dpot_afno_fused_pack(b.w1, b.b1, b.w2, b.b2, d.nb, d.bs, base + L.Wfus, stream));
    }
  }
  return 0;
}

extern "C" int dpot_forward(const dpot_config* cfg, const dpot_params* prm, const float* packed, const float* x,
                            int32_t B, float* y, float* cls, float* ws, int32_t engine, void* stream) {
  return dpot_forward_ring(cfg, prm, packed, x, 0, B, y, cls, ws, engine, stream);
}

static int forward_impl(const dpot_config* cfg, const dpot_params* prm, const float* packed, const float* x, int32_t t0,
                        int32_t B, float* y, float* cls, float* ws, int32_t engine, const dpot::RingOut* ro, void* stream);

extern "C" int dpot_forward_ring(const dpot_config* cfg, const dpot_params* prm, const float* packed, const float* x,
                                 int32_t t0, int32_t B, float* y, float* cls, float* ws, int32_t engine, void* stream) {
  return forward_impl(cfg, prm, packed, x, t0, B, y, cls, ws, engine, nullptr, stream);
}

// One autoregressive step on a ring window (evaluate.py:192-208): im = model(window at t0); the T_out new frames
// replace the oldest slots (t0 + j) % T of the ring and land in pred[..., step*T_out + j, :].  The caller advances
// t0 by T_out.  y: scratch [B,X,Y,T_out,C_out] (used only by geometries the fused tail kernel does not serve).
extern "C" int dpot_rollout_step(const dpot_config* cfg, const dpot_params* prm, const float* packed, float* ring,
                                 int32_t t0, int32_t B, float* y, float* cls, float* ws, int32_t engine, float* pred,
                                 int32_t pred_frames, int32_t step, void* stream) {
  DPOT_REQUIRE(cfg && ring && y, DPOT_E_BADARG, "dpot_rollout_step: null pointer");
  DPOT_REQUIRE(cfg->in_channels == cfg->out_channels, DPOT_E_BADARG, "dpot_rollout_step: needs out_channels == in_channels");
  DPOT_REQUIRE(!pred || (step >= 0 && (step + 1) * cfg->out_timesteps <= pred_frames), DPOT_E_BADARG,
               "dpot_rollout_step: step %d outside the prediction tensor (%d frames)", step, pred_frames);
  dpot::RingOut ro;
  ro.ring = ring; ro.pred = pred; ro.slot0 = t0; ro.Ttot = pred_frames; ro.step = step;
  return forward_impl(cfg, prm, packed, ring, t0, B, y, cls, ws, engine, &ro, stream);
}

static int forward_impl(const dpot_config* cfg, const dpot_params* prm, const float* packed, const float* x, int32_t t0,
                        int32_t B, float* y, float* cls, float* ws, int32_t engine, const dpot::RingOut* ro, void* stream) {
  Dims d;
  DPOT_CALL(make_dims(cfg, d));
  DPOT_REQUIRE(prm && packed && x && y && ws && prm->blocks && B > 0, DPOT_E_BADARG, "dpot_forward: null pointer / bad B");
  DPOT_REQUIRE(t0 >= 0 && t0 < d.T, DPOT_E_BADARG, "dpot_forward_ring: t0=%d outside [0,%d)", t0, d.T);
  const Packed PL = packed_layout(d);
  const Work WL = work_layout(d, cfg, B);
  cudaStream_t st = as_stream(stream);
  const int Mt = B * d.n, Ms = B * d.km1 * d.km2, act = cfg->act, R = cfg->img_size;
  const int groups = 8;

  // engine: AUTO / TC16 = the f16-split tensor-core pipeline when available; TC = fp32 operands with 3xTF32
  // tensor cores wherever a problem fits (CUDA cores elsewhere); SIMT = CUDA cores only.
  const bool tc16 = use_tc16(d, engine);
  if (engine == DPOT_GEMM_TC16 && !tc16) {
    set_error("dpot_forward: the f16-split engine is not available for this device / configuration");
    return DPOT_E_UNSUPPORTED;
  }
  if (engine == DPOT_GEMM_TC || engine == DPOT_GEMM_TC16) engine = DPOT_GEMM_AUTO;

  // ---- input normalisation (normalize=True only), models/dpot.py:366-370
  if (cfg->normalize) {
    DPOT_CALL(dpot_input_stats(x, B, (int64_t)R * R * d.T * d.C, d.C, d.P * d.P, ws + WL.musig, ws + WL.asc, ws + WL.ash, stream));
    dpot_gemm_args g = gemm_args(ws + WL.musig, 2 * d.C, prm->mu_w, 2 * d.C, ws + WL.smu, d.E, B, d.E, 2 * d.C, prm->mu_b, DPOT_ACT_NONE, engine);
    DPOT_CALL(dpot_gemm(&g, stream));
    g = gemm_args(ws + WL.musig, 2 * d.C, prm->sigma_w, 2 * d.C, ws + WL.ssg, d.E, B, d.E, 2 * d.C, prm->sigma_b, DPOT_ACT_NONE, engine);
    DPOT_CALL(dpot_gemm(&g, stream));
  }

  if (tc16) return forward_tc16(cfg, prm, packed, x, t0, B, y, cls, ws, d, PL, WL, ro, stream);

  // ---- PatchEmbed conv0 + act, coordinate channels folded into rowbias0
  if (d.Kp != d.T * d.mid) DPOT_CUDA(cudaMemsetAsync(ws + WL.z1, 0, sizeof(float) * (size_t)Mt * d.Kp, st));
  DPOT_CALL(dpot_patch_embed(x, t0, packed + PL.W0p, packed + PL.rowbias0, cfg->normalize ? ws + WL.asc : nullptr,
                             cfg->normalize ? ws + WL.ash : nullptr, B, R, R, d.T, d.C, d.P, d.mid, act, ws + WL.z1, d.Kp,
                             DPOT_FMT_F32, stream));
  // ---- folded (conv 1x1 + pos_embed + TimeAggregator) GEMM (+ AdaIN), models/dpot.py:201,378-387
  float* lat = ws + WL.lat0;
  float* lat_next = ws + WL.lat1;
  {
    dpot_gemm_args g = gemm_args(ws + WL.z1, d.Kp, packed + PL.WeffT, d.Kp, lat, d.E, Mt, d.E, d.Kp, nullptr, DPOT_ACT_NONE, engine);
    g.rowbias = packed + PL.bias_eff; g.rowbias_period = d.n; g.ldrb = d.E;
    if (cfg->normalize) { g.c_scale = ws + WL.ssg; g.c_shift = ws + WL.smu; g.c_rows_per_sample = d.n; }
    g.out_stats = reinterpret_cast<double*>(ws + WL.st1); g.stats_groups = 8; g.stats_rows_per_sample = d.n;
    DPOT_CALL(dpot_gemm(&g, stream));
  }

  // ---- blocks, models/dpot.py:165-180.  GroupNorm-1 statistics of the block input are produced by the
  // epilogue of the GEMM that wrote it (time aggregation above, fc2 of the previous block below).
  double* st1 = reinterpret_cast<double*>(ws + WL.st1);
  double* st2 = reinterpret_cast<double*>(ws + WL.st2);
  for (int i = 0; i < d.depth; ++i) {
    const dpot_block_params& bp = prm->blocks[i];
    const float* pk = packed + PL.blocks + (int64_t)i * PL.blk_stride;
    DPOT_CALL(dpot_gn_finalize(st1, bp.norm1_w, bp.norm1_b, B, d.n, d.E, groups, 1e-5f, ws + WL.sc1, ws + WL.sh1, stream));
    DPOT_CALL(dpot_afno_fft_fwd(lat, ws + WL.sc1, ws + WL.sh1, B, d.h, d.E, d.nb, d.km1, d.km2, ws + WL.S, 1.0f, stream));
    {
      dpot_gemm_args g = gemm_args(ws + WL.S, 2 * d.E, pk + PL.Wc1, 2 * d.bs, ws + WL.O1, 2 * d.E, Ms, 2 * d.bs, 2 * d.bs, pk + PL.bc1, act, engine);
      g.batch = d.nb; g.strideA = 2 * d.bs; g.strideW = (int64_t)4 * d.bs * d.bs; g.strideC = 2 * d.bs; g.strideBias = 2 * d.bs;
      DPOT_CALL(dpot_gemm(&g, stream));
      g.A = ws + WL.O1; g.W = pk + PL.Wc2; g.C = ws + WL.S; g.bias = pk + PL.bc2; g.act = DPOT_ACT_NONE;
      DPOT_CALL(dpot_gemm(&g, stream));
    }
    DPOT_CUDA(cudaMemsetAsync(st2, 0, sizeof(double) * 2 * groups * B, st));
    DPOT_CALL(dpot_afno_fft_inv(ws + WL.S, lat, ws + WL.sc1, ws + WL.sh1, B, d.h, d.E, d.nb, d.km1, d.km2, ws + WL.f, st2, groups, 1.0f, stream));
    DPOT_CALL(dpot_gn_finalize(st2, bp.norm2_w, bp.norm2_b, B, d.n, d.E, groups, 1e-5f, ws + WL.sc2, ws + WL.sh2, stream));
    {
      dpot_gemm_args g = gemm_args(ws + WL.f, d.E, bp.fc1_w, d.E, ws + WL.hid, d.hid, Mt, d.hid, d.E, bp.fc1_b, act, engine);
      g.a_scale = ws + WL.sc2; g.a_shift = ws + WL.sh2; g.a_rows_per_sample = d.n;
      DPOT_CALL(dpot_gemm(&g, stream));
      g = gemm_args(ws + WL.hid, d.hid, bp.fc2_w, d.hid, lat_next, d.E, Mt, d.E, d.hid, bp.fc2_b, DPOT_ACT_NONE, engine);
      g.residual = lat; g.ldr = d.E;
      if (i + 1 < d.depth) { g.out_stats = st1; g.stats_groups = groups; g.stats_rows_per_sample = d.n; }
      DPOT_CALL(dpot_gemm(&g, stream));
    }
    float* t = lat; lat = lat_next; lat_next = t;
  }

  // ---- classification head, models/dpot.py:394-395
  if (cls) {
    DPOT_CALL(dpot_spatial_mean(lat, B, d.n, d.E, ws + WL.tok, stream));
    dpot_gemm_args g = gemm_args(ws + WL.tok, d.E, prm->cls0_w, d.E, ws + WL.c1, d.E, B, d.E, d.E, prm->cls0_b, act, engine);
    DPOT_CALL(dpot_gemm(&g, stream));
    g = gemm_args(ws + WL.c1, d.E, prm->cls2_w, d.E, ws + WL.c2, d.E, B, d.E, d.E, prm->cls2_b, act, engine);
    DPOT_CALL(dpot_gemm(&g, stream));
    g = gemm_args(ws + WL.c2, d.E, prm->cls4_w, d.E, cls, d.ncls, B, d.ncls, d.E, prm->cls4_b, DPOT_ACT_NONE, engine);
    DPOT_CALL(dpot_gemm(&g, stream));
  }

  // ---- output head, models/dpot.py:315-321,397-401
  {
    dpot_gemm_args g = gemm_args(lat, d.E, packed + PL.WtT, d.E, ws + WL.Y1, d.NP, Mt, d.NP, d.E, packed + PL.bias_t, act, engine);
    DPOT_CALL(dpot_gemm(&g, stream));
    const float* mu = cfg->normalize ? ws + WL.musig : nullptr;
    const int nout = d.Co * d.To;
    const bool fused_tail = (d.old == 4 || d.old == 8 || d.old == 16 || d.old == 32);
    // musig is [B, 2C] = [mu | sigma]; the tail indexes [b*Co + c], so it needs C == Co and a compact copy
    if (cfg->normalize) DPOT_REQUIRE(d.C == d.Co, DPOT_E_UNSUPPORTED, "normalize=True needs in_channels == out_channels");
    const float* mu_c = nullptr; const float* sg_c = nullptr;
    if (mu) {
      // compact [B,C] views: reuse the (now free) smu/ssg buffers
      float* mc = ws + WL.smu; float* sc = ws + WL.ssg;
      DPOT_CUDA(cudaMemcpy2DAsync(mc, sizeof(float) * d.C, mu, sizeof(float) * 2 * d.C, sizeof(float) * d.C, B, cudaMemcpyDeviceToDevice, st));
      DPOT_CUDA(cudaMemcpy2DAsync(sc, sizeof(float) * d.C, mu + d.C, sizeof(float) * 2 * d.C, sizeof(float) * d.C, B, cudaMemcpyDeviceToDevice, st));
      mu_c = mc; sg_c = sc;
    }
    if (fused_tail) {
      DPOT_CALL(run_tail(ws + WL.Y1, prm->out2_w, prm->out2_b, prm, B, d, nout, act, mu_c, sg_c, y, ro, stream));
    } else {
      g = gemm_args(ws + WL.Y1, d.old, prm->out2_w, d.old, ws + WL.Y2, d.old, Mt * d.P * d.P, d.old, d.old, prm->out2_b, act, engine);
      DPOT_CALL(dpot_gemm(&g, stream));
      DPOT_CALL(run_tail(ws + WL.Y2, nullptr, nullptr, prm, B, d, nout, act, mu_c, sg_c, y, ro, stream));
    }
  }
  return 0;
}

// 1: the classification head of dpot_forward* runs on a library-owned side stream, concurrently with the output head (default 0)
extern "C" void dpot_set_cls_overlap(int32_t on) { dpot::g_cls_overlap = on ? 1 : 0; }
extern "C" int dpot_afno_set_fused_gn2(int32_t on) { const int prev = dpot::g_fused_gn2; if (on >= 0) dpot::g_fused_gn2 = on ? 1 : 0; return prev; }
extern "C" void dpot_set_cls_engine(int32_t tc) { dpot::g_cls_tc = tc ? 1 : 0; }
