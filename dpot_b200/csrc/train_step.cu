// The training step of DPOTNet as three stream-ordered calls (no allocation, no synchronisation):
//   dpot_train_prepare   -- once per optimizer step: packed / folded / split weights (models/dpot.py:183-232 folds)
//   dpot_train_forward   -- DPOTNet.forward (models/dpot.py:364-403) keeping what backward needs on a caller-owned tape
//   dpot_train_backward  -- its autograd (train_temporal.py:227 loss.backward()): every parameter gradient and dL/dx
// Every dense contraction of forward AND backward runs on the f16-split tcgen05 engine (gemm_tc16.cu): data gradients
// read the weights in their forward layout (w_trans), weight gradients read both token-major activations as stored
// (a_trans + w_trans) in <= 1024-token chunks whose partial results are summed by the un-packing kernels; the
// activation derivative rides in the data-gradient epilogue and its result leaves the engine already split.
// Gradients are carried scaled by a power of two S (S * max|dL/dy| in [1, 2)) so that small gradient entries keep full
// fp16 significands in the split operands; every final gradient is multiplied by 1 / S as it is written.
#include "common.cuh"
#include "model_layout.cuh"
#include "train_kernels.cuh"
#include <string.h>

namespace dpot {
namespace {

constexpr int GROUPS = 8;
constexpr float GN_EPS = 1e-5f;

struct Bump {   // bump allocator over a float arena (sizes in floats, 256-byte granules)
  int64_t o = 0;
  int64_t take(int64_t n) { const int64_t r = o; o += slot(n); return r; }
};

struct TapeL {
  int64_t z1, blocks, blk_stride, alat16, Y1pre, tok, c1pre, c2pre, c1, c2, Y1_16, Y2pre, Y2_16, total;   // Y1_16.. : generic tail only
  // offsets inside a block slab
  int64_t lat, st1, st2, S, O1pre, O1, f, n2, hpre, hid;
};
TapeL tape_layout(const Dims& d, int B) {
  TapeL L; Bump a;
  const int64_t Mt = (int64_t)B * d.n, Ms = (int64_t)B * d.km1 * d.km2;
  L.z1 = a.take(Mt * d.Kp);
  Bump b;
  L.lat = b.take(Mt * d.E);
  L.st1 = b.take((int64_t)B * GROUPS * 4);
  L.st2 = b.take((int64_t)B * GROUPS * 4);
  L.S = b.take(Ms * 2 * d.E);
  L.O1pre = b.take(Ms * 2 * d.E);
  L.O1 = b.take(Ms * 2 * d.E);
  L.f = b.take(Mt * d.E);
  L.n2 = b.take(Mt * d.E);
  L.hpre = b.take(Mt * d.hid);
  L.hid = b.take(Mt * d.hid);
  L.blk_stride = b.o;
  L.blocks = a.take(b.o * d.depth);
  L.alat16 = a.take(Mt * d.E);
  L.Y1pre = a.take(Mt * d.NP);
  L.tok = a.take((int64_t)B * d.E);
  L.c1pre = a.take((int64_t)B * d.E);
  L.c2pre = a.take((int64_t)B * d.E);
  L.c1 = a.take((int64_t)B * d.E);
  L.c2 = a.take((int64_t)B * d.E);
  const bool gen_tail = d.old != 32;
  L.Y1_16 = a.take(gen_tail ? Mt * d.NP : 0);
  L.Y2pre = a.take(gen_tail ? Mt * d.NP : 0);
  L.Y2_16 = a.take(gen_tail ? Mt * d.NP : 0);
  L.total = a.o;
  return L;
}

inline int midp_of(const Dims& d) { return (int)round_up(d.mid, 8); }
struct PrepL { int64_t wts16, Wsum16, bp16, W2p16, W2T16, o2_16, o4_16, o4p16, total; };   // split-fp16 fold operands (prepare), reused by backward
PrepL prep_layout(const Dims& d) {
  PrepL L; Bump a;
  L.wts16 = a.take((int64_t)d.T * d.E * d.E);         // [(t,i), hi E | lo E]
  L.Wsum16 = a.take((int64_t)d.E * d.E);
  L.bp16 = a.take((int64_t)d.E * d.n);
  L.W2p16 = a.take((int64_t)d.E * midp_of(d));        // conv 1x1 weight [E, mid] zero-padded to a multiple of 8 columns
  L.W2T16 = a.take((int64_t)midp_of(d) * d.E);        // its transpose [mid, E]
  L.o2_16 = a.take((int64_t)d.old * d.old);          // generic tail: out_layer[2] / [4] weights split, [4] also zero-padded to 8 rows
  L.o4_16 = a.take((int64_t)8 * d.old);
  L.o4p16 = a.take((int64_t)8 * d.old);
  L.total = a.o;
  return L;
}

int64_t kchunk_for(int64_t K, int64_t tiles) {   // contraction chunk: <= 1024 deep, smaller when the output alone cannot fill the SMs
  if (K <= 1024 && tiles >= 96) return 0;        // one chunk
  int64_t ch = 1024;
  while (ch > 256 && tiles * ceil_div(K, ch) < 148) ch /= 2;
  return ch;
}
int64_t nchunks(int64_t K, int64_t ch) { return ch > 0 ? ceil_div(K, ch) : 1; }

struct ScratchL {
  // prepare
  int64_t W2T;
  // forward
  int64_t O2, Y1g;
  // backward
  int64_t gA, gB, g16, g1_16, dn2, df, dO2, dO1, dS, dn1, slabs, g1t, z1pre, gz, dWeffT, Gp16, dbe, dbe16, dWsum, dwt,
      dW2s, dbp, dtemb, gn, cls, dbl, scale, pslabs, tparts, g3_16, g2_16, Y3;
  // double accumulators (offsets in DOUBLES from dbl): per block [db2 E | db1 hid | dbc2 2E | dbc1 2E], then the rest
  int64_t d_blk, d_blk_stride, d_tail, d_bias_t, d_be, d_rb, d_gt, d_total;
  int64_t total;
};
ScratchL scratch_layout(const Dims& d, int B) {
  ScratchL L; memset(&L, 0, sizeof(L));
  const int64_t Mt = (int64_t)B * d.n, Ms = (int64_t)B * d.km1 * d.km2, E = d.E;
  int64_t mx = 0;
  { Bump a;   // prepare
    L.W2T = a.take((int64_t)d.mid * E);
    mx = a.o; }
  { Bump a;   // forward
    L.O2 = a.take(Ms * 2 * E); L.Y1g = a.take(Mt * d.NP); L.Y3 = a.take(Mt * d.P * d.P * 8);
    if (a.o > mx) mx = a.o; }
  { Bump a;   // backward
    const int64_t W = E > d.hid ? E : d.hid;
    L.gA = a.take(Mt * E); L.gB = a.take(Mt * E); L.g16 = a.take(Mt * E); L.g1_16 = a.take(Mt * W);
    L.dn2 = a.take(Mt * E); L.df = a.take(Mt * E); L.dO2 = a.take(Ms * 2 * E); L.dO1 = a.take(Ms * 2 * E); L.dS = a.take(Ms * 2 * E);
    L.dn1 = a.take(Mt * E);
    const int64_t kt = 8 + ceil_div(Mt, 256), ks = 8 + ceil_div(Ms, 256);   // generous chunk counts (kchunk_for >= 256)
    int64_t sl = kt * (int64_t)d.hid * E;
    if (kt * (int64_t)d.NP * E > sl) sl = kt * (int64_t)d.NP * E;
    if (kt * E * d.Kp > sl) sl = kt * E * d.Kp;
    if (ks * (int64_t)d.nb * 4 * d.bs * d.bs > sl) sl = ks * (int64_t)d.nb * 4 * d.bs * d.bs;
    if (ceil_div(d.NP, 1024) * Mt * E > sl) sl = ceil_div(d.NP, 1024) * Mt * E;
    if (ceil_div(d.hid, 1024) * Mt * E > sl) sl = ceil_div(d.hid, 1024) * Mt * E;        // split-K partials of the head's data gradient
    L.slabs = a.take(sl);
    L.g1t = a.take(Mt * d.NP); L.z1pre = a.take(Mt * d.Kp); L.gz = a.take(Mt * d.Kp);
    L.dWeffT = a.take(E * d.Kp); L.Gp16 = a.take((int64_t)d.T * E * midp_of(d)); L.dbe = a.take((int64_t)d.n * E);
    L.dbe16 = a.take((int64_t)d.n * E);
    L.dWsum = a.take(E * E); L.dwt = a.take((int64_t)d.T * E * E); L.dW2s = a.take((int64_t)d.T * E * midp_of(d)); L.dbp = a.take(E * d.n);
    L.pslabs = a.take((int64_t)2 * 160 * d.mid * d.K0);         // per-block partials of the PatchEmbed weight gradient (<= 2 per SM)
    L.dtemb = a.take((int64_t)d.T * E);
    L.tparts = a.take(tk_tail_bwd_part_floats(3 * 160));
    L.g3_16 = a.take(d.old != 32 ? Mt * d.P * d.P * 8 : 0);
    L.g2_16 = a.take(d.old != 32 ? Mt * d.NP : 0);
    L.gn = a.take(5 * (int64_t)B * E + 2 * (int64_t)B * GROUPS);
    L.cls = a.take(8 * (int64_t)B * E + 3 * E * E + (int64_t)d.ncls * E + 4096);
    L.scale = a.take(64);
    // doubles
    int64_t q = 0;
    L.d_blk = q; L.d_blk_stride = E + d.hid + 4 * E; q += L.d_blk_stride * d.depth;
    L.d_tail = q; q += 1024 + 32 + 8 * 32 + 8 + 8;
    L.d_bias_t = q; q += d.NP;
    L.d_be = q; q += (int64_t)d.n * E;
    L.d_rb = q; q += (int64_t)d.n * d.Kp;
    L.d_gt = q; q += (int64_t)2 * d.NP + (int64_t)d.P * d.P * 8;     // generic tail: per-(u,v) column sums of g2, g1, g3
    L.d_total = q;
    L.dbl = a.take(2 * q + 2);
    if (a.o > mx) mx = a.o; }
  L.total = mx;
  return L;
}

bool train_ok(const dpot_config* c, const Dims& d) {
  if (c->normalize) return false;
  if (!use_tc16(d, DPOT_GEMM_AUTO)) return false;
  if (d.depth < 1) return false;
  if ((d.E / GROUPS) % 8 != 0 || d.E % GROUPS != 0) return false;
  const bool fast_tail = dpot_out_tail_tc_supported(d.old, d.Co * d.To, d.Co) && tk_tail_bwd_supported(d.old, d.Co * d.To);
  if (!fast_tail && !(d.old % 8 == 0 && d.old >= 8 && d.Co * d.To <= 8)) return false;   // else: the generic tail on batched contractions
  if (!tk_patch_bwd_supported(d.mid, d.T, d.K0)) return false;
  if (d.E % 8 != 0 || d.n % 8 != 0) return false;                   // fold contractions on the f16-split engine
  if ((2 * d.bs) % 8 != 0 || d.hid % 8 != 0 || d.NP % 8 != 0 || d.Kp % 8 != 0) return false;
  if (d.C != d.Co) return false;
  return true;
}

// split-fp16 matrix handle: `cols` columns, row = [hi cols halves | lo cols halves]
dpot_gemm_args g16(const float* A16, int64_t a_cols, const float* W16, int64_t w_cols, float* C, int64_t ldc, int M, int N, int K) {
  return gemm16_args(A16, a_cols, W16, w_cols, C, ldc, M, N, K, nullptr, DPOT_ACT_NONE);
}
void ksplit(dpot_gemm_args& g, int64_t chunk, int64_t slab) {
  if (chunk > 0 && chunk < g.K) { g.k_split = (int)ceil_div(g.K, chunk); g.k_chunk = chunk; g.strideC_split = slab; }
}

// weight gradient dW[N, K] = g[Mtok, N]^T x[Mtok, K] (both stored token-major, split) -> `*ns` partial results in slabs
int wgrad16(const float* g16p, int64_t g_cols, const float* x16p, int64_t x_cols, int N, int K, int Mtok, int batch, float* slabs,
            int* ns, void* stream) {
  dpot_gemm_args a = g16(g16p, g_cols, x16p, x_cols, slabs, K, N, K, Mtok);
  a.a_trans = 1; a.w_trans = 1;
  const int64_t tiles = ceil_div(N, 128) * ceil_div(K, 128) * batch;
  if (batch > 1) { a.batch = batch; a.strideA = N; a.strideW = K; a.strideC = (int64_t)N * K; }
  const int64_t ch = kchunk_for(Mtok, tiles);
  ksplit(a, ch, (int64_t)batch * N * K);
  *ns = a.k_split > 1 ? a.k_split : 1;
  return dpot_gemm(&a, stream);
}

}  // namespace
}  // namespace dpot

using namespace dpot;

extern "C" int dpot_train_supported(const dpot_config* cfg) {
  Dims d;
  if (make_dims(cfg, d) != 0) return 0;
  return train_ok(cfg, d) ? 1 : 0;
}
extern "C" int64_t dpot_train_tape_floats(const dpot_config* cfg, int32_t B) {
  Dims d;
  if (make_dims(cfg, d) != 0 || B <= 0) return -1;
  return tape_layout(d, B).total;
}
extern "C" int64_t dpot_train_scratch_floats(const dpot_config* cfg, int32_t B) {
  Dims d;
  if (make_dims(cfg, d) != 0 || B <= 0) return -1;
  return scratch_layout(d, B).total;
}
extern "C" int64_t dpot_train_wprep_floats(const dpot_config* cfg) {
  Dims d;
  if (make_dims(cfg, d) != 0) return -1;
  return prep_layout(d).total;
}

// ---------------------------------------------------------------------------------------------------------------------
extern "C" int dpot_train_prepare(const dpot_config* cfg, const dpot_params* prm, float* packed, float* wprep, float* scratch,
                                  void* stream) {
  Dims d;
  DPOT_CALL(make_dims(cfg, d));
  DPOT_REQUIRE(train_ok(cfg, d), DPOT_E_UNSUPPORTED, "dpot_train_prepare: configuration not served by the fused training step");
  DPOT_REQUIRE(prm && packed && wprep && scratch && prm->blocks, DPOT_E_BADARG, "dpot_train_prepare: null pointer");
  cudaStream_t st = as_stream(stream);
  const Packed L = packed_layout(d);
  const PrepL PL = prep_layout(d);
  const ScratchL SL = scratch_layout(d, 1);
  const int E = d.E;
  DPOT_CALL(tk_pack_patch(prm->pe0_w, prm->pe0_b, prm->grid_x, prm->grid_y, prm->grid_t, d.mid, d.C, d.P, d.h, d.h, d.T,
                          packed + L.W0p, packed + L.rowbias0, st));
  // fold conv 1x1 + pos_embed + time aggregation with the contraction engine (the double-precision fold kernels of the
  // inference packer take ~1 ms: fine once per checkpoint, not once per optimizer step)
  float* wts16 = wprep + PL.wts16; float* Wsum16 = wprep + PL.Wsum16; float* bp16 = wprep + PL.bp16;
  float* W2p16 = wprep + PL.W2p16; float* W2T16 = wprep + PL.W2T16;
  float* W2T = scratch + SL.W2T;
  const int midp = midp_of(d);
  DPOT_CALL(tk_tagg_scale16(prm->tagg_w, prm->temb, d.T, E, reinterpret_cast<__half*>(wts16), reinterpret_cast<__half*>(Wsum16), st));
  DPOT_CALL(tk_tagg_bp16(prm->pe2_b, prm->pos_embed, E, d.n, reinterpret_cast<__half*>(bp16), st));
  DPOT_CALL(tk_pad_split(prm->pe2_w, d.mid, E, d.mid, midp, reinterpret_cast<__half*>(W2p16), st));
  DPOT_CALL(tk_transpose(prm->pe2_w, d.mid, W2T, E, E, d.mid, st));
  DPOT_CALL(dpot_split_f16(W2T, E, d.mid, E, nullptr, nullptr, 0, W2T16, 2 * E, E, stream));
  DPOT_CUDA(cudaMemsetAsync(packed + L.WeffT, 0, sizeof(float) * (size_t)E * d.Kp, st));
  {   // WeffT[j, t*mid + m] = sum_i wts[t,i,j] W2[i,m]: T problems side by side, wts[t] read transposed, W2^T shared
    dpot_gemm_args g = gemm16_args(wts16, E, W2T16, E, packed + L.WeffT, d.Kp, E, d.mid, E, nullptr, DPOT_ACT_NONE);
    g.a_trans = 1; g.batch = d.T; g.strideA = 2 * (int64_t)E * E; g.strideW = 0; g.strideC = d.mid;
    DPOT_CALL(dpot_gemm(&g, stream));
  }
  {   // bias_eff[p, j] = sum_i (b2[i] + pos[i,p]) Wsum[i,j]
    dpot_gemm_args g = gemm16_args(bp16, d.n, Wsum16, E, packed + L.bias_eff, E, d.n, E, E, nullptr, DPOT_ACT_NONE);
    g.a_trans = 1; g.w_trans = 1;
    DPOT_CALL(dpot_gemm(&g, stream));
  }
  DPOT_CALL(dpot_pack_out(prm->out0_w, prm->out0_b, E, d.old, d.P, packed + L.WtT, packed + L.bias_t, stream));
  DPOT_CALL(dpot_split_f16(packed + L.WeffT, d.Kp, E, d.Kp, nullptr, nullptr, 0, packed + L.WeffT16, 2 * d.Kp, d.Kp, stream));
  DPOT_CALL(dpot_split_f16(packed + L.WtT, E, d.NP, E, nullptr, nullptr, 0, packed + L.WtT16, 2 * E, E, stream));
  if (d.old != 32) {   // generic tail: the 1x1 conv weights of out_layer[2] / [4] as split operands ([4] also zero-padded to 8 rows)
    const int nout = d.Co * d.To;
    DPOT_CALL(dpot_split_f16(prm->out2_w, d.old, d.old, d.old, nullptr, nullptr, 0, wprep + PL.o2_16, 2 * d.old, d.old, stream));
    DPOT_CUDA(cudaMemsetAsync(wprep + PL.o4p16, 0, sizeof(float) * 8 * (size_t)d.old, st));
    DPOT_CALL(dpot_split_f16(prm->out4_w, d.old, nout, d.old, nullptr, nullptr, 0, wprep + PL.o4_16, 2 * d.old, d.old, stream));
    DPOT_CALL(dpot_split_f16(prm->out4_w, d.old, nout, d.old, nullptr, nullptr, 0, wprep + PL.o4p16, 2 * d.old, d.old, stream));
  }
  for (int i = 0; i < d.depth; ++i) {
    float* base = packed + L.blocks + (int64_t)i * L.blk_stride;
    const dpot_block_params& b = prm->blocks[i];
    DPOT_CALL(tk_pack_afno16(b.w1, b.b1, d.nb, d.bs, reinterpret_cast<__half*>(base + L.Wc1_16), base + L.bc1, st));
    DPOT_CALL(tk_pack_afno16(b.w2, b.b2, d.nb, d.bs, reinterpret_cast<__half*>(base + L.Wc2_16), base + L.bc2, st));
    DPOT_CALL(dpot_split_f16(b.fc1_w, E, d.hid, E, nullptr, nullptr, 0, base + L.fc1_16, 2 * E, E, stream));
    DPOT_CALL(dpot_split_f16(b.fc2_w, d.hid, E, d.hid, nullptr, nullptr, 0, base + L.fc2_16, 2 * d.hid, d.hid, stream));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
extern "C" int dpot_train_forward(const dpot_config* cfg, const dpot_params* prm, const float* packed, const float* wprep,
                                  const float* x, int32_t B, float* y, float* cls, float* tape, float* scratch, void* stream) {
  Dims d;
  DPOT_CALL(make_dims(cfg, d));
  DPOT_REQUIRE(train_ok(cfg, d), DPOT_E_UNSUPPORTED, "dpot_train_forward: configuration not served by the fused training step");
  DPOT_REQUIRE(prm && packed && wprep && x && y && tape && scratch && prm->blocks && B > 0, DPOT_E_BADARG, "dpot_train_forward: null pointer / bad B");
  cudaStream_t st = as_stream(stream);
  const Packed PL = packed_layout(d);
  const TapeL TL = tape_layout(d, B);
  const ScratchL SL = scratch_layout(d, B);
  const int Mt = B * d.n, Ms = B * d.km1 * d.km2, act = cfg->act, R = cfg->img_size, E = d.E;
  const int64_t kb = 2 * d.bs;
  float* z1 = tape + TL.z1;
  if (d.Kp != d.T * d.mid) DPOT_CUDA(cudaMemsetAsync(z1, 0, sizeof(float) * (size_t)Mt * d.Kp, st));
  DPOT_CALL(dpot_patch_embed(x, 0, packed + PL.W0p, packed + PL.rowbias0, nullptr, nullptr, B, R, R, d.T, d.C, d.P, d.mid, act, z1,
                             d.Kp, DPOT_FMT_HL16, stream));
  auto blk = [&](int i) { return tape + TL.blocks + (int64_t)i * TL.blk_stride; };
  {
    dpot_gemm_args g = gemm16_args(z1, d.Kp, packed + PL.WeffT16, d.Kp, blk(0) + TL.lat, E, Mt, E, d.Kp, nullptr, DPOT_ACT_NONE);
    g.rowbias = packed + PL.bias_eff; g.rowbias_period = d.n; g.ldrb = E;
    g.out_stats = reinterpret_cast<double*>(blk(0) + TL.st1); g.stats_groups = GROUPS; g.stats_rows_per_sample = d.n;
    DPOT_CALL(dpot_gemm(&g, stream));
  }
  float* O2 = scratch + SL.O2;
  for (int i = 0; i < d.depth; ++i) {
    const dpot_block_params& bp = prm->blocks[i];
    const float* pk = packed + PL.blocks + (int64_t)i * PL.blk_stride;
    float* tb = blk(i);
    float* lat = tb + TL.lat;
    double* st1 = reinterpret_cast<double*>(tb + TL.st1);
    double* st2 = reinterpret_cast<double*>(tb + TL.st2);
    DPOT_CALL(dpot_afno_fft_fwd16_gn(lat, st1, bp.norm1_w, bp.norm1_b, GROUPS, GN_EPS, B, d.h, E, d.nb, d.km1, d.km2, tb + TL.S, stream));
    {
      dpot_gemm_args g = gemm16_args(tb + TL.S, 2 * E, pk + PL.Wc1_16, kb, tb + TL.O1, 0, Ms, (int)kb, (int)kb, pk + PL.bc1, act);
      g.batch = d.nb; g.strideA = kb; g.strideW = 2 * kb * kb; g.strideC = kb; g.strideBias = kb;
      out16(g, 2 * E);
      g.C_pre = tb + TL.O1pre; g.ld_pre = 2 * E; g.stride_pre = kb;
      DPOT_CALL(dpot_gemm(&g, stream));
      g = gemm16_args(tb + TL.O1, 2 * E, pk + PL.Wc2_16, kb, O2, 2 * E, Ms, (int)kb, (int)kb, pk + PL.bc2, DPOT_ACT_NONE);
      g.batch = d.nb; g.strideA = kb; g.strideW = 2 * kb * kb; g.strideC = kb; g.strideBias = kb;
      DPOT_CALL(dpot_gemm(&g, stream));
    }
    DPOT_CUDA(cudaMemsetAsync(st2, 0, sizeof(double) * 2 * GROUPS * B, st));
    DPOT_CALL(dpot_afno_fft_inv_gn(O2, lat, st1, bp.norm1_w, bp.norm1_b, GROUPS, GN_EPS, B, d.h, E, d.nb, d.km1, d.km2, tb + TL.f, st2, stream));
    DPOT_CALL(dpot_split_f16_gn(tb + TL.f, E, Mt, E, st2, bp.norm2_w, bp.norm2_b, GROUPS, GN_EPS, d.n, tb + TL.n2, 2 * E, E, stream));
    {
      dpot_gemm_args g = gemm16_args(tb + TL.n2, E, pk + PL.fc1_16, E, tb + TL.hid, 0, Mt, d.hid, E, bp.fc1_b, act);
      out16(g, d.hid);
      g.C_pre = tb + TL.hpre; g.ld_pre = d.hid;
      DPOT_CALL(dpot_gemm(&g, stream));
      const bool last = i + 1 == d.depth;
      float* dst = last ? tape + TL.alat16 : blk(i + 1) + TL.lat;
      g = gemm16_args(tb + TL.hid, d.hid, pk + PL.fc2_16, d.hid, dst, E, Mt, E, d.hid, bp.fc2_b, DPOT_ACT_NONE);
      g.residual = lat; g.ldr = E;
      if (!last) { g.out_stats = reinterpret_cast<double*>(blk(i + 1) + TL.st1); g.stats_groups = GROUPS; g.stats_rows_per_sample = d.n; }
      else out16(g, E);
      // long contraction (hid > 2048: DPOT-M / L / H): accumulation chains of <= dpot_tc16_set_chain halves, partial sums in
      // the result itself or, for the split result of the last block, in a backward-phase scratch slot (free until then)
      DPOT_CALL(dpot_gemm_chained(&g, dpot_tc16_set_chain(-1), last ? scratch + SL.dn2 : dst, E, stream));
    }
  }
  if (cls) {   // classification head; pre-activations kept for backward
    DPOT_CALL(dpot_spatial_mean16(tape + TL.alat16, B, d.n, E, tape + TL.tok, stream));
    dpot_gemm_args g = gemm_args(tape + TL.tok, E, prm->cls0_w, E, tape + TL.c1, E, B, E, E, prm->cls0_b, act, DPOT_GEMM_AUTO);
    g.C_pre = tape + TL.c1pre;
    DPOT_CALL(dpot_gemm(&g, stream));
    g = gemm_args(tape + TL.c1, E, prm->cls2_w, E, tape + TL.c2, E, B, E, E, prm->cls2_b, act, DPOT_GEMM_AUTO);
    g.C_pre = tape + TL.c2pre;
    DPOT_CALL(dpot_gemm(&g, stream));
    g = gemm_args(tape + TL.c2, E, prm->cls4_w, E, cls, d.ncls, B, d.ncls, E, prm->cls4_b, DPOT_ACT_NONE, DPOT_GEMM_AUTO);
    DPOT_CALL(dpot_gemm(&g, stream));
  }
  const bool fast_tail = dpot_out_tail_tc_supported(d.old, d.Co * d.To, d.Co) && tk_tail_bwd_supported(d.old, d.Co * d.To);
  if (fast_tail) {   // output head: ConvTranspose GEMM (per-pixel [hi 32 | lo 32] records for the tcgen05 tail, fp32 pre-activation for backward)
    float* Y1g = scratch + SL.Y1g;
    dpot_gemm_args g = gemm16_args(tape + TL.alat16, E, packed + PL.WtT16, E, Y1g, d.NP, Mt, d.NP, E, packed + PL.bias_t, act);
    g.c_fmt = DPOT_FMT_HL16G32; g.ldc = 2 * (int64_t)d.NP;
    g.C_pre = tape + TL.Y1pre; g.ld_pre = d.NP;
    DPOT_CALL(dpot_gemm(&g, stream));
    DPOT_CALL(dpot_out_tail_tc(Y1g, prm->out2_w, prm->out2_b, prm->out4_w, prm->out4_b, B, d.h, d.h, d.P, d.old, d.Co * d.To, act,
                               nullptr, nullptr, d.Co, y, nullptr, nullptr, d.T, 0, 0, 0, stream));
  } else {
    // generic tail (out_layer_dim = 128 of DPOT-L/H): the per-pixel layers are contractions batched over the P*P
    // intra-patch positions (u, v) -- pixel (tok, uv) is row tok of problem uv, column offset uv*old -- with the 1x1
    // conv weights shared by the batch; every result is kept split + its pre-activation for backward
    const PrepL WP = prep_layout(d);
    const int PP = d.P * d.P, old = d.old, nout = d.Co * d.To;
    dpot_gemm_args g = gemm16_args(tape + TL.alat16, E, packed + PL.WtT16, E, tape + TL.Y1_16, 0, Mt, d.NP, E, packed + PL.bias_t, act);
    out16(g, d.NP);
    g.C_pre = tape + TL.Y1pre; g.ld_pre = d.NP;
    DPOT_CALL(dpot_gemm(&g, stream));
    g = gemm16_args(tape + TL.Y1_16, d.NP, wprep + WP.o2_16, old, tape + TL.Y2_16, 0, Mt, old, old, prm->out2_b, act);
    g.batch = PP; g.strideA = old; g.strideW = 0; g.strideC = old; g.strideBias = 0;
    out16(g, d.NP);
    g.C_pre = tape + TL.Y2pre; g.ld_pre = d.NP; g.stride_pre = old;
    DPOT_CALL(dpot_gemm(&g, stream));
    float* Y3 = scratch + SL.Y3;                                  // rows (tok, uv) x nout
    g = gemm16_args(tape + TL.Y2_16, d.NP, wprep + WP.o4_16, old, Y3, (int64_t)PP * nout, Mt, nout, old, prm->out4_b, DPOT_ACT_NONE);
    g.batch = PP; g.strideA = old; g.strideW = 0; g.strideC = nout; g.strideBias = 0;
    DPOT_CALL(dpot_gemm(&g, stream));
    DPOT_CALL(dpot_pixel_shuffle(Y3, y, B, d.h, d.h, d.P, nout, 1, stream));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
namespace dpot {
namespace {

// classification head backward (M = B rows: CUDA-core kernels); gc = dcls * S.  Adds dtok / n into g (the gradient of the
// last latent).  Scratch `cs` as laid out by scratch_layout (cls).
int cls_backward(const dpot_config* cfg, const dpot_params* prm, const Dims& d, int B, const float* dcls, const float* tape,
                 const TapeL& TL, float* cs, const float* scale, const dpot_params* gr, float* g, void* stream) {
  cudaStream_t st = as_stream(stream);
  const int E = d.E, act = cfg->act;
  const float* inv = scale + 1;
  Bump a;
  float* gc = cs + a.take((int64_t)B * d.ncls);
  float* dc2 = cs + a.take((int64_t)B * E); float* g2 = cs + a.take((int64_t)B * E);
  float* dc1 = cs + a.take((int64_t)B * E); float* g1 = cs + a.take((int64_t)B * E);
  float* dtok = cs + a.take((int64_t)B * E);
  float* WT = cs + a.take((int64_t)E * E);
  float* W4T = cs + a.take((int64_t)d.ncls * E);
  DPOT_CALL(tk_scale_copy(dcls, (int64_t)B * d.ncls, scale, gc, st));
  auto wgrad = [&](const float* X, int ldx, const float* Y, int ldy, int N, int K, float* dW) -> int {
    dpot_wgrad_args w; memset(&w, 0, sizeof(w));
    w.X = X; w.ldx = ldx; w.Y = Y; w.ldy = ldy; w.dW = dW; w.ldw = K; w.M = B; w.N = N; w.K = K; w.batch = 1;
    DPOT_CALL(dpot_wgrad(&w, stream));
    return tk_scale_copy(dW, (int64_t)N * K, inv, dW, st);
  };
  auto bgrad = [&](const float* X, int N, float* db) -> int {
    DPOT_CALL(dpot_colsum(X, N, B, N, db, 0, stream));
    return tk_scale_copy(db, N, inv, db, st);
  };
  // layer 4: cls = c2 W4^T + b4
  DPOT_CALL(wgrad(gc, d.ncls, tape + TL.c2, E, d.ncls, E, const_cast<float*>(gr->cls4_w)));
  DPOT_CALL(bgrad(gc, d.ncls, const_cast<float*>(gr->cls4_b)));
  DPOT_CALL(tk_transpose(prm->cls4_w, E, W4T, d.ncls, d.ncls, E, st));          // [E, ncls]
  dpot_gemm_args m = gemm_args(gc, d.ncls, W4T, d.ncls, dc2, E, B, E, d.ncls, nullptr, DPOT_ACT_NONE, DPOT_GEMM_SIMT);
  DPOT_CALL(dpot_gemm(&m, stream));
  DPOT_CALL(dpot_act_bwd(dc2, tape + TL.c2pre, act, (int64_t)B * E, g2, stream));
  // layer 2
  DPOT_CALL(wgrad(g2, E, tape + TL.c1, E, E, E, const_cast<float*>(gr->cls2_w)));
  DPOT_CALL(bgrad(g2, E, const_cast<float*>(gr->cls2_b)));
  DPOT_CALL(tk_transpose(prm->cls2_w, E, WT, E, E, E, st));
  m = gemm_args(g2, E, WT, E, dc1, E, B, E, E, nullptr, DPOT_ACT_NONE, DPOT_GEMM_SIMT);
  DPOT_CALL(dpot_gemm(&m, stream));
  DPOT_CALL(dpot_act_bwd(dc1, tape + TL.c1pre, act, (int64_t)B * E, g1, stream));
  // layer 0
  DPOT_CALL(wgrad(g1, E, tape + TL.tok, E, E, E, const_cast<float*>(gr->cls0_w)));
  DPOT_CALL(bgrad(g1, E, const_cast<float*>(gr->cls0_b)));
  DPOT_CALL(tk_transpose(prm->cls0_w, E, WT, E, E, E, st));
  m = gemm_args(g1, E, WT, E, dtok, E, B, E, E, nullptr, DPOT_ACT_NONE, DPOT_GEMM_SIMT);
  DPOT_CALL(dpot_gemm(&m, stream));
  return tk_add_mean_grad(dtok, B, d.n, E, g, st);
}

}  // namespace
}  // namespace dpot

extern "C" int dpot_train_backward(const dpot_config* cfg, const dpot_params* prm, const float* packed, const float* wprep,
                                   const float* x, int32_t B, const float* dy, const float* dcls, const float* tape,
                                   float* scratch, const dpot_params* grads, float* dx, void* const* events, void* stream) {
  Dims d;
  DPOT_CALL(make_dims(cfg, d));
  DPOT_REQUIRE(train_ok(cfg, d), DPOT_E_UNSUPPORTED, "dpot_train_backward: configuration not served by the fused training step");
  DPOT_REQUIRE(prm && packed && wprep && x && dy && tape && scratch && grads && grads->blocks && prm->blocks && B > 0, DPOT_E_BADARG,
               "dpot_train_backward: null pointer / bad B");
  cudaStream_t st = as_stream(stream);
  const Packed PL = packed_layout(d);
  const PrepL WL = prep_layout(d);
  const TapeL TL = tape_layout(d, B);
  const ScratchL SL = scratch_layout(d, B);
  const int Mt = B * d.n, Ms = B * d.km1 * d.km2, act = cfg->act, R = cfg->img_size, E = d.E, H = d.hid, NP = d.NP;
  const int nout = d.Co * d.To;
  const int64_t kb = 2 * d.bs;
  auto G = [](const float* p) { return const_cast<float*>(p); };
  auto blk = [&](int i) { return tape + TL.blocks + (int64_t)i * TL.blk_stride; };
  float* scale = scratch + SL.scale;                 // [S, 1/S]
  const float* inv = scale + 1;
  double* dbl = reinterpret_cast<double*>(scratch + SL.dbl);
  DPOT_CUDA(cudaMemsetAsync(dbl, 0, sizeof(double) * (size_t)SL.d_total, st));
  DPOT_CALL(tk_grad_scale(dy, (int64_t)B * R * R * nout, reinterpret_cast<unsigned*>(scale + 8), scale, st));
  float* slabs = scratch + SL.slabs;
  int ns = 1;

  // ---- output head (models/dpot.py:315-321)
  float* g1t = scratch + SL.g1t;
  const bool fast_tail = dpot_out_tail_tc_supported(d.old, nout, d.Co) && tk_tail_bwd_supported(d.old, nout);
  if (fast_tail) {
    DPOT_CALL(tk_tail_bwd(tape + TL.Y1pre, dy, scale, prm->out2_w, prm->out2_b, prm->out4_w, B, d.h, d.h, d.P, nout, act,
                          reinterpret_cast<__half*>(g1t), scratch + SL.tparts, inv, G(grads->out2_w), G(grads->out2_b),
                          G(grads->out4_w), G(grads->out4_b), G(grads->out0_b), st));
  } else {
    // generic tail: the same chain as contractions batched over the P*P intra-patch positions (see dpot_train_forward)
    const int PP = d.P * d.P, old = d.old;
    float* g3 = scratch + SL.g3_16; float* g2 = scratch + SL.g2_16;
    double* cs_g2 = dbl + SL.d_gt; double* cs_g1 = cs_g2 + NP; double* cs_g3 = cs_g1 + NP;
    DPOT_CALL(tk_unshuffle_pad_split(dy, scale, B, d.h, d.h, d.P, nout, reinterpret_cast<__half*>(g3), st));   // [tok, (uv, 8)] split, x S
    // out_layer[4]: y3 = y2 W4^T + b4
    DPOT_CALL(tk_colsum(g3, true, 2 * (int64_t)PP * 8, (int64_t)PP * 8, Mt, PP * 8, cs_g3, st));
    DPOT_CALL(tk_sum_batches(cs_g3, PP, 8, nout, inv, G(grads->out4_b), st));
    DPOT_CALL(wgrad16(g3, (int64_t)PP * 8, tape + TL.Y2_16, NP, 8, old, Mt, PP, slabs, &ns, stream));            // [ks][uv][8, old]
    DPOT_CALL(tk_slab_reduce(slabs, ns * PP, (int64_t)8 * old, (int64_t)nout * old, inv, G(grads->out4_w), st));
    {   // g2 = (g3 W4) * act'(Y2pre), split; column sums per (u, v) -> db2
      dpot_gemm_args a = g16(g3, (int64_t)PP * 8, wprep + WL.o4p16, old, g2, 0, Mt, old, 8);
      a.w_trans = 1; a.batch = PP; a.strideA = 8; a.strideW = 0; a.strideC = old;
      out16(a, NP);
      a.dact_src = tape + TL.Y2pre; a.ld_dact = NP; a.stride_dact = old; a.dact = act;
      a.out_colsum = cs_g2;
      DPOT_CALL(dpot_gemm(&a, stream));
    }
    DPOT_CALL(tk_sum_batches(cs_g2, PP, old, old, inv, G(grads->out2_b), st));
    DPOT_CALL(wgrad16(g2, NP, tape + TL.Y1_16, NP, old, old, Mt, PP, slabs, &ns, stream));                         // [ks][uv][old, old]
    DPOT_CALL(tk_slab_reduce(slabs, ns * PP, (int64_t)old * old, (int64_t)old * old, inv, G(grads->out2_w), st));
    {   // g1 = (g2 W2) * act'(Y1pre), split; column sums -> the ConvTranspose bias gradient
      dpot_gemm_args a = g16(g2, NP, wprep + WL.o2_16, old, g1t, 0, Mt, old, old);
      a.w_trans = 1; a.batch = PP; a.strideA = old; a.strideW = 0; a.strideC = old;
      out16(a, NP);
      a.dact_src = tape + TL.Y1pre; a.ld_dact = NP; a.stride_dact = old; a.dact = act;
      a.out_colsum = cs_g1;
      DPOT_CALL(dpot_gemm(&a, stream));
    }
    DPOT_CALL(tk_sum_batches(cs_g1, PP, old, old, inv, G(grads->out0_b), st));
  }
  {
    DPOT_CALL(wgrad16(g1t, NP, tape + TL.alat16, E, NP, E, Mt, 1, slabs, &ns, stream));
    DPOT_CALL(tk_unpack_out_grad(slabs, ns, (int64_t)NP * E, nullptr, E, d.old, d.P, inv, G(grads->out0_w), G(grads->out0_b), st));
  }
  if (events) DPOT_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(events[0]), st));   // out_layer gradients are final
  float* g = scratch + SL.gA;          // dL/d(latent after the last block), fp32, scaled by S
  float* g_other = scratch + SL.gB;
  float* gs16 = scratch + SL.g16;      // the same, split
  {
    // dL/d(latent) = g1 WtT: the contraction runs over the P*P*out_layer_dim head columns -- 2048 for the 32-wide head,
    // 8192 / 32768 for DPOT-H / L.  A tensor-core accumulation chain is kept <= 1024 deep (measured 1e-5 /
    // 3.9e-5 on these two before the split), so long ones are cut into chunks summed in fp32
    dpot_gemm_args a = g16(g1t, NP, packed + PL.WtT16, E, g, E, Mt, E, NP);
    a.w_trans = 1;
    const bool split_k = NP > 2048;
    if (split_k) {
      a.C = slabs;
      ksplit(a, 1024, (int64_t)Mt * E);
    } else if (!dcls) {
      a.out_colsum = dbl + SL.d_blk + (int64_t)(d.depth - 1) * SL.d_blk_stride;   // = db2 of the last block
    }
    DPOT_CALL(dpot_gemm(&a, stream));
    if (split_k) DPOT_CALL(tk_slab_reduce(slabs, a.k_split, (int64_t)Mt * E, (int64_t)Mt * E, nullptr, g, st));
    if (dcls) DPOT_CALL(cls_backward(cfg, prm, d, B, dcls, tape, TL, scratch + SL.cls, scale, grads, g, stream));
    if (dcls || split_k)
      DPOT_CALL(tk_colsum(g, false, E, 0, Mt, E, dbl + SL.d_blk + (int64_t)(d.depth - 1) * SL.d_blk_stride, st));
  }
  DPOT_CALL(tk_split_scaled(g, Mt, E, nullptr, reinterpret_cast<__half*>(gs16), 2 * (int64_t)E, E, st));

  // ---- blocks, last to first (models/dpot.py:165-180)
  for (int i = d.depth - 1; i >= 0; --i) {
    const dpot_block_params& bp = prm->blocks[i];
    const dpot_block_params& gb = grads->blocks[i];
    const float* pk = packed + PL.blocks + (int64_t)i * PL.blk_stride;
    const float* tb = blk(i);
    double* dblk = dbl + SL.d_blk + (int64_t)i * SL.d_blk_stride;
    double* db2 = dblk; double* db1 = db2 + E; double* dbc2 = db1 + H; double* dbc1 = dbc2 + 2 * E;
    float* g1h = scratch + SL.g1_16;
    // channel MLP: lat_next = fc2(act(fc1(n2))) + lat
    DPOT_CALL(tk_finish_double(db2, E, inv, G(gb.fc2_b), st));      // accumulated by the producer of g (epilogue / GroupNorm-1 backward)
    DPOT_CALL(wgrad16(gs16, E, tb + TL.hid, H, E, H, Mt, 1, slabs, &ns, stream));
    DPOT_CALL(tk_slab_reduce(slabs, ns, (int64_t)E * H, (int64_t)E * H, inv, G(gb.fc2_w), st));
    {   // g1 = (g W2) * act'(hpre), split
      dpot_gemm_args a = g16(gs16, E, pk + PL.fc2_16, H, g1h, 0, Mt, H, E);
      a.w_trans = 1; out16(a, H);
      a.dact_src = tb + TL.hpre; a.ld_dact = H; a.dact = act;
      a.out_colsum = db1;
      DPOT_CALL(dpot_gemm(&a, stream));
    }
    DPOT_CALL(tk_finish_double(db1, H, inv, G(gb.fc1_b), st));
    DPOT_CALL(wgrad16(g1h, H, tb + TL.n2, E, H, E, Mt, 1, slabs, &ns, stream));
    DPOT_CALL(tk_slab_reduce(slabs, ns, (int64_t)E * H, (int64_t)E * H, inv, G(gb.fc1_w), st));
    float* dn2 = scratch + SL.dn2;
    {
      dpot_gemm_args a = g16(g1h, H, pk + PL.fc1_16, E, dn2, E, Mt, E, H);
      a.w_trans = 1;
      if (H > 2048) {        // mlp_ratio 4 (DPOT-M/L/H): keep the accumulation chain <= 1024 deep (see the head's data gradient)
        a.C = slabs;
        ksplit(a, 1024, (int64_t)Mt * E);
      }
      DPOT_CALL(dpot_gemm(&a, stream));
      if (a.k_split > 1) DPOT_CALL(tk_slab_reduce(slabs, a.k_split, (int64_t)Mt * E, (int64_t)Mt * E, nullptr, dn2, st));
    }
    // GroupNorm-2
    float* df = scratch + SL.df;
    DPOT_CALL(tk_gn_bwd(dn2, tb + TL.f, reinterpret_cast<const double*>(tb + TL.st2), bp.norm2_w, nullptr, B, d.n, E, GROUPS, GN_EPS,
                        inv, scratch + SL.gn, df, nullptr, G(gb.norm2_w), G(gb.norm2_b), nullptr, st));
    // AFNO mixer: f = irfft2(O2) + n1, O2 = O1 Wc2 + bc2, O1 = act(S Wc1 + bc1), S = rfft2(n1)
    float* dO2 = scratch + SL.dO2; float* dO1 = scratch + SL.dO1; float* dS = scratch + SL.dS;
    DPOT_CALL(dpot_afno_fft_fwd16w(df, B, d.h, E, d.nb, d.km1, d.km2, dO2, 2.0f, dbc2, stream));   // adjoint of the inverse transform (+ bias gradient)
    DPOT_CALL(wgrad16(dO2, 2 * E, tb + TL.O1, 2 * E, (int)kb, (int)kb, Ms, d.nb, slabs, &ns, stream));
    DPOT_CALL(tk_unpack_afno_grad(slabs, ns, (int64_t)d.nb * kb * kb, dbc2, d.nb, d.bs, inv, G(gb.w2), G(gb.b2), st));
    {
      dpot_gemm_args a = g16(dO2, 2 * E, pk + PL.Wc2_16, kb, dO1, 0, Ms, (int)kb, (int)kb);
      a.w_trans = 1; a.batch = d.nb; a.strideA = kb; a.strideW = 2 * kb * kb; a.strideC = kb;
      out16(a, 2 * E);
      a.dact_src = tb + TL.O1pre; a.ld_dact = 2 * E; a.stride_dact = kb; a.dact = act;
      a.out_colsum = dbc1;
      DPOT_CALL(dpot_gemm(&a, stream));
    }
    DPOT_CALL(wgrad16(dO1, 2 * E, tb + TL.S, 2 * E, (int)kb, (int)kb, Ms, d.nb, slabs, &ns, stream));
    DPOT_CALL(tk_unpack_afno_grad(slabs, ns, (int64_t)d.nb * kb * kb, dbc1, d.nb, d.bs, inv, G(gb.w1), G(gb.b1), st));
    {
      dpot_gemm_args a = g16(dO1, 2 * E, pk + PL.Wc1_16, kb, dS, 2 * E, Ms, (int)kb, (int)kb);
      a.w_trans = 1; a.batch = d.nb; a.strideA = kb; a.strideW = 2 * kb * kb; a.strideC = kb;
      DPOT_CALL(dpot_gemm(&a, stream));
    }
    float* dn1 = scratch + SL.dn1;
    DPOT_CALL(dpot_afno_fft_inv(dS, df, nullptr, nullptr, B, d.h, E, d.nb, d.km1, d.km2, dn1, nullptr, GROUPS, 0.5f, stream));   // adjoint of the forward transform + skip
    // GroupNorm-1 + the residual path
    DPOT_CALL(tk_gn_bwd(dn1, tb + TL.lat, reinterpret_cast<const double*>(tb + TL.st1), bp.norm1_w, g, B, d.n, E, GROUPS, GN_EPS, inv,
                        scratch + SL.gn, g_other, reinterpret_cast<__half*>(gs16), G(gb.norm1_w), G(gb.norm1_b),
                        i > 0 ? dbl + SL.d_blk + (int64_t)(i - 1) * SL.d_blk_stride : nullptr, st));   // db2 of the block below
    float* t = g; g = g_other; g_other = t;
    if (events) DPOT_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(events[1 + (d.depth - 1 - i)]), st));   // block i's gradients are final
  }

  // ---- front: lat0 = z1 WeffT^T + bias_eff  (conv 1x1 + pos_embed + time aggregation folded), z1 = act(conv0(x))
  float* dWeffT = scratch + SL.dWeffT; float* dbe = scratch + SL.dbe;
  DPOT_CALL(tk_colsum(g, false, (int64_t)d.n * E, 0, B, d.n * E, dbl + SL.d_be, st));
  DPOT_CALL(tk_finish_double(dbl + SL.d_be, (int64_t)d.n * E, nullptr, dbe, st));          // stays scaled by S
  DPOT_CALL(wgrad16(gs16, E, tape + TL.z1, d.Kp, E, d.Kp, Mt, 1, slabs, &ns, stream));
  DPOT_CALL(tk_slab_reduce(slabs, ns, (int64_t)E * d.Kp, (int64_t)E * d.Kp, nullptr, dWeffT, st));
  float* z1pre = scratch + SL.z1pre; float* gz = scratch + SL.gz;
  if (d.Kp != d.T * d.mid) DPOT_CUDA(cudaMemsetAsync(z1pre, 0, sizeof(float) * (size_t)Mt * d.Kp, st));
  DPOT_CALL(dpot_patch_embed(x, 0, packed + PL.W0p, packed + PL.rowbias0, nullptr, nullptr, B, R, R, d.T, d.C, d.P, d.mid, DPOT_ACT_NONE,
                             z1pre, d.Kp, DPOT_FMT_F32, stream));
  {
    dpot_gemm_args a = g16(gs16, E, packed + PL.WeffT16, d.Kp, gz, d.Kp, Mt, d.Kp, E);
    a.w_trans = 1;
    a.dact_src = z1pre; a.ld_dact = d.Kp; a.dact = act;
    DPOT_CALL(dpot_gemm(&a, stream));
  }
  DPOT_CALL(tk_colsum(gz, false, (int64_t)d.n * d.Kp, 0, B, d.n * d.Kp, dbl + SL.d_rb, st));
  DPOT_CALL(tk_patch_bwd(gz, x, packed + PL.W0p, B, R, R, d.T, d.C, d.P, d.mid, d.Kp, inv, scratch + SL.pslabs, dx, st));
  DPOT_CALL(tk_unpack_patch_grad(scratch + SL.pslabs, tk_patch_bwd_slabs(B, R, R, d.P), dbl + SL.d_rb, prm->grid_x, prm->grid_y, prm->grid_t, d.mid, d.C, d.P, d.h, d.h, d.T,
                                 d.Kp, inv, G(grads->pe0_w), G(grads->pe0_b), st));
  // fold backward: every contraction on the f16-split engine (WeffT = sum_i wts W2, bias_eff = bp^T Wsum, Wsum = sum_t wts)
  {
    const float* wts16 = wprep + WL.wts16; const float* Wsum16 = wprep + WL.Wsum16; const float* bp16 = wprep + WL.bp16;
    const float* W2p16 = wprep + WL.W2p16;
    float* Gp16 = scratch + SL.Gp16; float* dbe16 = scratch + SL.dbe16; float* dWsum = scratch + SL.dWsum;
    float* dwt = scratch + SL.dwt; float* dW2s = scratch + SL.dW2s; float* dbp = scratch + SL.dbp; float* dtemb = scratch + SL.dtemb;
    const int midp = midp_of(d);
    DPOT_CALL(tk_tagg_pad_g(dWeffT, E, d.Kp, d.T, d.mid, midp, reinterpret_cast<__half*>(Gp16), st));   // Gp[t][j][m]
    DPOT_CALL(dpot_split_f16(dbe, E, d.n, E, nullptr, nullptr, 0, dbe16, 2 * E, E, stream));
    {   // dWsum[i,j] = sum_p bp[i,p] dbe[p,j]
      dpot_gemm_args m = g16(bp16, d.n, dbe16, E, dWsum, E, E, E, d.n);
      m.w_trans = 1;
      DPOT_CALL(dpot_gemm(&m, stream));
    }
    {   // dwt[t][i,j] = sum_m W2[i,m] Gp[t][j][m] + dWsum[i,j]     (W2 shared by the T problems)
      dpot_gemm_args m = g16(W2p16, midp, Gp16, midp, dwt, E, E, E, midp);
      m.batch = d.T; m.strideA = 0; m.strideW = 2 * (int64_t)E * midp; m.strideC = (int64_t)E * E;
      m.residual = dWsum; m.ldr = E;
      DPOT_CALL(dpot_gemm(&m, stream));
    }
    {   // dW2 partials [t][i, m] = sum_j wts[t][i,j] Gp[t][j][m]
      dpot_gemm_args m = g16(wts16, E, Gp16, midp, dW2s, midp, E, midp, E);
      m.w_trans = 1; m.batch = d.T; m.strideA = 2 * (int64_t)E * E; m.strideW = 2 * (int64_t)E * midp; m.strideC = (int64_t)E * midp;
      DPOT_CALL(dpot_gemm(&m, stream));
    }
    DPOT_CALL(tk_tagg_dw2_finish(dW2s, d.T, E, d.mid, midp, inv, G(grads->pe2_w), st));
    {   // dbp[i,p] = sum_j Wsum[i,j] dbe[p,j]
      dpot_gemm_args m = g16(Wsum16, E, dbe16, E, dbp, d.n, E, d.n, E);
      DPOT_CALL(dpot_gemm(&m, stream));
    }
    DPOT_CALL(tk_rowsum_scale(dbp, E, d.n, inv, G(grads->pe2_b), G(grads->pos_embed), st));
    DPOT_CALL(tk_tagg_finish(dwt, prm->tagg_w, prm->temb, d.T, E, inv, G(grads->tagg_w), dtemb, st));
    if (cfg->time_agg == 1 && grads->tagg_gamma)
      DPOT_CALL(tk_tagg_gamma_grad(dtemb, prm->tagg_gamma, d.T, E, inv, G(grads->tagg_gamma), st));
  }
  if (events) DPOT_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(events[1 + d.depth]), st));
  return 0;
}
