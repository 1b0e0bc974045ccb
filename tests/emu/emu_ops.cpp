// tests/emu/emu_ops.cpp — TEST INFRASTRUCTURE ONLY: the encoder-side op KERNEL SOURCES
// (opus_b200/csrc/celt_ops.cuh) on the host SIMT emulator.
#define OB_HOST_EMU 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string.h>
#include "../../opus_b200/csrc/celt_types.h"
#include "../../opus_b200/csrc/celt_ops.cuh"
#include "../../opus_b200/csrc/celt_mdct_fast.cuh"

using namespace ob;

static void mdct_entry(void *a) { const MdctFwdParams &p = *(const MdctFwdParams *)a; mdct_forward_warp(p, block_id(), dyn_smem()); }
static void xcorr_entry(void *a) { const XcorrParams &p = *(const XcorrParams *)a; pitch_xcorr_warp(p, block_id(), dyn_smem()); }
static void rot_entry(void *a) { const VqParams &p = *(const VqParams *)a; exp_rotation_warp(p, block_id(), dyn_smem()); }
static void search_entry(void *a) { const VqParams &p = *(const VqParams *)a; pvq_search_warp(p, block_id(), dyn_smem()); }

// the register-FFT long-block MDCT (celt_mdct_fast.cuh): one warp, two transforms; shared memory = tables + 2 inputs + 2 point buffers
struct FastMdctArgs { const float *in; float *out; int n; };
static void mdct_fast_entry(void *a) {
  const FastMdctArgs &p = *(const FastMdctArgs *)a;
  unsigned char *sm = dyn_smem();
  Mdct480Tabs *t = (Mdct480Tabs *)sm;
  float *in0 = (float *)(sm + sizeof(Mdct480Tabs)), *in1 = in0 + 1080;
  cpx *F0 = (cpx *)(in1 + 1080), *F1 = F0 + kF480Cpx;
  const int lane = lane_id(), pair = block_id();
  for (int i = lane; i < 960; i += 32) { t->trig[i] = ob_mdct_trig[i]; t->tw[i] = ob_fft_tw[i]; }
  for (int i = lane; i < 120; i += 32) t->window[i] = ob_window[i];
  for (int i = lane; i < 480; i += 32) t->bitrev[i] = ob_fft_bitrev480[i];
  const bool two = 2 * pair + 1 < p.n;
  for (int i = lane; i < 1080; i += 32) { in0[i] = p.in[(size_t)(2 * pair) * 1080 + i]; if (two) in1[i] = p.in[(size_t)(2 * pair + 1) * 1080 + i]; }
  wsync();
  mdct480_forward_pair(in0, two ? in1 : nullptr, F0, F1, t, p.out + (size_t)(2 * pair) * 960, p.out + (size_t)(2 * pair + 1) * 960, lane);
}
extern "C" {
void emu_mdct_forward_fast(const float *in, float *out, int n) {
  FastMdctArgs p = {in, out, n};
  emu_launch(mdct_fast_entry, &p, (n + 1) / 2, 32, sizeof(Mdct480Tabs) + 2 * 1080 * sizeof(float) + 2 * kF480Cpx * sizeof(cpx) + 64);
}
void emu_mdct_forward(const float *in, float *out, int n, int LM, int transient) {
  MdctFwdParams p;
  p.in = in; p.out = out; p.n_items = n; p.in_stride = (120 << LM) + 120; p.out_stride = 120 << LM; p.LM = LM; p.transient = transient;
  emu_launch(mdct_entry, &p, n, 32, sizeof(MdctWS));
}
void emu_pitch_xcorr(const float *x, const float *y, float *xc, int n, int len, int max_pitch) {
  XcorrParams p;
  p.x = x; p.y = y; p.xcorr = xc; p.n_items = n; p.len = len; p.max_pitch = max_pitch;
  emu_launch(xcorr_entry, &p, n, 32, (((len + 3) & ~3) + len + max_pitch + 4) * sizeof(float));
}
void emu_exp_rotation(float *X, const int *K, int n, int N, int dir, int stride, int spread) {
  VqParams p;
  p.X = X; p.iy = nullptr; p.yy = nullptr; p.K = K; p.n_items = n; p.N = N; p.dir = dir; p.stride = stride; p.spread = spread;
  emu_launch(rot_entry, &p, (n + 31) / 32, 32, (size_t)N * 32 * sizeof(float));
}
void emu_pvq_search(float *X, int *iy, float *yy, const int *K, int n, int N) {
  VqParams p;
  p.X = X; p.iy = iy; p.yy = yy; p.K = K; p.n_items = n; p.N = N; p.dir = 1; p.stride = 1; p.spread = 0;
  emu_launch(search_entry, &p, (n + 31) / 32, 32, (size_t)N * 32 * (2 * sizeof(float) + 1));
}
}

// ---- SILK integer kernels ----
#include "../../opus_b200/csrc/silk_ops.cuh"
static void lpc_entry(void *a) {
  const SilkLpcParams &p = *(const SilkLpcParams *)a;
  int item = block_id() * 32 + lane_id();
  if (item < p.n_items) silk_lpc_analysis_filter_thread(p, item);
}
static void rs_entry(void *a) {
  const SilkResamplerParams &p = *(const SilkResamplerParams *)a;
  int s = block_id() * 32 + lane_id();
  if (s < p.n_streams) silk_resampler_thread(p, s);
}
extern "C" {
void emu_silk_lpc(int16_t *out, const int16_t *in, const int16_t *B, int n, int len, int d) {
  SilkLpcParams p;
  p.out = out; p.in = in; p.B = B; p.n_items = n; p.len = len; p.d = d;
  emu_launch(lpc_entry, &p, (n + 31) / 32, 32, 0);
}
int emu_silk_resampler_state_size(void) { return (int)sizeof(ObSilkResamplerState); }
// any rate pair silk_resampler_init accepts (silk_resampler_setup is the product's own host-side setup)
int emu_silk_resampler(void *state, const int16_t *in, int16_t *out, int n, int in_len, int fs_in, int fs_out, int for_enc) {
  SilkResamplerParams p;
  if (silk_resampler_setup(p, fs_in, fs_out, for_enc)) return -1;
  p.state = (ObSilkResamplerState *)state; p.in = in; p.out = out; p.n_streams = n; p.in_len = in_len; p.out_len = in_len / (fs_in / 1000) * (fs_out / 1000);
  emu_launch(rs_entry, &p, (n + 31) / 32, 32, 0);
  return 0;
}
void emu_silk_resampler_48_16(void *state, const int16_t *in, int16_t *out, int n, int in_len) {
  emu_silk_resampler(state, in, out, n, in_len, 48000, 16000, 1);
}
}

static void burg_entry(void *a) {
  const SilkBurgParams &p = *(const SilkBurgParams *)a;
  int item = block_id() * 32 + lane_id();
  if (item < p.n_items) silk_burg_thread(p, item);
}
extern "C" void emu_silk_burg(float *A, float *nrg, const float *x, float mig, int n, int subfr, int nb, int D) {
  SilkBurgParams p;
  p.A = A; p.nrg = nrg; p.x = x; p.minInvGain = mig; p.n_items = n; p.subfr_length = subfr; p.nb_subfr = nb; p.D = D;
  emu_launch(burg_entry, &p, (n + 31) / 32, 32, 0);
}

#include "../../opus_b200/csrc/silk_nsq.cuh"
static void nsq_dd_entry(void *a) {
  const SilkNsqParams &p = *(const SilkNsqParams *)a;
  if (p.cfg.shapingLPCOrder == 24) nsq_del_dec_warp<24>(p, block_id() * 8, (int32_t *)dyn_smem());
  else if (p.cfg.shapingLPCOrder == 16) nsq_del_dec_warp<16>(p, block_id() * 8, (int32_t *)dyn_smem());
  else if (p.cfg.shapingLPCOrder == 20) nsq_del_dec_warp<20>(p, block_id() * 8, (int32_t *)dyn_smem());
  else if (p.cfg.shapingLPCOrder == 12) nsq_del_dec_warp<12>(p, block_id() * 8, (int32_t *)dyn_smem());
  else nsq_del_dec_warp<0>(p, block_id() * 8, (int32_t *)dyn_smem());
}
static void nsq_plain_entry(void *a) {
  const SilkNsqParams &p = *(const SilkNsqParams *)a;
  int s = block_id() * 32 + lane_id();
  if (s < p.n_streams) nsq_plain_thread(p, s);
}
extern "C" int emu_nsq_state_size(void) { return (int)sizeof(ObNsqState); }
extern "C" void emu_nsq_state_init(ObNsqState *S, int n) {
  memset(S, 0, sizeof(ObNsqState) * (size_t)n);
  for (int i = 0; i < n; i++) { S[i].prev_gain_Q16 = 65536; S[i].lagPrev = 100; }
}
extern "C" void emu_silk_nsq(const int32_t *cfg, ObNsqState *state, ObNsqFrame *frames, const int16_t *x16, int8_t *pulses, int n) {
  SilkNsqParams p;
  p.state = state; p.frames = frames; p.x16 = x16; p.pulses = pulses; p.n_streams = n;
  memcpy(&p.cfg, cfg, sizeof(p.cfg));
  if (p.cfg.nStates > 1 || p.cfg.warping_Q16 > 0) emu_launch(nsq_dd_entry, &p, (n + 7) / 8, 32, kNsqWarpSmemWords * 4);
  else emu_launch(nsq_plain_entry, &p, (n + 31) / 32, 32, 0);
}
extern "C" void emu_nsq_state_export(const ObNsqState *S, int32_t *o) {
  for (int i = 0; i < 640; i++) *o++ = S->xq[i];
  for (int i = 0; i < 640; i++) *o++ = S->sLTP_shp_Q14[i];
  for (int i = 0; i < 16; i++) *o++ = S->sLPC_Q14[i];
  for (int i = 0; i < 24; i++) *o++ = S->sAR2_Q14[i];
  *o++ = S->sLF_AR_shp_Q14; *o++ = S->sDiff_shp_Q14; *o++ = S->lagPrev; *o++ = S->rand_seed; *o++ = S->prev_gain_Q16;
}

// ---- dnn primitives -------------------------------------------------------------------------------------------------
#include "../../opus_b200/csrc/nnet.cuh"
struct EmuNnetArgs { NnetLinearParams p; int nsb; };
static void nnet_quant_entry(void *a) { nnet_quant_thread(*(const NnetQuantParams *)a, (long)block_id() * 32 + lane_id()); }
static void nnet_mma_entry(void *a) {
  const EmuNnetArgs &g = *(const EmuNnetArgs *)a;
  const int sb = block_id() % g.nsb, ob_ = block_id() / g.nsb;
  if (g.p.L.su) nnet_linear_i8_mma_warp<true>(g.p, sb * 32, ob_ * 64);
  else nnet_linear_i8_mma_warp<false>(g.p, sb * 32, ob_ * 64);
}
static void nnet_i8s_entry(void *a) { nnet_linear_i8_sparse_thread(((const EmuNnetArgs *)a)->p, (long)block_id() * 32 + lane_id()); }
static void nnet_f32s_entry(void *a) { nnet_linear_f32_sparse_thread(((const EmuNnetArgs *)a)->p, (long)block_id() * 32 + lane_id()); }
static void nnet_f32_entry(void *a) {
  const EmuNnetArgs &g = *(const EmuNnetArgs *)a;
  const int sb = block_id() % g.nsb, ob_ = block_id() / g.nsb;
  nnet_linear_f32_block(g.p, sb * kNnF32Tile, ob_ * kNnF32Tile, *(NnetF32Smem *)dyn_smem(), warp_in_block() * 32 + lane_id());
}
static void nnet_gru_entry(void *a) { nnet_gru_thread(*(const NnetGruParams *)a, (long)block_id() * 32 + lane_id()); }
static void nnet_act_entry(void *a) { nnet_activation_thread(*(const NnetActParams *)a, (long)block_id() * 32 + lane_id()); }

// host-side LinearLayer as the tests pass it (same fields as dnn/nnet.h:66-76)
struct EmuLinearLayer { const float *bias, *subias; const int8_t *weights; const float *float_weights; const int *weights_idx; const float *diag, *scale; int nb_inputs, nb_outputs; };
extern "C" void emu_nnet_linear(const EmuLinearLayer *h, int su, float *out, const float *in, int n) {
  const int M = h->nb_inputs, N = h->nb_outputs;
  std::vector<int> rb_off, pos;
  EmuNnetArgs g;
  NnetLinear &L = g.p.L;
  memset(&L, 0, sizeof(L));
  L.M = M; L.N = N;
  const bool i8 = !h->float_weights && h->weights;
  L.su = (i8 && su) ? 1 : 0;
  L.bias = (i8 && su) ? h->subias : h->bias;
  L.diag = h->diag; L.scale = h->scale; L.w8 = i8 ? h->weights : nullptr; L.wf = h->float_weights;
  if (h->weights_idx) {
    const int *idx = h->weights_idx;
    int nb = 0;
    for (int rb = 0; rb < N / 8; rb++) { rb_off.push_back(nb); int cnt = *idx++; for (int j = 0; j < cnt; j++) pos.push_back(*idx++); nb += cnt; }
    rb_off.push_back(nb);
    pos.push_back(0);
    L.rb_off = rb_off.data(); L.pos = pos.data();
  }
  g.p.out = out; g.p.in = in; g.p.n = n; g.p.Mp = (M + 31) & ~31; g.p.xq = nullptr;
  std::vector<int8_t> xq;
  if (i8) {
    xq.resize((size_t)n * g.p.Mp);
    NnetQuantParams q;
    q.xq = xq.data(); q.in = in; q.n = n; q.M = M; q.Mp = g.p.Mp; q.su = L.su; q.in_stride = M;
    emu_launch(nnet_quant_entry, &q, (n * (g.p.Mp / 4) + 31) / 32, 32, 0);
    g.p.xq = xq.data();
    if (L.rb_off) emu_launch(nnet_i8s_entry, &g, (n * (N / 8) + 31) / 32, 32, 0);
    else { g.nsb = (n + 31) / 32; emu_launch(nnet_mma_entry, &g, g.nsb * ((N + 63) / 64), 32, 0); }
  } else if (L.wf) {
    if (L.rb_off) emu_launch(nnet_f32s_entry, &g, (n * (N / 8) + 31) / 32, 32, 0);
    else { g.nsb = (n + kNnF32Tile - 1) / kNnF32Tile; emu_launch(nnet_f32_entry, &g, g.nsb * ((N + kNnF32Tile - 1) / kNnF32Tile), 256, sizeof(NnetF32Smem)); }
  }
}
extern "C" void emu_nnet_gru_gates(float *state, const float *zrh, const float *recur, int n, int N) {
  NnetGruParams p;
  p.state = state; p.zrh = zrh; p.recur = recur; p.n = n; p.N = N;
  emu_launch(nnet_gru_entry, &p, (n * N + 31) / 32, 32, 0);
}
extern "C" void emu_nnet_activation(float *out, const float *in, int n, int N, int act) {
  NnetActParams p;
  p.out = out; p.in = in; p.n = n; p.N = N; p.act = act;
  emu_launch(nnet_act_entry, &p, ((act == NN_SOFTMAX ? n : n * N) + 31) / 32, 32, 0);
}

// ---- encoder front end ----
static ObEncFrameInfo *g_enc_info = nullptr;
extern "C" void emu_set_enc_info(void *p) { g_enc_info = (ObEncFrameInfo *)p; }
static void enc_analysis_entry(void *a) {
  const EncAnalysisParams &p = *(const EncAnalysisParams *)a;
  enc_analysis_block(p, block_id(), dyn_smem());
}
extern "C" int emu_enc_analysis_state_size(void) { return (int)sizeof(ObEncAnalysisState); }
extern "C" void emu_celt_enc_analysis(void *state, const float *pcm, const int32_t *sb, float *X, float *E, float *Lg, int n, int C, int LM) {
  EncAnalysisParams p;
  p.state = (ObEncAnalysisState *)state; p.pcm = pcm; p.short_blocks = sb; p.X = X; p.bandE = E; p.bandLogE = Lg; p.n_streams = n; p.C = C; p.LM = LM; p.info = g_enc_info; p.complexity = 0; p.lsb_depth = 24;
  emu_launch(enc_analysis_entry, &p, n, 32 * C, C * sizeof(MdctWS) + 32);
}

// ---- frame encoder (analysis block + thread-per-stream coder) ----
#include "../../opus_b200/csrc/celt_enc_frame.cuh"
#include "../../opus_b200/csrc/celt_enc_chains.cuh"
static std::vector<float> g_chain_scratch;
static int g_enc_chains = 1;   // run the recurrence kernel ahead of the analysis block (the product's default); 0 = the fused single-kernel form
extern "C" void emu_celt_set_chains(int on) { g_enc_chains = on; }
static void enc_chains_entry(void *a) {
  const EncAnalysisParams &p = *(const EncAnalysisParams *)a;
  enc_chains_warp(p, block_id() * 32, (float *)dyn_smem(), g_chain_scratch.data() + (size_t)block_id() * OB_CH_SCRATCH_FLOATS);
}
static std::vector<float> g_chain_in;
static void emu_run_chains(EncAnalysisParams &a, int n) {
  a.chains_done = g_enc_chains;
  a.in_buf = nullptr;
  if (!g_enc_chains) return;
  if (g_enc_chains >= 2) {   // 2: run_prefilter too on the thread-per-stream kernel (OPUSB200_ENC_SERIAL_PF)
    g_chain_in.assign((size_t)n * 2 * OB_CH_IN_STRIDE, 0.f);
    a.in_buf = g_chain_in.data();
  }
  g_chain_scratch.assign((size_t)((n + 31) / 32) * OB_CH_SCRATCH_FLOATS, 0.f);
  emu_launch(enc_chains_entry, &a, (n + 31) / 32, 32, 32 * OB_CH_TS * sizeof(float));
}
static void enc_frame_entry(void *a) {
  const EncFrameParams &p = *(const EncFrameParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s >= p.n_streams) return;
  if (p.complexity >= 8 && p.C == 2) enc_frame_thread<true>(p, s);
  else enc_frame_thread<false>(p, s);
}
// with the Opus layer in front a stereo encoder at complexity >= 8 runs the RDO instantiation for every stream (mono-coded ones take its plain branch)
static void enc_frame_entry_ctl(void *a) {
  const EncFrameParams &p = *(const EncFrameParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s >= p.n_streams) return;
  if (p.complexity >= 8 && p.CC == 2) enc_frame_thread<true>(p, s);
  else enc_frame_thread<false>(p, s);
}
// the barrier-per-band form of the frame coder (OPUSB200_ENC_SYNC): blocks of two warps; the emulator's barrier counts every thread of the
// block, so a thread without a stream keeps the barrier count instead of exiting
static int g_frame_sync = 0;
extern "C" void emu_celt_set_frame_sync(int on) { g_frame_sync = on; }
static void enc_frame_entry_sync(void *a) {
  const EncFrameParams &p = *(const EncFrameParams *)a;
  const int s = block_id() * 64 + warp_in_block() * 32 + lane_id();
  if (s >= p.n_streams) {
    for (int i = 0; i < OB_NBANDS; i++) bsync();
    return;
  }
  enc_frame_thread<true, true>(p, s);
}
extern "C" int emu_enc_state_size(void) { return (int)sizeof(ObEncState); }
extern "C" void emu_enc_state_init(ObEncState *st, int n) {
  memset(st, 0, sizeof(ObEncState) * (size_t)n);
  for (int s = 0; s < n; s++) {
    for (int i = 0; i < 2 * OB_NBANDS; i++) st[s].oldLogE[i] = st[s].oldLogE2[i] = -28.f;
    st[s].delayedIntra = 1.f; st[s].spread_decision = 2; st[s].tonal_average = 256;
  }
}
static int g_enc_end = OB_NBANDS;   // CELT_SET_END_BAND for the next emu_celt_encode calls
extern "C" void emu_celt_set_end_band(int end) { g_enc_end = end; }
static int g_enc_up = 1;   // 48000 / input rate for the next emu_celt_encode calls (pcm then holds N/up samples per channel)
extern "C" void emu_celt_set_upsample(int up) { g_enc_up = up; }
static int g_enc_pred = 2;   // CELT_SET_PREDICTION for the next emu_celt_encode calls
extern "C" void emu_celt_set_prediction(int v) { g_enc_pred = v; }
static int g_enc_coded = 0;   // CELT_SET_CHANNELS for the next emu_celt_encode calls (0 = as many as the input has)
extern "C" void emu_celt_set_channels(int c) { g_enc_coded = c; }
static int32_t g_enc_cov[3];   // frames that took the post-filter / patched-transient / transient paths (test coverage counters)
// one frame for n streams: an_state/enc_state carried by the caller
extern "C" void emu_celt_encode(void *an_state, ObEncState *enc_state, const float *pcm, int n, int C, int LM, int max_bytes, int bitrate, int lsb_depth,
                                unsigned char *packets, int32_t *nbytes, uint32_t *rng, int complexity, int vbr, int cvbr, float *prefilter_mem) {
  const int N = OB_SHORT_N << LM, CC = C;   // C becomes the coded channel count below
  if (g_enc_coded) C = g_enc_coded;
  std::vector<float> X((size_t)n * C * N), E((size_t)n * C * OB_NBANDS), Lg((size_t)n * C * OB_NBANDS), Lg2((size_t)n * C * OB_NBANDS);
  std::vector<ObEncFrameInfo> info(n);
  EncAnalysisParams a;
  a.state = (ObEncAnalysisState *)an_state; a.pcm = pcm; a.short_blocks = nullptr; a.X = X.data(); a.bandE = E.data(); a.bandLogE = Lg.data();
  a.info = info.data(); a.n_streams = n; a.C = C; a.CC = CC; a.LM = LM; a.complexity = complexity; a.lsb_depth = lsb_depth;
  a.enc_state = enc_state; a.prefilter_mem = prefilter_mem; a.bandLogE2 = Lg2.data(); a.max_bytes = max_bytes; a.bitrate = bitrate; a.vbr = vbr;
  a.constrained_vbr = cvbr; a.disable_pf = g_enc_pred <= 1; a.loss_rate = 0; a.end = g_enc_end; a.upsample = g_enc_up;
  emu_run_chains(a, n);
  emu_launch(enc_analysis_entry, &a, n, 32 * CC, CC * sizeof(MdctWS) + 32 + (complexity >= 5 ? sizeof(PfWS) : sizeof(float) * 64));
  EncFrameParams p;
  p.state = enc_state; p.info = info.data(); p.X = X.data(); p.bandE = E.data(); p.bandLogE = Lg.data(); p.bandLogE2 = Lg2.data(); p.packets = packets; p.nbytes = nbytes;
  p.n_streams = n; p.C = C; p.CC = CC; p.LM = LM; p.pkt_stride = max_bytes; p.max_bytes = max_bytes; p.complexity = complexity; p.bitrate = bitrate; p.lsb_depth = lsb_depth;
  p.force_intra = g_enc_pred == 0; p.disable_inv = 0; p.disable_pf = g_enc_pred <= 1; p.vbr = vbr; p.constrained_vbr = cvbr; p.loss_rate = 0; p.end = g_enc_end;
  emu_launch(enc_frame_entry, &p, (n + 31) / 32, 32, 0);
  for (int s = 0; s < n; s++) rng[s] = enc_state[s].rng;
  for (int s = 0; s < n; s++) { g_enc_cov[0] += info[s].pf_on; g_enc_cov[1] += info[s].patched; g_enc_cov[2] += info[s].is_transient; }
}
extern "C" void emu_enc_coverage(int32_t *out, int reset) { for (int i = 0; i < 3; i++) { out[i] = g_enc_cov[i]; if (reset) g_enc_cov[i] = 0; } }

// ---- the Opus layer in front of the CELT encoder (opus_b200/csrc/opus_enc_layer.cuh): opus_encode() for n streams ----
static void opus_pre_entry(void *a) {
  const OpusLayerParams &p = *(const OpusLayerParams *)a;
  opus_enc_pre_warp(p, block_id() * 32, (float *)dyn_smem());
}
static void opus_an_entry(void *a) {
  const OpusLayerParams &p = *(const OpusLayerParams *)a;
  if (block_id() < p.n_streams) opus_enc_analysis_warp(p, block_id(), (AnWS *)dyn_smem());
}
struct EmuOpusEncoder {
  int n, channels;
  OpusEncSettings cfg;
  int disable_inv, analysis_was_on;
  std::vector<ObOpusEncState> st;
  std::vector<ObTonality> tonal;
  std::vector<ObOpusPre> pre;
  std::vector<ObOpusFrameCtl> ctl;
  std::vector<float> pcm_buf, prefilter_mem;
  std::vector<ObEncAnalysisState> an_state;
  std::vector<ObEncState> enc_state;
};
extern "C" void *emu_opus_encoder_create(int n, int Fs, int channels, int application) {
  EmuOpusEncoder *e = new EmuOpusEncoder;
  e->n = n; e->channels = channels; e->disable_inv = 0; e->analysis_was_on = 0;
  opus_enc_settings_init(&e->cfg, Fs, channels, application);
  e->st.resize(n); e->tonal.resize(n); e->pre.resize(n); e->ctl.resize(n);
  e->pcm_buf.assign((size_t)n * channels * OB_MAXN, 0.f);
  e->prefilter_mem.assign((size_t)n * channels * kPfMaxPeriod, 0.f);
  e->an_state.resize(n); e->enc_state.resize(n);
  memset(e->tonal.data(), 0, sizeof(ObTonality) * n);
  memset(e->an_state.data(), 0, sizeof(ObEncAnalysisState) * n);
  memset(e->pre.data(), 0, sizeof(ObOpusPre) * n);
  memset(e->ctl.data(), 0, sizeof(ObOpusFrameCtl) * n);
  emu_enc_state_init(e->enc_state.data(), n);
  for (int s = 0; s < n; s++) opus_enc_state_init(&e->st[s], channels);
  return e;
}
extern "C" void emu_opus_encoder_destroy(void *h) { delete (EmuOpusEncoder *)h; }
extern "C" int emu_opus_encoder_ctl(void *h, int request, int value) {
  EmuOpusEncoder *e = (EmuOpusEncoder *)h;
  if (request == 4046) { e->disable_inv = value; return 0; }
  return opus_enc_settings_ctl(&e->cfg, request, value);
}
// the CELT kernels of one frame for every stream (celt_encode_launch of celt_enc_host.cuh): packets receive the payload behind the TOC byte
static int emu_opus_celt_frame(EmuOpusEncoder *e, const OpusCallPlan &pl, const float *pcm_buf, const ObOpusFrameCtl *ctl, int frame_size, unsigned char *packets,
                               int pkt_stride, int32_t *nbytes) {
  const int n = e->n, CC = e->channels, up = 48000 / e->cfg.Fs;
  int LM = -1;
  for (int l = 0; l <= 3; l++) if (frame_size * up == OB_SHORT_N << l) LM = l;
  if (LM < 0) return -1;
  const int N = OB_SHORT_N << LM, cx = e->cfg.complexity, max_bytes = pl.max_data_bytes - 1;
  std::vector<float> X((size_t)n * CC * N), E((size_t)n * CC * OB_NBANDS), Lg((size_t)n * CC * OB_NBANDS), Lg2((size_t)n * CC * OB_NBANDS);
  std::vector<ObEncFrameInfo> info(n);
  EncAnalysisParams a;
  a.state = e->an_state.data(); a.pcm = pcm_buf; a.short_blocks = nullptr; a.X = X.data(); a.bandE = E.data(); a.bandLogE = Lg.data();
  a.info = info.data(); a.n_streams = n; a.C = CC; a.CC = CC; a.LM = LM; a.complexity = cx; a.lsb_depth = pl.lsb;
  a.enc_state = e->enc_state.data(); a.prefilter_mem = e->prefilter_mem.data(); a.bandLogE2 = Lg2.data(); a.max_bytes = max_bytes; a.bitrate = pl.celt_bitrate;
  a.vbr = pl.celt_vbr; a.constrained_vbr = pl.celt_cvbr; a.disable_pf = pl.celt_disable_pf; a.loss_rate = e->cfg.packet_loss; a.end = OB_NBANDS; a.upsample = up;
  a.ctl = ctl;
  emu_run_chains(a, n);
  emu_launch(enc_analysis_entry, &a, n, 32 * CC, CC * sizeof(MdctWS) + 32 + (cx >= 5 ? sizeof(PfWS) : sizeof(float) * 64));
  EncFrameParams f;
  f.state = e->enc_state.data(); f.info = info.data(); f.X = X.data(); f.bandE = E.data(); f.bandLogE = Lg.data(); f.bandLogE2 = Lg2.data();
  f.packets = packets + 1; f.nbytes = nbytes; f.n_streams = n; f.C = CC; f.CC = CC; f.LM = LM; f.pkt_stride = pkt_stride; f.max_bytes = max_bytes;
  f.complexity = cx; f.bitrate = pl.celt_bitrate; f.lsb_depth = pl.lsb; f.force_intra = pl.celt_force_intra; f.disable_inv = e->disable_inv;
  f.disable_pf = pl.celt_disable_pf; f.vbr = pl.celt_vbr; f.constrained_vbr = pl.celt_cvbr; f.loss_rate = e->cfg.packet_loss; f.end = OB_NBANDS;
  f.ctl = ctl; f.toc_bytes = 1;
  if (g_frame_sync && f.complexity >= 8 && f.CC == 2) emu_launch(enc_frame_entry_sync, &f, (n + 63) / 64, 64, 0);
  else emu_launch(enc_frame_entry_ctl, &f, (n + 31) / 32, 32, 0);
  return 0;
}
static void opus_mf_decide_entry(void *a) {
  const OpusLayerParams &p = *(const OpusLayerParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n_streams) opus_multiframe_decide_thread(p, s);
}
static void opus_repack_entry(void *a) {
  const OpusLayerParams &p = *(const OpusLayerParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n_streams) opus_repacketize_thread(p, s);
}
extern "C" int emu_opus_encode(void *h, const void *pcm, int is16, int frame_size, unsigned char *packets, int pkt_stride, int max_data_bytes,
                               int32_t *nbytes, uint32_t *rng) {
  EmuOpusEncoder *e = (EmuOpusEncoder *)h;
  OpusCallPlan pl;
  const int pr = opus_enc_plan(e->cfg, frame_size, max_data_bytes, is16, &pl);
  if (pr) return pr;
  const int n = e->n, CC = e->channels;
  if (!pl.run_analysis && e->analysis_was_on) memset(e->tonal.data(), 0, sizeof(ObTonality) * n);
  e->analysis_was_on = pl.run_analysis;
  OpusLayerParams p;
  if (pl.nb_frames == 1) {
    if (pkt_stride < pl.max_data_bytes) return -1;
    opus_layer_params_init(&p, e->cfg, pl, frame_size, is16, n);
    p.st = e->st.data(); p.tonal = e->tonal.data(); p.pre = e->pre.data(); p.ctl = e->ctl.data(); p.pcm = pcm; p.pcm_buf = e->pcm_buf.data();
    p.packets = packets; p.pkt_stride = pkt_stride;
    emu_launch(opus_pre_entry, &p, (n + 31) / 32, 32, 32 * OB_PRE_TS * sizeof(float));
    emu_launch(opus_an_entry, &p, n, 32, sizeof(AnWS));
    const int r = emu_opus_celt_frame(e, pl, e->pcm_buf.data(), e->ctl.data(), frame_size, packets, pkt_stride, nbytes);
    if (r) return r;
  } else {   // opus_encode_multiframe of opus_enc_host.cuh
    const int nb = pl.nb_frames, fs = pl.enc_frame_size, stride = (pl.max_data_bytes + 15) & ~15;
    std::vector<ObOpusPre> pre((size_t)nb * n);
    std::vector<ObOpusFrameCtl> ctl((size_t)nb * n);
    std::vector<float> pcm_buf((size_t)nb * n * CC * OB_MAXN);
    std::vector<unsigned char> tmp((size_t)nb * n * stride);
    std::vector<int32_t> tnb((size_t)nb * n);
    memset(pre.data(), 0, sizeof(ObOpusPre) * pre.size());
    memset(ctl.data(), 0, sizeof(ObOpusFrameCtl) * ctl.size());
    opus_layer_params_init(&p, e->cfg, pl, fs, is16, n);
    p.st = e->st.data(); p.tonal = e->tonal.data(); p.pcm = pcm; p.packets = packets; p.pkt_stride = pkt_stride;
    p.multi = 1; p.pcm_row = frame_size * CC; p.nb_frames = nb; p.packet_frame_size = frame_size; p.pre_all = pre.data(); p.ctl_all = ctl.data();
    p.tmp_packets = tmp.data(); p.tmp_stride = stride; p.tmp_nbytes = tnb.data(); p.out_nbytes = nbytes; p.repacketize_len = pl.repacketize_len;
    p.pad = !e->cfg.use_vbr;
    for (int i = 0; i < nb; i++) {
      p.pre = pre.data() + (size_t)i * n; p.ctl = ctl.data() + (size_t)i * n; p.pcm_buf = pcm_buf.data() + (size_t)i * n * CC * OB_MAXN; p.pcm_off = i * fs * CC;
      emu_launch(opus_pre_entry, &p, (n + 31) / 32, 32, 32 * OB_PRE_TS * sizeof(float));
      if (pl.run_analysis) emu_launch(opus_an_entry, &p, n, 32, sizeof(AnWS));
    }
    emu_launch(opus_mf_decide_entry, &p, (n + 31) / 32, 32, 0);
    for (int i = 0; i < nb; i++) {
      const int r = emu_opus_celt_frame(e, pl, pcm_buf.data() + (size_t)i * n * CC * OB_MAXN, ctl.data() + (size_t)i * n, fs, tmp.data() + (size_t)i * n * stride, stride,
                                        tnb.data() + (size_t)i * n);
      if (r) return r;
    }
    emu_launch(opus_repack_entry, &p, (n + 31) / 32, 32, 0);
    for (int s = 0; s < n; s++) e->ctl[s] = ctl[(size_t)(nb - 1) * n + s];
  }
  for (int s = 0; s < n; s++) rng[s] = e->enc_state[s].rng;
  return 0;
}
// the analysis alone (kernels P1 + P2), as 30 floats per stream like ref_harness.c's refh_run_analysis_stream
extern "C" void emu_opus_last_analysis(void *h, float *out) {
  EmuOpusEncoder *e = (EmuOpusEncoder *)h;
  for (int s = 0; s < e->n; s++) {
    const ObAnalysisInfo &i = e->ctl[s].analysis;
    float *o = out + 30 * s;
    o[0] = (float)i.valid; o[1] = i.tonality; o[2] = i.tonality_slope; o[3] = i.noisiness; o[4] = i.activity; o[5] = i.music_prob; o[6] = i.music_prob_min;
    o[7] = i.music_prob_max; o[8] = (float)i.bandwidth; o[9] = i.activity_probability; o[10] = i.max_pitch_ratio;
    for (int k = 0; k < 19; k++) o[11 + k] = (float)i.leak_boost[k];
  }
}

// ---- SILK excitation coder (silk_coder.cuh): thread per stream on the emulator ----
#include "../../opus_b200/csrc/silk_coder.cuh"
static void sc_init_entry(void *a) { const SilkCoderParams &p = *(const SilkCoderParams *)a; int s = block_id() * 32 + lane_id(); if (s < p.n) silk_coder_init_thread(p, s); }
static void sc_icdf_entry(void *a) { const SilkCoderParams &p = *(const SilkCoderParams *)a; int s = block_id() * 32 + lane_id(); if (s < p.n) silk_coder_icdf_thread(p, s); }
static void sc_pulses_entry(void *a) { const SilkCoderParams &p = *(const SilkCoderParams *)a; int s = block_id() * 32 + lane_id(); if (s < p.n) silk_encode_pulses_thread(p, s); }
static void sc_done_entry(void *a) { const SilkCoderParams &p = *(const SilkCoderParams *)a; int s = block_id() * 32 + lane_id(); if (s < p.n) silk_coder_done_thread(p, s); }
static void sc_indices_entry(void *a) { const SilkCoderParams &p = *(const SilkCoderParams *)a; int s = block_id() * 32 + lane_id(); if (s < p.n) silk_encode_indices_thread(p, s); }
extern "C" {
void emu_silk_encode_indices(void *state, unsigned char *buf, int n, int stride, int storage, int fs_kHz, int nb_subfr, const void *indices, const int32_t *cond,
                             int32_t *prev) {
  SilkCoderParams p;
  memset(&p, 0, sizeof(p));
  p.state = (ObRangeEncState *)state; p.buf = buf; p.n = n; p.buf_stride = stride; p.storage = storage;
  p.indices = (const ObSilkIndices *)indices; p.cond_coding = cond; p.silk_prev = prev; p.fs_kHz = fs_kHz; p.nb_subfr = nb_subfr;
  emu_launch(sc_indices_entry, &p, (n + 31) / 32, 32, 0);
}
int emu_silk_coder_state_size(void) { return (int)sizeof(ObRangeEncState); }
// op: 0 init, 1 icdf (symbols, icdf table, ftb), 2 encode_pulses, 3 done
void emu_silk_coder(int op, void *state, unsigned char *buf, int n, int stride, int storage, const int32_t *a0, const int32_t *a1, const int8_t *pulses,
                    int pulses_stride, int frame_length, const unsigned char *icdf, int icdf_len, int ftb, uint32_t *tell_frac, uint32_t *final_range) {
  SilkCoderParams p;
  memset(&p, 0, sizeof(p));
  p.state = (ObRangeEncState *)state; p.buf = buf; p.n = n; p.buf_stride = stride; p.storage = storage;
  p.signal_type = a0; p.quant_offset_type = a1; p.symbols = a0; p.pulses = pulses; p.pulses_stride = pulses_stride; p.frame_length = frame_length;
  if (icdf) memcpy(p.icdf, icdf, icdf_len);
  p.ftb = ftb; p.tell_frac = tell_frac; p.final_range = final_range;
  emu_launch(op == 0 ? sc_init_entry : op == 1 ? sc_icdf_entry : op == 2 ? sc_pulses_entry : sc_done_entry, &p, (n + 31) / 32, 32, 0);
}
}

// ---- LPCNet feature extraction (lpcnet_enc.cuh): the two warp kernels around the pitch network, and the Burg cepstrum ----
#include "../../opus_b200/csrc/lpcnet_enc.cuh"
static void le_analysis_entry(void *a) { lpcnet_enc_analysis_warp(*(const LpcnetEncParams *)a, block_id(), (LeWS *)dyn_smem()); }
static void le_finish_entry(void *a) { lpcnet_enc_finish_warp(*(const LpcnetEncParams *)a, block_id(), (float *)dyn_smem()); }
static void le_burg_entry(void *a) { lpcnet_burg_cepstrum_warp(*(const LpcnetBurgParams *)a, block_id(), (LeBurgWS *)dyn_smem()); }
extern "C" int emu_lpcnet_enc_state_size(void) { return (int)sizeof(ObLpcnetEncState); }
extern "C" void emu_lpcnet_enc_analysis(void *state, const void *pcm, int is16, float *if_features, float *conv_in, int n) {
  LpcnetEncParams p = {};
  p.state = (ObLpcnetEncState *)state; p.pcm = pcm; p.pcm_is_int16 = is16; p.if_features = if_features; p.conv_in = conv_in; p.n = n;
  emu_launch(le_analysis_entry, &p, n, 32, sizeof(LeWS));
}
extern "C" void emu_lpcnet_enc_finish(void *state, const float *pitch_out, float *features, int n) {
  LpcnetEncParams p = {};
  p.state = (ObLpcnetEncState *)state; p.pitch_out = pitch_out; p.features = features; p.n = n;
  emu_launch(le_finish_entry, &p, n, 32, kLeBuf * sizeof(float));
}
extern "C" void emu_lpcnet_burg(float *ceps, const float *x, int n) {
  LpcnetBurgParams p;
  p.ceps = ceps; p.x = x; p.n = n;
  emu_launch(le_burg_entry, &p, n, 32, sizeof(LeBurgWS));
}
static void dred_conv_entry(void *a) {
  const DredConvParams &p = *(const DredConvParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n) dred_convert_thread(p, s);
}
extern "C" void emu_dred_convert_16k(const float *pcm, float *input_buffer, float *resample_mem, int n, int Fs, int channels, int in_len, int out_len, int fill,
                                     int pcm_off, int pcm_stride) {
  DredConvParams p;
  p.pcm = pcm; p.input_buffer = input_buffer; p.resample_mem = resample_mem; p.n = n; p.Fs = Fs; p.channels = channels; p.in_len = in_len; p.out_len = out_len;
  p.fill = fill; p.pcm_off = pcm_off; p.pcm_stride = pcm_stride;
  emu_launch(dred_conv_entry, &p, (n + 31) / 32, 32, 0);
}

// ---- DRED entropy coding (dred_coding.cuh), one thread per stream ----
#include "../../opus_b200/csrc/dred_coding.cuh"
static void dred_code_entry(void *a) {
  const DredCodeParams &p = *(const DredCodeParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n) p.nbytes[s] = dred_encode_silk_frame_thread(p, s);
}
extern "C" void emu_dred_encode_silk_frame(const float *lat_ring, const float *state_ring, const uint8_t *const *stats, const uint8_t *activity, int32_t *last_extra,
                                           unsigned char *buf, int32_t *nbytes, const int32_t *ints /* n buf_stride head latent_dim state_dim fill latent_offset
                                           dred_offset max_chunks max_bytes q0 dQ qmax */) {
  DredCodeParams p;
  p.lat_ring = lat_ring; p.state_ring = state_ring;
  for (int k = 0; k < 8; k++) p.stats[k] = stats[k];
  p.activity = activity; p.last_extra = last_extra; p.buf = buf; p.nbytes = nbytes;
  p.n = ints[0]; p.buf_stride = ints[1]; p.head = ints[2]; p.latent_dim = ints[3]; p.state_dim = ints[4]; p.latents_buffer_fill = ints[5]; p.latent_offset = ints[6];
  p.dred_offset = ints[7]; p.max_chunks = ints[8]; p.max_bytes = ints[9]; p.q0 = ints[10]; p.dQ = ints[11]; p.qmax = ints[12];
  emu_launch(dred_code_entry, &p, (p.n + 31) / 32, 32, 0);
}

// ---- silk_process_gains_FLP (silk_ops.cuh), thread per stream ----
static void silk_gains_entry(void *a) {
  const SilkGainsParams &p = *(const SilkGainsParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n) silk_process_gains_thread(p, s);
}
extern "C" void emu_silk_process_gains(void *frames, int n, int nb_subfr, int subfr_length, int nstates) {
  SilkGainsParams p;
  p.frames = (ObSilkGains *)frames; p.n = n; p.nb_subfr = nb_subfr; p.subfr_length = subfr_length; p.nStatesDelayedDecision = nstates;
  emu_launch(silk_gains_entry, &p, (n + 31) / 32, 32, 0);
}

// ---- silk_process_NLSFs (silk_nlsf.cuh), thread per stream ----
#include "../../opus_b200/csrc/silk_nlsf.cuh"
static void silk_nlsf_entry(void *a) {
  const SilkNlsfParams &p = *(const SilkNlsfParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n) silk_process_nlsfs_thread(p, s);
}
extern "C" void emu_silk_process_nlsfs(void *frames, int n, int wb, int nb_subfr) {
  SilkNlsfParams p;
  p.frames = (ObSilkNlsf *)frames; p.n = n; p.wb = wb; p.nb_subfr = nb_subfr;
  emu_launch(silk_nlsf_entry, &p, (n + 31) / 32, 32, 0);
}
static void silk_find_lpc_entry(void *a) {
  const SilkFindLpcParams &p = *(const SilkFindLpcParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n) silk_find_lpc_thread(p, s);
}
extern "C" void emu_silk_find_lpc(void *frames, const float *x, int n, int order, int nb_subfr, int subfr_length) {
  SilkFindLpcParams p;
  p.frames = (ObSilkFindLpc *)frames; p.x = x; p.n = n; p.order = order; p.nb_subfr = nb_subfr; p.subfr_length = subfr_length;
  emu_launch(silk_find_lpc_entry, &p, (n + 31) / 32, 32, 0);
}
static void silk_pred_coefs_entry(void *a) {
  const SilkPredCoefsParams &p = *(const SilkPredCoefsParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n) silk_find_pred_coefs_thread(p, s);
}
extern "C" void emu_silk_find_pred_coefs(void *frames, const float *res_pitch, const float *x, int n, int sig_stride, int frame_off, int order, int nb_subfr,
                                         int subfr_length) {
  SilkPredCoefsParams p;
  p.frames = (ObSilkPredCoefs *)frames; p.res_pitch = res_pitch; p.x = x; p.n = n; p.sig_stride = sig_stride; p.frame_off = frame_off; p.order = order;
  p.nb_subfr = nb_subfr; p.subfr_length = subfr_length;
  emu_launch(silk_pred_coefs_entry, &p, (n + 31) / 32, 32, 0);
}
static void silk_noise_shape_entry(void *a) {
  const SilkNoiseShapeParams &p = *(const SilkNoiseShapeParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n) silk_noise_shape_thread(p, s);
}
extern "C" void emu_silk_noise_shape(void *frames, const float *pitch_res, int res_stride, const float *x, int x_stride, int x_off, int n, int fs_kHz, int nb_subfr,
                                     int order) {
  SilkNoiseShapeParams p;
  p.frames = (ObSilkNoiseShape *)frames; p.pitch_res = pitch_res; p.x = x; p.n = n; p.res_stride = res_stride; p.x_stride = x_stride; p.x_off = x_off;
  p.fs_kHz = fs_kHz; p.nb_subfr = nb_subfr; p.shapingLPCOrder = order;
  emu_launch(silk_noise_shape_entry, &p, (n + 31) / 32, 32, 0);
}

// ---- silk_find_pitch_lags_FLP (silk_pitch.cuh), thread per stream ----
#include "../../opus_b200/csrc/silk_pitch.cuh"
static void silk_pitch_entry(void *a) {
  const SilkPitchParams &p = *(const SilkPitchParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n) silk_find_pitch_lags_thread(p, s);
}
extern "C" void emu_silk_find_pitch_lags(void *frames, float *res, int res_stride, const float *x, int x_stride, int x_off, int n, int fs_kHz, int nb_subfr, int order,
                                         int complexity) {
  SilkPitchParams p;
  p.frames = (ObSilkPitch *)frames; p.res = res; p.x = x; p.n = n; p.x_stride = x_stride; p.x_off = x_off; p.res_stride = res_stride; p.fs_kHz = fs_kHz;
  p.nb_subfr = nb_subfr; p.lpc_order = order; p.complexity = complexity;
  emu_launch(silk_pitch_entry, &p, (n + 31) / 32, 32, 0);
}

// ---- silk_encode_do_VAD_FLP / silk_HP_variable_cutoff (silk_vad.cuh), thread per stream ----
#include "../../opus_b200/csrc/silk_vad.cuh"
static void silk_vad_entry(void *a) {
  const SilkVadParams &p = *(const SilkVadParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n) silk_encode_do_vad_thread(p, s);
}
static void silk_hp_cutoff_entry(void *a) {
  const SilkHpCutoffParams &p = *(const SilkHpCutoffParams *)a;
  const int s = block_id() * 32 + lane_id();
  if (s < p.n) silk_hp_variable_cutoff_thread(p, s);
}
extern "C" void emu_silk_encode_do_vad(void *frames, const int16_t *pcm, int pcm_stride, const int32_t *activity, int n, int frame_length, int fs_kHz) {
  SilkVadParams p;
  p.frames = (ObSilkVad *)frames; p.pcm = pcm; p.activity = activity; p.n = n; p.pcm_stride = pcm_stride; p.frame_length = frame_length; p.fs_kHz = fs_kHz;
  emu_launch(silk_vad_entry, &p, (n + 31) / 32, 32, 0);
}
extern "C" void emu_silk_hp_variable_cutoff(void *frames, int n, int fs_kHz) {
  SilkHpCutoffParams p;
  p.frames = (ObSilkVad *)frames; p.n = n; p.fs_kHz = fs_kHz;
  emu_launch(silk_hp_cutoff_entry, &p, (n + 31) / 32, 32, 0);
}
