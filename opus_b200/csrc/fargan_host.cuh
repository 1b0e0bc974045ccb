// fargan_host.cuh — the FARGAN vocoder graph (dnn/fargan.c:41-217) for a batch of streams: the layers run through the batched
// dnn primitives of nnet_host.cuh (int8 layers on the tensor cores), the glue between them (pitch prediction, gating,
// concatenations, the pitch-buffer update and de-emphasis) in the small kernels below.  Activations live in HBM as
// [n_streams][width] row-major buffers; every kernel is one launch over the whole batch (included at the end of opus_b200.cu).
#pragma once

namespace ob {
constexpr int kFgFeatures = 20, kFgBands = 18, kFgPitchMax = 256, kFgSub = 40, kFgNbSub = 4, kFgFrame = 160, kFgCont = 320, kFgCond = 80,
              kFgSigIn = kFgCond + 2 * kFgSub + 4, kFgPembed = 12, kFgFwc0 = 192, kFgGru1 = 160, kFgGru2 = 128, kFgGru3 = 128, kFgSkip = 128,
              kFgSkipCat = kFgGru1 + kFgGru2 + kFgGru3 + kFgFwc0 + 2 * kFgSub;
enum { FG_PEMBED, FG_FDENSE1, FG_FCONV1, FG_FDENSE2, FG_COND_GAIN, FG_FWC0_CONV, FG_FWC0_GLU, FG_GRU1_IN, FG_GRU1_REC, FG_GRU2_IN, FG_GRU2_REC,
       FG_GRU3_IN, FG_GRU3_REC, FG_GRU1_GLU, FG_GRU2_GLU, FG_GRU3_GLU, FG_SKIP_GLU, FG_SKIP_DENSE, FG_SIG_OUT, FG_GAIN_OUT, FG_COUNT };

struct FgStreamState {   // FARGANState minus the layer-owned memories (dnn/fargan.h:45-57)
  float pitch_buf[kFgPitchMax];
  float deemph_mem;
  int32_t last_period, period;   // period: of the frame being synthesised (becomes last_period at its end)
  int32_t pad;
};
struct FgParams {
  FgStreamState *st;
  const float *features;   // [n][20]
  const float *pembed;     // [224][12]
  float *dense_in;         // [n][32]
  const float *cond;       // [n][320]
  const float *gain_w, *gain_b;   // sig_net_cond_gain_dense: [80] float weights, [1] bias
  float *gain;             // [n]
  float *pred, *prev;      // [n][44], [n][40]
  float *fwc0_in;          // [n][164]
  const float *x192;       // [n][192]  fwc0 conv + GLU output
  const float *pgate;      // [n][4]
  const float *g1, *g2, *g3;   // GLU outputs of the three GRUs
  const float *sig;        // [n][40]
  float *pcm;              // [n][160]
  const float *pcm0;       // [n][320] (continuation)
  int n, sub, which, use_last;
};
OB_DEV int fg_period(float f) {   // dnn/fargan.c:148,198
  return (int)floor(.5 + 256. / pow((double)2.f, ((1. / 60.) * (((double)f + 1.5) * 60))));
}
}  // namespace ob
using namespace ob;

// compute_fargan_cond's input: [features | pembed row of the period] (dnn/fargan.c:51-52); also latches the period
extern "C" __global__ void ob_fargan_cond_in(FgParams p) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int s = (int)(t / 32), j = (int)(t % 32);
  if (s >= p.n) return;
  const int period = fg_period(p.features[(size_t)s * kFgFeatures + kFgBands]);
  if (j < kFgFeatures) p.dense_in[(size_t)s * 32 + j] = p.features[(size_t)s * kFgFeatures + j];
  else p.dense_in[(size_t)s * 32 + j] = p.pembed[OB_IMAX(0, OB_IMIN(period - 32, 223)) * kFgPembed + (j - kFgFeatures)];
  if (j == 0) {
    if (p.which) p.st[s].last_period = p.st[s].period;   // fargan_cont's pre-load loop: last_period trails by one vector
    p.st[s].period = period;
  }
}
// gain, pitch prediction, previous subframe and the first conv's input (dnn/fargan.c:89-104)
extern "C" __global__ void ob_fargan_sub_in(FgParams p) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int s = (int)(t / 64), j = (int)(t % 64);
  if (s >= p.n) return;
  const FgStreamState &st = p.st[s];
  const float *cond = p.cond + (size_t)s * 4 * kFgCond + p.sub * kFgCond;
  // sig_net_cond_gain_dense (80 -> 1, float weights): sgemv's column order (dnn/vec.h:104-121), then the bias; one lane per warp
  float gl = 0.f;
  if ((threadIdx.x & 31) == 0) {   // a warp serves one stream (64 threads = 2 warps): one lane sums, the others receive
#pragma unroll 8
    for (int k = 0; k < kFgCond; k++) gl = __fadd_rn(gl, __fmul_rn(p.gain_w[k], cond[k]));
    if (p.gain_b) gl = __fadd_rn(gl, p.gain_b[0]);
  }
  gl = __shfl_sync(0xffffffffu, gl, 0);
  const float g = (float)exp((double)gl);
  const float gain_1 = 1.f / (1e-5f + g);
  const int period = st.last_period;
  float *fin = p.fwc0_in + (size_t)s * kFgSigIn;
  if (j == 0) p.gain[s] = g;
  if (j < kFgSub + 4) {
    int pos = kFgPitchMax - period - 2 + j;
    while (pos >= kFgPitchMax) pos -= period;   // the wrap of :93-94 after j steps
    const float v = gain_1 * st.pitch_buf[OB_IMAX(0, pos)];
    const float pr = v > 1.f ? 1.f : (v < -1.f ? -1.f : v);
    p.pred[(size_t)s * (kFgSub + 4) + j] = pr;
    fin[kFgCond + j] = pr;
  }
  if (j < kFgSub) {
    const float v = gain_1 * st.pitch_buf[kFgPitchMax - kFgSub + j];
    const float pv = v < 1.f ? (v > -1.f ? v : -1.f) : 1.f;
    p.prev[(size_t)s * kFgSub + j] = pv;
    fin[kFgCond + kFgSub + 4 + j] = pv;
  }
  for (int k = j; k < kFgCond; k += 64) fin[k] = cond[k];
}
// scale by the gain, push into the pitch buffer, de-emphasis (dnn/fargan.c:136-140, :61-67).  64 threads per stream move the
// pitch buffer (values read, barrier, values written); one of them runs the 40-step de-emphasis recurrence.
// pcm0 != null: fargan_cont's loop, where the known signal replaces the synthesised subframe afterwards (:165).
extern "C" __global__ void __launch_bounds__(256) ob_fargan_sub_out(FgParams p) {
  const int s = blockIdx.x * 4 + (threadIdx.x >> 6), j = threadIdx.x & 63;
  const bool ok = s < p.n;
  FgStreamState *st = ok ? p.st + s : p.st;
  const float gain = ok ? p.gain[s] : 0.f;
  float nv[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int idx = j + 64 * k;
    if (!ok) nv[k] = 0.f;
    else if (idx < kFgPitchMax - kFgSub) nv[k] = st->pitch_buf[idx + kFgSub];
    else {
      const int i = idx - (kFgPitchMax - kFgSub);
      if (p.pcm0) {
        const float *pc = p.pcm0 + (size_t)s * kFgCont;
        const int kk = kFgFrame + p.sub * kFgSub + i;
        nv[k] = pc[kk] - 0.85f * pc[kk - 1];
      } else nv[k] = p.sig[(size_t)s * kFgSub + i] * gain;
    }
  }
  __syncthreads();
  if (!ok) return;
#pragma unroll
  for (int k = 0; k < 4; k++) st->pitch_buf[j + 64 * k] = nv[k];
  if (j == 0) {
    if (p.pcm0) {
      if (p.sub == kFgNbSub - 1) st->deemph_mem = p.pcm0[(size_t)s * kFgCont + kFgCont - 1];
    } else {
      float mem = st->deemph_mem;
#pragma unroll 8
      for (int i = 0; i < kFgSub; i++) {
        float x = p.sig[(size_t)s * kFgSub + i] * gain;
        x += 0.85f * mem;
        mem = x;
        p.pcm[(size_t)s * kFgFrame + p.sub * kFgSub + i] = x;
      }
      st->deemph_mem = mem;
    }
  }
}
// fargan_cont's prologue (:157-161): x0 = pre-emphasised pcm0, first frame into the pitch buffer
extern "C" __global__ void ob_fargan_cont_in(FgParams p) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int s = (int)(t / kFgFrame), i = (int)(t % kFgFrame);
  if (s >= p.n) return;
  const float *pc = p.pcm0 + (size_t)s * kFgCont;
  p.st[s].pitch_buf[kFgPitchMax - kFgFrame + i] = i == 0 ? 0.f : pc[i] - 0.85f * pc[i - 1];
}
extern "C" __global__ void ob_fargan_end_frame(FgParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) p.st[s].last_period = p.st[s].period;
}

struct OpusB200Fargan {
  OpusB200Ctx *ctx;
  int n, cont_initialized;
  OpusB200Linear *l[FG_COUNT];
  FgStreamState *st;
  float *cond_conv1_state, *fwc0_buf[2], *gru1, *gru2, *gru3;       // layer memories [n][..]; fwc0_buf: this / the previous subframe's conv input
  int fwc0_cur;
  float *dense_in, *conv1_in, *fdense2_in, *cond, *gain, *pred, *prev, *x192, *pgate, *g1, *g2, *g3, *skip, *sig;
  float *d_feat, *d_pcm, *d_pcm0;   // staging for host callers
  ObGraphCache graph;               // one synthesised frame as a CUDA graph
};
static void fg_free(OpusB200Fargan *f) {
  if (!f) return;
  if (f->graph.exec) cudaGraphExecDestroy(f->graph.exec);
  for (int i = 0; i < FG_COUNT; i++) opusb200_nnet_linear_destroy(f->l[i]);
  float **bufs[] = {&f->cond_conv1_state, &f->fwc0_buf[0], &f->fwc0_buf[1], &f->gru1, &f->gru2, &f->gru3, &f->dense_in, &f->conv1_in, &f->fdense2_in, &f->cond,
                    &f->gain, &f->pred, &f->prev, &f->x192, &f->pgate, &f->g1, &f->g2, &f->g3, &f->skip, &f->sig, &f->d_feat, &f->d_pcm,
                    &f->d_pcm0};
  for (auto b : bufs) cudaFree(*b);
  cudaFree(f->st);
  delete f;
}
// fargan_init + init_fargan (dnn/fargan.c:173-184): the model is the 20 LinearLayers of the generated FARGAN struct, in its order
extern "C" int opusb200_fargan_create(OpusB200Ctx *c, const OpusB200LinearLayer *layers, int n_layers, int su_bias, int n_streams, OpusB200Fargan **out) {
  if (!out) return OPUSB200_BAD_ARG;
  *out = nullptr;
  if (!c || !layers || n_layers != FG_COUNT || n_streams <= 0) return OPUSB200_BAD_ARG;
  static const int dims[FG_COUNT][2] = {{224, 12}, {32, 64}, {192, 128}, {128, 320}, {80, 1}, {328, 192}, {192, 192}, {272, 480}, {160, 480}, {240, 384},
                                        {128, 384}, {208, 384}, {128, 384}, {160, 160}, {128, 128}, {128, 128}, {128, 128}, {688, 128}, {128, 40}, {192, 4}};
  for (int i = 0; i < FG_COUNT; i++)
    if (layers[i].nb_inputs != dims[i][0] || layers[i].nb_outputs != dims[i][1]) return OPUSB200_BAD_ARG;   // linear_init's size checks
  OB_CUDA(cudaSetDevice(c->device));
  OpusB200Fargan *f = new (std::nothrow) OpusB200Fargan;
  if (!f) return OPUSB200_ALLOC_FAIL;
  memset(f, 0, sizeof(*f));
  f->ctx = c; f->n = n_streams;
  for (int i = 0; i < FG_COUNT; i++) {
    const int r = opusb200_nnet_linear_create(c, &layers[i], su_bias, &f->l[i]);
    if (r != OPUSB200_OK) { fg_free(f); return r; }
  }
  const size_t n = (size_t)n_streams;
  struct { float **p; size_t w; } al[] = {{&f->cond_conv1_state, 128}, {&f->fwc0_buf[0], kFgSigIn}, {&f->fwc0_buf[1], kFgSigIn}, {&f->gru1, kFgGru1}, {&f->gru2, kFgGru2}, {&f->gru3, kFgGru3},
                                          {&f->dense_in, 32}, {&f->conv1_in, 64}, {&f->fdense2_in, 128}, {&f->cond, 320}, {&f->gain, 1},
                                          {&f->pred, kFgSub + 4}, {&f->prev, kFgSub}, {&f->x192, kFgFwc0}, {&f->pgate, 4},
                                          {&f->g1, kFgGru1}, {&f->g2, kFgGru2}, {&f->g3, kFgGru3}, {&f->skip, kFgSkip}, {&f->sig, kFgSub},
                                          {&f->d_feat, 5 * kFgFeatures}, {&f->d_pcm, kFgFrame}, {&f->d_pcm0, kFgCont}};
  for (auto &a : al)
    if (cudaMalloc(a.p, sizeof(float) * n * a.w) != cudaSuccess) { fg_free(f); return OPUSB200_ALLOC_FAIL; }
  if (cudaMalloc(&f->st, sizeof(FgStreamState) * n) != cudaSuccess) { fg_free(f); return OPUSB200_ALLOC_FAIL; }
  cudaMemsetAsync(f->st, 0, sizeof(FgStreamState) * n, c->stream);
  cudaMemsetAsync(f->cond_conv1_state, 0, sizeof(float) * n * 128, c->stream);
  cudaMemsetAsync(f->fwc0_buf[0], 0, sizeof(float) * n * kFgSigIn, c->stream);
  cudaMemsetAsync(f->fwc0_buf[1], 0, sizeof(float) * n * kFgSigIn, c->stream);
  cudaMemsetAsync(f->gru1, 0, sizeof(float) * n * kFgGru1, c->stream);
  cudaMemsetAsync(f->gru2, 0, sizeof(float) * n * kFgGru2, c->stream);
  cudaMemsetAsync(f->gru3, 0, sizeof(float) * n * kFgGru3, c->stream);
  OB_CUDA(cudaStreamSynchronize(c->stream));
  *out = f;
  return OPUSB200_OK;
}
extern "C" void opusb200_fargan_destroy(OpusB200Fargan *f) {
  if (!f) return;
  cudaSetDevice(f->ctx->device);
  cudaStreamSynchronize(f->ctx->stream);
  fg_free(f);
}
#define FG_TRY(x) do { const int r_ = (x); if (r_ != OPUSB200_OK) return r_; } while (0)
static inline unsigned fg_blk(long total) { return (unsigned)((total + 255) / 256); }
// compute_fargan_cond (dnn/fargan.c:41-59) on device features [n][20]; trail: the pre-load loop of fargan_cont
static int fg_cond(OpusB200Fargan *f, const float *d_features, int trail) {
  OpusB200Ctx *c = f->ctx;
  const int n = f->n;
  FgParams p = {};
  p.st = f->st; p.features = d_features; p.pembed = f->l[FG_PEMBED]->L.wf; p.dense_in = f->dense_in; p.n = n; p.which = trail;
  ob_fargan_cond_in<<<fg_blk((long)n * 32), 256, 0, c->stream>>>(p);
  c->launches++;
  FG_TRY(opusb200_nnet_compute_dense(c, f->l[FG_FDENSE1], f->conv1_in, f->dense_in, n, NN_TANH, 1));
  FG_TRY(opusb200_nnet_compute_conv1d(c, f->l[FG_FCONV1], f->fdense2_in, f->cond_conv1_state, f->conv1_in, n, 64, 1, NN_TANH, 1));
  FG_TRY(opusb200_nnet_compute_dense(c, f->l[FG_FDENSE2], f->cond, f->fdense2_in, n, NN_TANH, 1));
  return OPUSB200_OK;
}
// run_fargan_subframe (dnn/fargan.c:69-141)
static int fg_subframe(OpusB200Fargan *f, int sub, float *d_pcm, const float *d_pcm0) {
  OpusB200Ctx *c = f->ctx;
  const int n = f->n;
  FgParams p = {};
  p.st = f->st; p.cond = f->cond; p.gain = f->gain; p.pred = f->pred; p.prev = f->prev; p.x192 = f->x192;
  p.pgate = f->pgate; p.g1 = f->g1; p.g2 = f->g2; p.g3 = f->g3; p.sig = f->sig; p.pcm = d_pcm; p.pcm0 = d_pcm0; p.n = n; p.sub = sub;
  // No float concatenation is ever built: every wide int8 layer gets its operand from a segmented quantiser that reads the pieces
  // where they are ([head | pitch_gate[k] * pred[2..42) | prev], dnn/fargan.c:112-132); the first conv's window is
  // [previous subframe's input | this one's] (kernel size 2), i.e. two buffers that swap roles.
  float *fin_prev = f->fwc0_buf[f->fwc0_cur ^ 1], *fin = f->fwc0_buf[f->fwc0_cur];
  f->fwc0_cur ^= 1;
  p.fwc0_in = fin;
  p.gain_w = f->l[FG_COND_GAIN]->L.wf; p.gain_b = f->l[FG_COND_GAIN]->L.bias;
  ob_fargan_sub_in<<<fg_blk((long)n * 64), 256, 0, c->stream>>>(p);
  c->launches++;
  const int8_t *xq;
  auto seg = [](const float *ptr, int width, int stride, const float *gate = nullptr, int gate_stride = 0) {
    UmmaSeg sg; sg.ptr = ptr; sg.gate = gate; sg.width = width; sg.stride = stride; sg.gate_stride = gate_stride; return sg;
  };
  auto gated = [&](int k) { return seg(f->pred + 2, kFgSub, kFgSub + 4, f->pgate + k, 4); };
  const UmmaSeg s_prev = seg(f->prev, kFgSub, kFgSub);
  {  // compute_generic_conv1d(sig_net_fwc0_conv) + tanh
    const UmmaSeg sg[2] = {seg(fin_prev, kFgSigIn, kFgSigIn), seg(fin, kFgSigIn, kFgSigIn)};
    FG_TRY(nnet_umma_quant_seg(c, f->l[FG_FWC0_CONV], sg, 2, n, &xq));
    FG_TRY(nnet_umma_gemm(c, f->l[FG_FWC0_CONV], f->x192, xq, nullptr, n));
    FG_TRY(nnet_act_dev(c, f->x192, f->x192, n, kFgFwc0, NN_TANH));
  }
  FG_TRY(opusb200_nnet_compute_glu(c, f->l[FG_FWC0_GLU], f->x192, f->x192, n, 1));
  FG_TRY(opusb200_nnet_compute_dense(c, f->l[FG_GAIN_OUT], f->pgate, f->x192, n, NN_SIGMOID, 1));
  // the three GRUs: input matrix on the segmented operand, recurrent matrix on the state, then the gate kernel (dnn/nnet.c:76-105)
  auto gru = [&](int li, int lr, float *state, const UmmaSeg *sg, int N) -> int {
    float *zrh;
    if (nnet_scratch(c, 1, sizeof(float) * (size_t)n * 6 * N, (void **)&zrh)) return OPUSB200_ALLOC_FAIL;
    float *recur = zrh + (size_t)n * 3 * N;
    const int8_t *q;
    FG_TRY(nnet_umma_quant_seg(c, f->l[li], sg, 3, n, &q));
    FG_TRY(nnet_umma_gemm(c, f->l[li], zrh, q, nullptr, n));
    FG_TRY(nnet_linear_dev(c, f->l[lr], recur, state, n));
    NnetGruParams g;
    g.state = state; g.zrh = zrh; g.recur = recur; g.n = n; g.N = N;
    ob_nnet_gru<<<nblk((long)n * N, 256), 256, 0, c->stream>>>(g);
    c->launches++;
    return OPUSB200_OK;
  };
  {
    const UmmaSeg sg[3] = {seg(f->x192, kFgFwc0, kFgFwc0), gated(0), s_prev};
    FG_TRY(gru(FG_GRU1_IN, FG_GRU1_REC, f->gru1, sg, kFgGru1));
    FG_TRY(opusb200_nnet_compute_glu(c, f->l[FG_GRU1_GLU], f->g1, f->gru1, n, 1));
  }
  {
    const UmmaSeg sg[3] = {seg(f->g1, kFgGru1, kFgGru1), gated(1), s_prev};
    FG_TRY(gru(FG_GRU2_IN, FG_GRU2_REC, f->gru2, sg, kFgGru2));
    FG_TRY(opusb200_nnet_compute_glu(c, f->l[FG_GRU2_GLU], f->g2, f->gru2, n, 1));
  }
  {
    const UmmaSeg sg[3] = {seg(f->g2, kFgGru2, kFgGru2), gated(2), s_prev};
    FG_TRY(gru(FG_GRU3_IN, FG_GRU3_REC, f->gru3, sg, kFgGru3));
    FG_TRY(opusb200_nnet_compute_glu(c, f->l[FG_GRU3_GLU], f->g3, f->gru3, n, 1));
  }
  {  // skip_cat = [g1 | g2 | g3 | fwc0 | gated pred | prev] -> sig_net_skip_dense + tanh
    const UmmaSeg sg[6] = {seg(f->g1, kFgGru1, kFgGru1), seg(f->g2, kFgGru2, kFgGru2), seg(f->g3, kFgGru3, kFgGru3), seg(f->x192, kFgFwc0, kFgFwc0), gated(3),
                           s_prev};
    FG_TRY(nnet_umma_quant_seg(c, f->l[FG_SKIP_DENSE], sg, 6, n, &xq));
    FG_TRY(nnet_umma_gemm(c, f->l[FG_SKIP_DENSE], f->skip, xq, nullptr, n));
    FG_TRY(nnet_act_dev(c, f->skip, f->skip, n, kFgSkip, NN_TANH));
  }
  FG_TRY(opusb200_nnet_compute_glu(c, f->l[FG_SKIP_GLU], f->skip, f->skip, n, 1));
  FG_TRY(opusb200_nnet_compute_dense(c, f->l[FG_SIG_OUT], f->sig, f->skip, n, NN_TANH, 1));
  ob_fargan_sub_out<<<(unsigned)((n + 3) / 4), 256, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
// fargan_cont (dnn/fargan.c:143-170): pcm0 [n][320], features0 [n][5*20]
extern "C" int opusb200_fargan_cont(OpusB200Fargan *f, const float *pcm0, const float *features0, int on_device) {
  if (!f || !pcm0 || !features0) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = f->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)f->n;
  const float *dp = pcm0, *df = features0;
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(f->d_pcm0, pcm0, sizeof(float) * n * kFgCont, cudaMemcpyHostToDevice, c->stream));
    OB_CUDA(cudaMemcpyAsync(f->d_feat, features0, sizeof(float) * n * 5 * kFgFeatures, cudaMemcpyHostToDevice, c->stream));
    dp = f->d_pcm0; df = f->d_feat;
  }
  float *feat1;   // one feature vector per stream, contiguous
  if (nnet_scratch(c, 1, sizeof(float) * n * 1024, (void **)&feat1)) return OPUSB200_ALLOC_FAIL;   // (also pre-sizes the shared scratch)
  float *fv = f->skip;   // [n][20] fits in the [n][128] skip buffer, unused until the first subframe
  for (int i = 0; i < 5; i++) {
    OB_CUDA(cudaMemcpy2DAsync(fv, sizeof(float) * kFgFeatures, df + i * kFgFeatures, sizeof(float) * 5 * kFgFeatures, sizeof(float) * kFgFeatures, n,
                              cudaMemcpyDeviceToDevice, c->stream));
    FG_TRY(fg_cond(f, fv, 1));
  }
  FgParams p = {};
  p.st = f->st; p.pcm0 = dp; p.n = f->n;
  ob_fargan_cont_in<<<fg_blk((long)n * kFgFrame), 256, 0, c->stream>>>(p);
  c->launches++;
  f->cont_initialized = 1;
  for (int i = 0; i < kFgNbSub; i++) FG_TRY(fg_subframe(f, i, nullptr, dp));
  OB_CUDA(cudaGetLastError());
  if (!on_device) OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}
// fargan_synthesize (dnn/fargan.c:191-213): features [n][20] -> pcm [n][160] float
extern "C" int opusb200_fargan_synthesize(OpusB200Fargan *f, float *pcm, const float *features, int on_device) {
  if (!f || !pcm || !features) return OPUSB200_BAD_ARG;
  if (!f->cont_initialized) return OPUSB200_BAD_ARG;   // celt_assert(st->cont_initialized)
  OpusB200Ctx *c = f->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)f->n;
  const float *df = features;
  float *dp = pcm;
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(f->d_feat, features, sizeof(float) * n * kFgFeatures, cudaMemcpyHostToDevice, c->stream));
    df = f->d_feat; dp = f->d_pcm;
  }
  // a frame is ~200 launches whose arguments only depend on (features, pcm) pointers: replayed as a CUDA graph from the third call
  FG_TRY(ob_graph_run(c, f->graph, df, dp, f->n, [&]() -> int {
    FG_TRY(fg_cond(f, df, 0));
    for (int i = 0; i < kFgNbSub; i++) FG_TRY(fg_subframe(f, i, dp, nullptr));
    FgParams p = {};
    p.st = f->st; p.n = f->n;
    ob_fargan_end_frame<<<fg_blk(n), 256, 0, c->stream>>>(p);
    c->launches++;
    OB_CUDA(cudaGetLastError());
    return OPUSB200_OK;
  }));
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(pcm, dp, sizeof(float) * n * kFgFrame, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}
