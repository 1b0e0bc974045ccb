// plc_host.cuh — the neural packet-loss concealment of the reference's decoder, batched: dnn/lpcnet_plc.c:44-207 (lpcnet_plc_reset, lpcnet_plc_fec_add /
// _clear, compute_plc_pred, get_fec_or_pred, queue_features, lpcnet_plc_update, lpcnet_plc_conceal) for n streams that move through update /
// conceal calls in lockstep (the recovery decode of BASELINE config 5: every stream of the batch conceals frame k at the same time), while what
// a stream has in its FEC / DRED feature queue is its own.  The predictor (dense 57 -> 128 tanh, two 128-unit int8 GRUs, dense 128 -> 20) runs on
// the dnn/ layer kernels of nnet_host.cuh (int8 GRU matrices on the tensor cores, exact integers), the vocoder is the FARGAN object of
// fargan_host.cuh; the glue kernels below are the per-stream bookkeeping.  Included at the end of opus_b200.cu after fargan_host.cuh.
// The history analysis a first concealed frame runs when good audio came before it (:160-177: burg_cepstral_analysis,
// lpcnet_compute_single_frame_features_float and the predictor over the decoded frames not yet analysed) uses the feature extractor of
// lpcnet_enc_host.cuh; it needs the pitch model (opusb200_lpcnet_plc_load_pitch_model), without it such a call answers OPUSB200_INVALID_STATE.
#pragma once

namespace ob {
constexpr int kPlcFeat = 20, kPlcBands = 18, kPlcIn = 2 * kPlcBands + kPlcFeat + 1, kPlcMaxFec = 104, kPlcContVec = 5, kPlcFrame = 160;
constexpr int kPlcBuf = (kPlcContVec + 10) * kPlcFrame, kPlcGru = 128;
struct PlcParams {
  float *fec;             // [n][104][20]
  int32_t *cnt;           // [n][4]: fec_read_pos, fec_fill_pos, fec_skip, loss_count
  int32_t *used_fec;      // [n]: this step took its features from the queue
  float *in57, *out20, *features, *cont_features, *gru1, *gru2, *bak, *hist;
  const float *add_features;   // fec_add: [n][20]
  const int32_t *add_have;     // fec_add: [n] 1 = features, 0 = NULL (a gap), < 0 = no call for this stream; null = features for all
  const float *fpcm;           // finish: the vocoder's frame [n][160]
  int16_t *pcm;                // finish: [n][160]
  const int16_t *upd_pcm;      // update: [n][160]
  float *xf;                   // history analysis: the frame at analysis_pos, int16 range [n][160]
  const float *feat36, *ceps36;   // history analysis: this frame's features [n][36] and Burg cepstra [n][36]
  int n, final_step, restore, analysis_pos, predict;
};
}  // namespace ob

// lpcnet_plc_fec_add (:88-96)
extern "C" __global__ void ob_plc_fec_add(ob::PlcParams p) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= p.n) return;
  const int have = p.add_have ? p.add_have[s] : 1;
  if (have < 0) return;
  int32_t *c = p.cnt + 4 * (size_t)s;
  if (have == 0) { if (lane == 0) c[2]++; return; }
  const int fill = c[1];
  if (fill >= ob::kPlcMaxFec) return;   // celt_assert(st->fec_fill_pos < PLC_MAX_FEC)
  if (lane < ob::kPlcFeat) p.fec[((size_t)s * ob::kPlcMaxFec + fill) * ob::kPlcFeat + lane] = p.add_features[(size_t)s * ob::kPlcFeat + lane];
  __syncwarp();
  if (lane == 0) c[1] = fill + 1;
}
// plc_bak[0] = plc_bak[1]; plc_bak[1] = plc_net (:178-179 ...), after plc_net = plc_bak[0] when `restore` (:159)
extern "C" __global__ void ob_plc_bak(ob::PlcParams p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)p.n * 2 * ob::kPlcGru) return;
  const size_t s = i / (2 * ob::kPlcGru), k = i % (2 * ob::kPlcGru);
  float *net = (k < ob::kPlcGru ? p.gru1 + s * ob::kPlcGru : p.gru2 + s * ob::kPlcGru - ob::kPlcGru) + k;
  float *b0 = p.bak + s * 4 * ob::kPlcGru + k, *b1 = b0 + 2 * ob::kPlcGru;
  float v = *net;
  if (p.restore) { v = *b0; *net = v; }
  if (p.restore == 2) return;   // plc_net = plc_bak[0] alone (:159), the rotations follow inside the history loop
  *b0 = *b1;
  *b1 = v;
}
// history analysis (:163-166): the decoded frame at analysis_pos back in int16 range
extern "C" __global__ void ob_plc_hist_x(ob::PlcParams p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)p.n * ob::kPlcFrame) return;
  const size_t s = i / ob::kPlcFrame, k = i % ob::kPlcFrame;
  p.xf[i] = 32768.f * p.hist[s * ob::kPlcBuf + p.analysis_pos + k];
}
// history analysis (:169-176): features <- this frame's; when the frame also feeds the predictor: queue_features and plc_features = cepstra, features, 1
extern "C" __global__ void ob_plc_hist_glue(ob::PlcParams p) {
  const int s = blockIdx.x * (blockDim.x >> 6) + (threadIdx.x >> 6), k = threadIdx.x & 63;
  if (s >= p.n) return;
  const float *f36 = p.feat36 + (size_t)s * 36;
  if (k < ob::kPlcFeat) p.features[(size_t)s * ob::kPlcFeat + k] = f36[k];
  if (!p.predict) return;
  float *cf = p.cont_features + (size_t)s * ob::kPlcContVec * ob::kPlcFeat;
  float keep[ob::kPlcContVec - 1];
#pragma unroll
  for (int v = 0; v < ob::kPlcContVec - 1; v++) keep[v] = k < ob::kPlcFeat ? cf[(v + 1) * ob::kPlcFeat + k] : 0.f;
  __syncwarp();
  if (k < ob::kPlcFeat) {
#pragma unroll
    for (int v = 0; v < ob::kPlcContVec - 1; v++) cf[v * ob::kPlcFeat + k] = keep[v];
    cf[(ob::kPlcContVec - 1) * ob::kPlcFeat + k] = f36[k];
  }
  if (k < ob::kPlcIn) p.in57[(size_t)s * ob::kPlcIn + k] = k < 2 * ob::kPlcBands ? p.ceps36[(size_t)s * 36 + k] : k < ob::kPlcIn - 1 ? f36[k - 2 * ob::kPlcBands] : 1.f;
}
// get_fec_or_pred, first half (:115-133): the predictor's input — all zeros, or the queued features with the "from FEC" flag
extern "C" __global__ void ob_plc_pred_in(ob::PlcParams p) {
  const int s = blockIdx.x * (blockDim.x >> 6) + (threadIdx.x >> 6), k = threadIdx.x & 63;
  if (s >= p.n || k >= ob::kPlcIn) return;
  const int32_t *c = p.cnt + 4 * (size_t)s;
  const int use = c[0] != c[1] && c[2] == 0;
  float v = 0.f;
  if (use) {
    if (k >= 2 * ob::kPlcBands && k < 2 * ob::kPlcBands + ob::kPlcFeat) v = p.fec[((size_t)s * ob::kPlcMaxFec + c[0]) * ob::kPlcFeat + (k - 2 * ob::kPlcBands)];
    else if (k == ob::kPlcIn - 1) v = -1.f;
  }
  p.in57[(size_t)s * ob::kPlcIn + k] = v;
  if (k == 0) p.used_fec[s] = use;
}
// get_fec_or_pred, second half, and what lpcnet_plc_conceal does with the result: features <- queue or prediction, the counters, on the final step
// of a call the loss count and the energy attenuation (:183-187), then queue_features (:135-138)
extern "C" __global__ void ob_plc_pred_out(ob::PlcParams p) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= p.n) return;
  int32_t *c = p.cnt + 4 * (size_t)s;
  const int use = p.used_fec[s];
  float f = 0.f;
  if (lane < ob::kPlcFeat) f = use ? p.fec[((size_t)s * ob::kPlcMaxFec + c[0]) * ob::kPlcFeat + lane] : p.out20[(size_t)s * ob::kPlcFeat + lane];
  int loss = c[3];
  if (p.final_step) loss = use ? 0 : loss + 1;
  if (p.final_step && lane == 0) {
    const float att[10] = {0.f, 0.f, -.2f, -.2f, -.4f, -.4f, -.8f, -.8f, -1.6f, -1.6f};
    const float a = loss >= 10 ? f + att[9] - 2 * (loss - 9) : f + att[loss];
    f = a > -15.f ? a : -15.f;
  }
  float *cf = p.cont_features + (size_t)s * ob::kPlcContVec * ob::kPlcFeat;
  float keep[ob::kPlcContVec - 1];
#pragma unroll
  for (int v = 0; v < ob::kPlcContVec - 1; v++) keep[v] = lane < ob::kPlcFeat ? cf[(v + 1) * ob::kPlcFeat + lane] : 0.f;
  __syncwarp();
  if (lane < ob::kPlcFeat) {
#pragma unroll
    for (int v = 0; v < ob::kPlcContVec - 1; v++) cf[v * ob::kPlcFeat + lane] = keep[v];
    cf[(ob::kPlcContVec - 1) * ob::kPlcFeat + lane] = f;
    p.features[(size_t)s * ob::kPlcFeat + lane] = f;
  }
  if (lane == 0) {
    if (use) c[0]++;
    else if (c[2] > 0) c[2]--;
    if (p.final_step) c[3] = loss;
  }
}
// fargan_synthesize_int's conversion (dnn/fargan.c:219-225) and the PCM history of lpcnet_plc_conceal (:202-204) / lpcnet_plc_update (:143-153)
extern "C" __global__ void ob_plc_finish(ob::PlcParams p) {
  const int s = blockIdx.x;
  if (s >= p.n) return;
  float *h = p.hist + (size_t)s * ob::kPlcBuf;
  __shared__ float sh[ob::kPlcBuf];
  for (int i = threadIdx.x; i < ob::kPlcBuf - ob::kPlcFrame; i += blockDim.x) sh[i] = h[i + ob::kPlcFrame];
  for (int i = threadIdx.x; i < ob::kPlcFrame; i += blockDim.x) {
    int16_t q;
    if (p.upd_pcm) q = p.upd_pcm[(size_t)s * ob::kPlcFrame + i];
    else {
      float v = 32768.f * p.fpcm[(size_t)s * ob::kPlcFrame + i];
      v = v > -32767.f ? v : -32767.f;
      v = v < 32767.f ? v : 32767.f;
      q = (int16_t)(int)floor(.5 + (double)v);
      p.pcm[(size_t)s * ob::kPlcFrame + i] = q;
    }
    sh[ob::kPlcBuf - ob::kPlcFrame + i] = (1.f / 32768.f) * q;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ob::kPlcBuf; i += blockDim.x) h[i] = sh[i];
  if (p.upd_pcm && threadIdx.x == 0) p.cnt[4 * (size_t)s + 3] = 0;   // lpcnet_plc_update: loss_count = 0
}

struct OpusB200LpcnetPlc {
  OpusB200Ctx *ctx;
  int n;
  OpusB200Fargan *fargan;
  OpusB200Linear *l[6];   // PLCModel order (plc_data.h): dense_in, dense_out, gru1 input / recurrent, gru2 input / recurrent
  float *fec, *in57, *tmp128, *out20, *features, *cont_features, *gru1, *gru2, *bak, *hist, *fpcm, *pcm0, *d_feat, *xf, *feat36, *ceps36;
  OpusB200LpcnetEnc *enc;   // LPCNetPLCState.enc: the feature extractor of the history analysis (null until the pitch model is loaded)
  int32_t *cnt, *used_fec, *d_have;
  int16_t *d_pcm;
  int blend, analysis_pos, predict_pos, analysis_gap;   // lockstep state
};
static void plc_free(OpusB200LpcnetPlc *q) {
  if (!q) return;
  opusb200_fargan_destroy(q->fargan);
  opusb200_lpcnet_encoder_destroy(q->enc);
  for (int i = 0; i < 6; i++) opusb200_nnet_linear_destroy(q->l[i]);
  float **fb[] = {&q->fec, &q->in57, &q->tmp128, &q->out20, &q->features, &q->cont_features, &q->gru1, &q->gru2, &q->bak, &q->hist, &q->fpcm, &q->pcm0, &q->d_feat,
                  &q->xf, &q->feat36, &q->ceps36};
  for (auto b : fb) cudaFree(*b);
  cudaFree(q->cnt); cudaFree(q->used_fec); cudaFree(q->d_have); cudaFree(q->d_pcm);
  delete q;
}
static ob::PlcParams plc_params(OpusB200LpcnetPlc *q) {
  ob::PlcParams p = {};
  p.fec = q->fec; p.cnt = q->cnt; p.used_fec = q->used_fec; p.in57 = q->in57; p.out20 = q->out20; p.features = q->features;
  p.cont_features = q->cont_features; p.gru1 = q->gru1; p.gru2 = q->gru2; p.bak = q->bak; p.hist = q->hist; p.n = q->n;
  return p;
}
// lpcnet_plc_reset (:44-56): everything but the models
extern "C" int opusb200_lpcnet_plc_reset(OpusB200LpcnetPlc *q) {
  if (!q) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = q->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)q->n;
  struct { void *p; size_t b; } z[] = {{q->fec, sizeof(float) * n * ob::kPlcMaxFec * ob::kPlcFeat}, {q->cnt, sizeof(int32_t) * n * 4}, {q->hist, sizeof(float) * n * ob::kPlcBuf},
                                       {q->features, sizeof(float) * n * ob::kPlcFeat}, {q->cont_features, sizeof(float) * n * ob::kPlcContVec * ob::kPlcFeat},
                                       {q->gru1, sizeof(float) * n * ob::kPlcGru}, {q->gru2, sizeof(float) * n * ob::kPlcGru}, {q->bak, sizeof(float) * n * 4 * ob::kPlcGru}};
  for (auto &a : z) OB_CUDA(cudaMemsetAsync(a.p, 0, a.b, c->stream));
  if (q->enc) FG_TRY(opusb200_lpcnet_encoder_reset(q->enc));   // lpcnet_encoder_init(&st->enc), :48
  q->blend = 0; q->analysis_gap = 1; q->analysis_pos = ob::kPlcBuf; q->predict_pos = ob::kPlcBuf;
  return OPUSB200_OK;
}
// lpcnet_plc_init (:58-73) with init_plcmodel's size checks; plc_layers: the 6 LinearLayers of PLCModel, fargan_layers: the 20 of FARGAN
extern "C" int opusb200_lpcnet_plc_create(OpusB200Ctx *c, const OpusB200LinearLayer *plc_layers, int n_plc_layers, const OpusB200LinearLayer *fargan_layers,
                                          int n_fargan_layers, int su_bias, int n_streams, OpusB200LpcnetPlc **out) {
  if (!out) return OPUSB200_BAD_ARG;
  *out = nullptr;
  if (!c || !plc_layers || n_plc_layers != 6 || !fargan_layers || n_streams <= 0) return OPUSB200_BAD_ARG;
  static const int dims[6][2] = {{ob::kPlcIn, 128}, {128, ob::kPlcFeat}, {128, 384}, {128, 384}, {128, 384}, {128, 384}};
  for (int i = 0; i < 6; i++)
    if (plc_layers[i].nb_inputs != dims[i][0] || plc_layers[i].nb_outputs != dims[i][1]) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  OpusB200LpcnetPlc *q = new (std::nothrow) OpusB200LpcnetPlc;
  if (!q) return OPUSB200_ALLOC_FAIL;
  memset(q, 0, sizeof(*q));
  q->ctx = c; q->n = n_streams;
  int r = opusb200_fargan_create(c, fargan_layers, n_fargan_layers, su_bias, n_streams, &q->fargan);
  if (r != OPUSB200_OK) { plc_free(q); return r; }
  for (int i = 0; i < 6; i++)
    if ((r = opusb200_nnet_linear_create(c, &plc_layers[i], su_bias, &q->l[i])) != OPUSB200_OK) { plc_free(q); return r; }
  const size_t n = (size_t)n_streams;
  struct { float **p; size_t w; } al[] = {{&q->fec, (size_t)ob::kPlcMaxFec * ob::kPlcFeat}, {&q->in57, ob::kPlcIn}, {&q->tmp128, 128}, {&q->out20, ob::kPlcFeat}, {&q->features, ob::kPlcFeat},
                                          {&q->cont_features, (size_t)ob::kPlcContVec * ob::kPlcFeat}, {&q->gru1, ob::kPlcGru}, {&q->gru2, ob::kPlcGru}, {&q->bak, 4 * ob::kPlcGru},
                                          {&q->hist, ob::kPlcBuf}, {&q->fpcm, ob::kPlcFrame}, {&q->pcm0, 320}, {&q->d_feat, ob::kPlcFeat},
                                          {&q->xf, ob::kPlcFrame}, {&q->feat36, 36}, {&q->ceps36, 36}};
  for (auto &a : al)
    if (cudaMalloc(a.p, sizeof(float) * n * a.w) != cudaSuccess) { plc_free(q); return OPUSB200_ALLOC_FAIL; }
  if (cudaMalloc(&q->cnt, sizeof(int32_t) * n * 4) != cudaSuccess || cudaMalloc(&q->used_fec, sizeof(int32_t) * n) != cudaSuccess ||
      cudaMalloc(&q->d_have, sizeof(int32_t) * n) != cudaSuccess || cudaMalloc(&q->d_pcm, sizeof(int16_t) * n * ob::kPlcFrame) != cudaSuccess) {
    plc_free(q);
    return OPUSB200_ALLOC_FAIL;
  }
  if ((r = opusb200_lpcnet_plc_reset(q)) != OPUSB200_OK) { plc_free(q); return r; }
  OB_CUDA(cudaStreamSynchronize(c->stream));
  *out = q;
  return OPUSB200_OK;
}
extern "C" void opusb200_lpcnet_plc_destroy(OpusB200LpcnetPlc *q) {
  if (!q) return;
  cudaSetDevice(q->ctx->device);
  cudaStreamSynchronize(q->ctx->stream);
  plc_free(q);
}
// lpcnet_plc_fec_add for every stream at once: features [n][20]; have [n] (or NULL = features for all): 1 = this stream's vector, 0 = the call with
// features == NULL (a gap in the redundancy: fec_skip++), < 0 = no call for this stream
extern "C" int opusb200_lpcnet_plc_fec_add(OpusB200LpcnetPlc *q, const float *features, const int32_t *have, int on_device) {
  if (!q || !features) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = q->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  ob::PlcParams p = plc_params(q);
  p.add_features = features; p.add_have = have;
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(q->d_feat, features, sizeof(float) * (size_t)q->n * ob::kPlcFeat, cudaMemcpyHostToDevice, c->stream));
    p.add_features = q->d_feat;
    if (have) { OB_CUDA(cudaMemcpyAsync(q->d_have, have, sizeof(int32_t) * (size_t)q->n, cudaMemcpyHostToDevice, c->stream)); p.add_have = q->d_have; }
  }
  ob_plc_fec_add<<<(q->n + 7) / 8, 256, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}
extern "C" int opusb200_lpcnet_plc_fec_clear(OpusB200LpcnetPlc *q) {   // :98-100: fec_read_pos = fec_fill_pos = fec_skip = 0 (loss_count stays)
  if (!q) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(q->ctx->device));
  OB_CUDA(cudaMemset2DAsync(q->cnt, sizeof(int32_t) * 4, 0, sizeof(int32_t) * 3, (size_t)q->n, q->ctx->stream));
  return OPUSB200_OK;
}
// lpcnet_plc_update (:143-153): a good frame of every stream, pcm [n][160]
extern "C" int opusb200_lpcnet_plc_update(OpusB200LpcnetPlc *q, const int16_t *pcm, int on_device) {
  if (!q || !pcm) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = q->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  if (q->analysis_pos - ob::kPlcFrame >= 0) q->analysis_pos -= ob::kPlcFrame;
  else q->analysis_gap = 1;
  if (q->predict_pos - ob::kPlcFrame >= 0) q->predict_pos -= ob::kPlcFrame;
  ob::PlcParams p = plc_params(q);
  p.upd_pcm = pcm;
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(q->d_pcm, pcm, sizeof(int16_t) * (size_t)q->n * ob::kPlcFrame, cudaMemcpyHostToDevice, c->stream));
    p.upd_pcm = q->d_pcm;
  }
  ob_plc_finish<<<q->n, 256, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  q->blend = 0;
  if (!on_device) OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}
// one get_fec_or_pred + queue_features round (with the plc_bak rotation in front of it)
static int plc_step(OpusB200LpcnetPlc *q, int restore, int final_step) {
  OpusB200Ctx *c = q->ctx;
  const int n = q->n;
  ob::PlcParams p = plc_params(q);
  p.restore = restore; p.final_step = final_step;
  ob_plc_bak<<<(unsigned)(((size_t)n * 2 * ob::kPlcGru + 255) / 256), 256, 0, c->stream>>>(p);
  ob_plc_pred_in<<<(n + 3) / 4, 256, 0, c->stream>>>(p);
  c->launches += 2;
  FG_TRY(opusb200_nnet_compute_dense(c, q->l[0], q->tmp128, q->in57, n, NN_TANH, 1));       // compute_plc_pred (:103-113)
  FG_TRY(opusb200_nnet_compute_gru(c, q->l[2], q->l[3], q->gru1, q->tmp128, n, 1));
  FG_TRY(opusb200_nnet_compute_gru(c, q->l[4], q->l[5], q->gru2, q->gru1, n, 1));
  FG_TRY(opusb200_nnet_compute_dense(c, q->l[1], q->out20, q->gru2, n, NN_LINEAR, 1));
  ob_plc_pred_out<<<(n + 7) / 8, 256, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
// the pitch-model part of lpcnet_plc_load_model (:75-86: lpcnet_encoder_load_model): linear = the 6 LinearLayers, conv = the 2 Conv2dLayers of PitchDNN
extern "C" int opusb200_lpcnet_plc_load_pitch_model(OpusB200LpcnetPlc *q, const OpusB200LinearLayer *linear, int n_linear, const OpusB200Conv2dLayer *conv, int n_conv,
                                                    int su_bias) {
  if (!q) return OPUSB200_BAD_ARG;
  OpusB200LpcnetEnc *e = nullptr;
  FG_TRY(opusb200_lpcnet_encoder_create(q->ctx, linear, n_linear, conv, n_conv, su_bias, q->n, &e));
  opusb200_lpcnet_encoder_destroy(q->enc);
  q->enc = e;
  return OPUSB200_OK;
}
// lpcnet_plc_conceal (:156-207) for every stream: pcm [n][160] int16 out
extern "C" int opusb200_lpcnet_plc_conceal(OpusB200LpcnetPlc *q, int16_t *pcm, int on_device) {
  if (!q || !pcm) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = q->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)q->n;
  if (q->blend == 0) {
    if (q->analysis_pos + ob::kPlcFrame <= ob::kPlcBuf) {   // good audio came before: analyse what has not been analysed yet (:160-177)
      if (!q->enc) return OPUSB200_INVALID_STATE;             // the pitch model was not loaded
      ob::PlcParams p = plc_params(q);
      p.restore = 2;
      ob_plc_bak<<<(unsigned)((n * 2 * ob::kPlcGru + 255) / 256), 256, 0, c->stream>>>(p);
      c->launches++;
      int count = 0;
      while (q->analysis_pos + ob::kPlcFrame <= ob::kPlcBuf) {
        p = plc_params(q);
        p.xf = q->xf; p.feat36 = q->feat36; p.ceps36 = q->ceps36; p.analysis_pos = q->analysis_pos;
        p.predict = (!q->analysis_gap || count > 0) && q->analysis_pos >= q->predict_pos;
        ob_plc_hist_x<<<(unsigned)((n * ob::kPlcFrame + 255) / 256), 256, 0, c->stream>>>(p);
        c->launches++;
        OB_CUDA(cudaGetLastError());
        if (p.predict) FG_TRY(lpcnet_burg_launch(c, q->ceps36, q->xf, q->n));   // burg_cepstral_analysis: only the predictor reads it
        FG_TRY(lpcnet_enc_frame(q->enc, q->xf, 0, q->feat36));
        ob_plc_hist_glue<<<(q->n + 3) / 4, 256, 0, c->stream>>>(p);
        c->launches++;
        OB_CUDA(cudaGetLastError());
        if (p.predict) {
          p.restore = 0;
          ob_plc_bak<<<(unsigned)((n * 2 * ob::kPlcGru + 255) / 256), 256, 0, c->stream>>>(p);
          c->launches++;
          FG_TRY(opusb200_nnet_compute_dense(c, q->l[0], q->tmp128, q->in57, q->n, NN_TANH, 1));       // compute_plc_pred(st, st->features, plc_features)
          FG_TRY(opusb200_nnet_compute_gru(c, q->l[2], q->l[3], q->gru1, q->tmp128, q->n, 1));
          FG_TRY(opusb200_nnet_compute_gru(c, q->l[4], q->l[5], q->gru2, q->gru1, q->n, 1));
          FG_TRY(opusb200_nnet_compute_dense(c, q->l[1], q->features, q->gru2, q->n, NN_LINEAR, 1));
        }
        q->analysis_pos += ob::kPlcFrame;
        count++;
      }
      FG_TRY(plc_step(q, 0, 0));
    } else {
      FG_TRY(plc_step(q, 1, 0));
    }
    FG_TRY(plc_step(q, 0, 0));
    OB_CUDA(cudaMemcpy2DAsync(q->pcm0, sizeof(float) * 320, q->hist + (ob::kPlcBuf - 320), sizeof(float) * ob::kPlcBuf, sizeof(float) * 320, n, cudaMemcpyDeviceToDevice,
                              c->stream));
    FG_TRY(opusb200_fargan_cont(q->fargan, q->pcm0, q->cont_features, 1));
    q->analysis_gap = 0;
  }
  FG_TRY(plc_step(q, 0, 1));
  FG_TRY(opusb200_fargan_synthesize(q->fargan, q->fpcm, q->features, 1));
  ob::PlcParams p = plc_params(q);
  p.fpcm = q->fpcm; p.pcm = on_device ? pcm : q->d_pcm;
  ob_plc_finish<<<q->n, 256, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (q->analysis_pos - ob::kPlcFrame >= 0) q->analysis_pos -= ob::kPlcFrame;
  else q->analysis_gap = 1;
  q->predict_pos = ob::kPlcBuf;
  q->blend = 1;
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(pcm, q->d_pcm, sizeof(int16_t) * n * ob::kPlcFrame, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}
// parity aid: the predictor state and the feature queue of every stream (features [n][20], cont_features [n][100], gru1 / gru2 [n][128], counters [n][4])
extern "C" int opusb200_lpcnet_plc_get_state(OpusB200LpcnetPlc *q, float *features, float *cont_features, float *gru1, float *gru2, int32_t *counters) {
  if (!q || !features || !cont_features || !gru1 || !gru2 || !counters) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = q->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)q->n;
  OB_CUDA(cudaMemcpyAsync(features, q->features, sizeof(float) * n * ob::kPlcFeat, cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaMemcpyAsync(cont_features, q->cont_features, sizeof(float) * n * ob::kPlcContVec * ob::kPlcFeat, cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaMemcpyAsync(gru1, q->gru1, sizeof(float) * n * ob::kPlcGru, cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaMemcpyAsync(gru2, q->gru2, sizeof(float) * n * ob::kPlcGru, cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaMemcpyAsync(counters, q->cnt, sizeof(int32_t) * n * 4, cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}
