// dred_host.cuh — the DRED decoder network (dnn/dred_rdovae_dec.c:32-145) for a batch of streams: DRED_rdovae_decode_all turns one
// initial state + nb_latents latent vectors per stream into 4*nb_latents feature frames.  The network is DenseNet-shaped: every
// layer reads the leading columns of one growing activation row (576 floats per stream) — kept as one [n][576] matrix in HBM; the
// int8 layers' quantiser reads it with the row stride, outputs land in their column ranges.  Layers run through the batched dnn
// primitives (block-sparse int8 layers: dp4a kernel; the small dense ones: tcgen05).  Included at the end of opus_b200.cu.
#pragma once

namespace ob {
constexpr int kDrFeat = 20, kDrGru = 64, kDrConv = 32, kDrDense1 = 96, kDrBuf = kDrDense1 + 5 * (kDrGru + kDrConv), kDrOut = 80;
enum { DR_DENSE1 = 0, DR_GLU1 = 1, DR_HIDDEN_INIT = 6, DR_OUTPUT = 7, DR_CONV_DENSE1 = 8, DR_GRU_INIT = 13, DR_GRU1_IN = 14, DR_CONV1 = 24, DR_COUNT = 29 };
}
using namespace ob;

// dst[s*dst_stride + j] = src[s*src_stride + j], j < w (column placement in / extraction from the activation rows; the copy
// engine's 2-D copies are slow for rows this short)
extern "C" __global__ void __launch_bounds__(256) ob_dred_copy2d(float *dst, int dst_stride, const float *src, int src_stride, int w, int n) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)n * w) return;
  const int s = (int)(t / w), j = (int)(t - (long)s * w);
  dst[(size_t)s * dst_stride + j] = src[(size_t)s * src_stride + j];
}
static int dr_copy2d(OpusB200Ctx *c, float *dst, int dst_stride, const float *src, int src_stride, int w, int n) {
  ob_dred_copy2d<<<(unsigned)(((long)n * w + 255) / 256), 256, 0, c->stream>>>(dst, dst_stride, src, src_stride, w, n);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}

struct OpusB200DredDec {
  OpusB200Ctx *ctx;
  int n, latent_dim, state_dim;
  OpusB200Linear *l[DR_COUNT];
  float *buf;                 // [n][576] activation rows
  float *gru[5], *conv[5];    // [n][64], [n][32] layer memories
  float *tmp, *tmp2;          // [n][320] scratch
  float *d_state, *d_lat, *d_feat;   // staging for host callers
  size_t lat_cap, feat_cap;
  ObGraphCache graph;         // one decode_all (fixed nb_latents and pointers) as a CUDA graph
};
static void dr_free(OpusB200DredDec *d) {
  if (!d) return;
  if (d->graph.exec) cudaGraphExecDestroy(d->graph.exec);
  for (int i = 0; i < DR_COUNT; i++) opusb200_nnet_linear_destroy(d->l[i]);
  cudaFree(d->buf); cudaFree(d->tmp); cudaFree(d->tmp2); cudaFree(d->d_state); cudaFree(d->d_lat); cudaFree(d->d_feat);
  for (int k = 0; k < 5; k++) { cudaFree(d->gru[k]); cudaFree(d->conv[k]); }
  delete d;
}
// init_rdovaedec: the 29 LinearLayers of the generated RDOVAEDec struct (dred_rdovae_dec_data.h) in declaration order
extern "C" int opusb200_dred_decoder_create(OpusB200Ctx *c, const OpusB200LinearLayer *layers, int n_layers, int su_bias, int n_streams, OpusB200DredDec **out) {
  if (!out) return OPUSB200_BAD_ARG;
  *out = nullptr;
  if (!c || !layers || n_layers != DR_COUNT || n_streams <= 0) return OPUSB200_BAD_ARG;
  // sizes as linear_init checks them; the latent / state widths come from the model
  static const int outs[DR_COUNT] = {96, 64, 64, 64, 64, 64, 128, 80, 32, 32, 32, 32, 32, 320, 192, 192, 192, 192, 192, 192, 192, 192, 192, 192, 32, 32, 32, 32, 32};
  static const int ins[DR_COUNT] = {0, 64, 64, 64, 64, 64, 0, 576, 160, 256, 352, 448, 544, 128, 96, 64, 192, 64, 288, 64, 384, 64, 480, 64, 64, 64, 64, 64, 64};
  for (int i = 0; i < DR_COUNT; i++)
    if (layers[i].nb_outputs != outs[i] || (ins[i] && layers[i].nb_inputs != ins[i])) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  OpusB200DredDec *d = new (std::nothrow) OpusB200DredDec;
  if (!d) return OPUSB200_ALLOC_FAIL;
  memset(d, 0, sizeof(*d));
  d->ctx = c; d->n = n_streams; d->latent_dim = layers[DR_DENSE1].nb_inputs - 1; d->state_dim = layers[DR_HIDDEN_INIT].nb_inputs;
  for (int i = 0; i < DR_COUNT; i++) {
    const int r = opusb200_nnet_linear_create(c, &layers[i], su_bias, &d->l[i]);
    if (r != OPUSB200_OK) { dr_free(d); return r; }
  }
  const size_t n = (size_t)n_streams;
  bool ok = cudaMalloc(&d->buf, sizeof(float) * n * kDrBuf) == cudaSuccess && cudaMalloc(&d->tmp, sizeof(float) * n * 320) == cudaSuccess &&
            cudaMalloc(&d->tmp2, sizeof(float) * n * 320) == cudaSuccess && cudaMalloc(&d->d_state, sizeof(float) * n * d->state_dim) == cudaSuccess;
  for (int k = 0; k < 5 && ok; k++)
    ok = cudaMalloc(&d->gru[k], sizeof(float) * n * kDrGru) == cudaSuccess && cudaMalloc(&d->conv[k], sizeof(float) * n * kDrConv) == cudaSuccess;
  if (!ok) { dr_free(d); return OPUSB200_ALLOC_FAIL; }
  *out = d;
  return OPUSB200_OK;
}
extern "C" void opusb200_dred_decoder_destroy(OpusB200DredDec *d) {
  if (!d) return;
  cudaSetDevice(d->ctx->device);
  cudaStreamSynchronize(d->ctx->stream);
  dr_free(d);
}
#define DR_TRY(x) do { const int r_ = (x); if (r_ != OPUSB200_OK) return r_; } while (0)
// a [n][w] contiguous matrix into columns [col, col+w) of the activation rows
static int dr_put(OpusB200DredDec *d, int col, const float *src, int w) {
  return dr_copy2d(d->ctx, d->buf + col, kDrBuf, src, w, w, d->n);
}
// DRED_rdovae_decode_all (dnn/dred_rdovae_dec.c:44-61): state [n][state_dim], latents [n][nb_latents][latent_dim+1]
// -> features [n][nb_latents][4*20] (each quadruple frame in the reference's order)
extern "C" int opusb200_dred_decode_all(OpusB200DredDec *d, float *features, const float *state, const float *latents, int nb_latents, int on_device) {
  if (!d || !features || !state || !latents || nb_latents <= 0) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = d->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const int n = d->n, LW = d->latent_dim + 1;
  const size_t lb = sizeof(float) * (size_t)n * nb_latents * LW, fb = sizeof(float) * (size_t)n * nb_latents * kDrOut;
  const float *dst = state, *dl = latents;
  float *df = features;
  if (!on_device) {
    if (d->lat_cap < lb) { cudaFree(d->d_lat); d->d_lat = nullptr; d->lat_cap = 0; OB_CUDA(cudaMalloc(&d->d_lat, lb)); d->lat_cap = lb; }
    if (d->feat_cap < fb) { cudaFree(d->d_feat); d->d_feat = nullptr; d->feat_cap = 0; OB_CUDA(cudaMalloc(&d->d_feat, fb)); d->feat_cap = fb; }
    OB_CUDA(cudaMemcpyAsync(d->d_state, state, sizeof(float) * (size_t)n * d->state_dim, cudaMemcpyHostToDevice, c->stream));
    OB_CUDA(cudaMemcpyAsync(d->d_lat, latents, lb, cudaMemcpyHostToDevice, c->stream));
    dst = d->d_state; dl = d->d_lat; df = d->d_feat;
  }
  // the time loop is launch-bound (tiny layers): replayed as a CUDA graph from the third call with the same pointers and nb_latents
  DR_TRY(ob_graph_run(c, d->graph, dl, df, (long)nb_latents, [&]() -> int {
    // dred_rdovae_dec_init_states (:63-87)
    DR_TRY(opusb200_nnet_compute_dense(c, d->l[DR_HIDDEN_INIT], d->tmp, dst, n, NN_TANH, 1));
    DR_TRY(opusb200_nnet_compute_dense(c, d->l[DR_GRU_INIT], d->tmp2, d->tmp, n, NN_TANH, 1));
    for (int k = 0; k < 5; k++)
      DR_TRY(dr_copy2d(c, d->gru[k], kDrGru, d->tmp2 + k * kDrGru, 5 * kDrGru, kDrGru, n));
    for (int k = 0; k < 5; k++) OB_CUDA(cudaMemsetAsync(d->conv[k], 0, sizeof(float) * (size_t)n * kDrConv, c->stream));   // conv1_cond_init on the first frame
    for (int q = 0; q < nb_latents; q++) {   // dred_rdovae_decode_qframe (:90-145)
      // this step's latent vectors, contiguous
      DR_TRY(dr_copy2d(c, d->tmp, LW, dl + (size_t)q * LW, nb_latents * LW, LW, n));
      DR_TRY(opusb200_nnet_compute_dense(c, d->l[DR_DENSE1], d->tmp2, d->tmp, n, NN_TANH, 1));
      DR_TRY(dr_put(d, 0, d->tmp2, kDrDense1));
      int col = kDrDense1;
      for (int k = 0; k < 5; k++) {
        // compute_generic_gru on the leading `col` columns of the activation rows (dnn/nnet.c:76-105)
        const OpusB200Linear *wi = d->l[DR_GRU1_IN + 2 * k], *wr = d->l[DR_GRU1_IN + 2 * k + 1];
        float *zrh;
        if (nnet_scratch(c, 1, sizeof(float) * (size_t)n * 6 * kDrGru, (void **)&zrh)) return (int)OPUSB200_ALLOC_FAIL;
        float *recur = zrh + (size_t)n * 3 * kDrGru;
        DR_TRY(nnet_linear_dev(c, wi, zrh, d->buf, n, NN_LINEAR, nullptr, nullptr, kDrBuf));
        DR_TRY(nnet_linear_dev(c, wr, recur, d->gru[k], n));
        NnetGruParams g;
        g.state = d->gru[k]; g.zrh = zrh; g.recur = recur; g.n = n; g.N = kDrGru;
        ob_nnet_gru<<<nblk((long)n * kDrGru, 256), 256, 0, c->stream>>>(g);
        c->launches++;
        DR_TRY(opusb200_nnet_compute_glu(c, d->l[DR_GLU1 + k], d->tmp, d->gru[k], n, 1));
        DR_TRY(dr_put(d, col, d->tmp, kDrGru));
        col += kDrGru;
        // conv_dense on the leading columns + tanh, then the 2-tap conv1d + tanh
        DR_TRY(nnet_linear_dev(c, d->l[DR_CONV_DENSE1 + k], d->tmp, d->buf, n, NN_LINEAR, nullptr, nullptr, kDrBuf));
        DR_TRY(nnet_act_dev(c, d->tmp, d->tmp, n, kDrConv, NN_TANH));
        DR_TRY(opusb200_nnet_compute_conv1d(c, d->l[DR_CONV1 + k], d->tmp2, d->conv[k], d->tmp, n, kDrConv, 1, NN_TANH, 1));
        DR_TRY(dr_put(d, col, d->tmp2, kDrConv));
        col += kDrConv;
      }
      DR_TRY(nnet_linear_dev(c, d->l[DR_OUTPUT], d->tmp, d->buf, n, NN_LINEAR, nullptr, nullptr, kDrBuf));
      DR_TRY(dr_copy2d(c, df + (size_t)q * kDrOut, nb_latents * kDrOut, d->tmp, kDrOut, kDrOut, n));
    }
    return OPUSB200_OK;
  }, dst));   // the full state pointer is part of the key (not a residue of it)
  OB_CUDA(cudaGetLastError());
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(features, df, fb, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}

// ---- the DRED ENCODER network (dnn/dred_rdovae_enc.c:47-116): dred_rdovae_encode_dframe for a batch of streams in lockstep.  Same DenseNet shape as
// the decoder: one [n][544] activation matrix (dense1 64 + 5 x (GRU 32 + conv 64)), every layer reads its leading columns; conv1 has dilation 1,
// conv2..5 dilation 2.  One call = one double feature frame (2 x 20 features) per stream -> 25 latents + the 50-value initial state. ----
namespace ob {
constexpr int kDeGru = 32, kDeConv = 64, kDeDense1 = 64, kDeBuf = kDeDense1 + 5 * (kDeGru + kDeConv), kDeLatPad = 32, kDeStatePad = 56;
enum { DE_DENSE1 = 0, DE_ZDENSE = 1, DE_GDENSE1 = 2, DE_GDENSE2 = 3, DE_CONV_DENSE1 = 4, DE_GRU1_IN = 9, DE_CONV1 = 19, DE_COUNT = 24 };
}
constexpr int kDredMaxFrames = 104, kDredRedundancy = 52, kDredSilkDelay = 79 + 12 - 80;   // DRED_MAX_FRAMES, DRED_NUM_REDUNDANCY_FRAMES, DRED_SILK_ENCODER_DELAY
struct OpusB200DredEnc {
  OpusB200Ctx *ctx;
  int n, latent_dim, state_dim, initialized;
  OpusB200Linear *l[DE_COUNT];
  float *buf;                 // [n][544] activation rows
  float *gru[5], *conv[5];    // [n][32], [n][64 * dilation] layer memories
  float *tmp, *tmp2;          // [n][128] scratch
  float *d_in, *d_lat, *d_state;   // staging for host callers
  // front end (DREDEnc of dnn/dred_encoder.h:40-60): lockstep bookkeeping on the host, per-stream buffers on the device
  OpusB200LpcnetEnc *lpcnet;
  int Fs, channels, input_buffer_fill, dred_offset, latent_offset, latents_buffer_fill, head;
  float *input_buffer, *resample_mem, *lat_ring, *state_ring, *feat_a, *feat_b, *in40, *d_pcm;
  size_t d_pcm_elems;
  // entropy coding (dred_coding_host.cuh)
  uint8_t *stats, *activity, *d_buf;
  int32_t *last_extra, *d_nbytes;
  size_t d_buf_bytes;
};
static void de_free(OpusB200DredEnc *d) {
  if (!d) return;
  for (int i = 0; i < DE_COUNT; i++) opusb200_nnet_linear_destroy(d->l[i]);
  cudaFree(d->buf); cudaFree(d->tmp); cudaFree(d->tmp2); cudaFree(d->d_in); cudaFree(d->d_lat); cudaFree(d->d_state);
  opusb200_lpcnet_encoder_destroy(d->lpcnet);
  cudaFree(d->input_buffer); cudaFree(d->resample_mem); cudaFree(d->lat_ring); cudaFree(d->state_ring); cudaFree(d->feat_a); cudaFree(d->feat_b); cudaFree(d->in40);
  cudaFree(d->d_pcm); cudaFree(d->stats); cudaFree(d->activity); cudaFree(d->d_buf); cudaFree(d->last_extra); cudaFree(d->d_nbytes);
  for (int k = 0; k < 5; k++) { cudaFree(d->gru[k]); cudaFree(d->conv[k]); }
  delete d;
}
// init_rdovaeenc: the 24 LinearLayers of the generated RDOVAEEnc struct (dred_rdovae_enc_data.h) in declaration order: enc_dense1, enc_zdense,
// gdense1, gdense2, enc_conv_dense1..5, enc_gru1..5 (input, recurrent), enc_conv1..5; latent_dim / state_dim: DRED_LATENT_DIM / DRED_STATE_DIM
// (the z / state layers are padded to a multiple of 8 outputs, the call returns the leading latent_dim / state_dim values like the reference)
extern "C" int opusb200_dred_encoder_create(OpusB200Ctx *c, const OpusB200LinearLayer *layers, int n_layers, int su_bias, int latent_dim, int state_dim,
                                            int n_streams, OpusB200DredEnc **out) {
  if (!out) return OPUSB200_BAD_ARG;
  *out = nullptr;
  if (!c || !layers || n_layers != DE_COUNT || n_streams <= 0 || latent_dim <= 0 || latent_dim > kDeLatPad || state_dim <= 0 || state_dim > kDeStatePad)
    return OPUSB200_BAD_ARG;
  static const int ins[DE_COUNT] = {40, 544, 544, 128, 96, 192, 288, 384, 480, 64, 32, 160, 32, 256, 32, 352, 32, 448, 32, 128, 128, 128, 128, 128};
  static const int outs[DE_COUNT] = {64, kDeLatPad, 128, kDeStatePad, 64, 64, 64, 64, 64, 96, 96, 96, 96, 96, 96, 96, 96, 96, 96, 64, 64, 64, 64, 64};
  for (int i = 0; i < DE_COUNT; i++)
    if (layers[i].nb_inputs != ins[i] || layers[i].nb_outputs != outs[i]) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  OpusB200DredEnc *d = new (std::nothrow) OpusB200DredEnc;
  if (!d) return OPUSB200_ALLOC_FAIL;
  memset(d, 0, sizeof(*d));
  d->ctx = c; d->n = n_streams; d->latent_dim = latent_dim; d->state_dim = state_dim;
  for (int i = 0; i < DE_COUNT; i++) {
    const int r = opusb200_nnet_linear_create(c, &layers[i], su_bias, &d->l[i]);
    if (r != OPUSB200_OK) { de_free(d); return r; }
  }
  const size_t n = (size_t)n_streams;
  bool ok = cudaMalloc(&d->buf, sizeof(float) * n * kDeBuf) == cudaSuccess && cudaMalloc(&d->tmp, sizeof(float) * n * 128) == cudaSuccess &&
            cudaMalloc(&d->tmp2, sizeof(float) * n * 128) == cudaSuccess && cudaMalloc(&d->d_in, sizeof(float) * n * 40) == cudaSuccess &&
            cudaMalloc(&d->d_lat, sizeof(float) * n * kDeLatPad) == cudaSuccess && cudaMalloc(&d->d_state, sizeof(float) * n * kDeStatePad) == cudaSuccess;
  for (int k = 0; k < 5 && ok; k++)
    ok = cudaMalloc(&d->gru[k], sizeof(float) * n * kDeGru) == cudaSuccess && cudaMalloc(&d->conv[k], sizeof(float) * n * kDeConv * 2) == cudaSuccess;
  if (!ok) { de_free(d); return OPUSB200_ALLOC_FAIL; }
  for (int k = 0; k < 5; k++) cudaMemsetAsync(d->gru[k], 0, sizeof(float) * n * kDeGru, c->stream);
  OB_CUDA(cudaStreamSynchronize(c->stream));
  *out = d;
  return OPUSB200_OK;
}
extern "C" void opusb200_dred_encoder_destroy(OpusB200DredEnc *d) {
  if (!d) return;
  cudaSetDevice(d->ctx->device);
  cudaStreamSynchronize(d->ctx->stream);
  de_free(d);
}
extern "C" int opusb200_dred_encoder_reset(OpusB200DredEnc *d) {   // a zeroed RDOVAEEncState
  if (!d) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(d->ctx->device));
  for (int k = 0; k < 5; k++) OB_CUDA(cudaMemsetAsync(d->gru[k], 0, sizeof(float) * (size_t)d->n * kDeGru, d->ctx->stream));
  d->initialized = 0;
  if (d->lpcnet) {   // dred_encoder_reset (dnn/dred_encoder.c:70-78): buffers cleared, input_buffer_fill = DRED_SILK_ENCODER_DELAY, the feature extractor re-initialised
    const size_t n = (size_t)d->n;
    OB_CUDA(cudaMemsetAsync(d->input_buffer, 0, sizeof(float) * n * 640, d->ctx->stream));
    OB_CUDA(cudaMemsetAsync(d->resample_mem, 0, sizeof(float) * n * 9, d->ctx->stream));
    OB_CUDA(cudaMemsetAsync(d->lat_ring, 0, sizeof(float) * n * kDredMaxFrames * d->latent_dim, d->ctx->stream));
    OB_CUDA(cudaMemsetAsync(d->state_ring, 0, sizeof(float) * n * kDredMaxFrames * d->state_dim, d->ctx->stream));
    DR_TRY(opusb200_lpcnet_encoder_reset(d->lpcnet));
    d->input_buffer_fill = kDredSilkDelay; d->dred_offset = 0; d->latent_offset = 0; d->latents_buffer_fill = 0; d->head = 0;
    if (d->last_extra) OB_CUDA(cudaMemsetAsync(d->last_extra, 0, sizeof(int32_t) * n, d->ctx->stream));
  }
  return OPUSB200_OK;
}
// dred_rdovae_encode_dframe for every stream: input [n][40] -> latents [n][latent_dim], initial_state [n][state_dim]
static int de_encode_dev(OpusB200DredEnc *d, float *d_latents, int lat_stride, float *d_state, int state_stride, const float *din) {
  OpusB200Ctx *c = d->ctx;
  const int n = d->n;
  DR_TRY(opusb200_nnet_compute_dense(c, d->l[DE_DENSE1], d->tmp, din, n, NN_TANH, 1));
  DR_TRY(dr_copy2d(c, d->buf, kDeBuf, d->tmp, kDeDense1, kDeDense1, n));
  int col = kDeDense1;
  for (int k = 0; k < 5; k++) {
    const int dilation = k == 0 ? 1 : 2;
    const OpusB200Linear *wi = d->l[DE_GRU1_IN + 2 * k], *wr = d->l[DE_GRU1_IN + 2 * k + 1];
    float *zrh;
    if (nnet_scratch(c, 1, sizeof(float) * (size_t)n * 6 * kDeGru, (void **)&zrh)) return (int)OPUSB200_ALLOC_FAIL;
    float *recur = zrh + (size_t)n * 3 * kDeGru;
    DR_TRY(nnet_linear_dev(c, wi, zrh, d->buf, n, NN_LINEAR, nullptr, nullptr, kDeBuf));
    DR_TRY(nnet_linear_dev(c, wr, recur, d->gru[k], n));
    NnetGruParams g;
    g.state = d->gru[k]; g.zrh = zrh; g.recur = recur; g.n = n; g.N = kDeGru;
    ob_nnet_gru<<<nblk((long)n * kDeGru, 256), 256, 0, c->stream>>>(g);
    c->launches++;
    DR_TRY(dr_copy2d(c, d->buf + col, kDeBuf, d->gru[k], kDeGru, kDeGru, n));
    col += kDeGru;
    if (!d->initialized) OB_CUDA(cudaMemsetAsync(d->conv[k], 0, sizeof(float) * (size_t)n * kDeConv * dilation, c->stream));   // conv1_cond_init (:39-45)
    DR_TRY(nnet_linear_dev(c, d->l[DE_CONV_DENSE1 + k], d->tmp, d->buf, n, NN_LINEAR, nullptr, nullptr, kDeBuf));
    DR_TRY(nnet_act_dev(c, d->tmp, d->tmp, n, kDeConv, NN_TANH));
    DR_TRY(opusb200_nnet_compute_conv1d(c, d->l[DE_CONV1 + k], d->tmp2, d->conv[k], d->tmp, n, kDeConv, dilation, NN_TANH, 1));
    DR_TRY(dr_copy2d(c, d->buf + col, kDeBuf, d->tmp2, kDeConv, kDeConv, n));
    col += kDeConv;
  }
  DR_TRY(nnet_linear_dev(c, d->l[DE_ZDENSE], d->tmp, d->buf, n, NN_LINEAR, nullptr, nullptr, kDeBuf));
  DR_TRY(dr_copy2d(c, d_latents, lat_stride, d->tmp, kDeLatPad, d->latent_dim, n));
  DR_TRY(nnet_linear_dev(c, d->l[DE_GDENSE1], d->tmp, d->buf, n, NN_LINEAR, nullptr, nullptr, kDeBuf));
  DR_TRY(nnet_act_dev(c, d->tmp, d->tmp, n, 128, NN_TANH));
  DR_TRY(opusb200_nnet_compute_dense(c, d->l[DE_GDENSE2], d->tmp2, d->tmp, n, NN_LINEAR, 1));
  DR_TRY(dr_copy2d(c, d_state, state_stride, d->tmp2, kDeStatePad, d->state_dim, n));
  d->initialized = 1;
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
extern "C" int opusb200_dred_encode_dframe(OpusB200DredEnc *d, float *latents, float *initial_state, const float *input, int on_device) {
  if (!d || !latents || !initial_state || !input) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = d->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const int n = d->n;
  if (on_device) return de_encode_dev(d, latents, d->latent_dim, initial_state, d->state_dim, input);
  OB_CUDA(cudaMemcpyAsync(d->d_in, input, sizeof(float) * (size_t)n * 40, cudaMemcpyHostToDevice, c->stream));
  DR_TRY(de_encode_dev(d, d->d_lat, d->latent_dim, d->d_state, d->state_dim, d->d_in));
  OB_CUDA(cudaMemcpyAsync(latents, d->d_lat, sizeof(float) * (size_t)n * d->latent_dim, cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaMemcpyAsync(initial_state, d->d_state, sizeof(float) * (size_t)n * d->state_dim, cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}

// ---- the front end of the DRED encoder: dred_encoder_init + dred_compute_latents (dnn/dred_encoder.c:80-82, 215-250) ----
extern "C" __global__ void __launch_bounds__(128) ob_dred_convert_16k(ob::DredConvParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::dred_convert_thread(p, s);
}
// dred_encoder_init(Fs, channels) + the pitch-model half of dred_encoder_load_model: linear / conv = the PitchDNN layers (as opusb200_lpcnet_encoder_create)
extern "C" int opusb200_dred_encoder_init_frontend(OpusB200DredEnc *d, int Fs, int channels, const OpusB200LinearLayer *linear, int n_linear,
                                                   const OpusB200Conv2dLayer *conv, int n_conv, int su_bias) {
  if (!d || (Fs != 8000 && Fs != 12000 && Fs != 16000 && Fs != 24000 && Fs != 48000) || (channels != 1 && channels != 2)) return OPUSB200_BAD_ARG;
  if (d->lpcnet) return OPUSB200_INVALID_STATE;
  OpusB200Ctx *c = d->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  OpusB200LpcnetEnc *e = nullptr;
  DR_TRY(opusb200_lpcnet_encoder_create(c, linear, n_linear, conv, n_conv, su_bias, d->n, &e));
  const size_t n = (size_t)d->n;
  struct { float **p; size_t w; } al[] = {{&d->input_buffer, 640}, {&d->resample_mem, 9}, {&d->lat_ring, (size_t)kDredMaxFrames * d->latent_dim},
                                          {&d->state_ring, (size_t)kDredMaxFrames * d->state_dim}, {&d->feat_a, 36}, {&d->feat_b, 36}, {&d->in40, 40}};
  for (auto &a : al)
    if (cudaMalloc(a.p, sizeof(float) * n * a.w) != cudaSuccess) {
      opusb200_lpcnet_encoder_destroy(e);
      for (auto &b : al) { cudaFree(*b.p); *b.p = nullptr; }
      return OPUSB200_ALLOC_FAIL;
    }
  d->lpcnet = e; d->Fs = Fs; d->channels = channels;
  DR_TRY(opusb200_dred_encoder_reset(d));
  OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}
// dred_process_frame (:84-107): two feature frames of the 320 buffered samples, one latent / state pair at the front of the ring buffers
static int de_process_frame(OpusB200DredEnc *d) {
  OpusB200Ctx *c = d->ctx;
  const size_t n = (size_t)d->n;
  DR_TRY(lpcnet_enc_frame(d->lpcnet, d->input_buffer, 0, d->feat_a, 640));
  DR_TRY(lpcnet_enc_frame(d->lpcnet, d->input_buffer + 160, 0, d->feat_b, 640));
  OB_CUDA(cudaMemcpy2DAsync(d->in40, sizeof(float) * 40, d->feat_a, sizeof(float) * 36, sizeof(float) * 20, n, cudaMemcpyDeviceToDevice, c->stream));
  OB_CUDA(cudaMemcpy2DAsync(d->in40 + 20, sizeof(float) * 40, d->feat_b, sizeof(float) * 36, sizeof(float) * 20, n, cudaMemcpyDeviceToDevice, c->stream));
  d->head = (d->head + kDredMaxFrames - 1) % kDredMaxFrames;   // the reference moves the buffers up by one entry; here the head moves down
  DR_TRY(de_encode_dev(d, d->lat_ring + (size_t)d->head * d->latent_dim, kDredMaxFrames * d->latent_dim, d->state_ring + (size_t)d->head * d->state_dim,
                       kDredMaxFrames * d->state_dim, d->in40));
  d->latents_buffer_fill = d->latents_buffer_fill + 1 < kDredRedundancy ? d->latents_buffer_fill + 1 : kDredRedundancy;
  return OPUSB200_OK;
}
// dred_compute_latents for every stream: pcm [n][frame_size * channels] float in [-1, 1) at Fs
extern "C" int opusb200_dred_compute_latents(OpusB200DredEnc *d, const float *pcm, int frame_size, int extra_delay, int on_device) {
  if (!d || !pcm || frame_size <= 0) return OPUSB200_BAD_ARG;
  if (!d->lpcnet) return OPUSB200_INVALID_STATE;
  if ((long)frame_size * 16000 % d->Fs) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = d->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)d->n, row = (size_t)frame_size * d->channels;
  const float *dp = pcm;
  if (!on_device) {
    if (d->d_pcm_elems < n * row) {
      cudaFree(d->d_pcm); d->d_pcm = nullptr; d->d_pcm_elems = 0;
      if (cudaMalloc(&d->d_pcm, sizeof(float) * n * row) != cudaSuccess) return OPUSB200_ALLOC_FAIL;
      d->d_pcm_elems = n * row;
    }
    OB_CUDA(cudaMemcpyAsync(d->d_pcm, pcm, sizeof(float) * n * row, cudaMemcpyHostToDevice, c->stream));
    dp = d->d_pcm;
  }
  int frame_size16k = (int)((long)frame_size * 16000 / d->Fs), off = 0;
  int curr_offset16k = 40 + extra_delay * 16000 / d->Fs - d->input_buffer_fill;
  d->dred_offset = (int)floor((curr_offset16k + 20.f) / 40.f);
  d->latent_offset = 0;
  while (frame_size16k > 0) {
    const int process_size16k = 320 < frame_size16k ? 320 : frame_size16k, process_size = (int)((long)process_size16k * d->Fs / 16000);
    ob::DredConvParams p;
    p.pcm = dp; p.input_buffer = d->input_buffer; p.resample_mem = d->resample_mem; p.n = d->n; p.Fs = d->Fs; p.channels = d->channels;
    p.in_len = process_size; p.out_len = process_size16k; p.fill = d->input_buffer_fill; p.pcm_off = off; p.pcm_stride = (int)row;
    ob_dred_convert_16k<<<(d->n + 127) / 128, 128, 0, c->stream>>>(p);
    c->launches++;
    OB_CUDA(cudaGetLastError());
    d->input_buffer_fill += process_size16k;
    if (d->input_buffer_fill >= 320) {
      curr_offset16k += 320;
      DR_TRY(de_process_frame(d));
      d->input_buffer_fill -= 320;
      if (d->input_buffer_fill > 0)
        OB_CUDA(cudaMemcpy2DAsync(d->input_buffer, sizeof(float) * 640, d->input_buffer + 320, sizeof(float) * 640, sizeof(float) * d->input_buffer_fill, n,
                                  cudaMemcpyDeviceToDevice, c->stream));
      if (d->dred_offset < 6) d->dred_offset += 8;
      else d->latent_offset++;
    }
    off += process_size * d->channels;
    frame_size16k -= process_size16k;
  }
  if (!on_device) OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}
// the encoder's buffers in the reference's order (index 0 = the newest double frame): latents [n][104 * latent_dim], state [n][104 * state_dim] (host),
// info [4] = input_buffer_fill, dred_offset, latent_offset, latents_buffer_fill
extern "C" int opusb200_dred_encoder_get_latents(OpusB200DredEnc *d, float *latents, float *state, int32_t *info) {
  if (!d || !latents || !state || !info) return OPUSB200_BAD_ARG;
  if (!d->lpcnet) return OPUSB200_INVALID_STATE;
  OpusB200Ctx *c = d->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)d->n;
  struct { float *dst; const float *src; int w; } bufs[2] = {{latents, d->lat_ring, d->latent_dim}, {state, d->state_ring, d->state_dim}};
  for (auto &b : bufs) {
    const size_t rowb = sizeof(float) * kDredMaxFrames * b.w, first = sizeof(float) * (size_t)(kDredMaxFrames - d->head) * b.w;
    OB_CUDA(cudaMemcpy2DAsync(b.dst, rowb, b.src + (size_t)d->head * b.w, rowb, first, n, cudaMemcpyDeviceToHost, c->stream));
    if (d->head) OB_CUDA(cudaMemcpy2DAsync((char *)b.dst + first, rowb, b.src, rowb, rowb - first, n, cudaMemcpyDeviceToHost, c->stream));
  }
  OB_CUDA(cudaStreamSynchronize(c->stream));
  info[0] = d->input_buffer_fill; info[1] = d->dred_offset; info[2] = d->latent_offset; info[3] = d->latents_buffer_fill;
  return OPUSB200_OK;
}
