// lpcnet_enc_host.cuh — the batched LPCNet feature extractor: the reference's LPCNetEncState object (dnn/lpcnet_enc.c:49-71, lpcnet_private.h:25-43)
// and lpcnet_compute_single_frame_features[_float] (:217-230) for n streams in lockstep, plus burg_cepstral_analysis (dnn/freq.c:179-190).
// Per frame: ob_lpcnet_enc_analysis (warp per stream, lpcnet_enc.cuh) -> the pitch network of dnn/pitchdnn.c:12-44 on the layer kernels of
// nnet_host.cuh (int8 dense layers and GRU, two float 3x3 conv2d) -> ob_lpcnet_enc_finish (warp per stream).  Included after fargan_host.cuh.
#pragma once
#include "lpcnet_enc.cuh"

extern "C" __global__ void __launch_bounds__(128) ob_lpcnet_enc_analysis(ob::LpcnetEncParams p) {
  const int s = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (s < p.n) ob::lpcnet_enc_analysis_warp(p, s, (ob::LeWS *)ob::dyn_smem() + (threadIdx.x >> 5));
}
extern "C" __global__ void __launch_bounds__(128) ob_lpcnet_enc_finish(ob::LpcnetEncParams p) {
  const int s = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (s < p.n) ob::lpcnet_enc_finish_warp(p, s, (float *)ob::dyn_smem() + (threadIdx.x >> 5) * ob::kLeBuf);
}
extern "C" __global__ void __launch_bounds__(128) ob_lpcnet_burg_cepstrum(ob::LpcnetBurgParams p) {
  const int s = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (s < p.n) ob::lpcnet_burg_cepstrum_warp(p, s, (ob::LeBurgWS *)ob::dyn_smem() + (threadIdx.x >> 5));
}

struct OpusB200LpcnetEnc {
  OpusB200Ctx *ctx;
  int n;
  OpusB200Linear *l[6];    // dense_if_upsampler_1 / _2, dense_downsampler, dense_final_upsampler, gru_1 input / recurrent (pitchdnn_data.h)
  OpusB200Conv2d *cv[2];   // conv2d_1, conv2d_2
  ob::ObLpcnetEncState *state;
  float *if_feat, *if1, *if2, *conv_in, *conv_mid, *conv_out, *down_in, *down_out, *gru, *pitch_out, *mem1, *mem2, *d_feat;
  void *d_pcm;
};
static void lpcnet_enc_free(OpusB200LpcnetEnc *q) {
  if (!q) return;
  for (int i = 0; i < 6; i++) opusb200_nnet_linear_destroy(q->l[i]);
  for (int i = 0; i < 2; i++) opusb200_nnet_conv2d_destroy(q->cv[i]);
  float **fb[] = {&q->if_feat, &q->if1, &q->if2, &q->conv_in, &q->conv_mid, &q->conv_out, &q->down_in, &q->down_out, &q->gru, &q->pitch_out, &q->mem1, &q->mem2, &q->d_feat};
  for (auto b : fb) cudaFree(*b);
  cudaFree(q->state); cudaFree(q->d_pcm);
  delete q;
}
static const size_t kLeWidths[13] = {ob::kLeIf, 64, 64, ob::kLeConvIn, 4 * ob::kLeConvIn, ob::kLeXcorr, ob::kLeXcorr + 64, 64, 64, ob::kLePitchOut,
                                     2 * ob::kLeConvIn, 2 * 4 * ob::kLeConvIn, ob::kLeFeat};
// lpcnet_encoder_init (:53-57): the state cleared, the model kept
extern "C" int opusb200_lpcnet_encoder_reset(OpusB200LpcnetEnc *q) {
  if (!q) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = q->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)q->n;
  float *fb[] = {q->if_feat, q->if1, q->if2, q->conv_in, q->conv_mid, q->conv_out, q->down_in, q->down_out, q->gru, q->pitch_out, q->mem1, q->mem2};
  for (int i = 0; i < 12; i++) OB_CUDA(cudaMemsetAsync(fb[i], 0, sizeof(float) * n * kLeWidths[i], c->stream));
  OB_CUDA(cudaMemsetAsync(q->state, 0, sizeof(ob::ObLpcnetEncState) * n, c->stream));
  return OPUSB200_OK;
}
extern "C" int opusb200_lpcnet_encoder_create(OpusB200Ctx *c, const OpusB200LinearLayer *linear, int n_linear, const OpusB200Conv2dLayer *conv, int n_conv,
                                              int su_bias, int n_streams, OpusB200LpcnetEnc **out) {
  if (!out) return OPUSB200_BAD_ARG;
  *out = nullptr;
  if (!c || !linear || n_linear != 6 || !conv || n_conv != 2 || n_streams <= 0) return OPUSB200_BAD_ARG;
  static const int dims[6][2] = {{ob::kLeIf, 64}, {64, 64}, {ob::kLeXcorr + 64, 64}, {64, ob::kLePitchOut}, {64, 192}, {64, 192}};
  for (int i = 0; i < 6; i++)
    if (linear[i].nb_inputs != dims[i][0] || linear[i].nb_outputs != dims[i][1]) return OPUSB200_BAD_ARG;
  if (conv[0].in_channels != 1 || conv[0].out_channels != 4 || conv[1].in_channels != 4 || conv[1].out_channels != 1) return OPUSB200_BAD_ARG;
  for (int i = 0; i < 2; i++)
    if (conv[i].ktime != 3 || conv[i].kheight != 3) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  OpusB200LpcnetEnc *q = new (std::nothrow) OpusB200LpcnetEnc;
  if (!q) return OPUSB200_ALLOC_FAIL;
  memset(q, 0, sizeof(*q));
  q->ctx = c; q->n = n_streams;
  int r;
  for (int i = 0; i < 6; i++)
    if ((r = opusb200_nnet_linear_create(c, &linear[i], su_bias, &q->l[i])) != OPUSB200_OK) { lpcnet_enc_free(q); return r; }
  for (int i = 0; i < 2; i++)
    if ((r = opusb200_nnet_conv2d_create(c, &conv[i], &q->cv[i])) != OPUSB200_OK) { lpcnet_enc_free(q); return r; }
  const size_t n = (size_t)n_streams;
  float **fb[] = {&q->if_feat, &q->if1, &q->if2, &q->conv_in, &q->conv_mid, &q->conv_out, &q->down_in, &q->down_out, &q->gru, &q->pitch_out, &q->mem1, &q->mem2, &q->d_feat};
  for (int i = 0; i < 13; i++)
    if (cudaMalloc(fb[i], sizeof(float) * n * kLeWidths[i]) != cudaSuccess) { lpcnet_enc_free(q); return OPUSB200_ALLOC_FAIL; }
  if (cudaMalloc(&q->state, sizeof(ob::ObLpcnetEncState) * n) != cudaSuccess || cudaMalloc(&q->d_pcm, sizeof(float) * n * ob::kLeFrame) != cudaSuccess) {
    lpcnet_enc_free(q);
    return OPUSB200_ALLOC_FAIL;
  }
  if ((r = opusb200_lpcnet_encoder_reset(q)) != OPUSB200_OK) { lpcnet_enc_free(q); return r; }
  OB_CUDA(cudaStreamSynchronize(c->stream));
  *out = q;
  return OPUSB200_OK;
}
extern "C" void opusb200_lpcnet_encoder_destroy(OpusB200LpcnetEnc *q) {
  if (!q) return;
  cudaSetDevice(q->ctx->device);
  cudaStreamSynchronize(q->ctx->stream);
  lpcnet_enc_free(q);
}
// one frame of every stream from DEVICE pcm; d_features [n][36] on the device
static int lpcnet_enc_frame(OpusB200LpcnetEnc *q, const void *d_pcm, int pcm_is_int16, float *d_features, int pcm_stride = 0) {
  OpusB200Ctx *c = q->ctx;
  const int n = q->n;
  ob::LpcnetEncParams p = {};
  p.state = q->state; p.pcm = d_pcm; p.if_features = q->if_feat; p.conv_in = q->conv_in; p.pitch_out = q->pitch_out; p.features = d_features;
  p.n = n; p.pcm_is_int16 = pcm_is_int16; p.pcm_stride = pcm_stride;
  const unsigned blocks = (unsigned)((n + 3) / 4);
  ob_lpcnet_enc_analysis<<<blocks, 128, 4 * sizeof(ob::LeWS), c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  // compute_pitchdnn (pitchdnn.c:12-44)
  FG_TRY(opusb200_nnet_compute_dense(c, q->l[0], q->if1, q->if_feat, n, NN_TANH, 1));
  FG_TRY(opusb200_nnet_compute_dense(c, q->l[1], q->if2, q->if1, n, NN_TANH, 1));
  FG_TRY(opusb200_nnet_compute_conv2d(c, q->cv[0], q->conv_mid + 1, q->mem1, q->conv_in, n, ob::kLeXcorr, ob::kLeConvIn, NN_TANH, 1));
  FG_TRY(opusb200_nnet_compute_conv2d(c, q->cv[1], q->conv_out, q->mem2, q->conv_mid, n, ob::kLeXcorr, ob::kLeXcorr, NN_TANH, 1));
  const size_t dw = sizeof(float) * (ob::kLeXcorr + 64);
  OB_CUDA(cudaMemcpy2DAsync(q->down_in, dw, q->conv_out, sizeof(float) * ob::kLeXcorr, sizeof(float) * ob::kLeXcorr, (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
  OB_CUDA(cudaMemcpy2DAsync(q->down_in + ob::kLeXcorr, dw, q->if2, sizeof(float) * 64, sizeof(float) * 64, (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
  FG_TRY(opusb200_nnet_compute_dense(c, q->l[2], q->down_out, q->down_in, n, NN_TANH, 1));
  FG_TRY(opusb200_nnet_compute_gru(c, q->l[4], q->l[5], q->gru, q->down_out, n, 1));
  FG_TRY(opusb200_nnet_compute_dense(c, q->l[3], q->pitch_out, q->gru, n, NN_LINEAR, 1));
  ob_lpcnet_enc_finish<<<blocks, 128, 4 * ob::kLeBuf * sizeof(float), c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
// lpcnet_compute_single_frame_features / _float (:217-230) for every stream: pcm [n][160] (int16 or float, int16 range), features [n][36]
extern "C" int opusb200_lpcnet_compute_single_frame_features(OpusB200LpcnetEnc *q, const void *pcm, int pcm_is_int16, float *features, int on_device) {
  if (!q || !pcm || !features) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = q->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)q->n;
  const void *dp = pcm;
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(q->d_pcm, pcm, (pcm_is_int16 ? sizeof(int16_t) : sizeof(float)) * n * ob::kLeFrame, cudaMemcpyHostToDevice, c->stream));
    dp = q->d_pcm;
  }
  FG_TRY(lpcnet_enc_frame(q, dp, pcm_is_int16, on_device ? features : q->d_feat));
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(features, q->d_feat, sizeof(float) * n * ob::kLeFeat, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}
// parity aid: if_features [n][88], xcorr_features [n][224], lpc [n][16], misc [n][3] = dnn_pitch, pitch_filt, mem_preemph (host buffers)
extern "C" int opusb200_lpcnet_encoder_get_state(OpusB200LpcnetEnc *q, float *if_features, float *xcorr_features, float *lpc, float *misc) {
  if (!q || !if_features || !xcorr_features || !lpc || !misc) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = q->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)q->n;
  std::vector<ob::ObLpcnetEncState> h(n);
  OB_CUDA(cudaMemcpyAsync(if_features, q->if_feat, sizeof(float) * n * ob::kLeIf, cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaMemcpy2DAsync(xcorr_features, sizeof(float) * ob::kLeXcorr, q->conv_in + 1, sizeof(float) * ob::kLeConvIn, sizeof(float) * ob::kLeXcorr, n,
                            cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaMemcpyAsync(h.data(), q->state, sizeof(ob::ObLpcnetEncState) * n, cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaStreamSynchronize(c->stream));
  for (size_t s = 0; s < n; s++) {
    memcpy(lpc + s * ob::kLeOrder, h[s].lpc, sizeof(float) * ob::kLeOrder);
    misc[3 * s] = h[s].dnn_pitch; misc[3 * s + 1] = h[s].pitch_filt; misc[3 * s + 2] = h[s].mem_preemph;
  }
  return OPUSB200_OK;
}
// burg_cepstral_analysis (dnn/freq.c:179-190): x [n][160] float (int16 range), ceps [n][36]
static int lpcnet_burg_launch(OpusB200Ctx *c, float *d_ceps, const float *d_x, int n) {
  ob::LpcnetBurgParams p;
  p.ceps = d_ceps; p.x = d_x; p.n = n;
  ob_lpcnet_burg_cepstrum<<<(unsigned)((n + 3) / 4), 128, 4 * sizeof(ob::LeBurgWS), c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
extern "C" int opusb200_burg_cepstral_analysis(OpusB200Ctx *c, float *ceps, const float *x, int n, int on_device) {
  if (!c || !ceps || !x || n <= 0) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  if (on_device) return lpcnet_burg_launch(c, ceps, x, n);
  float *dx = nullptr, *dc = nullptr;
  if (cudaMalloc(&dx, sizeof(float) * (size_t)n * ob::kLeFrame) != cudaSuccess || cudaMalloc(&dc, sizeof(float) * (size_t)n * 36) != cudaSuccess) {
    cudaFree(dx); cudaFree(dc);
    return OPUSB200_ALLOC_FAIL;
  }
  int r = OPUSB200_OK;
  if (cudaMemcpyAsync(dx, x, sizeof(float) * (size_t)n * ob::kLeFrame, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
  if (r == OPUSB200_OK) r = lpcnet_burg_launch(c, dc, dx, n);
  if (r == OPUSB200_OK && cudaMemcpyAsync(ceps, dc, sizeof(float) * (size_t)n * 36, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
  if (cudaStreamSynchronize(c->stream) != cudaSuccess && r == OPUSB200_OK) r = OPUSB200_INTERNAL_ERROR;
  cudaFree(dx); cudaFree(dc);
  return r;
}
