// dred_coding_host.cuh — the entropy coding of the DRED encoder, one thread per stream: dred_encode_silk_frame (dnn/dred_encoder.c:281-383) with
// dred_encode_latents (:252-270), dred_voice_active (:272-279), compute_quantizer (dnn/dred_coding.c:39-44), ec_laplace_encode_p0
// (celt/laplace.c:136-162) and ec_enc_icdf16 (celt/entenc.c:175-184) on the range encoder of celt_enc_frame.cuh.  Every stream has its own
// activity history, so which latents it sends and how many bytes they take are its own; the buffers it reads are the latent / state rings of
// opusb200_dred_compute_latents (dred_host.cuh).  Integer work + one tanh per coefficient: bit-exact.  Included after celt_enc_host.cuh.
#pragma once

#include "dred_coding.cuh"

extern "C" __global__ void __launch_bounds__(64) ob_dred_encode_silk_frame(ob::DredCodeParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) p.nbytes[s] = ob::dred_encode_silk_frame_thread(p, s);
}

// The statistical model (dred_rdovae_stats_data.h): eight uint8 tables of 16 levels x dim, in the order latent quant_scales / dead_zone / r / p0, then
// the state's four.
extern "C" int opusb200_dred_encoder_load_stats(OpusB200DredEnc *d, const unsigned char *const *tables, int n_tables) {
  if (!d || !tables || n_tables != 8) return OPUSB200_BAD_ARG;
  for (int k = 0; k < 8; k++) if (!tables[k]) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = d->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)d->n, lb = 16 * (size_t)d->latent_dim, sb = 16 * (size_t)d->state_dim;
  if (!d->stats && cudaMalloc(&d->stats, 4 * (lb + sb)) != cudaSuccess) return OPUSB200_ALLOC_FAIL;
  if (!d->activity && (cudaMalloc(&d->activity, n * ob::kDredFrames * 4) != cudaSuccess || cudaMalloc(&d->last_extra, sizeof(int32_t) * n) != cudaSuccess ||
                       cudaMalloc(&d->d_nbytes, sizeof(int32_t) * n) != cudaSuccess))
    return OPUSB200_ALLOC_FAIL;
  for (int k = 0; k < 8; k++)
    OB_CUDA(cudaMemcpyAsync(d->stats + (k < 4 ? k * lb : 4 * lb + (k - 4) * sb), tables[k], k < 4 ? lb : sb, cudaMemcpyHostToDevice, c->stream));
  OB_CUDA(cudaMemsetAsync(d->last_extra, 0, sizeof(int32_t) * n, c->stream));
  OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}
// dred_encode_silk_frame for every stream: buf [n][buf_stride] (max_bytes <= buf_stride), activity_mem [n][416], nbytes [n] = the reference's return
// value per stream (0: this stream sends no DRED with this packet).  q0 / dQ / qmax / max_chunks / max_bytes are the batch's.
extern "C" int opusb200_dred_encode_silk_frame(OpusB200DredEnc *d, unsigned char *buf, int buf_stride, int max_chunks, int max_bytes, int q0, int dQ, int qmax,
                                               const unsigned char *activity_mem, int32_t *nbytes, int on_device) {
  if (!d || !buf || !activity_mem || !nbytes || max_bytes <= 0 || buf_stride < max_bytes || max_chunks < 0 || q0 < 0 || q0 > 15 || dQ < 0 || dQ > 7 || qmax < q0 ||
      qmax > 15)
    return OPUSB200_BAD_ARG;
  if (!d->lpcnet || !d->stats) return OPUSB200_INVALID_STATE;
  OpusB200Ctx *c = d->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)d->n, lb = 16 * (size_t)d->latent_dim, sb = 16 * (size_t)d->state_dim;
  ob::DredCodeParams p;
  p.lat_ring = d->lat_ring; p.state_ring = d->state_ring;
  for (int k = 0; k < 8; k++) p.stats[k] = d->stats + (k < 4 ? k * lb : 4 * lb + (k - 4) * sb);
  p.last_extra = d->last_extra; p.n = d->n; p.buf_stride = buf_stride; p.head = d->head; p.latent_dim = d->latent_dim; p.state_dim = d->state_dim;
  p.latents_buffer_fill = d->latents_buffer_fill; p.latent_offset = d->latent_offset; p.dred_offset = d->dred_offset; p.max_chunks = max_chunks;
  p.max_bytes = max_bytes; p.q0 = q0; p.dQ = dQ; p.qmax = qmax;
  unsigned char *dbuf = buf;
  if (on_device) { p.activity = activity_mem; p.nbytes = nbytes; }
  else {
    if (d->d_buf_bytes < n * buf_stride) {
      cudaFree(d->d_buf); d->d_buf = nullptr; d->d_buf_bytes = 0;
      if (cudaMalloc(&d->d_buf, n * buf_stride) != cudaSuccess) return OPUSB200_ALLOC_FAIL;
      d->d_buf_bytes = n * buf_stride;
    }
    OB_CUDA(cudaMemcpyAsync(d->activity, activity_mem, n * ob::kDredFrames * 4, cudaMemcpyHostToDevice, c->stream));
    p.activity = d->activity; p.nbytes = d->d_nbytes; dbuf = d->d_buf;
  }
  p.buf = dbuf;
  ob_dred_encode_silk_frame<<<(d->n + 63) / 64, 64, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(buf, dbuf, n * buf_stride, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaMemcpyAsync(nbytes, d->d_nbytes, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}
