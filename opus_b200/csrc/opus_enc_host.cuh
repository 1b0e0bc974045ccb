// opus_enc_host.cuh — kernel entry points and C ABI of the batched Opus encoder (the public opus_encode() surface of a CELT-only
// encoder, n independent streams per call; included at the end of opus_b200.cu, after celt_enc_host.cuh).
#pragma once
#include "opus_enc_layer.cuh"

extern "C" __global__ void __launch_bounds__(32) ob_opus_enc_pre(ob::OpusLayerParams p) {
  extern __shared__ float ob_pre_tile[];
  ob::opus_enc_pre_warp(p, blockIdx.x * 32, ob_pre_tile);
}
extern "C" __global__ void __launch_bounds__(128) ob_opus_enc_analysis(ob::OpusLayerParams p) {
  extern __shared__ __align__(16) unsigned char ob_an_smem[];
  const int s = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (s < p.n_streams) ob::opus_enc_analysis_warp(p, s, (ob::AnWS *)ob_an_smem + (threadIdx.x >> 5));
}
extern "C" __global__ void __launch_bounds__(64) ob_opus_enc_multiframe_decide(ob::OpusLayerParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n_streams) ob::opus_multiframe_decide_thread(p, s);
}
extern "C" __global__ void __launch_bounds__(64) ob_opus_repacketize(ob::OpusLayerParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n_streams) ob::opus_repacketize_thread(p, s);
}
extern "C" __global__ void ob_opus_enc_reset(ob::ObOpusEncState *st, int n, int channels) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) ob::opus_enc_state_init(st + s, channels);
}

struct OpusB200OpusEncoder {
  OpusB200Encoder *celt;
  int device, channels, n;
  ob::OpusEncSettings cfg;   // opus_encoder_ctl settings
  ob::ObOpusEncState *st;
  ob::ObTonality *tonal;
  ob::ObOpusPre *pre;
  ob::ObOpusFrameCtl *ctl;
  float *pcm_buf;
  int analysis_was_on;
  // staging of the host-buffer calls
  void *d_pcm; size_t d_pcm_cap;
  unsigned char *d_pkt; size_t d_pkt_cap;
  int32_t *d_nbytes;
  int64_t launches;
  // 40..120 ms packets: per-frame hand-off records, dc-rejected PCM, one-frame packets and their sizes (allocated at the first such call)
  ob::ObOpusPre *mf_pre; ob::ObOpusFrameCtl *mf_ctl; float *mf_pcm_buf; unsigned char *mf_pkt; int32_t *mf_nbytes;
  int mf_frames, mf_stride;
  // the pipelined host-buffer call (opusb200_opus_encode_submit / _wait): two staging slots, PCM upload and packet read-back on their own
  // streams, so the upload of call k+1 and the read-back of call k-1 run under the kernels of call k
  cudaStream_t st_in, st_out;
  cudaEvent_t ev_in[2], ev_cmp[2], ev_out[2];
  void *s_pcm[2]; size_t s_pcm_cap[2];
  unsigned char *s_pkt[2]; size_t s_pkt_cap[2];
  int32_t *s_nb[2], *h_nb[2], *user_nb[2];
  int slot, pending[2];
};

extern "C" int opusb200_opus_encoder_reset(OpusB200OpusEncoder *e) {   // OPUS_RESET_STATE
  if (!e) return OPUSB200_BAD_ARG;
  const int r = opusb200_encoder_reset(e->celt);
  if (r != OPUSB200_OK) return r;
  cudaStream_t s = e->celt->stream;
  OB_CUDA(cudaMemsetAsync(e->st, 0, sizeof(ob::ObOpusEncState) * (size_t)e->n, s));
  OB_CUDA(cudaMemsetAsync(e->tonal, 0, sizeof(ob::ObTonality) * (size_t)e->n, s));
  ob_opus_enc_reset<<<(e->n + 127) / 128, 128, 0, s>>>(e->st, e->n, e->channels);
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
extern "C" int opusb200_opus_encoder_create(int device, int32_t Fs, int channels, int application, int n_streams, OpusB200OpusEncoder **out) {
  if (!out) return OPUSB200_BAD_ARG;
  *out = nullptr;
  // the applications that can pick SILK (VOIP 2048, AUDIO 2049, RESTRICTED_SILK 2052) need the SILK encoder: not built
  if (application == 2048 || application == 2049 || application == 2052) return OPUSB200_UNIMPLEMENTED;
  if (application != 2051 && application != 2053) return OPUSB200_BAD_ARG;   // RESTRICTED_LOWDELAY, RESTRICTED_CELT
  OpusB200Encoder *celt = nullptr;
  int r = opusb200_encoder_create(device, Fs, channels, n_streams, &celt);
  if (r != OPUSB200_OK) return r;
  OpusB200OpusEncoder *e = new (std::nothrow) OpusB200OpusEncoder;
  if (!e) { opusb200_encoder_destroy(celt); return OPUSB200_ALLOC_FAIL; }
  memset(e, 0, sizeof(*e));
  e->celt = celt; e->device = device; e->channels = channels; e->n = n_streams;
  ob::opus_enc_settings_init(&e->cfg, Fs, channels, application);
  const size_t n = (size_t)n_streams;
  OB_CUDA(cudaMalloc(&e->st, sizeof(ob::ObOpusEncState) * n));
  OB_CUDA(cudaMalloc(&e->tonal, sizeof(ob::ObTonality) * n));
  OB_CUDA(cudaMalloc(&e->pre, sizeof(ob::ObOpusPre) * n));
  OB_CUDA(cudaMalloc(&e->ctl, sizeof(ob::ObOpusFrameCtl) * n));
  OB_CUDA(cudaMalloc(&e->pcm_buf, sizeof(float) * n * channels * OB_MAXN));
  OB_CUDA(cudaMalloc(&e->d_nbytes, sizeof(int32_t) * n));
  OB_CUDA(cudaFuncSetAttribute(ob_opus_enc_analysis, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (int)sizeof(ob::AnWS)));
  r = opusb200_opus_encoder_reset(e);
  if (r != OPUSB200_OK) return r;
  OB_CUDA(cudaStreamSynchronize(celt->stream));
  *out = e;
  return OPUSB200_OK;
}
extern "C" void opusb200_opus_encoder_destroy(OpusB200OpusEncoder *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->celt->stream);
  cudaFree(e->st); cudaFree(e->tonal); cudaFree(e->pre); cudaFree(e->ctl); cudaFree(e->pcm_buf); cudaFree(e->d_pcm); cudaFree(e->d_pkt); cudaFree(e->d_nbytes);
  cudaFree(e->mf_pre); cudaFree(e->mf_ctl); cudaFree(e->mf_pcm_buf); cudaFree(e->mf_pkt); cudaFree(e->mf_nbytes);
  if (e->st_in) {
    cudaStreamSynchronize(e->st_in); cudaStreamSynchronize(e->st_out);
    for (int b = 0; b < 2; b++) {
      cudaFree(e->s_pcm[b]); cudaFree(e->s_pkt[b]); cudaFree(e->s_nb[b]); cudaFreeHost(e->h_nb[b]);
      cudaEventDestroy(e->ev_in[b]); cudaEventDestroy(e->ev_cmp[b]); cudaEventDestroy(e->ev_out[b]);
    }
    cudaStreamDestroy(e->st_in); cudaStreamDestroy(e->st_out);
  }
  opusb200_encoder_destroy(e->celt);
  delete e;
}
// opus_encoder_ctl: the integer setters of the CELT-only path (src/opus_encoder.c:2783-3290; request numbers of include/opus_defines.h)
extern "C" int opusb200_opus_encoder_ctl(OpusB200OpusEncoder *e, int request, int32_t value) {
  if (!e) return OPUSB200_BAD_ARG;
  if (request == 4046) return (value < 0 || value > 1) ? OPUSB200_BAD_ARG : opusb200_encoder_ctl_set(e->celt, 4046, value);   // OPUS_SET_PHASE_INVERSION_DISABLED
  if (request == 4028) return opusb200_opus_encoder_reset(e);                                                                   // OPUS_RESET_STATE
  return ob::opus_enc_settings_ctl(&e->cfg, request, value);
}
// 40..120 ms (src/opus_encoder.c:1698-1838): P1 + the analysis chunk of P2 for every 20 ms frame, the packet's decisions and the per-frame controls,
// the CELT kernels frame by frame into one-frame packets, then the repacketizer
static int opus_encode_multiframe(OpusB200OpusEncoder *e, const ob::OpusCallPlan &pl, const void *d_pcm, int pcm_is_int16, int frame_size,
                                  unsigned char *d_packets, int pkt_stride, int32_t *d_nbytes) {
  cudaStream_t s = e->celt->stream;
  const size_t n = (size_t)e->n;
  const int nb = pl.nb_frames, fs = pl.enc_frame_size, CC = e->channels, stride = (pl.max_data_bytes + 15) & ~15;
  if (e->mf_frames < nb || e->mf_stride < stride) {
    cudaFree(e->mf_pre); cudaFree(e->mf_ctl); cudaFree(e->mf_pcm_buf); cudaFree(e->mf_pkt); cudaFree(e->mf_nbytes);
    e->mf_pre = nullptr; e->mf_ctl = nullptr; e->mf_pcm_buf = nullptr; e->mf_pkt = nullptr; e->mf_nbytes = nullptr; e->mf_frames = 0;
    const int st2 = stride > e->mf_stride ? stride : e->mf_stride;
    if (cudaMalloc(&e->mf_pre, sizeof(ob::ObOpusPre) * n * 6) != cudaSuccess || cudaMalloc(&e->mf_ctl, sizeof(ob::ObOpusFrameCtl) * n * 6) != cudaSuccess ||
        cudaMalloc(&e->mf_pcm_buf, sizeof(float) * n * 6 * CC * OB_MAXN) != cudaSuccess || cudaMalloc(&e->mf_pkt, n * 6 * st2) != cudaSuccess ||
        cudaMalloc(&e->mf_nbytes, sizeof(int32_t) * n * 6) != cudaSuccess)
      return OPUSB200_ALLOC_FAIL;
    e->mf_frames = 6; e->mf_stride = st2;
  }
  if (!pl.run_analysis && e->analysis_was_on) OB_CUDA(cudaMemsetAsync(e->tonal, 0, sizeof(ob::ObTonality) * n, s));
  e->analysis_was_on = pl.run_analysis;
  ob::OpusLayerParams p;
  ob::opus_layer_params_init(&p, e->cfg, pl, fs, pcm_is_int16, e->n);
  p.st = e->st; p.tonal = e->tonal; p.pcm = d_pcm; p.packets = d_packets; p.pkt_stride = pkt_stride;
  p.multi = 1; p.pcm_row = frame_size * CC; p.nb_frames = nb; p.packet_frame_size = frame_size; p.pre_all = e->mf_pre; p.ctl_all = e->mf_ctl;
  p.tmp_packets = e->mf_pkt; p.tmp_stride = e->mf_stride; p.tmp_nbytes = e->mf_nbytes; p.out_nbytes = d_nbytes; p.repacketize_len = pl.repacketize_len;
  p.pad = !e->cfg.use_vbr;
  for (int i = 0; i < nb; i++) {
    p.pre = e->mf_pre + (size_t)i * n; p.ctl = e->mf_ctl + (size_t)i * n; p.pcm_buf = e->mf_pcm_buf + (size_t)i * n * CC * OB_MAXN; p.pcm_off = i * fs * CC;
    ob_opus_enc_pre<<<(e->n + 31) / 32, 32, 32 * OB_PRE_TS * sizeof(float), s>>>(p);
    e->launches++;
    if (pl.run_analysis) { ob_opus_enc_analysis<<<(e->n + 3) / 4, 128, 4 * sizeof(ob::AnWS), s>>>(p); e->launches++; }
  }
  ob_opus_enc_multiframe_decide<<<(e->n + 63) / 64, 64, 0, s>>>(p);
  e->launches++;
  OB_CUDA(cudaGetLastError());
  OpusB200Encoder *c = e->celt;
  c->complexity = e->cfg.complexity; c->lsb_depth = pl.lsb; c->loss_rate = e->cfg.packet_loss;
  c->disable_pf = pl.celt_disable_pf; c->force_intra = pl.celt_force_intra;
  c->vbr = pl.celt_vbr; c->bitrate = pl.celt_bitrate; c->constrained_vbr = pl.celt_cvbr;
  c->stream_channels = e->channels; c->end = OB_NBANDS;
  for (int i = 0; i < nb; i++) {
    const int64_t l0 = c->launches;
    const int r = celt_encode_launch(c, e->mf_pcm_buf + (size_t)i * n * CC * OB_MAXN, fs, e->mf_pkt + (size_t)i * n * e->mf_stride + 1, e->mf_stride,
                                     pl.max_data_bytes - 1, e->mf_nbytes + (size_t)i * n, e->mf_ctl + (size_t)i * n);
    e->launches += c->launches - l0;
    if (r != OPUSB200_OK) return r;
  }
  ob_opus_repacketize<<<(e->n + 63) / 64, 64, 0, s>>>(p);
  e->launches++;
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
// opus_encode_native for n streams (device buffers).  d_pcm: [n][frame_size*channels] int16 or float; d_packets [n][pkt_stride]; d_nbytes [n]
extern "C" int opusb200_opus_encode_device(OpusB200OpusEncoder *e, const void *d_pcm, int pcm_is_int16, int frame_size, unsigned char *d_packets,
                                           int pkt_stride, int max_data_bytes_in, int32_t *d_nbytes) {
  if (!e || !d_pcm || !d_packets || !d_nbytes) return OPUSB200_BAD_ARG;
  ob::OpusCallPlan pl;
  const int pr = ob::opus_enc_plan(e->cfg, frame_size, max_data_bytes_in, pcm_is_int16, &pl);
  if (pr) return pr;
  if (pl.nb_frames == 1 && pkt_stride < pl.max_data_bytes) return OPUSB200_BAD_ARG;
  if (pl.nb_frames > 1) {   // CBR pads the packet to repacketize_len; VBR needs room for what the frames can produce
    const int need = e->cfg.use_vbr ? pl.nb_frames * (pl.max_data_bytes + 2) : pl.repacketize_len;
    if (pkt_stride < (need < pl.repacketize_len ? need : pl.repacketize_len)) return OPUSB200_BAD_ARG;
  }
  OB_CUDA(cudaSetDevice(e->device));
  cudaStream_t s = e->celt->stream;
  if (pl.nb_frames > 1) return opus_encode_multiframe(e, pl, d_pcm, pcm_is_int16, frame_size, d_packets, pkt_stride, d_nbytes);
  if (!pl.run_analysis && e->analysis_was_on) OB_CUDA(cudaMemsetAsync(e->tonal, 0, sizeof(ob::ObTonality) * (size_t)e->n, s));   // tonality_analysis_reset (:1261-1263)
  e->analysis_was_on = pl.run_analysis;
  ob::OpusLayerParams p;
  ob::opus_layer_params_init(&p, e->cfg, pl, frame_size, pcm_is_int16, e->n);
  p.st = e->st; p.tonal = e->tonal; p.pre = e->pre; p.ctl = e->ctl; p.pcm = d_pcm; p.pcm_buf = e->pcm_buf; p.packets = d_packets; p.pkt_stride = pkt_stride;
  if (e->celt->timing) cudaEventRecord(e->celt->tev[0], s);
  ob_opus_enc_pre<<<(e->n + 31) / 32, 32, 32 * OB_PRE_TS * sizeof(float), s>>>(p);
  if (e->celt->timing) cudaEventRecord(e->celt->tev[1], s);
  ob_opus_enc_analysis<<<(e->n + 3) / 4, 128, 4 * sizeof(ob::AnWS), s>>>(p);
  e->launches += 2;
  OB_CUDA(cudaGetLastError());
  // the CELT encoder as opus_encode_frame_native configures it (:2266-2294, 2445-2463); channels and end band go per stream through ctl
  OpusB200Encoder *c = e->celt;
  c->complexity = e->cfg.complexity; c->lsb_depth = pl.lsb; c->loss_rate = e->cfg.packet_loss;
  c->disable_pf = pl.celt_disable_pf; c->force_intra = pl.celt_force_intra;
  c->vbr = pl.celt_vbr; c->bitrate = pl.celt_bitrate; c->constrained_vbr = pl.celt_cvbr;
  c->stream_channels = e->channels; c->end = OB_NBANDS;
  const int64_t l0 = c->launches;
  const int r = celt_encode_launch(c, e->pcm_buf, frame_size, d_packets + 1, pkt_stride, pl.max_data_bytes - 1, d_nbytes, e->ctl);
  e->launches += c->launches - l0;
  return r;
}
static int opus_encode_host(OpusB200OpusEncoder *e, const void *pcm, int is16, int frame_size, unsigned char *packets, int pkt_stride, int max_data_bytes,
                            int32_t *nbytes) {
  if (!e || !pcm || !packets || !nbytes || frame_size <= 0) return OPUSB200_BAD_ARG;
  {   // the call's own verdict first (BAD_ARG for a frame size Opus does not have, UNIMPLEMENTED for the 40..120 ms multi-frame packets)
    ob::OpusCallPlan pl;
    const int pr = ob::opus_enc_plan(e->cfg, frame_size, max_data_bytes, is16, &pl);
    if (pr) return pr;
    if (frame_size > 6 * OB_MAXN) return OPUSB200_BAD_ARG;
  }
  OB_CUDA(cudaSetDevice(e->device));
  cudaStream_t s = e->celt->stream;
  const size_t pb = (is16 ? sizeof(int16_t) : sizeof(float)) * (size_t)e->n * frame_size * e->channels, kb = (size_t)e->n * pkt_stride;
  if (e->d_pcm_cap < pb) { cudaFree(e->d_pcm); e->d_pcm = nullptr; e->d_pcm_cap = 0; OB_CUDA(cudaMalloc(&e->d_pcm, pb)); e->d_pcm_cap = pb; }
  if (e->d_pkt_cap < kb) { cudaFree(e->d_pkt); e->d_pkt = nullptr; e->d_pkt_cap = 0; OB_CUDA(cudaMalloc(&e->d_pkt, kb)); e->d_pkt_cap = kb; }
  OB_CUDA(cudaMemcpyAsync(e->d_pcm, pcm, pb, cudaMemcpyHostToDevice, s));
  const int r = opusb200_opus_encode_device(e, e->d_pcm, is16, frame_size, e->d_pkt, pkt_stride, max_data_bytes, e->d_nbytes);
  if (r != OPUSB200_OK) return r;
  OB_CUDA(cudaMemcpyAsync(packets, e->d_pkt, kb, cudaMemcpyDeviceToHost, s));
  OB_CUDA(cudaMemcpyAsync(nbytes, e->d_nbytes, sizeof(int32_t) * (size_t)e->n, cudaMemcpyDeviceToHost, s));
  OB_CUDA(cudaStreamSynchronize(s));
  return OPUSB200_OK;
}
extern "C" int opusb200_opus_encode(OpusB200OpusEncoder *e, const int16_t *pcm, int frame_size, unsigned char *packets, int pkt_stride, int max_data_bytes,
                                    int32_t *nbytes) {
  return opus_encode_host(e, pcm, 1, frame_size, packets, pkt_stride, max_data_bytes, nbytes);
}
extern "C" int opusb200_opus_encode_float(OpusB200OpusEncoder *e, const float *pcm, int frame_size, unsigned char *packets, int pkt_stride,
                                          int max_data_bytes, int32_t *nbytes) {
  return opus_encode_host(e, pcm, 0, frame_size, packets, pkt_stride, max_data_bytes, nbytes);
}
// The batch-server form of the host-buffer call.  submit: upload the PCM (own stream), encode, read the packets back (own stream), all
// asynchronous; staging is double-buffered, so a caller that submits call k+1 before waiting for call k gets k+1's upload and k's read-back
// under the kernels.  The host buffers must stay valid (and, for overlap, be pinned: opusb200_host_alloc) until opusb200_opus_encode_wait
// returns; nbytes[] of every submitted call is filled in by the wait.  At most two calls in flight: a third submit first completes the oldest.
extern "C" int opusb200_opus_encode_wait(OpusB200OpusEncoder *e) {
  if (!e) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(e->device));
  if (!e->st_in) return OPUSB200_OK;
  OB_CUDA(cudaStreamSynchronize(e->st_out));
  OB_CUDA(cudaStreamSynchronize(e->celt->stream));
  for (int b = 0; b < 2; b++)
    if (e->pending[b]) {
      memcpy(e->user_nb[b], e->h_nb[b], sizeof(int32_t) * (size_t)e->n);
      e->pending[b] = 0;
    }
  return OPUSB200_OK;
}
extern "C" int opusb200_opus_encode_submit(OpusB200OpusEncoder *e, const void *pcm, int pcm_is_int16, int frame_size, unsigned char *packets,
                                           int pkt_stride, int max_data_bytes, int32_t *nbytes) {
  if (!e || !pcm || !packets || !nbytes || frame_size <= 0) return OPUSB200_BAD_ARG;
  const int is16 = pcm_is_int16 ? 1 : 0;
  {
    ob::OpusCallPlan pl;
    const int pr = ob::opus_enc_plan(e->cfg, frame_size, max_data_bytes, is16, &pl);
    if (pr) return pr;
    if (frame_size > 6 * OB_MAXN || (pl.nb_frames == 1 && pkt_stride < pl.max_data_bytes)) return OPUSB200_BAD_ARG;
  }
  OB_CUDA(cudaSetDevice(e->device));
  cudaStream_t s = e->celt->stream;
  if (!e->st_in) {
    OB_CUDA(cudaStreamCreateWithFlags(&e->st_in, cudaStreamNonBlocking));
    OB_CUDA(cudaStreamCreateWithFlags(&e->st_out, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
      OB_CUDA(cudaEventCreateWithFlags(&e->ev_in[b], cudaEventDisableTiming));
      OB_CUDA(cudaEventCreateWithFlags(&e->ev_cmp[b], cudaEventDisableTiming));
      OB_CUDA(cudaEventCreateWithFlags(&e->ev_out[b], cudaEventDisableTiming));
      OB_CUDA(cudaMalloc(&e->s_nb[b], sizeof(int32_t) * (size_t)e->n));
      OB_CUDA(cudaMallocHost(&e->h_nb[b], sizeof(int32_t) * (size_t)e->n));
    }
  }
  const int b = e->slot;
  if (e->pending[b]) {   // a third call in flight: finish the one that owns this slot
    OB_CUDA(cudaEventSynchronize(e->ev_out[b]));
    memcpy(e->user_nb[b], e->h_nb[b], sizeof(int32_t) * (size_t)e->n);
    e->pending[b] = 0;
  }
  const size_t pb = (is16 ? sizeof(int16_t) : sizeof(float)) * (size_t)e->n * frame_size * e->channels, kb = (size_t)e->n * pkt_stride;
  if (e->s_pcm_cap[b] < pb || e->s_pkt_cap[b] < kb) {
    OB_CUDA(cudaStreamSynchronize(s)); OB_CUDA(cudaStreamSynchronize(e->st_in)); OB_CUDA(cudaStreamSynchronize(e->st_out));
    if (e->s_pcm_cap[b] < pb) { cudaFree(e->s_pcm[b]); e->s_pcm[b] = nullptr; e->s_pcm_cap[b] = 0; OB_CUDA(cudaMalloc(&e->s_pcm[b], pb)); e->s_pcm_cap[b] = pb; }
    if (e->s_pkt_cap[b] < kb) { cudaFree(e->s_pkt[b]); e->s_pkt[b] = nullptr; e->s_pkt_cap[b] = 0; OB_CUDA(cudaMalloc(&e->s_pkt[b], kb)); e->s_pkt_cap[b] = kb; }
  }
  OB_CUDA(cudaStreamWaitEvent(e->st_in, e->ev_cmp[b], 0));   // the kernels that last read this slot's PCM (no-op if none)
  OB_CUDA(cudaMemcpyAsync(e->s_pcm[b], pcm, pb, cudaMemcpyHostToDevice, e->st_in));
  OB_CUDA(cudaEventRecord(e->ev_in[b], e->st_in));
  OB_CUDA(cudaStreamWaitEvent(s, e->ev_in[b], 0));
  OB_CUDA(cudaStreamWaitEvent(s, e->ev_out[b], 0));          // the read-back that last used this slot's packet buffer
  const int r = opusb200_opus_encode_device(e, e->s_pcm[b], is16, frame_size, e->s_pkt[b], pkt_stride, max_data_bytes, e->s_nb[b]);
  if (r != OPUSB200_OK) return r;
  OB_CUDA(cudaEventRecord(e->ev_cmp[b], s));
  OB_CUDA(cudaStreamWaitEvent(e->st_out, e->ev_cmp[b], 0));
  OB_CUDA(cudaMemcpyAsync(packets, e->s_pkt[b], kb, cudaMemcpyDeviceToHost, e->st_out));
  OB_CUDA(cudaMemcpyAsync(e->h_nb[b], e->s_nb[b], sizeof(int32_t) * (size_t)e->n, cudaMemcpyDeviceToHost, e->st_out));
  OB_CUDA(cudaEventRecord(e->ev_out[b], e->st_out));
  e->user_nb[b] = nbytes;
  e->pending[b] = 1;
  e->slot ^= 1;
  return OPUSB200_OK;
}
extern "C" int opusb200_opus_encoder_get_final_range(OpusB200OpusEncoder *e, uint32_t *final_range) {
  return e ? opusb200_encoder_get_final_range(e->celt, final_range) : OPUSB200_BAD_ARG;
}
// Profiling aid: CUDA events around each of the five kernels of a call (the call then waits for them: not for throughput runs).
// get_timing: accumulated device milliseconds ms_kernels[5] = [Opus-layer pre-pass, Opus-layer analysis, CELT recurrences, CELT analysis, CELT frame coder]
// and the calls they cover
extern "C" int opusb200_opus_encoder_set_timing(OpusB200OpusEncoder *e, int enable) {
  if (!e) return OPUSB200_BAD_ARG;
  OpusB200Encoder *c = e->celt;
  OB_CUDA(cudaSetDevice(e->device));
  if (enable && !c->tev[0]) for (int i = 0; i < 6; i++) OB_CUDA(cudaEventCreate(&c->tev[i]));
  c->timing = enable != 0;
  for (int i = 0; i < 5; i++) c->t_ms[i] = 0.0;
  c->t_calls = 0;
  return OPUSB200_OK;
}
extern "C" int opusb200_opus_encoder_get_timing(OpusB200OpusEncoder *e, double *ms_kernels, int64_t *calls) {
  if (!e || !ms_kernels || !calls) return OPUSB200_BAD_ARG;
  for (int i = 0; i < 5; i++) ms_kernels[i] = e->celt->t_ms[i];
  *calls = e->celt->t_calls;
  return OPUSB200_OK;
}
extern "C" int opusb200_opus_encoder_sync(OpusB200OpusEncoder *e) { return e ? opusb200_encoder_sync(e->celt) : OPUSB200_BAD_ARG; }
extern "C" void *opusb200_opus_encoder_cuda_stream(OpusB200OpusEncoder *e) { return e ? (void *)e->celt->stream : nullptr; }
extern "C" int64_t opusb200_opus_encoder_launch_count(OpusB200OpusEncoder *e) { return e ? e->launches : 0; }
