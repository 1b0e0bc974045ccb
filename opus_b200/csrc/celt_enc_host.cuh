// celt_enc_host.cuh — kernel entry points and C ABI of the batched CELT encoder (included at the end of opus_b200.cu).
#pragma once
#include "celt_enc_frame.cuh"
#include "celt_enc_chains.cuh"

extern "C" __global__ void __launch_bounds__(32) ob_celt_enc_chains(EncAnalysisParams p, float *scratch) {
  extern __shared__ float ob_chain_tile[];
  enc_chains_warp(p, p.first_stream + blockIdx.x * 32, ob_chain_tile, scratch + (size_t)blockIdx.x * OB_CH_SCRATCH_FLOATS);
}

// Frame-type binning ahead of the thread-per-stream frame coder.  The coder's lanes are independent streams, and a frame coded with short blocks
// (transient) walks a different path through quant_all_bands than a long-block frame (Hadamard / time-divide reordering, other band shapes): a warp
// that holds both runs both.  Measured (profiles/r02_enc_stream_grouping.txt): warps of like streams 19.7 ms instead of 22.3, identical streams 8.3.
// So each frame the streams are ranked by the analysis kernel's block-size decision and the coder takes stream perm[thread]: long-block frames
// first, short-block frames packed at the end (order inside a bin = arrival order of the atomics: streams are independent, any order is correct).
extern "C" __global__ void ob_celt_enc_bin_count(const ObEncFrameInfo *info, int32_t *slot, int32_t *counters, int n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int key = info[s].short_blocks != 0;
  const unsigned peers = __match_any_sync(__activemask(), key);
  const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
  int base = 0;
  if (lane == leader) base = atomicAdd(&counters[key], __popc(peers));
  base = __shfl_sync(peers, base, leader);
  slot[s] = (base + __popc(peers & ((1u << lane) - 1))) * 2 + key;
}
extern "C" __global__ void ob_celt_enc_bin_place(const int32_t *slot, int32_t *perm, int n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int v = slot[s], key = v & 1, pos = v >> 1;
  perm[key ? n - 1 - pos : pos] = s;
}
extern "C" __global__ void __launch_bounds__(64) ob_celt_enc_frame(EncFrameParams p) {
  const int t = p.first_stream + blockIdx.x * blockDim.x + threadIdx.x;
  if (t < p.n_streams) enc_frame_thread<false>(p, p.perm ? p.perm[t] : t);
}
extern "C" __global__ void __launch_bounds__(64, 8) ob_celt_enc_frame_rdo(EncFrameParams p) {   // complexity >= 8, stereo; 128 registers
  const int t = p.first_stream + blockIdx.x * blockDim.x + threadIdx.x;
  if (t < p.n_streams) enc_frame_thread<true>(p, p.perm ? p.perm[t] : t);
}
// (Round-2 experiments, measured and removed: the same kernel at register caps 96 / 80 / 64, twice the streams per GPU, and a variant with one
// barrier per band for blocks of 8 warps — no gain; the barrier variant deadlocks under divergent control flow.  DESIGN.md section 2.2.)
extern "C" __global__ void ob_celt_enc_reset(ObEncState *st, ObEncAnalysisState *an, int n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  // OPUS_RESET_STATE (celt/celt_encoder.c:3076-3094); everything else was memset to 0
  for (int i = 0; i < 2 * OB_NBANDS; i++) st[s].oldLogE[i] = st[s].oldLogE2[i] = -28.f;
  st[s].delayedIntra = 1.f;
  st[s].spread_decision = 2;
  st[s].tonal_average = 256;
  (void)an;
}
extern "C" __global__ void ob_celt_enc_get_range(const ObEncState *st, uint32_t *out, int n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) out[s] = st[s].rng;
}

struct OpusB200Encoder {
  int device, channels, n;
  cudaStream_t stream;
  ObEncState *state;
  ObEncAnalysisState *an_state;
  ObEncFrameInfo *info;
  float *X, *bandE, *bandLogE, *bandLogE2;
  float *prefilter_mem;   // [n][C][1024]: the unfiltered history run_prefilter searches (CELTEncoder::prefilter_mem)
  int32_t *d_nbytes;
  // settings (celt_encoder_ctl)
  int complexity, bitrate, vbr, constrained_vbr, lsb_depth, force_intra, disable_inv, disable_pf, loss_rate;
  int end;   // CELT_SET_END_BAND
  int stream_channels;   // CELT_SET_CHANNELS: coded channels (<= channels)
  int upsample;   // 48000 / Fs of opusb200_encoder_create: frame_size and the pcm are in Fs samples
  // staging for the host-buffer call
  float *d_pcm; size_t d_pcm_cap;
  unsigned char *d_pkt; size_t d_pkt_cap;
  int64_t launches;
  // the frame coder runs on a second stream, chunk by chunk behind the analysis kernel (see opusb200_encode_device)
  cudaStream_t stream_f[4];   // one per chunk: the frame coder is latency-bound per warp, so the chunks' coders must overlap each other too
  cudaEvent_t ev_an[4], ev_fr[4];
  // profiling aid (opusb200_opus_encoder_set_timing): CUDA events around each kernel of a call, on the launching stream
  int32_t *bin_slot, *bin_perm, *bin_counters;   // frame-type binning (ob_celt_enc_bin_*): [n], [n], [2]
  float *chain_scratch;   // [ceil(n/32)][OB_CH_SCRATCH_FLOATS]: the recurrence kernel's transposed tracks (celt_enc_chains.cuh)
  float *in_buf;          // [n][channels][1080]: the pre-emphasised, pre-filtered analysis window it hands to the transform kernel
  int timing;
  cudaEvent_t tev[6];
  double t_ms[5];   // Opus-layer pre-pass, Opus-layer analysis, CELT recurrences (thread per stream), CELT analysis, CELT frame coder
  int64_t t_calls;
};

extern "C" int opusb200_encoder_reset(OpusB200Encoder *e) {
  if (!e) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(e->device));
  OB_CUDA(cudaMemsetAsync(e->state, 0, sizeof(ObEncState) * (size_t)e->n, e->stream));
  OB_CUDA(cudaMemsetAsync(e->an_state, 0, sizeof(ObEncAnalysisState) * (size_t)e->n, e->stream));
  OB_CUDA(cudaMemsetAsync(e->prefilter_mem, 0, sizeof(float) * (size_t)e->n * e->channels * kPfMaxPeriod, e->stream));
  ob_celt_enc_reset<<<(e->n + 127) / 128, 128, 0, e->stream>>>(e->state, e->an_state, e->n);
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
extern "C" int opusb200_encoder_create(int device, int32_t Fs, int channels, int n_streams, OpusB200Encoder **out) {
  if (!out) return OPUSB200_BAD_ARG;
  *out = nullptr;
  if ((Fs != 48000 && Fs != 24000 && Fs != 16000 && Fs != 12000 && Fs != 8000) || (channels != 1 && channels != 2) || n_streams <= 0)
    return OPUSB200_BAD_ARG;   // celt_encoder_init: resampling_factor(Fs) must exist (celt/celt_encoder.c:255)
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    fprintf(stderr, "opus_b200: no usable CUDA device (there is no CPU fallback)\n");
    return OPUSB200_INTERNAL_ERROR;
  }
  OB_CUDA(cudaSetDevice(device));
  OpusB200Encoder *e = new (std::nothrow) OpusB200Encoder;
  if (!e) return OPUSB200_ALLOC_FAIL;
  memset(e, 0, sizeof(*e));
  e->device = device; e->channels = channels; e->n = n_streams;
  e->complexity = 5; e->bitrate = -1; e->vbr = 0; e->constrained_vbr = 1; e->lsb_depth = 24;   // celt_encoder_init defaults
  e->end = OB_NBANDS; e->upsample = 48000 / Fs; e->stream_channels = channels;
  OB_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // the frame coder's few, long-running CTAs go ahead of the analysis kernel's many short ones
    for (int i = 0; i < 4; i++) {
      OB_CUDA(cudaStreamCreateWithPriority(&e->stream_f[i], cudaStreamNonBlocking, hi));
      OB_CUDA(cudaEventCreateWithFlags(&e->ev_an[i], cudaEventDisableTiming));
      OB_CUDA(cudaEventCreateWithFlags(&e->ev_fr[i], cudaEventDisableTiming));
    }
  }
  const size_t n = (size_t)n_streams;
  OB_CUDA(cudaMalloc(&e->state, sizeof(ObEncState) * n));
  OB_CUDA(cudaMalloc(&e->an_state, sizeof(ObEncAnalysisState) * n));
  OB_CUDA(cudaMalloc(&e->info, sizeof(ObEncFrameInfo) * n));
  OB_CUDA(cudaMalloc(&e->X, sizeof(float) * n * channels * OB_MAXN));
  OB_CUDA(cudaMalloc(&e->bandE, sizeof(float) * n * channels * OB_NBANDS));
  OB_CUDA(cudaMalloc(&e->bandLogE, sizeof(float) * n * channels * OB_NBANDS));
  OB_CUDA(cudaMalloc(&e->bandLogE2, sizeof(float) * n * channels * OB_NBANDS));
  OB_CUDA(cudaMalloc(&e->prefilter_mem, sizeof(float) * n * channels * kPfMaxPeriod));
  OB_CUDA(cudaMalloc(&e->d_nbytes, sizeof(int32_t) * n));
  OB_CUDA(cudaMalloc(&e->bin_slot, sizeof(int32_t) * n));
  OB_CUDA(cudaMalloc(&e->bin_perm, sizeof(int32_t) * n));
  OB_CUDA(cudaMalloc(&e->bin_counters, sizeof(int32_t) * 2));
  OB_CUDA(cudaMalloc(&e->chain_scratch, sizeof(float) * ((n + 31) / 32) * OB_CH_SCRATCH_FLOATS));
  OB_CUDA(cudaMalloc(&e->in_buf, sizeof(float) * n * channels * OB_CH_IN_STRIDE));
  OB_CUDA(cudaFuncSetAttribute(ob_celt_enc_analysis, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)sizeof(MdctWS) + 32 + (int)sizeof(PfWS)));
  // the frame coder keeps its band vectors in thread-private local memory and uses no shared memory: give the L1 everything
  OB_CUDA(cudaFuncSetAttribute(ob_celt_enc_frame_rdo, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
  OB_CUDA(cudaFuncSetAttribute(ob_celt_enc_frame, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
  // the analysis kernel lives on shared memory (25 KB per stream): ask for the largest carveout whatever the launch's own size is
  OB_CUDA(cudaFuncSetAttribute(ob_celt_enc_analysis, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int r = opusb200_encoder_reset(e);
  if (r != OPUSB200_OK) return r;
  OB_CUDA(cudaStreamSynchronize(e->stream));
  *out = e;
  return OPUSB200_OK;
}
extern "C" void opusb200_encoder_destroy(OpusB200Encoder *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  cudaFree(e->state); cudaFree(e->an_state); cudaFree(e->info); cudaFree(e->X); cudaFree(e->bandE); cudaFree(e->bandLogE); cudaFree(e->bandLogE2); cudaFree(e->prefilter_mem);
  cudaFree(e->d_nbytes); cudaFree(e->d_pcm); cudaFree(e->d_pkt); cudaFree(e->chain_scratch); cudaFree(e->in_buf); cudaFree(e->bin_slot); cudaFree(e->bin_perm); cudaFree(e->bin_counters);
  if (e->tev[0]) for (int i = 0; i < 6; i++) cudaEventDestroy(e->tev[i]);
  for (int i = 0; i < 4; i++) {
    cudaStreamSynchronize(e->stream_f[i]);
    cudaStreamDestroy(e->stream_f[i]);
    cudaEventDestroy(e->ev_an[i]);
    cudaEventDestroy(e->ev_fr[i]);
  }
  cudaStreamDestroy(e->stream);
  delete e;
}
// the subset of opus_encoder_ctl / celt_encoder_ctl requests the device path implements (include/opus_defines.h request numbers)
extern "C" int opusb200_encoder_ctl_set(OpusB200Encoder *e, int request, int32_t value) {
  if (!e) return OPUSB200_BAD_ARG;
  switch (request) {
    case 4002: if (value != -1 && (value < 500 || value > 750000 * e->channels)) return OPUSB200_BAD_ARG; e->bitrate = value; return OPUSB200_OK;  // OPUS_SET_BITRATE (-1 = OPUS_BITRATE_MAX)
    case 4006: e->vbr = value != 0; return OPUSB200_OK;                                                       // OPUS_SET_VBR
    case 4020: e->constrained_vbr = value != 0; return OPUSB200_OK;                                            // OPUS_SET_VBR_CONSTRAINT
    case 4014: if (value < 0 || value > 100) return OPUSB200_BAD_ARG; e->loss_rate = value; return OPUSB200_OK;   // OPUS_SET_PACKET_LOSS_PERC
    case 4010: if (value < 0 || value > 10) return OPUSB200_BAD_ARG; e->complexity = value; return OPUSB200_OK;  // OPUS_SET_COMPLEXITY
    case 4036: if (value < 8 || value > 24) return OPUSB200_BAD_ARG; e->lsb_depth = value; return OPUSB200_OK;  // OPUS_SET_LSB_DEPTH
    case 4046: e->disable_inv = value != 0; return OPUSB200_OK;                                                // OPUS_SET_PHASE_INVERSION_DISABLED
    case 10002: if (value < 0 || value > 2) return OPUSB200_BAD_ARG; e->disable_pf = value <= 1; e->force_intra = value == 0; return OPUSB200_OK;   // CELT_SET_PREDICTION (celt/celt.h:97)
    case 10008: if (value < 1 || value > e->channels) return OPUSB200_BAD_ARG; e->stream_channels = value; return OPUSB200_OK;   // CELT_SET_CHANNELS (celt/celt.h:111): mono coding of a stereo input
    case 10012: if (value < 1 || value > OB_NBANDS) return OPUSB200_BAD_ARG; e->end = value; return OPUSB200_OK;   // CELT_SET_END_BAND (celt/celt.h:119): 13 NB, 17 WB, 19 SWB, 21 FB
    default: return OPUSB200_UNIMPLEMENTED;
  }
}
extern "C" void *opusb200_encoder_cuda_stream(OpusB200Encoder *e) { return e ? (void *)e->stream : nullptr; }
extern "C" int opusb200_encoder_sync(OpusB200Encoder *e) {
  if (!e) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaStreamSynchronize(e->stream));
  return OPUSB200_OK;
}
// device pointers: pcm [n][frame_size*channels] float (+-1.0), packets [n][pkt_stride], nbytes [n]
// ctl != nullptr: the Opus layer in front (opus_enc_host.cuh): coded channels / end band / AnalysisInfo per stream, one TOC byte ahead of the payload
static int celt_encode_launch(OpusB200Encoder *e, const float *d_pcm, int frame_size, unsigned char *d_packets, int pkt_stride, int max_data_bytes,
                              int32_t *d_nbytes, const ob::ObOpusFrameCtl *ctl) {
  if (!e || !d_pcm || !d_packets || !d_nbytes || pkt_stride < max_data_bytes || max_data_bytes < 2) return OPUSB200_BAD_ARG;
  int LM = -1;
  for (int l = 0; l <= 3; l++) if (frame_size * e->upsample == OB_SHORT_N << l) LM = l;   // frame_size counts input-rate samples (celt_encoder.c:1838)
  if (LM < 0) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(e->device));
  EncAnalysisParams a;
  a.state = e->an_state; a.pcm = d_pcm; a.short_blocks = nullptr; a.X = e->X; a.bandE = e->bandE; a.bandLogE = e->bandLogE; a.info = e->info;
  a.n_streams = e->n; a.C = e->stream_channels; a.CC = e->channels; a.LM = LM; a.complexity = e->complexity; a.lsb_depth = e->lsb_depth;
  a.enc_state = e->state; a.prefilter_mem = e->prefilter_mem; a.bandLogE2 = e->bandLogE2; a.max_bytes = max_data_bytes; a.bitrate = e->bitrate;
  a.vbr = e->vbr; a.constrained_vbr = e->constrained_vbr; a.disable_pf = e->disable_pf; a.loss_rate = e->loss_rate; a.end = e->end; a.upsample = e->upsample;
  a.ctl = ctl;
  // the pitch-search buffers exist from complexity 5 only (PfWS: sc always, buf/xlp2 for the search)
  const size_t pf_smem = e->complexity >= 5 ? sizeof(PfWS) : sizeof(float) * 64;
  EncFrameParams p;
  p.state = e->state; p.info = e->info; p.X = e->X; p.bandE = e->bandE; p.bandLogE = e->bandLogE; p.bandLogE2 = e->bandLogE2; p.packets = d_packets; p.nbytes = d_nbytes;
  p.C = e->stream_channels; p.CC = e->channels; p.LM = LM; p.pkt_stride = pkt_stride; p.max_bytes = max_data_bytes; p.complexity = e->complexity;
  p.bitrate = e->bitrate; p.lsb_depth = e->lsb_depth; p.force_intra = e->force_intra; p.disable_inv = e->disable_inv; p.disable_pf = e->disable_pf; p.end = e->end;
  p.vbr = e->vbr; p.constrained_vbr = e->constrained_vbr; p.loss_rate = e->loss_rate;
  p.ctl = ctl; p.toc_bytes = ctl ? 1 : 0;
  // Chunking the batch (analysis of chunk k+1 on the library stream under the frame coder of chunk k on its own higher-priority
  // stream) was measured and is OFF: 9.6 vs 5.9 ms at complexity 0, 56 vs 29 ms at complexity 10 for 65 536 streams.  Both
  // kernels are limited by how many of their warps are resident at once, and splitting the batch only lowers that.  The plumbing
  // (first_stream, per-chunk streams) stays for experiments: OPUSB200_ENC_CHUNKS=4.
  static const int want_chunks = getenv("OPUSB200_ENC_CHUNKS") ? atoi(getenv("OPUSB200_ENC_CHUNKS")) : 1;
  const int chunks = (e->n >= 8192 && want_chunks > 1) ? OB_IMIN(want_chunks, 4) : 1;
  const int per = ((e->n + chunks - 1) / chunks + 63) & ~63;
  for (int k = 0; k < chunks; k++) {
    const int s0 = k * per, s1 = OB_IMIN(e->n, s0 + per);
    if (s0 >= s1) break;
    a.first_stream = s0; a.n_streams = s1;
    if (e->timing) cudaEventRecord(e->tev[2], e->stream);
    // the transient / tone recurrences: one thread per stream ahead of the CTA-per-stream kernel (OPUSB200_ENC_NO_CHAINS=1: the old fused form)
    static const bool no_chains = getenv("OPUSB200_ENC_NO_CHAINS") != nullptr;
    a.chains_done = !no_chains && chunks == 1;
    // OPUSB200_ENC_SERIAL_PF=1: run_prefilter too on the thread-per-stream kernel.  Measured (32 768 streams, complexity 10): the CTA-per-stream
    // kernel drops from 8.4 to 0.95 ms but the serial kernel needs 21.6 ms — 2 M instructions per warp on 7 warps per SM is latency-bound, and
    // the pitch correlations and the comb filter are lane-parallel work that belongs in the CTA-per-stream kernel.  Off by default.
    static const bool serial_pf = getenv("OPUSB200_ENC_SERIAL_PF") != nullptr;
    a.in_buf = (a.chains_done && serial_pf) ? e->in_buf : nullptr;
    if (a.chains_done) {
      ob_celt_enc_chains<<<(s1 - s0 + 31) / 32, 32, 32 * OB_CH_TS * sizeof(float), e->stream>>>(a, e->chain_scratch);
      e->launches++;
    }
    if (e->timing) cudaEventRecord(e->tev[5], e->stream);
    ob_celt_enc_analysis<<<s1 - s0, 32 * e->channels, e->channels * sizeof(MdctWS) + 32 + pf_smem, e->stream>>>(a);
    if (e->timing) cudaEventRecord(e->tev[3], e->stream);
    cudaStream_t sf = chunks > 1 ? e->stream_f[k] : e->stream;
    if (chunks > 1) {
      OB_CUDA(cudaEventRecord(e->ev_an[k], e->stream));
      OB_CUDA(cudaStreamWaitEvent(sf, e->ev_an[k], 0));
    }
    p.first_stream = s0; p.n_streams = s1;
    static const bool no_bin = getenv("OPUSB200_ENC_NO_BIN") != nullptr;
    p.perm = nullptr;
    if (!no_bin && chunks == 1 && e->n >= 64) {
      OB_CUDA(cudaMemsetAsync(e->bin_counters, 0, sizeof(int32_t) * 2, e->stream));
      ob_celt_enc_bin_count<<<(e->n + 255) / 256, 256, 0, e->stream>>>(e->info, e->bin_slot, e->bin_counters, e->n);
      ob_celt_enc_bin_place<<<(e->n + 255) / 256, 256, 0, e->stream>>>(e->bin_slot, e->bin_perm, e->n);
      e->launches += 2;
      p.perm = e->bin_perm;
    }
    if (e->complexity >= 8 && (ctl ? e->channels : e->stream_channels) == 2) ob_celt_enc_frame_rdo<<<(s1 - s0 + 63) / 64, 64, 0, sf>>>(p);
    else ob_celt_enc_frame<<<(s1 - s0 + 63) / 64, 64, 0, sf>>>(p);
    e->launches += 2;
    if (e->timing) cudaEventRecord(e->tev[4], sf);
    if (chunks > 1) OB_CUDA(cudaEventRecord(e->ev_fr[k], sf));
  }
  if (chunks > 1)
    for (int k = 0; k < chunks; k++) OB_CUDA(cudaStreamWaitEvent(e->stream, e->ev_fr[k], 0));
  OB_CUDA(cudaGetLastError());
  if (e->timing) {   // (profiling runs only: the call then waits for its kernels)
    OB_CUDA(cudaEventSynchronize(e->tev[4]));
    float ms = 0.f;
    if (ctl) { cudaEventElapsedTime(&ms, e->tev[0], e->tev[1]); e->t_ms[0] += ms; cudaEventElapsedTime(&ms, e->tev[1], e->tev[2]); e->t_ms[1] += ms; }
    cudaEventElapsedTime(&ms, e->tev[2], e->tev[5]); e->t_ms[2] += ms;
    cudaEventElapsedTime(&ms, e->tev[5], e->tev[3]); e->t_ms[3] += ms;
    cudaEventElapsedTime(&ms, e->tev[3], e->tev[4]); e->t_ms[4] += ms;
    e->t_calls++;
  }
  return OPUSB200_OK;
}
extern "C" int opusb200_encode_device(OpusB200Encoder *e, const float *d_pcm, int frame_size, unsigned char *d_packets, int pkt_stride,
                                      int max_data_bytes, int32_t *d_nbytes) {
  return celt_encode_launch(e, d_pcm, frame_size, d_packets, pkt_stride, max_data_bytes, d_nbytes, nullptr);
}
// host buffers; returns when the packets are on the host.  nbytes[s] = payload size (CBR: the same for all) or a negative code.
extern "C" int opusb200_encode_float(OpusB200Encoder *e, const float *pcm, int frame_size, unsigned char *packets, int pkt_stride, int max_data_bytes,
                                     int32_t *nbytes) {
  if (!e || !pcm || !packets || !nbytes || frame_size <= 0) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(e->device));
  const size_t pb = sizeof(float) * (size_t)e->n * frame_size * e->channels, kb = (size_t)e->n * pkt_stride;
  if (e->d_pcm_cap < pb) { cudaFree(e->d_pcm); e->d_pcm = nullptr; e->d_pcm_cap = 0; OB_CUDA(cudaMalloc(&e->d_pcm, pb)); e->d_pcm_cap = pb; }
  if (e->d_pkt_cap < kb) { cudaFree(e->d_pkt); e->d_pkt = nullptr; e->d_pkt_cap = 0; OB_CUDA(cudaMalloc(&e->d_pkt, kb)); e->d_pkt_cap = kb; }
  OB_CUDA(cudaMemcpyAsync(e->d_pcm, pcm, pb, cudaMemcpyHostToDevice, e->stream));
  const int r = opusb200_encode_device(e, e->d_pcm, frame_size, e->d_pkt, pkt_stride, max_data_bytes, e->d_nbytes);
  if (r != OPUSB200_OK) return r;
  OB_CUDA(cudaMemcpyAsync(packets, e->d_pkt, kb, cudaMemcpyDeviceToHost, e->stream));
  OB_CUDA(cudaMemcpyAsync(nbytes, e->d_nbytes, sizeof(int32_t) * (size_t)e->n, cudaMemcpyDeviceToHost, e->stream));
  OB_CUDA(cudaStreamSynchronize(e->stream));
  return OPUSB200_OK;
}
extern "C" int opusb200_encoder_get_final_range(OpusB200Encoder *e, uint32_t *final_range) {
  if (!e || !final_range) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(e->device));
  DevBuf d;
  if (d.alloc(sizeof(uint32_t) * (size_t)e->n)) return OPUSB200_ALLOC_FAIL;
  ob_celt_enc_get_range<<<(e->n + 127) / 128, 128, 0, e->stream>>>(e->state, (uint32_t *)d.p, e->n);
  OB_CUDA(cudaMemcpyAsync(final_range, d.p, sizeof(uint32_t) * (size_t)e->n, cudaMemcpyDeviceToHost, e->stream));
  OB_CUDA(cudaStreamSynchronize(e->stream));
  return OPUSB200_OK;
}
extern "C" int64_t opusb200_encoder_launch_count(OpusB200Encoder *e) { return e ? e->launches : 0; }
