// opus_b200.cu — kernels' __global__ entry points and the C ABI (include/opus_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false ... (see
// __graft_entry__.build()).  -fmad=false: the CELT float path keeps the reference's separate
// multiply/add roundings (the reference oracle is built with -ffp-contract=off).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>
#include <math.h>

#include "../../include/opus_b200.h"
#include "celt_dec_kernels.cuh"
#include "celt_dec_plc.cuh"
#include "celt_ops.cuh"
#include "celt_mdct_tma.cuh"
#include "silk_ops.cuh"
#include "silk_nsq.cuh"

using namespace ob;

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// lpw = streams (active lanes) per warp.  With few streams per GPU the entropy stage is one wave of
// long-running warps; using fewer lanes per warp trades idle lanes for more, less divergent warps.
extern "C" __global__ void __launch_bounds__(128) ob_celt_dec_entropy(DecEntropyParams p, int lpw) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int stream = warp * lpw + lane;
  if (lane >= lpw || stream >= p.n_streams) return;
  dec_entropy_thread(p, stream);
}

// Persistent: one CTA per SM, the PVQ table staged once per CTA; every warp keeps fetching the next stream from a global counter until the
// batch is exhausted, so a warp that drew a cheap frame does not idle until the CTA's slowest stream is done (the shared-memory slice
// per warp allows only one 13-14 warp CTA per SM).
extern "C" __global__ void __launch_bounds__(512) ob_celt_dec_pvq(DecPvqParams p, int *next_stream) {
  // the PVQ U(n,k) table (5 KB) comes into shared memory once per persistent CTA by cp.async.bulk (the TMA unit's bulk copy) on an mbarrier
  PvqShared *tab = (PvqShared *)dyn_smem();
  const uint32_t bar = tma::saddr(&tab->bar);
  if (threadIdx.x == 0) {
    tma::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tma::mbar_expect_tx(bar, (uint32_t)sizeof(tab->u));
    tma::bulk_g2s(tma::saddr(tab->u), ob_pvq_u, (uint32_t)sizeof(tab->u), bar);
  }
  if (threadIdx.x < 15) tab->rowoff[threadIdx.x] = ob_pvq_rowoff[threadIdx.x];
  tma::mbar_wait(bar, 0);
  __syncthreads();
  unsigned char *ws = dyn_smem() + ((sizeof(PvqShared) + 15) & ~(size_t)15) + (size_t)(threadIdx.x >> 5) * sizeof(PvqWS);
  for (;;) {
    int stream = 0;
    if ((threadIdx.x & 31) == 0) stream = atomicAdd(next_stream, 1);
    stream = __shfl_sync(0xffffffffu, stream, 0);
    if (stream >= p.n_streams) break;
    dec_pvq_warp(p, stream, tab, ws);
    __syncwarp();
  }
}

extern "C" __global__ void __launch_bounds__(64) ob_celt_dec_synth(DecSynthParams p) {
  dec_synth_block(p, blockIdx.x, dyn_smem());
}

// Pitch-based packet-loss concealment.  CTA b looks at streams [8b, 8b+8): eight lanes read their side records and ballot; the CTA
// then conceals the lost ones one after the other.  With nothing lost the kernel is 2 loads per stream (measured: 8 us per launch at
// 65 536 streams).  Small chunks on purpose: a concealed frame takes ~0.2 ms of mostly serial work (the 24-tap synthesis recursion),
// so the lost streams of a batch should spread over as many CTAs as are resident (8 per SM) instead of queueing inside a few.
#define OB_PLC_CHUNK 8
extern "C" __global__ void __launch_bounds__(64) ob_celt_dec_softclip_packet(DecClipParams p) {
  if ((int)blockIdx.x < p.n_streams) dec_softclip_packet_block(p, blockIdx.x);
}
extern "C" __global__ void __launch_bounds__(64) ob_celt_dec_plc(DecPlcParams p) {
  __shared__ unsigned lost_mask;
  if (threadIdx.x < 32) {
    const int s = blockIdx.x * OB_PLC_CHUNK + threadIdx.x;
    int lost = 0;
    if (threadIdx.x < OB_PLC_CHUNK && s < p.n_streams) {
      const ObFrameSide *sd = p.side + s;
      lost = sd->status > 0 && OB_PLC_TYPE(sd->plc) >= OB_PLC_PITCH;
    }
    const unsigned m = __ballot_sync(0xffffffffu, lost);
    if (threadIdx.x == 0) lost_mask = m;
  }
  __syncthreads();
  unsigned m = lost_mask;
  while (m) {
    const int j = __ffs(m) - 1;
    m &= m - 1;
    dec_plc_block(p, blockIdx.x * OB_PLC_CHUNK + j, dyn_smem());
    __syncthreads();
  }
}

// ---- encoder-side kernel operations (celt_ops.cuh) ----
extern "C" __global__ void __launch_bounds__(128) ob_mdct_forward(MdctFwdParams p) {
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= p.n_items) return;
  mdct_forward_warp(p, item, dyn_smem() + (size_t)(threadIdx.x >> 5) * sizeof(MdctWS));
}
// long-block forward MDCT, FFT in registers, inputs and tables staged by cp.async.bulk (celt_mdct_fast.cuh); persistent, 2 CTAs per SM
extern "C" __global__ void __launch_bounds__(kMdctFastWarps * 32, 2) ob_mdct_forward_tma(MdctFwdParams p) { mdct_forward_tma_block(p, dyn_smem()); }
extern "C" __global__ void __launch_bounds__(64) ob_celt_enc_analysis(EncAnalysisParams p) {
  enc_analysis_block(p, p.first_stream + blockIdx.x, dyn_smem());
}
extern "C" __global__ void __launch_bounds__(128) ob_pitch_xcorr(XcorrParams p, int smem_per_warp) {
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= p.n_items) return;
  pitch_xcorr_warp(p, item, dyn_smem() + (size_t)(threadIdx.x >> 5) * smem_per_warp);
}
extern "C" __global__ void __launch_bounds__(32) ob_exp_rotation(VqParams p) { exp_rotation_warp(p, blockIdx.x, dyn_smem()); }
extern "C" __global__ void __launch_bounds__(32) ob_pvq_search(VqParams p) { pvq_search_warp(p, blockIdx.x, dyn_smem()); }

// ---- SILK integer kernels (silk_ops.cuh): one thread per stream ----
extern "C" __global__ void __launch_bounds__(128) ob_silk_lpc_analysis_filter(SilkLpcParams p) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item < p.n_items) silk_lpc_analysis_filter_thread(p, item);
}
extern "C" __global__ void __launch_bounds__(128) ob_silk_resampler_down(SilkResamplerParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n_streams) silk_resampler_thread(p, s);
}

extern "C" __global__ void __launch_bounds__(128) ob_silk_burg_modified(SilkBurgParams p) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item < p.n_items) silk_burg_thread(p, item);
}
extern "C" __global__ void __launch_bounds__(128) ob_silk_autocorrelation(SilkAutocorrParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < p.n_items * p.count) silk_autocorr_thread(p, idx);
}

// SILK noise-shaping quantizers.  Delayed decision: one warp per block, 8 streams x 4 states per warp; the decision rings
// take 26.9 KB of shared memory per warp, so 8 blocks are resident per SM.
template <int ORDER> __device__ __forceinline__ void nsq_del_dec_entry(const SilkNsqParams &p) {
  nsq_del_dec_warp<ORDER>(p, blockIdx.x * 8, (int32_t *)dyn_smem());
}
// one instance per shapingLPCOrder the reference configures (silk/control_codec.c:316-386) + a generic one
extern "C" __global__ void __launch_bounds__(32) ob_silk_nsq_del_dec(SilkNsqParams p) { nsq_del_dec_entry<0>(p); }
extern "C" __global__ void __launch_bounds__(32) ob_silk_nsq_del_dec_o12(SilkNsqParams p) { nsq_del_dec_entry<12>(p); }
extern "C" __global__ void __launch_bounds__(32) ob_silk_nsq_del_dec_o14(SilkNsqParams p) { nsq_del_dec_entry<14>(p); }
extern "C" __global__ void __launch_bounds__(32) ob_silk_nsq_del_dec_o16(SilkNsqParams p) { nsq_del_dec_entry<16>(p); }
extern "C" __global__ void __launch_bounds__(32) ob_silk_nsq_del_dec_o20(SilkNsqParams p) { nsq_del_dec_entry<20>(p); }
extern "C" __global__ void __launch_bounds__(32) ob_silk_nsq_del_dec_o24(SilkNsqParams p) { nsq_del_dec_entry<24>(p); }
typedef void (*NsqDdKernel)(SilkNsqParams);
static NsqDdKernel nsq_dd_kernel(int order) {
  switch (order) {
    case 12: return ob_silk_nsq_del_dec_o12;
    case 14: return ob_silk_nsq_del_dec_o14;
    case 16: return ob_silk_nsq_del_dec_o16;
    case 20: return ob_silk_nsq_del_dec_o20;
    case 24: return ob_silk_nsq_del_dec_o24;
    default: return ob_silk_nsq_del_dec;
  }
}
extern "C" __global__ void __launch_bounds__(64) ob_silk_nsq(SilkNsqParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n_streams) nsq_plain_thread(p, s);
}
extern "C" __global__ void ob_silk_nsq_reset(ObNsqState *st, int n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) {  // silk/init_encoder.c + control_codec.c: prev_gain 1.0, lagPrev 100 (the arrays were memset to 0)
    st[s].prev_gain_Q16 = 65536;
    st[s].lagPrev = 100;
  }
}
extern "C" __global__ void ob_silk_nsq_export(const ObNsqState *st, int32_t *out, int n) {
  const int s = blockIdx.x, t = threadIdx.x;
  const ObNsqState &S = st[s];
  int32_t *o = out + (size_t)s * 1325;
  for (int i = t; i < 640; i += blockDim.x) { o[i] = S.xq[i]; o[640 + i] = S.sLTP_shp_Q14[i]; }
  if (t < 16) o[1280 + t] = S.sLPC_Q14[t];
  if (t < 24) o[1296 + t] = S.sAR2_Q14[t];
  if (t == 0) { o[1320] = S.sLF_AR_shp_Q14; o[1321] = S.sDiff_shp_Q14; o[1322] = S.lagPrev; o[1323] = S.rand_seed; o[1324] = S.prev_gain_Q16; }
}

// fresh decoder state (celt/celt_decoder.c:1794-1813 OPUS_RESET_STATE)
extern "C" __global__ void ob_celt_dec_reset(ObDecBandState *bs, ObDecSynthState *ss, int n) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  const size_t nb = (size_t)n * (sizeof(ObDecBandState) / 4), ns = (size_t)n * (sizeof(ObDecSynthState) / 4);
  uint32_t *b = (uint32_t *)bs, *s = (uint32_t *)ss;
  for (size_t i = tid; i < ns; i += nth) s[i] = 0u;
  for (size_t i = tid; i < nb; i += nth) {
    size_t w = i % (sizeof(ObDecBandState) / 4);
    // words [42, 126) are oldLogE and oldLogE2
    b[i] = (w >= 2 * OB_NBANDS && w < 6 * OB_NBANDS) ? __float_as_uint(-28.f) : 0u;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct OpusB200Decoder {
  int device, channels, n;
  int downsample;   // 48000 / Fs of opusb200_decoder_create
  int disable_inv, decode_gain;   // OPUS_SET_PHASE_INVERSION_DISABLED, OPUS_SET_GAIN (Q8 dB)
  float gain;                     // 2^(decode_gain / (256 * 6.0206))
  int16_t *h_last_N; // host mirror of ObDecBandState.last_N (host-buffer API: sizes the frame-launch loop of a lost packet)
  float *d_clip_stage; size_t d_clip_stage_cap;   // float PCM of multi-frame packets with int16 output (soft-clipped per packet)
  int *d_pvq_next;   // work counter of the persistent PVQ kernel
  int sm_count;
  cudaStream_t stream;
  ObDecBandState *bstate;
  ObDecSynthState *sstate;
  // side info and spectra are double-buffered: the synthesis of frame t (own stream) overlaps the bitstream parsing of t+1
  ObFrameSide *side[2];
  ObScript *script[2];
  float *X[2];
  int par;                  // buffer parity of the next frame
  // three stages, three streams: entropy (library stream) -> vector stage (stream_v) -> synthesis (stream_b).  The entropy kernel
  // owns the energy state, so in pipelined mode frame t+1 is parsed while frame t's vector and synthesis kernels run.
  cudaStream_t stream_v, stream_b, stream_e;   // stream_e: high-priority parsing stream of the pipelined device mode
  cudaEvent_t ev_in;
  cudaEvent_t ev_a1[2], ev_a2[2], ev_b[2];
  int last_b;               // parity whose ev_b marks the most recent synthesis (-1: none)
  int pipelined;            // device API: leave the synthesis un-joined at the end of a call (opusb200_decoder_join)
  int pvq_warps;  // warps (= streams) per CTA of the PVQ kernel
  int ent_lpw;    // streams per warp of the entropy kernel
  // staging for the host-buffer API
  // (two slots: opusb200_decode_submit double-buffers, so one call's PCM read-back runs under the next call's kernels)
  unsigned char *d_packets[2]; size_t d_packets_cap[2];
  int32_t *d_len[2], *d_out[2];
  void *d_pcm[2]; size_t d_pcm_cap[2];
  int slot;
  cudaEvent_t ev_ready[2], ev_done[2];
  int32_t *h_len[2], *h_out[2];   // pinned staging of the two small per-stream arrays (callers rarely pin those)
  int32_t *user_out[2];
  int pending[2];
  int64_t launches;
  int max_smem_optin;
  cudaStream_t stream2;   // read-back stream of the host-buffer API
  // optional per-kernel timing (CUDA events on the decoder's stream)
  int timing;
  cudaEvent_t ev[4];
  double kernel_ms[3];
  int64_t timed_launches;
};

#define OB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "opus_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return OPUSB200_INTERNAL_ERROR; } } while (0)

extern "C" {

const char *opusb200_version(void) { return "opus_b200 0.1 (sm_100a; CELT decode path)"; }

int opusb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int opusb200_decoder_reset(OpusB200Decoder *d) {
  if (!d) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(d->device));
  OB_CUDA(cudaStreamSynchronize(d->stream_e));
  OB_CUDA(cudaStreamSynchronize(d->stream_v));
  OB_CUDA(cudaStreamSynchronize(d->stream_b));   // the reset rewrites the synthesis state too
  ob_celt_dec_reset<<<592, 256, 0, d->stream>>>(d->bstate, d->sstate, d->n);
  for (int s = 0; s < d->n; s++) d->h_last_N[s] = OB_SHORT_N;
  d->launches++;
  OB_CUDA(cudaGetLastError());
  OB_CUDA(cudaStreamSynchronize(d->stream));
  return OPUSB200_OK;
}

int opusb200_decoder_create(int device, int32_t Fs, int channels, int n_streams, OpusB200Decoder **out) {
  if (!out) return OPUSB200_BAD_ARG;
  *out = nullptr;
  if ((Fs != 48000 && Fs != 24000 && Fs != 16000 && Fs != 12000 && Fs != 8000) || (channels != 1 && channels != 2) || n_streams <= 0)
    return OPUSB200_BAD_ARG;  // opus_decoder_init, src/opus_decoder.c:139-141
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    fprintf(stderr, "opus_b200: no usable CUDA device (there is no CPU fallback)\n");
    return OPUSB200_INTERNAL_ERROR;
  }
  OB_CUDA(cudaSetDevice(device));
  OpusB200Decoder *d = new (std::nothrow) OpusB200Decoder;
  if (!d) return OPUSB200_ALLOC_FAIL;
  memset(d, 0, sizeof(*d));
  d->device = device; d->channels = channels; d->n = n_streams;
  d->downsample = 48000 / Fs;  // resampling_factor(), celt/celt.c:66-97
  d->h_last_N = (int16_t *)calloc((size_t)n_streams, sizeof(int16_t));
  if (!d->h_last_N) { delete d; return OPUSB200_ALLOC_FAIL; }
  OB_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  {
    // pipelined device calls parse on a high-priority stream: the entropy kernel's few long-running CTAs then win free SM resources
    // over the vector / synthesis kernels' many short ones and frame t+1's parsing really overlaps frame t's vector stage
    // (device-resident: 6.41 vs 6.22 M frames/s).  The host-buffer path keeps the normal priority: there the delayed synthesis
    // would delay the PCM read-back (end to end 5.36 vs 5.84 M frames/s, measured).
    int lo = 0, hi = 0;   // (numerically: lo = 0 is the default and the LOWEST priority, hi < 0 the highest)
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    OB_CUDA(cudaStreamCreateWithPriority(&d->stream_e, cudaStreamNonBlocking, hi));
    OB_CUDA(cudaEventCreateWithFlags(&d->ev_in, cudaEventDisableTiming));
  }
  OB_CUDA(cudaStreamCreateWithFlags(&d->stream2, cudaStreamNonBlocking));
  {
    // the entropy kernel (few long-running CTAs, on the library stream) should win free SM resources over the vector and synthesis
    // kernels' many short CTAs: those two streams get the lowest priority
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    OB_CUDA(cudaStreamCreateWithPriority(&d->stream_b, cudaStreamNonBlocking, lo));
    OB_CUDA(cudaStreamCreateWithPriority(&d->stream_v, cudaStreamNonBlocking, lo));
  }
  d->last_b = -1;
  for (int b = 0; b < 2; b++) {
    OB_CUDA(cudaEventCreateWithFlags(&d->ev_ready[b], cudaEventDisableTiming));
    OB_CUDA(cudaEventCreateWithFlags(&d->ev_done[b], cudaEventDisableTiming));
    OB_CUDA(cudaEventCreateWithFlags(&d->ev_a1[b], cudaEventDisableTiming));
    OB_CUDA(cudaEventCreateWithFlags(&d->ev_a2[b], cudaEventDisableTiming));
    OB_CUDA(cudaEventCreateWithFlags(&d->ev_b[b], cudaEventDisableTiming));
  }
  OB_CUDA(cudaMalloc(&d->bstate, sizeof(ObDecBandState) * (size_t)n_streams));
  OB_CUDA(cudaMalloc(&d->sstate, sizeof(ObDecSynthState) * (size_t)n_streams));
  for (int b = 0; b < 2; b++) {
    OB_CUDA(cudaMalloc(&d->script[b], sizeof(ObScript) * (size_t)n_streams));
    OB_CUDA(cudaMalloc(&d->side[b], sizeof(ObFrameSide) * (size_t)n_streams));
    OB_CUDA(cudaMalloc(&d->X[b], sizeof(float) * 2 * OB_MAXN * (size_t)n_streams));
  }
  for (int b = 0; b < 2; b++) {
    OB_CUDA(cudaMalloc(&d->d_len[b], sizeof(int32_t) * (size_t)n_streams));
    OB_CUDA(cudaMalloc(&d->d_out[b], sizeof(int32_t) * (size_t)n_streams));
    OB_CUDA(cudaMallocHost(&d->h_len[b], sizeof(int32_t) * (size_t)n_streams));
    OB_CUDA(cudaMallocHost(&d->h_out[b], sizeof(int32_t) * (size_t)n_streams));
  }
  OB_CUDA(cudaDeviceGetAttribute(&d->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  OB_CUDA(cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, device));
  OB_CUDA(cudaMalloc(&d->d_pvq_next, sizeof(int)));
  OB_CUDA(cudaFuncSetAttribute(ob_celt_dec_pvq, cudaFuncAttributeMaxDynamicSharedMemorySize, d->max_smem_optin));
  OB_CUDA(cudaFuncSetAttribute(ob_celt_dec_pvq, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  OB_CUDA(cudaFuncSetAttribute(ob_celt_dec_entropy, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
  {
    const char *e = getenv("OPUSB200_PVQ_WARPS");
    // default: one CTA per SM holding as many stream-warps as its shared memory allows
    const int fit = (int)((d->max_smem_optin - ((sizeof(PvqShared) + 15) & ~(size_t)15)) / sizeof(PvqWS));
    d->pvq_warps = e ? atoi(e) : fit;
    if (d->pvq_warps > fit) d->pvq_warps = fit;
    if (d->pvq_warps < 1) d->pvq_warps = 1;
    if (d->pvq_warps > 16) d->pvq_warps = 16;
    e = getenv("OPUSB200_ENT_LPW");
    d->ent_lpw = e ? atoi(e) : 32;
    if (d->ent_lpw < 1) d->ent_lpw = 1;
    if (d->ent_lpw > 32) d->ent_lpw = 32;
  }
  OB_CUDA(cudaFuncSetAttribute(ob_celt_dec_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)sizeof(SynthWS)));
  OB_CUDA(cudaFuncSetAttribute(ob_celt_dec_plc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PlcWS)));
  OB_CUDA(cudaFuncSetAttribute(ob_celt_dec_plc, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));   // 8 CTAs of 26 KB per SM

  OB_CUDA(cudaFuncSetAttribute(ob_celt_dec_synth, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // the band decoder recurses over partitions (depth <= 5)
  size_t stk = 0;
  cudaDeviceGetLimit(&stk, cudaLimitStackSize);
  if (stk < 4096) OB_CUDA(cudaDeviceSetLimit(cudaLimitStackSize, 4096));
  int r = opusb200_decoder_reset(d);
  if (r != OPUSB200_OK) return r;
  OB_CUDA(cudaStreamSynchronize(d->stream));
  *out = d;
  return OPUSB200_OK;
}

void opusb200_decoder_destroy(OpusB200Decoder *d) {
  if (!d) return;
  cudaSetDevice(d->device);
  cudaStreamSynchronize(d->stream);
  cudaStreamSynchronize(d->stream_e);
  cudaStreamSynchronize(d->stream_v);
  cudaStreamSynchronize(d->stream_b);
  cudaStreamSynchronize(d->stream2);
  cudaStreamDestroy(d->stream_e);
  cudaEventDestroy(d->ev_in);
  cudaFree(d->bstate); cudaFree(d->sstate); cudaFree(d->d_clip_stage); cudaFree(d->d_pvq_next);
  free(d->h_last_N);
  for (int b = 0; b < 2; b++) { cudaFree(d->script[b]); cudaFree(d->side[b]); cudaFree(d->X[b]); cudaEventDestroy(d->ev_a1[b]); cudaEventDestroy(d->ev_a2[b]); cudaEventDestroy(d->ev_b[b]); }
  cudaStreamDestroy(d->stream_v);
  cudaStreamDestroy(d->stream_b);
  for (int b = 0; b < 2; b++) { cudaFree(d->d_len[b]); cudaFree(d->d_out[b]); cudaFree(d->d_packets[b]); cudaFree(d->d_pcm[b]); cudaFreeHost(d->h_len[b]); cudaFreeHost(d->h_out[b]); }
  cudaStreamDestroy(d->stream);
  cudaStreamDestroy(d->stream2);
  for (int b = 0; b < 2; b++) { cudaEventDestroy(d->ev_ready[b]); cudaEventDestroy(d->ev_done[b]); }
  delete d;
}

int opusb200_decoder_n_streams(const OpusB200Decoder *d) { return d ? d->n : OPUSB200_BAD_ARG; }
void *opusb200_decoder_cuda_stream(OpusB200Decoder *d) { return d ? (void *)d->stream : nullptr; }
int64_t opusb200_decoder_launch_count(const OpusB200Decoder *d) { return d ? d->launches : 0; }

// Make the library stream wait for the most recent synthesis: work enqueued on it afterwards sees the PCM.
int opusb200_decoder_join(OpusB200Decoder *d) {
  if (!d) return OPUSB200_BAD_ARG;
  if (d->last_b >= 0) OB_CUDA(cudaStreamWaitEvent(d->stream, d->ev_b[d->last_b], 0));
  return OPUSB200_OK;
}
int opusb200_decoder_set_pipelined(OpusB200Decoder *d, int enable) {
  if (!d) return OPUSB200_BAD_ARG;
  d->pipelined = enable != 0;
  return opusb200_decoder_join(d);
}
int opusb200_decoder_sync(OpusB200Decoder *d) {
  if (!d) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaStreamSynchronize(d->stream));
  OB_CUDA(cudaStreamSynchronize(d->stream_e));
  OB_CUDA(cudaStreamSynchronize(d->stream_v));
  OB_CUDA(cudaStreamSynchronize(d->stream_b));
  return OPUSB200_OK;
}

// Launch the three decode kernels for streams [s0, s0+cn) on stream `st`.  Every per-stream array is indexed by stream id,
// so a sub-range is just an offset of the base pointers.
static int decode_launch(OpusB200Decoder *d, cudaStream_t st, int s0, int cn, const unsigned char *d_packets, int pkt_stride,
                         const int32_t *d_len, void *d_pcm, int pcm_stride, int pcm_is_int16, int frame_size, int max_frames_per_packet,
                         int32_t *d_out_samples) {
  const int pw = d->pvq_warps;
  const size_t pvq_smem = ((sizeof(PvqShared) + 15) & ~(size_t)15) + (size_t)pw * sizeof(PvqWS);
  if ((int)pvq_smem > d->max_smem_optin) return OPUSB200_INTERNAL_ERROR;
  const size_t esz = pcm_is_int16 ? 2 : 4;
  // int16 output of packets that may hold several frames: opus_decode() soft-clips the PACKET, so the frames go to a float staging
  // area first and one more kernel clips and converts (one-frame packets keep the fused per-frame path, which is the same thing)
  const bool defer_clip = pcm_is_int16 && max_frames_per_packet > 1;
  const int stage_stride = frame_size * d->channels;
  if (defer_clip) {
    const size_t need = sizeof(float) * (size_t)d->n * stage_stride;
    if (d->d_clip_stage_cap < need) {
      OB_CUDA(cudaDeviceSynchronize());
      cudaFree(d->d_clip_stage); d->d_clip_stage = nullptr; d->d_clip_stage_cap = 0;
      OB_CUDA(cudaMalloc(&d->d_clip_stage, need));
      d->d_clip_stage_cap = need;
    }
  }
  cudaStream_t last_sb = st;
  for (int k = 0; k < max_frames_per_packet; k++) {
    const int par = d->par;
    d->par ^= 1;
    const bool timing = d->timing && st == d->stream;
    cudaStream_t sv = timing ? st : d->stream_v, sb = timing ? st : d->stream_b;   // profiling mode keeps everything on one stream
    cudaStream_t se = st;   // parsing stream
    if (d->pipelined && st == d->stream && !timing) {
      se = d->stream_e;
      OB_CUDA(cudaEventRecord(d->ev_in, st));          // the caller's inputs are ordered on the library stream
      OB_CUDA(cudaStreamWaitEvent(se, d->ev_in, 0));
    }
    // this parity's script / side / X buffers were last used two frames ago
    if (!timing) {
      OB_CUDA(cudaStreamWaitEvent(se, d->ev_a2[par], 0));
      OB_CUDA(cudaStreamWaitEvent(se, d->ev_b[par], 0));
    }
    DecEntropyParams ep;
    ep.packets = d_packets + (size_t)s0 * pkt_stride; ep.pkt_len = d_len + s0; ep.pkt_stride = pkt_stride; ep.n_streams = cn;
    ep.channels = d->channels; ep.frame_index = k; ep.state = d->bstate + s0; ep.side = d->side[par] + s0; ep.script = d->script[par] + s0;
    ep.out_samples = (d_out_samples ? d_out_samples : d->d_out[0]) + s0; ep.max_samples = frame_size; ep.downsample = d->downsample; ep.disable_inv = d->disable_inv;
    DecPvqParams pp;
    pp.side = d->side[par] + s0; pp.script = d->script[par] + s0; pp.state = d->bstate + s0; pp.X = d->X[par] + (size_t)s0 * 2 * OB_MAXN; pp.n_streams = cn;
    if (timing) cudaEventRecord(d->ev[0], st);
    {
      const int warps_needed = (cn + d->ent_lpw - 1) / d->ent_lpw;
      ob_celt_dec_entropy<<<(warps_needed + 3) / 4, 128, 0, se>>>(ep, d->ent_lpw);
    }
    if (timing) cudaEventRecord(d->ev[1], st);
    else {
      OB_CUDA(cudaEventRecord(d->ev_a1[par], se));
      // the parse reads the caller's packets / lengths and writes out_samples: whatever the caller enqueues next on the library stream
      // (an upload of the next batch into the same buffers, a read of out_samples) is ordered after it.  This costs no overlap:
      // consecutive parses serialise on the parsing stream anyway.
      if (se != st) OB_CUDA(cudaStreamWaitEvent(st, d->ev_a1[par], 0));
      OB_CUDA(cudaStreamWaitEvent(sv, d->ev_a1[par], 0));
      OB_CUDA(cudaStreamWaitEvent(sv, d->ev_b[par], 0));   // X[par] is free once the synthesis of two frames ago has read it
    }
    {
      const int ctas = OB_IMIN((cn + pw - 1) / pw, d->sm_count);
      OB_CUDA(cudaMemsetAsync(d->d_pvq_next, 0, sizeof(int), sv));
      ob_celt_dec_pvq<<<ctas, pw * 32, pvq_smem, sv>>>(pp, d->d_pvq_next);
    }
    if (timing) cudaEventRecord(d->ev[2], st);
    else {
      OB_CUDA(cudaEventRecord(d->ev_a2[par], sv));
      OB_CUDA(cudaStreamWaitEvent(sb, d->ev_a2[par], 0));
    }
    DecSynthParams sp;
    static const int force_smem_fft = getenv("OPUSB200_SYNTH_SMEM_FFT") ? 1 : 0;
    sp.force_smem_fft = force_smem_fft;
    static const int no_prefetch = getenv("OPUSB200_SYNTH_NO_PREFETCH") ? 1 : 0;
    sp.no_prefetch = no_prefetch;
    sp.side = d->side[par] + s0; sp.X = d->X[par] + (size_t)s0 * 2 * OB_MAXN; sp.state = d->sstate + s0;
    sp.pcm = (char *)d_pcm + (size_t)s0 * pcm_stride * esz; sp.pcm_stride = pcm_stride; sp.frame_index = k;
    sp.pcm_is_int16 = pcm_is_int16; sp.n_streams = cn; sp.channels = d->channels; sp.downsample = d->downsample;
    sp.apply_gain = d->decode_gain != 0; sp.gain = d->gain;
    if (defer_clip) { sp.pcm = d->d_clip_stage + (size_t)s0 * stage_stride; sp.pcm_stride = stage_stride; sp.pcm_is_int16 = 0; sp.defer_clip = 1; }
    last_sb = sb;
    ob_celt_dec_synth<<<cn, 32 * d->channels, d->channels * sizeof(SynthWS), sb>>>(sp);
    {  // lost frames (len == 0, or a DTX frame inside a packet): concealed from the history ring; a no-op scan otherwise
      DecPlcParams lp;
      lp.side = sp.side; lp.state = sp.state; lp.pcm = sp.pcm; lp.pcm_stride = sp.pcm_stride; lp.pcm_is_int16 = sp.pcm_is_int16; lp.defer_clip = sp.defer_clip;
      lp.n_streams = cn; lp.channels = d->channels; lp.downsample = d->downsample; lp.apply_gain = sp.apply_gain; lp.gain = sp.gain;
      ob_celt_dec_plc<<<(cn + OB_PLC_CHUNK - 1) / OB_PLC_CHUNK, 64, sizeof(PlcWS), sb>>>(lp);
    }
    d->launches += 4;
    if (!timing) {
      OB_CUDA(cudaEventRecord(d->ev_b[par], sb));
      d->last_b = par;
    }
    if (timing) {  // profiling mode: serialises the stream, never used for throughput numbers
      cudaEventRecord(d->ev[3], st);
      cudaEventSynchronize(d->ev[3]);
      float a = 0.f, b = 0.f, c = 0.f;
      cudaEventElapsedTime(&a, d->ev[0], d->ev[1]);
      cudaEventElapsedTime(&b, d->ev[1], d->ev[2]);
      cudaEventElapsedTime(&c, d->ev[2], d->ev[3]);
      d->kernel_ms[0] += a; d->kernel_ms[1] += b; d->kernel_ms[2] += c; d->timed_launches++;
    }
  }
  if (defer_clip) {   // after the last frame's synthesis / concealment, on their stream
    DecClipParams cp;
    cp.stage = d->d_clip_stage + (size_t)s0 * stage_stride; cp.stage_stride = stage_stride; cp.pcm = (int16_t *)d_pcm + (size_t)s0 * pcm_stride;
    cp.pcm_stride = pcm_stride; cp.out_samples = (d_out_samples ? d_out_samples : d->d_out[0]) + s0; cp.pkt_len = d_len + s0; cp.state = d->sstate + s0;
    cp.n_streams = cn; cp.channels = d->channels;
    const bool timing = d->timing && st == d->stream;
    // (out_samples is final: the last frame's synthesis on this stream already waited, through the vector stage, for the last parse)
    ob_celt_dec_softclip_packet<<<cn, 32 * d->channels, 0, last_sb>>>(cp);
    d->launches++;
    if (!timing && d->last_b >= 0) OB_CUDA(cudaEventRecord(d->ev_b[d->last_b], last_sb));   // the join waits for the clip too
  }
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}

int opusb200_decode_device(OpusB200Decoder *d, const unsigned char *d_packets, int pkt_stride, const int32_t *d_len,
                           void *d_pcm, int pcm_stride, int pcm_is_int16, int frame_size, int max_frames_per_packet,
                           int32_t *d_out_samples) {
  if (!d || !d_packets || !d_len || !d_pcm || pkt_stride <= 0 || frame_size <= 0) return OPUSB200_BAD_ARG;
  if (pcm_stride < frame_size * d->channels) return OPUSB200_BAD_ARG;
  if (max_frames_per_packet < 1) max_frames_per_packet = 1;
  if (max_frames_per_packet > 48) max_frames_per_packet = 48;
  OB_CUDA(cudaSetDevice(d->device));
  const int r = decode_launch(d, d->stream, 0, d->n, d_packets, pkt_stride, d_len, d_pcm, pcm_stride, pcm_is_int16, frame_size,
                              max_frames_per_packet, d_out_samples);
  if (r != OPUSB200_OK) return r;
  // stream-ordered by default: the PCM is complete for whatever is enqueued on the library stream next.  In pipelined mode
  // the synthesis stays on its own stream so that the next call's parsing overlaps it (opusb200_decoder_join / _sync).
  return d->pipelined ? OPUSB200_OK : opusb200_decoder_join(d);
}

// frames that conceal a lost packet of frame_size output-rate samples (src/opus_decoder.c:316-365), for a stream whose last accepted
// packet had frames of last_N 48 kHz samples; the same walk as ent_lost_packet
static int host_plc_frames(int frame_size, int ds, int last_N) {
  const int F20 = 960 / ds, F10 = F20 >> 1, F5 = F10 >> 1, cap = last_N / ds;
  int n = 0;
  while (frame_size > 0 && n < 48) {
    int a = frame_size < cap ? frame_size : cap;
    if (a > F20) a = F20;
    else if (a < F20) {
      if (a > F10) a = F10;
      else if (a > F5 && a < F10) a = F5;
    }
    frame_size -= a;
    n++;
  }
  return n;
}

// Enqueue one batched decode with HOST buffers: H2D + kernels on the main stream, PCM read-back on the second stream.
// Staging is double-buffered, so a caller that submits call k+1 before waiting for call k gets k's read-back (the largest
// transfer by far) for free under k+1's kernels.  The host buffers must stay valid until opusb200_decode_wait returns.
static int decode_submit(OpusB200Decoder *d, const unsigned char *packets, int pkt_stride, const int32_t *len, void *pcm,
                         int pcm_stride, int is_int16, int frame_size, int32_t *out_samples) {
  if (!d || !packets || !len || !pcm || !out_samples || pkt_stride <= 0 || frame_size <= 0) return OPUSB200_BAD_ARG;
  if (pcm_stride < frame_size * d->channels) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(d->device));
  const int b = d->slot;
  const size_t pbytes = (size_t)d->n * pkt_stride;
  const size_t esz = is_int16 ? 2 : 4;
  const size_t obytes = (size_t)d->n * pcm_stride * esz;
  if (d->d_packets_cap[b] < pbytes || d->d_pcm_cap[b] < obytes) {
    OB_CUDA(cudaStreamSynchronize(d->stream));
    OB_CUDA(cudaStreamSynchronize(d->stream2));
    if (d->d_packets_cap[b] < pbytes) {
      cudaFree(d->d_packets[b]); d->d_packets[b] = nullptr; d->d_packets_cap[b] = 0;
      OB_CUDA(cudaMalloc(&d->d_packets[b], pbytes));
      d->d_packets_cap[b] = pbytes;
    }
    if (d->d_pcm_cap[b] < obytes) {
      cudaFree(d->d_pcm[b]); d->d_pcm[b] = nullptr; d->d_pcm_cap[b] = 0;
      OB_CUDA(cudaMalloc(&d->d_pcm[b], obytes));
      d->d_pcm_cap[b] = obytes;
    }
  }
  // How many frame launches does the call need?  The host walks the packet framing exactly as the entropy kernel will (the same
  // function), so it also knows each stream's frame size when a later packet is lost.
  int maxf = 1;
  const int ds = d->downsample;
  for (int s = 0; s < d->n; s++) {
    const int l = len[s];
    __builtin_prefetch(packets + (size_t)(s + 24) * pkt_stride);   // the TOC bytes are one cache miss per stream (measured: 0.48 -> 0.38 ms per 65 536 streams)
    if (l > pkt_stride) continue;   // would run into the next stream's row: the kernel answers OPUS_BAD_ARG for this stream
    if (l > 0) {
      const unsigned char *pk = packets + (size_t)s * pkt_stride;
      const int toc = pk[0];
      if (!(toc & 0x80)) continue;   // not CELT-only: the kernel rejects it
      const int N = OB_SHORT_N << ((toc >> 3) & 3);
      int c;
      if ((toc & 3) == 0) c = l - 1 <= OB_MAX_PACKET ? 1 : -4;   // one frame: locate_frame's code-0 case, without the call
      else { int off = 0, fl = 0; c = ob::locate_frame(pk, l, 0, &off, &fl); }
      if (c > 0 && c * (N / ds) <= frame_size) {   // what the entropy kernel accepts
        d->h_last_N[s] = (int16_t)N;
        if (c > maxf) maxf = c;
      }
    } else if (l == 0) {  // lost: concealed over the whole frame_size, possibly in several frames
      const int c = host_plc_frames(frame_size, ds, d->h_last_N[s]);
      if (c > maxf) maxf = c;
    }
  }
  if (maxf > 48) maxf = 48;
  if (d->pending[b]) {  // a third call in flight: finish the one that owns this slot first
    OB_CUDA(cudaEventSynchronize(d->ev_done[b]));
    memcpy(d->user_out[b], d->h_out[b], sizeof(int32_t) * (size_t)d->n);
    d->pending[b] = 0;
  }
  memcpy(d->h_len[b], len, sizeof(int32_t) * (size_t)d->n);
  OB_CUDA(cudaStreamWaitEvent(d->stream, d->ev_done[b], 0));  // the read-back that last used this slot (no-op if none)
  OB_CUDA(cudaMemcpyAsync(d->d_packets[b], packets, pbytes, cudaMemcpyHostToDevice, d->stream));
  OB_CUDA(cudaMemcpyAsync(d->d_len[b], d->h_len[b], sizeof(int32_t) * (size_t)d->n, cudaMemcpyHostToDevice, d->stream));
  const int r = decode_launch(d, d->stream, 0, d->n, d->d_packets[b], pkt_stride, d->d_len[b], d->d_pcm[b], pcm_stride, is_int16,
                              frame_size, maxf, d->d_out[b]);
  if (r != OPUSB200_OK) return r;
  OB_CUDA(cudaEventRecord(d->ev_ready[b], d->stream));   // out_samples (entropy kernel)
  OB_CUDA(cudaStreamWaitEvent(d->stream2, d->ev_ready[b], 0));
  if (d->last_b >= 0) OB_CUDA(cudaStreamWaitEvent(d->stream2, d->ev_b[d->last_b], 0));   // the PCM (synthesis stream)
  // (splitting this read-back over two streams / copy engines was measured: 6.06 vs 6.07 M frames/s end to end — the PCIe link is the bound)
  OB_CUDA(cudaMemcpyAsync(pcm, d->d_pcm[b], obytes, cudaMemcpyDeviceToHost, d->stream2));
  OB_CUDA(cudaMemcpyAsync(d->h_out[b], d->d_out[b], sizeof(int32_t) * (size_t)d->n, cudaMemcpyDeviceToHost, d->stream2));
  OB_CUDA(cudaEventRecord(d->ev_done[b], d->stream2));
  d->user_out[b] = out_samples;
  d->pending[b] = 1;
  d->slot ^= 1;
  return OPUSB200_OK;
}
int opusb200_decode_wait(OpusB200Decoder *d) {
  if (!d) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(d->device));
  OB_CUDA(cudaStreamSynchronize(d->stream));
  OB_CUDA(cudaStreamSynchronize(d->stream_v));
  OB_CUDA(cudaStreamSynchronize(d->stream_b));
  OB_CUDA(cudaStreamSynchronize(d->stream2));
  for (int b = 0; b < 2; b++)
    if (d->pending[b]) {
      memcpy(d->user_out[b], d->h_out[b], sizeof(int32_t) * (size_t)d->n);
      d->pending[b] = 0;
    }
  return OPUSB200_OK;
}
int opusb200_decode_submit(OpusB200Decoder *d, const unsigned char *packets, int pkt_stride, const int32_t *len, void *pcm,
                           int pcm_stride, int pcm_is_int16, int frame_size, int32_t *out_samples) {
  return decode_submit(d, packets, pkt_stride, len, pcm, pcm_stride, pcm_is_int16 ? 1 : 0, frame_size, out_samples);
}
static int decode_host(OpusB200Decoder *d, const unsigned char *packets, int pkt_stride, const int32_t *len, void *pcm,
                       int pcm_stride, int is_int16, int frame_size, int32_t *out_samples) {
  const int r = decode_submit(d, packets, pkt_stride, len, pcm, pcm_stride, is_int16, frame_size, out_samples);
  if (r != OPUSB200_OK) return r;
  return opusb200_decode_wait(d);
}

int opusb200_decode(OpusB200Decoder *d, const unsigned char *packets, int pkt_stride, const int32_t *len, int16_t *pcm,
                    int pcm_stride, int frame_size, int32_t *out_samples) {
  return decode_host(d, packets, pkt_stride, len, pcm, pcm_stride, 1, frame_size, out_samples);
}
int opusb200_decode_float(OpusB200Decoder *d, const unsigned char *packets, int pkt_stride, const int32_t *len, float *pcm,
                          int pcm_stride, int frame_size, int32_t *out_samples) {
  return decode_host(d, packets, pkt_stride, len, pcm, pcm_stride, 0, frame_size, out_samples);
}

int opusb200_decoder_get_final_range(OpusB200Decoder *d, uint32_t *final_range) {
  if (!d || !final_range) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(d->device));
  OB_CUDA(cudaStreamWaitEvent(d->stream, d->ev_in, 0));   // no-op unless a pipelined call parsed on the high-priority stream:
  if (d->last_b >= 0) OB_CUDA(cudaStreamWaitEvent(d->stream, d->ev_a1[d->last_b], 0));   // its entropy kernel wrote the ranges
  OB_CUDA(cudaMemcpy2DAsync(final_range, sizeof(uint32_t), &d->bstate[0].final_range, sizeof(ObDecBandState),
                            sizeof(uint32_t), d->n, cudaMemcpyDeviceToHost, d->stream));
  OB_CUDA(cudaStreamSynchronize(d->stream));
  return OPUSB200_OK;
}

// opus_decoder_ctl() for the integer requests that touch the decode path (src/opus_decoder.c:1057-1215), applied to every stream.
int opusb200_decoder_ctl(OpusB200Decoder *d, int request, int32_t value) {
  if (!d) return OPUSB200_BAD_ARG;
  switch (request) {
    case 4034:  // OPUS_SET_GAIN: Q8 dB, pcm *= 2^(gain / (256 * 6.0206))  (celt_exp2 of the float build)
      if (value < -32768 || value > 32767) return OPUSB200_BAD_ARG;
      d->decode_gain = value;
      d->gain = (float)exp(0.6931471805599453094 * (6.48814081e-4f * (float)value));
      return OPUSB200_OK;
    case 4046:  // OPUS_SET_PHASE_INVERSION_DISABLED
      if (value < 0 || value > 1) return OPUSB200_BAD_ARG;
      d->disable_inv = value;
      return OPUSB200_OK;
    default:
      return OPUSB200_UNIMPLEMENTED;
  }
}

// The getters, one value per stream: 4033 OPUS_GET_PITCH (the post-filter period of a CELT-only stream, 0 before its first packet),
// 4039 OPUS_GET_LAST_PACKET_DURATION, 4045 OPUS_GET_GAIN, 4047 OPUS_GET_PHASE_INVERSION_DISABLED, 4029 OPUS_GET_SAMPLE_RATE.
int opusb200_decoder_get_ctl(OpusB200Decoder *d, int request, int32_t *values) {
  if (!d || !values) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(d->device));
  if (request == 4045 || request == 4047 || request == 4029) {
    const int32_t v = request == 4045 ? d->decode_gain : request == 4047 ? d->disable_inv : 48000 / d->downsample;
    for (int s = 0; s < d->n; s++) values[s] = v;
    return OPUSB200_OK;
  }
  if (request != 4033 && request != 4039) return OPUSB200_UNIMPLEMENTED;
  OB_CUDA(cudaStreamWaitEvent(d->stream, d->ev_in, 0));
  if (d->last_b >= 0) OB_CUDA(cudaStreamWaitEvent(d->stream, d->ev_a1[d->last_b], 0));   // the entropy kernel owns this state
  const int32_t *src = request == 4033 ? &d->bstate[0].pf_period : &d->bstate[0].last_status;
  OB_CUDA(cudaMemcpy2DAsync(values, sizeof(int32_t), src, sizeof(ObDecBandState), sizeof(int32_t), d->n, cudaMemcpyDeviceToHost, d->stream));
  OB_CUDA(cudaStreamSynchronize(d->stream));
  if (request == 4033) {   // a stream that has not decoded a packet reports 0 (src/opus_decoder.c:1143-1146)
    std::vector<uint32_t> fl((size_t)d->n);
    OB_CUDA(cudaMemcpy2D(fl.data(), sizeof(uint32_t), &d->bstate[0].plc_flags, sizeof(ObDecBandState), sizeof(uint32_t), d->n, cudaMemcpyDeviceToHost));
    for (int s = 0; s < d->n; s++) if (!(fl[s] & OB_PF_HAS_PACKET)) values[s] = 0;
  }
  return OPUSB200_OK;
}

int opusb200_decoder_set_timing(OpusB200Decoder *d, int enable) {
  if (!d) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(d->device));
  if (enable && !d->ev[0]) for (int i = 0; i < 4; i++) OB_CUDA(cudaEventCreate(&d->ev[i]));
  d->timing = enable ? 1 : 0;
  d->kernel_ms[0] = d->kernel_ms[1] = d->kernel_ms[2] = 0.0;
  d->timed_launches = 0;
  return OPUSB200_OK;
}
int opusb200_decoder_get_timing(OpusB200Decoder *d, double *ms_kernels, int64_t *frames) {
  if (!d || !ms_kernels || !frames) return OPUSB200_BAD_ARG;
  ms_kernels[0] = d->kernel_ms[0]; ms_kernels[1] = d->kernel_ms[1]; ms_kernels[2] = d->kernel_ms[2];
  *frames = d->timed_launches;
  return OPUSB200_OK;
}

// ------------------------------------------------------------------------------------------
// kernel-level operations (B3-level surface, batched)
// ------------------------------------------------------------------------------------------
struct OpusB200Ctx {
  int device; cudaStream_t stream; int max_smem_optin; int64_t launches;
  void *scratch[2]; size_t scratch_bytes[2];   // dnn temporaries (nnet_host.cuh)
  uint64_t scratch_gen;                        // bumped whenever a scratch buffer is reallocated: captured CUDA graphs hold the old pointers
  size_t umma_smem_set; int sms;               // per-device kernel attribute / SM count of the tcgen05 layer kernel (one context = one device)
};

int opusb200_ctx_create(int device, OpusB200Ctx **out) {
  if (!out) return OPUSB200_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    fprintf(stderr, "opus_b200: no usable CUDA device (there is no CPU fallback)\n");
    return OPUSB200_INTERNAL_ERROR;
  }
  OB_CUDA(cudaSetDevice(device));
  OpusB200Ctx *c = new (std::nothrow) OpusB200Ctx;
  if (!c) return OPUSB200_ALLOC_FAIL;
  c->device = device; c->launches = 0;
  c->scratch[0] = c->scratch[1] = nullptr; c->scratch_bytes[0] = c->scratch_bytes[1] = 0; c->scratch_gen = 0; c->umma_smem_set = 0; c->sms = 0;
  OB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  OB_CUDA(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  OB_CUDA(cudaFuncSetAttribute(ob_mdct_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
  OB_CUDA(cudaFuncSetAttribute(ob_mdct_forward_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MdctFastSmem)));
  OB_CUDA(cudaFuncSetAttribute(ob_mdct_forward_tma, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  OB_CUDA(cudaFuncSetAttribute(ob_celt_enc_analysis, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
  OB_CUDA(cudaFuncSetAttribute(ob_pitch_xcorr, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
  OB_CUDA(cudaFuncSetAttribute(ob_exp_rotation, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
  OB_CUDA(cudaFuncSetAttribute(ob_pvq_search, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
  *out = c;
  return OPUSB200_OK;
}
void opusb200_ctx_destroy(OpusB200Ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaFree(c->scratch[0]); cudaFree(c->scratch[1]);
  cudaStreamDestroy(c->stream);
  delete c;
}
int opusb200_ctx_sync(OpusB200Ctx *c) {
  if (!c) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}
void *opusb200_ctx_cuda_stream(OpusB200Ctx *c) { return c ? (void *)c->stream : nullptr; }

// device staging helper for the host-buffer variants of the ops
struct DevBuf {
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 4) == cudaSuccess ? 0 : -1; }
};

int opusb200_mdct_forward(OpusB200Ctx *c, const float *in, float *out, int n_items, int LM, int transient, int on_device) {
  if (!c || !in || !out || n_items <= 0 || LM < 0 || LM > 3) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  const int N = OB_SHORT_N << LM, in_stride = N + OB_OVERLAP, out_stride = N;
  DevBuf di, dout;
  const float *pin = in;
  float *pout = out;
  if (!on_device) {
    if (di.alloc(sizeof(float) * (size_t)n_items * in_stride) || dout.alloc(sizeof(float) * (size_t)n_items * out_stride)) return OPUSB200_ALLOC_FAIL;
    OB_CUDA(cudaMemcpyAsync(di.p, in, sizeof(float) * (size_t)n_items * in_stride, cudaMemcpyHostToDevice, c->stream));
    pin = (const float *)di.p; pout = (float *)dout.p;
  }
  MdctFwdParams p;
  p.in = pin; p.out = pout; p.n_items = n_items; p.in_stride = in_stride; p.out_stride = out_stride; p.LM = LM; p.transient = transient;
  // the long block (LM = 3, not transient) has the register-FFT kernel with bulk-copy staging; 16-byte aligned inputs (cp.async.bulk) and
  // 8-byte aligned outputs hold for every cudaMalloc'ed buffer (strides 1080 / 960 floats).  OPUSB200_MDCT_SMEM_FFT=1 forces the general kernel.
  static const bool force_general = getenv("OPUSB200_MDCT_SMEM_FFT") != nullptr;
  if (LM == 3 && !transient && !force_general && ((uintptr_t)pin & 15) == 0 && ((uintptr_t)pout & 7) == 0) {
    if (!c->sms) cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, c->device);
    const int npairs = (n_items + 1) / 2, want = (npairs + kMdctFastWarps - 1) / kMdctFastWarps;
    ob_mdct_forward_tma<<<OB_IMIN(want, 2 * c->sms), kMdctFastWarps * 32, sizeof(MdctFastSmem), c->stream>>>(p);
  } else {
    const int warps = 4;
    ob_mdct_forward<<<(n_items + warps - 1) / warps, warps * 32, warps * sizeof(MdctWS), c->stream>>>(p);
  }
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(out, pout, sizeof(float) * (size_t)n_items * out_stride, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}

struct OpusB200CeltAnalysis {
  OpusB200Ctx *ctx;
  int n, channels;
  ObEncAnalysisState *state;
};
int opusb200_celt_analysis_reset(OpusB200CeltAnalysis *a) {
  if (!a) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(a->ctx->device));
  OB_CUDA(cudaMemsetAsync(a->state, 0, sizeof(ObEncAnalysisState) * (size_t)a->n, a->ctx->stream));
  return OPUSB200_OK;
}
int opusb200_celt_analysis_create(OpusB200Ctx *c, int n_streams, int channels, OpusB200CeltAnalysis **out) {
  if (!c || !out || n_streams <= 0 || (channels != 1 && channels != 2)) return OPUSB200_BAD_ARG;
  *out = nullptr;
  OB_CUDA(cudaSetDevice(c->device));
  OpusB200CeltAnalysis *a = new (std::nothrow) OpusB200CeltAnalysis;
  if (!a) return OPUSB200_ALLOC_FAIL;
  a->ctx = c; a->n = n_streams; a->channels = channels; a->state = nullptr;
  if (cudaMalloc(&a->state, sizeof(ObEncAnalysisState) * (size_t)n_streams) != cudaSuccess) { delete a; return OPUSB200_ALLOC_FAIL; }
  const int r = opusb200_celt_analysis_reset(a);
  if (r != OPUSB200_OK) { cudaFree(a->state); delete a; return r; }
  *out = a;
  return OPUSB200_OK;
}
void opusb200_celt_analysis_destroy(OpusB200CeltAnalysis *a) {
  if (!a) return;
  cudaSetDevice(a->ctx->device);
  cudaStreamSynchronize(a->ctx->stream);
  cudaFree(a->state);
  delete a;
}
int opusb200_celt_analysis_process(OpusB200CeltAnalysis *a, const float *pcm, const int32_t *short_blocks, float *X, float *bandE,
                                   float *bandLogE, int frame_size, int on_device) {
  if (!a || !pcm || !X || !bandE || !bandLogE) return OPUSB200_BAD_ARG;
  int LM = -1;
  for (int l = 0; l <= 3; l++) if (frame_size == OB_SHORT_N << l) LM = l;
  if (LM < 0) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = a->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  const int C = a->channels;
  const size_t pb = sizeof(float) * (size_t)a->n * frame_size * C, eb = sizeof(float) * (size_t)a->n * C * OB_NBANDS, sb = sizeof(int32_t) * (size_t)a->n;
  DevBuf dp, ds, dX, dE, dL;
  EncAnalysisParams p;
  p.state = a->state; p.pcm = pcm; p.short_blocks = short_blocks; p.X = X; p.bandE = bandE; p.bandLogE = bandLogE;
  p.n_streams = a->n; p.C = C; p.LM = LM;
  if (!on_device) {
    if (dp.alloc(pb) || dX.alloc(pb) || dE.alloc(eb) || dL.alloc(eb) || (short_blocks && ds.alloc(sb))) return OPUSB200_ALLOC_FAIL;
    OB_CUDA(cudaMemcpyAsync(dp.p, pcm, pb, cudaMemcpyHostToDevice, c->stream));
    if (short_blocks) OB_CUDA(cudaMemcpyAsync(ds.p, short_blocks, sb, cudaMemcpyHostToDevice, c->stream));
    p.pcm = (const float *)dp.p; p.short_blocks = short_blocks ? (const int32_t *)ds.p : nullptr;
    p.X = (float *)dX.p; p.bandE = (float *)dE.p; p.bandLogE = (float *)dL.p;
  }
  p.info = nullptr; p.complexity = 0; p.lsb_depth = 24;
  ob_celt_enc_analysis<<<a->n, 32 * C, C * sizeof(MdctWS) + 32, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(X, p.X, pb, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaMemcpyAsync(bandE, p.bandE, eb, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaMemcpyAsync(bandLogE, p.bandLogE, eb, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}

int opusb200_pitch_xcorr(OpusB200Ctx *c, const float *x, const float *y, float *xcorr, int n_items, int len, int max_pitch,
                         int on_device) {
  if (!c || !x || !y || !xcorr || n_items <= 0 || len <= 0 || max_pitch <= 0) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  const int smem_per_warp = (((len + 3) & ~3) + len + max_pitch + 4) * (int)sizeof(float);
  if (4 * smem_per_warp > c->max_smem_optin) return OPUSB200_BAD_ARG;
  DevBuf dx, dy, dc;
  const float *px = x, *py = y;
  float *pc = xcorr;
  if (!on_device) {
    if (dx.alloc(sizeof(float) * (size_t)n_items * len) || dy.alloc(sizeof(float) * (size_t)n_items * (len + max_pitch)) ||
        dc.alloc(sizeof(float) * (size_t)n_items * max_pitch)) return OPUSB200_ALLOC_FAIL;
    OB_CUDA(cudaMemcpyAsync(dx.p, x, sizeof(float) * (size_t)n_items * len, cudaMemcpyHostToDevice, c->stream));
    OB_CUDA(cudaMemcpyAsync(dy.p, y, sizeof(float) * (size_t)n_items * (len + max_pitch), cudaMemcpyHostToDevice, c->stream));
    px = (const float *)dx.p; py = (const float *)dy.p; pc = (float *)dc.p;
  }
  XcorrParams p;
  p.x = px; p.y = py; p.xcorr = pc; p.n_items = n_items; p.len = len; p.max_pitch = max_pitch;
  ob_pitch_xcorr<<<(n_items + 3) / 4, 128, 4 * smem_per_warp, c->stream>>>(p, smem_per_warp);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(xcorr, pc, sizeof(float) * (size_t)n_items * max_pitch, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}

static int vq_common(OpusB200Ctx *c, int which, float *X, int *iy, float *yy, const int *K, int n_items, int N, int dir,
                     int stride, int spread, int on_device) {
  if (!c || !X || !K || n_items <= 0 || N < 2 || N > 176) return OPUSB200_BAD_ARG;
  if (which == 1 && (!iy || !yy)) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  DevBuf dX, dI, dY, dK;
  float *pX = X, *pY = yy;
  int *pI = iy;
  const int *pK = K;
  const size_t xb = sizeof(float) * (size_t)n_items * N;
  if (!on_device) {
    if (dX.alloc(xb) || dK.alloc(sizeof(int) * (size_t)n_items) || dI.alloc(xb) || dY.alloc(sizeof(float) * (size_t)n_items)) return OPUSB200_ALLOC_FAIL;
    OB_CUDA(cudaMemcpyAsync(dX.p, X, xb, cudaMemcpyHostToDevice, c->stream));
    OB_CUDA(cudaMemcpyAsync(dK.p, K, sizeof(int) * (size_t)n_items, cudaMemcpyHostToDevice, c->stream));
    pX = (float *)dX.p; pK = (const int *)dK.p; pI = (int *)dI.p; pY = (float *)dY.p;
  }
  VqParams p;
  p.X = pX; p.iy = pI; p.yy = pY; p.K = pK; p.n_items = n_items; p.N = N; p.dir = dir; p.stride = stride; p.spread = spread;
  const int blocks = (n_items + 31) / 32;
  if (which == 0) ob_exp_rotation<<<blocks, 32, (size_t)N * 32 * sizeof(float), c->stream>>>(p);
  else ob_pvq_search<<<blocks, 32, (size_t)N * 32 * (2 * sizeof(float) + 1), c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(X, pX, xb, cudaMemcpyDeviceToHost, c->stream));
    if (which == 1) {
      OB_CUDA(cudaMemcpyAsync(iy, pI, xb, cudaMemcpyDeviceToHost, c->stream));
      OB_CUDA(cudaMemcpyAsync(yy, pY, sizeof(float) * (size_t)n_items, cudaMemcpyDeviceToHost, c->stream));
    }
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}
int opusb200_exp_rotation(OpusB200Ctx *c, float *X, const int *K, int n_items, int N, int dir, int stride, int spread, int on_device) {
  if (stride < 1 || (N % stride) || (dir != 1 && dir != -1) || spread < 0 || spread > 3) return OPUSB200_BAD_ARG;
  return vq_common(c, 0, X, nullptr, nullptr, K, n_items, N, dir, stride, spread, on_device);
}
int opusb200_pvq_search(OpusB200Ctx *c, float *X, int *iy, float *yy, const int *K, int n_items, int N, int on_device) {
  return vq_common(c, 1, X, iy, yy, K, n_items, N, 1, 1, 0, on_device);
}

int opusb200_silk_lpc_analysis_filter(OpusB200Ctx *c, int16_t *out, const int16_t *in, const int16_t *B, int n_items, int len,
                                      int d, int on_device) {
  if (!c || !out || !in || !B || n_items <= 0 || d < 6 || (d & 1) || d > 24 || d > len) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  DevBuf di, dob, db;
  const int16_t *pin = in, *pB = B;
  int16_t *pout = out;
  const size_t nb = sizeof(int16_t) * (size_t)n_items * len, bb = sizeof(int16_t) * (size_t)n_items * d;
  if (!on_device) {
    if (di.alloc(nb) || dob.alloc(nb) || db.alloc(bb)) return OPUSB200_ALLOC_FAIL;
    OB_CUDA(cudaMemcpyAsync(di.p, in, nb, cudaMemcpyHostToDevice, c->stream));
    OB_CUDA(cudaMemcpyAsync(db.p, B, bb, cudaMemcpyHostToDevice, c->stream));
    pin = (const int16_t *)di.p; pB = (const int16_t *)db.p; pout = (int16_t *)dob.p;
  }
  SilkLpcParams p;
  p.out = pout; p.in = pin; p.B = pB; p.n_items = n_items; p.len = len; p.d = d;
  ob_silk_lpc_analysis_filter<<<(n_items + 127) / 128, 128, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(out, pout, nb, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}

int opusb200_silk_burg_modified(OpusB200Ctx *c, float *A, float *nrg, const float *x, float minInvGain, int n_items,
                                int subfr_length, int nb_subfr, int D, int on_device) {
  if (!c || !A || !nrg || !x || n_items <= 0 || D < 2 || D > 24 || nb_subfr < 1 || subfr_length <= D ||
      subfr_length * nb_subfr > 384) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  DevBuf dA, dn, dx;
  float *pA = A, *pn = nrg;
  const float *px = x;
  const size_t xb = sizeof(float) * (size_t)n_items * subfr_length * nb_subfr, ab = sizeof(float) * (size_t)n_items * D;
  if (!on_device) {
    if (dA.alloc(ab) || dn.alloc(sizeof(float) * (size_t)n_items) || dx.alloc(xb)) return OPUSB200_ALLOC_FAIL;
    OB_CUDA(cudaMemcpyAsync(dx.p, x, xb, cudaMemcpyHostToDevice, c->stream));
    pA = (float *)dA.p; pn = (float *)dn.p; px = (const float *)dx.p;
  }
  SilkBurgParams p;
  p.A = pA; p.nrg = pn; p.x = px; p.minInvGain = minInvGain; p.n_items = n_items; p.subfr_length = subfr_length;
  p.nb_subfr = nb_subfr; p.D = D;
  ob_silk_burg_modified<<<(n_items + 127) / 128, 128, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(A, pA, ab, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaMemcpyAsync(nrg, pn, sizeof(float) * (size_t)n_items, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}

int opusb200_silk_autocorrelation(OpusB200Ctx *c, float *results, const float *x, int n_items, int size, int count, int on_device) {
  if (!c || !results || !x || n_items <= 0 || size <= 0 || count <= 0) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  DevBuf dr, dx;
  float *pr = results;
  const float *px = x;
  const size_t xb = sizeof(float) * (size_t)n_items * size, rb = sizeof(float) * (size_t)n_items * count;
  if (!on_device) {
    if (dr.alloc(rb) || dx.alloc(xb)) return OPUSB200_ALLOC_FAIL;
    OB_CUDA(cudaMemcpyAsync(dx.p, x, xb, cudaMemcpyHostToDevice, c->stream));
    pr = (float *)dr.p; px = (const float *)dx.p;
  }
  SilkAutocorrParams p;
  p.results = pr; p.x = px; p.n_items = n_items; p.size = size; p.count = count;
  const long total = (long)n_items * count;
  ob_silk_autocorrelation<<<(unsigned)((total + 127) / 128), 128, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(results, pr, rb, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}

struct OpusB200SilkNsq {
  OpusB200Ctx *ctx;
  int n, frame_length;
  ObNsqCfg cfg;
  ObNsqState *state;
};
static_assert(sizeof(OpusB200NsqFrame) == sizeof(ObNsqFrame), "OpusB200NsqFrame layout");
static_assert(sizeof(OpusB200NsqConfig) == sizeof(ObNsqCfg), "OpusB200NsqConfig layout");

int opusb200_silk_nsq_reset(OpusB200SilkNsq *q) {
  if (!q) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(q->ctx->device));
  OB_CUDA(cudaMemsetAsync(q->state, 0, sizeof(ObNsqState) * (size_t)q->n, q->ctx->stream));
  ob_silk_nsq_reset<<<(q->n + 127) / 128, 128, 0, q->ctx->stream>>>(q->state, q->n);
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
int opusb200_silk_nsq_create(OpusB200Ctx *c, int n_streams, const OpusB200NsqConfig *cfg, OpusB200SilkNsq **out) {
  if (!c || !out || !cfg || n_streams <= 0) return OPUSB200_BAD_ARG;
  *out = nullptr;
  // the ranges silk/control_codec.c can set
  if ((cfg->fs_khz != 8 && cfg->fs_khz != 12 && cfg->fs_khz != 16) || (cfg->nb_subfr != 2 && cfg->nb_subfr != 4) ||
      (cfg->predictLPCOrder != 10 && cfg->predictLPCOrder != 16) || cfg->shapingLPCOrder < 2 || cfg->shapingLPCOrder > 24 ||
      (cfg->shapingLPCOrder & 1) || cfg->nStatesDelayedDecision < 1 || cfg->nStatesDelayedDecision > 4 || cfg->warping_Q16 < 0 ||
      cfg->warping_Q16 > 32767)
    return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  OpusB200SilkNsq *q = new (std::nothrow) OpusB200SilkNsq;
  if (!q) return OPUSB200_ALLOC_FAIL;
  q->ctx = c; q->n = n_streams; q->frame_length = 5 * cfg->fs_khz * cfg->nb_subfr;
  memcpy(&q->cfg, cfg, sizeof(q->cfg));
  q->state = nullptr;
  if (cudaMalloc(&q->state, sizeof(ObNsqState) * (size_t)n_streams) != cudaSuccess) { delete q; return OPUSB200_ALLOC_FAIL; }
  cudaFuncSetAttribute(nsq_dd_kernel(cfg->shapingLPCOrder), cudaFuncAttributeMaxDynamicSharedMemorySize, kNsqWarpSmemWords * 4);
  const int rc = opusb200_silk_nsq_reset(q);
  if (rc != OPUSB200_OK) { cudaFree(q->state); delete q; return rc; }
  *out = q;
  return OPUSB200_OK;
}
void opusb200_silk_nsq_destroy(OpusB200SilkNsq *q) {
  if (!q) return;
  cudaSetDevice(q->ctx->device);
  cudaStreamSynchronize(q->ctx->stream);
  cudaFree(q->state);
  delete q;
}
int opusb200_silk_nsq_process(OpusB200SilkNsq *q, OpusB200NsqFrame *frames, const int16_t *x16, int8_t *pulses, int on_device) {
  if (!q || !frames || !x16 || !pulses) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = q->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  DevBuf df, dx, dp;
  const size_t fb = sizeof(ObNsqFrame) * (size_t)q->n, xb = sizeof(int16_t) * (size_t)q->n * q->frame_length, pb = (size_t)q->n * q->frame_length;
  SilkNsqParams p;
  p.state = q->state; p.n_streams = q->n; p.cfg = q->cfg;
  p.frames = (ObNsqFrame *)frames; p.x16 = x16; p.pulses = pulses;
  if (!on_device) {
    if (df.alloc(fb) || dx.alloc(xb) || dp.alloc(pb)) return OPUSB200_ALLOC_FAIL;
    OB_CUDA(cudaMemcpyAsync(df.p, frames, fb, cudaMemcpyHostToDevice, c->stream));
    OB_CUDA(cudaMemcpyAsync(dx.p, x16, xb, cudaMemcpyHostToDevice, c->stream));
    p.frames = (ObNsqFrame *)df.p; p.x16 = (const int16_t *)dx.p; p.pulses = (int8_t *)dp.p;
  }
  if (q->cfg.nStates > 1 || q->cfg.warping_Q16 > 0)  // silk/float/wrappers_FLP.c:163
    nsq_dd_kernel(q->cfg.shapingLPCOrder)<<<(q->n + 7) / 8, 32, kNsqWarpSmemWords * 4, c->stream>>>(p);
  else
    ob_silk_nsq<<<(q->n + 63) / 64, 64, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(pulses, dp.p, pb, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaMemcpyAsync(frames, df.p, fb, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}
int opusb200_silk_nsq_get_state(OpusB200SilkNsq *q, int32_t *out) {
  if (!q || !out) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = q->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  DevBuf d;
  const size_t b = sizeof(int32_t) * 1325 * (size_t)q->n;
  if (d.alloc(b)) return OPUSB200_ALLOC_FAIL;
  ob_silk_nsq_export<<<q->n, 128, 0, c->stream>>>(q->state, (int32_t *)d.p, q->n);
  OB_CUDA(cudaGetLastError());
  OB_CUDA(cudaMemcpyAsync(out, d.p, b, cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}

struct OpusB200SilkResampler {
  OpusB200Ctx *ctx;
  int n, fs_in, fs_out;
  ObSilkResamplerState *state;
  SilkResamplerParams proto;
};

// silk_resampler_init (silk/resampler.c:77-180): for_enc = 1: 8/12/16/24/48 kHz -> 8/12/16 kHz; for_enc = 0: 8/12/16 kHz -> 8/12/16/24/48 kHz
int opusb200_silk_resampler_create_ex(OpusB200Ctx *c, int n_streams, int32_t Fs_in, int32_t Fs_out, int for_enc, OpusB200SilkResampler **out) {
  if (!c || !out || n_streams <= 0) return OPUSB200_BAD_ARG;
  *out = nullptr;
  SilkResamplerParams proto;
  if (silk_resampler_setup(proto, Fs_in, Fs_out, for_enc)) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  OpusB200SilkResampler *r = new (std::nothrow) OpusB200SilkResampler;
  if (!r) return OPUSB200_ALLOC_FAIL;
  r->ctx = c; r->n = n_streams; r->fs_in = Fs_in; r->fs_out = Fs_out;
  OB_CUDA(cudaMalloc(&r->state, sizeof(ObSilkResamplerState) * (size_t)n_streams));
  OB_CUDA(cudaMemsetAsync(r->state, 0, sizeof(ObSilkResamplerState) * (size_t)n_streams, c->stream));
  r->proto = proto;
  r->proto.state = r->state; r->proto.n_streams = n_streams;
  *out = r;
  return OPUSB200_OK;
}
int opusb200_silk_resampler_create(OpusB200Ctx *c, int n_streams, int32_t Fs_in, int32_t Fs_out, OpusB200SilkResampler **out) {
  return opusb200_silk_resampler_create_ex(c, n_streams, Fs_in, Fs_out, 1, out);
}
void opusb200_silk_resampler_destroy(OpusB200SilkResampler *r) {
  if (!r) return;
  cudaSetDevice(r->ctx->device);
  cudaStreamSynchronize(r->ctx->stream);
  cudaFree(r->state);
  delete r;
}
int opusb200_silk_resampler_reset(OpusB200SilkResampler *r) {
  if (!r) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(r->ctx->device));
  OB_CUDA(cudaMemsetAsync(r->state, 0, sizeof(ObSilkResamplerState) * (size_t)r->n, r->ctx->stream));
  return OPUSB200_OK;
}
int opusb200_silk_resampler_process(OpusB200SilkResampler *r, const int16_t *in, int16_t *out, int in_len, int on_device) {
  if (!r || !in || !out) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = r->ctx;
  const int khz = r->fs_in / 1000;
  if (in_len < khz || in_len % khz || in_len > 480 * 4) return OPUSB200_BAD_ARG;
  const int out_len = in_len / khz * (r->fs_out / 1000);
  OB_CUDA(cudaSetDevice(c->device));
  DevBuf di, dob;
  const int16_t *pin = in;
  int16_t *pout = out;
  const size_t ib = sizeof(int16_t) * (size_t)r->n * in_len, ob_ = sizeof(int16_t) * (size_t)r->n * out_len;
  if (!on_device) {
    if (di.alloc(ib) || dob.alloc(ob_)) return OPUSB200_ALLOC_FAIL;
    OB_CUDA(cudaMemcpyAsync(di.p, in, ib, cudaMemcpyHostToDevice, c->stream));
    pin = (const int16_t *)di.p; pout = (int16_t *)dob.p;
  }
  SilkResamplerParams p = r->proto;
  p.in = pin; p.out = pout; p.in_len = in_len; p.out_len = out_len;
  ob_silk_resampler_down<<<(r->n + 127) / 128, 128, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(out, pout, ob_, cudaMemcpyDeviceToHost, c->stream));
    OB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return OPUSB200_OK;
}

void *opusb200_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void opusb200_host_free(void *p) { if (p) cudaFreeHost(p); }

}  // extern "C"

#include "nnet_host.cuh"
#include "fargan_host.cuh"
#include "lpcnet_enc_host.cuh"
#include "dred_host.cuh"
#include "plc_host.cuh"
#include "silk_coder_host.cuh"
#include "celt_enc_host.cuh"
#include "dred_coding_host.cuh"
#include "opus_enc_host.cuh"
