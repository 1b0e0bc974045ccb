// silk_coder_host.cuh — kernel entry points and C ABI of the batched range-encoder object and the SILK excitation coder on it
// (silk_coder.cuh).  Included at the end of opus_b200.cu.
#pragma once
#include "silk_coder.cuh"

extern "C" __global__ void __launch_bounds__(128) ob_silk_coder_init(ob::SilkCoderParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_coder_init_thread(p, s);
}
extern "C" __global__ void __launch_bounds__(128) ob_silk_coder_icdf(ob::SilkCoderParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_coder_icdf_thread(p, s);
}
extern "C" __global__ void __launch_bounds__(128) ob_silk_encode_pulses(ob::SilkCoderParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_encode_pulses_thread(p, s);
}
extern "C" __global__ void __launch_bounds__(128) ob_silk_encode_indices(ob::SilkCoderParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_encode_indices_thread(p, s);
}
extern "C" __global__ void __launch_bounds__(128) ob_silk_coder_done(ob::SilkCoderParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_coder_done_thread(p, s);
}

struct OpusB200RangeEncoder {
  OpusB200Ctx *ctx;
  int n, stride, storage;
  ob::ObRangeEncState *state;
  unsigned char *buf;
  int32_t *d_i32[2];
  int8_t *d_pulses; size_t d_pulses_cap;
  uint32_t *d_u32[2];
  int32_t *silk_prev;          // [n][2]: ec_prevSignalType, ec_prevLagIndex
  ob::ObSilkIndices *d_idx;
};
static ob::SilkCoderParams rc_params(OpusB200RangeEncoder *r) {
  ob::SilkCoderParams p;
  memset(&p, 0, sizeof(p));
  p.state = r->state; p.buf = r->buf; p.n = r->n; p.buf_stride = r->stride; p.storage = r->storage;
  return p;
}
// n independent range encoders (ec_enc, celt/entcode.h:62-88) with buffers of buf_bytes (<= 1275 + slack) each
extern "C" int opusb200_range_encoder_create(OpusB200Ctx *c, int n_streams, int buf_bytes, OpusB200RangeEncoder **out) {
  if (!out) return OPUSB200_BAD_ARG;
  *out = nullptr;
  if (!c || n_streams <= 0 || buf_bytes < 2 || buf_bytes > 4096) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  OpusB200RangeEncoder *r = new (std::nothrow) OpusB200RangeEncoder;
  if (!r) return OPUSB200_ALLOC_FAIL;
  memset(r, 0, sizeof(*r));
  r->ctx = c; r->n = n_streams; r->stride = (buf_bytes + 15) & ~15; r->storage = buf_bytes;
  const size_t n = (size_t)n_streams;
  if (cudaMalloc(&r->state, sizeof(ob::ObRangeEncState) * n) != cudaSuccess || cudaMalloc(&r->buf, n * r->stride) != cudaSuccess ||
      cudaMalloc(&r->d_i32[0], sizeof(int32_t) * n) != cudaSuccess || cudaMalloc(&r->d_i32[1], sizeof(int32_t) * n) != cudaSuccess ||
      cudaMalloc(&r->d_u32[0], sizeof(uint32_t) * n) != cudaSuccess || cudaMalloc(&r->d_u32[1], sizeof(uint32_t) * n) != cudaSuccess ||
      cudaMalloc(&r->silk_prev, sizeof(int32_t) * 2 * n) != cudaSuccess || cudaMalloc(&r->d_idx, sizeof(ob::ObSilkIndices) * n) != cudaSuccess) {
    cudaFree(r->silk_prev); cudaFree(r->d_idx);
    cudaFree(r->state); cudaFree(r->buf); cudaFree(r->d_i32[0]); cudaFree(r->d_i32[1]); cudaFree(r->d_u32[0]); cudaFree(r->d_u32[1]);
    delete r;
    return OPUSB200_ALLOC_FAIL;
  }
  cudaMemsetAsync(r->silk_prev, 0, sizeof(int32_t) * 2 * n, c->stream);
  *out = r;
  return OPUSB200_OK;
}
extern "C" void opusb200_range_encoder_destroy(OpusB200RangeEncoder *r) {
  if (!r) return;
  cudaSetDevice(r->ctx->device);
  cudaStreamSynchronize(r->ctx->stream);
  cudaFree(r->silk_prev); cudaFree(r->d_idx);
  cudaFree(r->state); cudaFree(r->buf); cudaFree(r->d_i32[0]); cudaFree(r->d_i32[1]); cudaFree(r->d_u32[0]); cudaFree(r->d_u32[1]); cudaFree(r->d_pulses);
  delete r;
}
// ec_enc_init(&enc, buf, storage) for every stream (celt/entenc.c:108)
extern "C" int opusb200_range_encoder_init(OpusB200RangeEncoder *r, int storage) {
  if (!r || storage < 2 || storage > r->stride) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(r->ctx->device));
  r->storage = storage;
  ob::SilkCoderParams p = rc_params(r);
  ob_silk_coder_init<<<(r->n + 127) / 128, 128, 0, r->ctx->stream>>>(p);
  r->ctx->launches++;
  OB_CUDA(cudaGetLastError());
  return OPUSB200_OK;
}
// ec_enc_icdf(&enc, symbols[s], icdf, ftb) (celt/entenc.c:195): one symbol per stream from a table of icdf_len bytes; symbols[s] < 0 skips the stream
extern "C" int opusb200_range_encoder_icdf(OpusB200RangeEncoder *r, const int32_t *symbols, const unsigned char *icdf, int icdf_len, int ftb, int on_device) {
  if (!r || !symbols || !icdf || icdf_len <= 0 || icdf_len > 64 || ftb < 1 || ftb > 15) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = r->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkCoderParams p = rc_params(r);
  memcpy(p.icdf, icdf, icdf_len);
  p.ftb = ftb; p.symbols = symbols;
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(r->d_i32[0], symbols, sizeof(int32_t) * (size_t)r->n, cudaMemcpyHostToDevice, c->stream));
    p.symbols = r->d_i32[0];
  }
  ob_silk_coder_icdf<<<(r->n + 127) / 128, 128, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}
// silk_encode_pulses(&enc, signalType, quantOffsetType, pulses, frame_length) (silk/encode_pulses.c:59): signal_type / quant_offset_type [n],
// pulses [n][pulses_stride] int8 (pulses_stride >= frame_length), frame_length 80..320 (a multiple of 16, or 120 = 10 ms at 12 kHz)
extern "C" int opusb200_silk_encode_pulses(OpusB200RangeEncoder *r, const int32_t *signal_type, const int32_t *quant_offset_type, const int8_t *pulses,
                                           int pulses_stride, int frame_length, int on_device) {
  if (!r || !signal_type || !quant_offset_type || !pulses || frame_length < 16 || frame_length > 320 || pulses_stride < frame_length) return OPUSB200_BAD_ARG;
  if ((frame_length & 15) && frame_length != 120) return OPUSB200_BAD_ARG;   // celt_assert(frame_length == 12 * 10)
  OpusB200Ctx *c = r->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkCoderParams p = rc_params(r);
  p.signal_type = signal_type; p.quant_offset_type = quant_offset_type; p.pulses = pulses; p.pulses_stride = pulses_stride; p.frame_length = frame_length;
  if (!on_device) {
    const size_t pb = (size_t)r->n * pulses_stride;
    if (r->d_pulses_cap < pb) { cudaFree(r->d_pulses); r->d_pulses = nullptr; r->d_pulses_cap = 0; OB_CUDA(cudaMalloc(&r->d_pulses, pb)); r->d_pulses_cap = pb; }
    OB_CUDA(cudaMemcpyAsync(r->d_i32[0], signal_type, sizeof(int32_t) * (size_t)r->n, cudaMemcpyHostToDevice, c->stream));
    OB_CUDA(cudaMemcpyAsync(r->d_i32[1], quant_offset_type, sizeof(int32_t) * (size_t)r->n, cudaMemcpyHostToDevice, c->stream));
    OB_CUDA(cudaMemcpyAsync(r->d_pulses, pulses, pb, cudaMemcpyHostToDevice, c->stream));
    p.signal_type = r->d_i32[0]; p.quant_offset_type = r->d_i32[1]; p.pulses = r->d_pulses;
  }
  ob_silk_encode_pulses<<<(r->n + 127) / 128, 128, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}
// silk_encode_indices(psEncC, &enc, 0, 0, condCoding) (silk/encode_indices.c:35): indices [n] = SideInfoIndices of every stream, cond_coding [n];
// fs_kHz 8 / 12 / 16 and nb_subfr 2 / 4 select the tables as silk_control_encoder does.  The encoder-state fields the function carries from frame
// to frame (ec_prevSignalType, ec_prevLagIndex) live in the object; opusb200_silk_coder_reset_state clears them (silk_init_encoder).
extern "C" int opusb200_silk_encode_indices(OpusB200RangeEncoder *r, int fs_kHz, int nb_subfr, const OpusB200SilkIndices *indices, const int32_t *cond_coding,
                                            int on_device) {
  if (!r || !indices || !cond_coding || (fs_kHz != 8 && fs_kHz != 12 && fs_kHz != 16) || (nb_subfr != 2 && nb_subfr != 4)) return OPUSB200_BAD_ARG;
  static_assert(sizeof(OpusB200SilkIndices) == sizeof(ob::ObSilkIndices), "OpusB200SilkIndices mirrors SideInfoIndices");
  OpusB200Ctx *c = r->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkCoderParams p = rc_params(r);
  p.indices = (const ob::ObSilkIndices *)indices; p.cond_coding = cond_coding; p.silk_prev = r->silk_prev; p.fs_kHz = fs_kHz; p.nb_subfr = nb_subfr;
  if (!on_device) {
    OB_CUDA(cudaMemcpyAsync(r->d_idx, indices, sizeof(ob::ObSilkIndices) * (size_t)r->n, cudaMemcpyHostToDevice, c->stream));
    OB_CUDA(cudaMemcpyAsync(r->d_i32[0], cond_coding, sizeof(int32_t) * (size_t)r->n, cudaMemcpyHostToDevice, c->stream));
    p.indices = r->d_idx; p.cond_coding = r->d_i32[0];
  }
  ob_silk_encode_indices<<<(r->n + 127) / 128, 128, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  if (!on_device) OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}
extern "C" int opusb200_silk_coder_reset_state(OpusB200RangeEncoder *r) {
  if (!r) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(r->ctx->device));
  OB_CUDA(cudaMemsetAsync(r->silk_prev, 0, sizeof(int32_t) * 2 * (size_t)r->n, r->ctx->stream));
  return OPUSB200_OK;
}
// ec_tell_frac, then ec_enc_done (celt/entenc.c:256): bytes [n][storage] (host), tell_frac [n] (bits in 1/8 units before the flush), final_range [n]
// (OPUS_GET_FINAL_RANGE), error [n] (the encoder's error flag: buffer too small)
extern "C" int opusb200_range_encoder_done(OpusB200RangeEncoder *r, unsigned char *bytes, uint32_t *tell_frac, uint32_t *final_range, int32_t *error) {
  if (!r || !bytes) return OPUSB200_BAD_ARG;
  OpusB200Ctx *c = r->ctx;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkCoderParams p = rc_params(r);
  p.tell_frac = r->d_u32[0]; p.final_range = r->d_u32[1];
  ob_silk_coder_done<<<(r->n + 127) / 128, 128, 0, c->stream>>>(p);
  c->launches++;
  OB_CUDA(cudaGetLastError());
  OB_CUDA(cudaMemcpy2DAsync(bytes, r->storage, r->buf, r->stride, r->storage, r->n, cudaMemcpyDeviceToHost, c->stream));
  if (tell_frac) OB_CUDA(cudaMemcpyAsync(tell_frac, r->d_u32[0], sizeof(uint32_t) * (size_t)r->n, cudaMemcpyDeviceToHost, c->stream));
  if (final_range) OB_CUDA(cudaMemcpyAsync(final_range, r->d_u32[1], sizeof(uint32_t) * (size_t)r->n, cudaMemcpyDeviceToHost, c->stream));
  if (error) OB_CUDA(cudaMemcpy2DAsync(error, sizeof(int32_t), &r->state->error, sizeof(ob::ObRangeEncState), sizeof(int32_t), r->n, cudaMemcpyDeviceToHost, c->stream));
  OB_CUDA(cudaStreamSynchronize(c->stream));
  return OPUSB200_OK;
}

// ---- silk_process_gains_FLP for n streams (silk_ops.cuh: silk_process_gains_thread): frames [n] records, in place ----
extern "C" __global__ void __launch_bounds__(128) ob_silk_process_gains(ob::SilkGainsParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_process_gains_thread(p, s);
}
extern "C" int opusb200_silk_process_gains(OpusB200Ctx *c, OpusB200SilkGains *frames, int n, int nb_subfr, int subfr_length, int nStatesDelayedDecision,
                                           int on_device) {
  static_assert(sizeof(OpusB200SilkGains) == sizeof(ob::ObSilkGains), "OpusB200SilkGains mirrors ObSilkGains");
  if (!c || !frames || n <= 0 || (nb_subfr != 2 && nb_subfr != 4) || subfr_length <= 0 || nStatesDelayedDecision < 1 || nStatesDelayedDecision > 4)
    return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkGainsParams p;
  p.n = n; p.nb_subfr = nb_subfr; p.subfr_length = subfr_length; p.nStatesDelayedDecision = nStatesDelayedDecision;
  const size_t bytes = sizeof(ob::ObSilkGains) * (size_t)n;
  ob::ObSilkGains *d = (ob::ObSilkGains *)frames;
  if (!on_device) {
    if (cudaMalloc(&d, bytes) != cudaSuccess) return OPUSB200_ALLOC_FAIL;
    if (cudaMemcpyAsync(d, frames, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { cudaFree(d); return OPUSB200_INTERNAL_ERROR; }
  }
  p.frames = d;
  ob_silk_process_gains<<<(n + 127) / 128, 128, 0, c->stream>>>(p);
  c->launches++;
  int r = cudaGetLastError() == cudaSuccess ? OPUSB200_OK : OPUSB200_INTERNAL_ERROR;
  if (!on_device) {
    if (r == OPUSB200_OK && cudaMemcpyAsync(frames, d, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && r == OPUSB200_OK) r = OPUSB200_INTERNAL_ERROR;
    cudaFree(d);
  }
  return r;
}

// ---- silk_process_NLSFs for n streams (silk_nlsf.cuh: silk_process_nlsfs_thread): frames [n] records, in place ----
#include "silk_nlsf.cuh"
extern "C" __global__ void __launch_bounds__(64) ob_silk_process_nlsfs(ob::SilkNlsfParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_process_nlsfs_thread(p, s);
}
extern "C" int opusb200_silk_process_nlsfs(OpusB200Ctx *c, OpusB200SilkNlsf *frames, int n, int fs_kHz, int nb_subfr, int on_device) {
  static_assert(sizeof(OpusB200SilkNlsf) == sizeof(ob::ObSilkNlsf), "OpusB200SilkNlsf mirrors ObSilkNlsf");
  if (!c || !frames || n <= 0 || (fs_kHz != 8 && fs_kHz != 12 && fs_kHz != 16) || (nb_subfr != 2 && nb_subfr != 4)) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkNlsfParams p;
  p.n = n; p.wb = fs_kHz == 16; p.nb_subfr = nb_subfr;
  const size_t bytes = sizeof(ob::ObSilkNlsf) * (size_t)n;
  ob::ObSilkNlsf *d = (ob::ObSilkNlsf *)frames;
  if (!on_device) {
    if (cudaMalloc(&d, bytes) != cudaSuccess) return OPUSB200_ALLOC_FAIL;
    if (cudaMemcpyAsync(d, frames, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { cudaFree(d); return OPUSB200_INTERNAL_ERROR; }
  }
  p.frames = d;
  ob_silk_process_nlsfs<<<(n + 63) / 64, 64, 0, c->stream>>>(p);
  c->launches++;
  int r = cudaGetLastError() == cudaSuccess ? OPUSB200_OK : OPUSB200_INTERNAL_ERROR;
  if (!on_device) {
    if (r == OPUSB200_OK && cudaMemcpyAsync(frames, d, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && r == OPUSB200_OK) r = OPUSB200_INTERNAL_ERROR;
    cudaFree(d);
  }
  return r;
}

// ---- silk_find_LPC_FLP for n streams (silk_nlsf.cuh: silk_find_lpc_thread): frames [n] records in place, x [n][nb_subfr * (subfr_length + order)] ----
extern "C" __global__ void __launch_bounds__(64) ob_silk_find_lpc(ob::SilkFindLpcParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_find_lpc_thread(p, s);
}
extern "C" int opusb200_silk_find_lpc(OpusB200Ctx *c, OpusB200SilkFindLpc *frames, const float *x, int n, int order, int nb_subfr, int subfr_length,
                                      int on_device) {
  static_assert(sizeof(OpusB200SilkFindLpc) == sizeof(ob::ObSilkFindLpc), "OpusB200SilkFindLpc mirrors ObSilkFindLpc");
  if (!c || !frames || !x || n <= 0 || (order != 10 && order != 16) || (nb_subfr != 2 && nb_subfr != 4) || subfr_length <= order || subfr_length > 80)
    return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkFindLpcParams p;
  p.n = n; p.order = order; p.nb_subfr = nb_subfr; p.subfr_length = subfr_length;
  const size_t fb = sizeof(ob::ObSilkFindLpc) * (size_t)n, xb = sizeof(float) * (size_t)n * nb_subfr * (subfr_length + order);
  ob::ObSilkFindLpc *d = (ob::ObSilkFindLpc *)frames;
  float *dx = nullptr;
  if (!on_device) {
    if (cudaMalloc(&d, fb) != cudaSuccess) return OPUSB200_ALLOC_FAIL;
    if (cudaMalloc(&dx, xb) != cudaSuccess) { cudaFree(d); return OPUSB200_ALLOC_FAIL; }
    if (cudaMemcpyAsync(d, frames, fb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess || cudaMemcpyAsync(dx, x, xb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
      cudaFree(d); cudaFree(dx);
      return OPUSB200_INTERNAL_ERROR;
    }
  }
  p.frames = d; p.x = on_device ? x : dx;
  ob_silk_find_lpc<<<(n + 63) / 64, 64, 0, c->stream>>>(p);
  c->launches++;
  int r = cudaGetLastError() == cudaSuccess ? OPUSB200_OK : OPUSB200_INTERNAL_ERROR;
  if (!on_device) {
    if (r == OPUSB200_OK && cudaMemcpyAsync(frames, d, fb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && r == OPUSB200_OK) r = OPUSB200_INTERNAL_ERROR;
    cudaFree(d); cudaFree(dx);
  }
  return r;
}

// ---- silk_find_pred_coefs_FLP for n streams (silk_nlsf.cuh: silk_find_pred_coefs_thread) ----
extern "C" __global__ void __launch_bounds__(64) ob_silk_find_pred_coefs(ob::SilkPredCoefsParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_find_pred_coefs_thread(p, s);
}
extern "C" int opusb200_silk_find_pred_coefs(OpusB200Ctx *c, OpusB200SilkPredCoefs *frames, const float *res_pitch, const float *x, int n, int sig_stride,
                                             int frame_off, int fs_kHz, int nb_subfr, int on_device) {
  static_assert(sizeof(OpusB200SilkPredCoefs) == sizeof(ob::ObSilkPredCoefs), "OpusB200SilkPredCoefs mirrors ObSilkPredCoefs");
  if (!c || !frames || !res_pitch || !x || n <= 0 || (fs_kHz != 8 && fs_kHz != 12 && fs_kHz != 16) || (nb_subfr != 2 && nb_subfr != 4)) return OPUSB200_BAD_ARG;
  const int order = fs_kHz == 16 ? 16 : 10, subfr_length = 5 * fs_kHz;
  // history: the longest pitch lag (18 ms) + LTP_ORDER / 2 + order in front of the frame; behind it LTP_ORDER samples of the residual are read
  if (frame_off < 18 * fs_kHz + 2 + order || sig_stride < frame_off + nb_subfr * subfr_length + 5) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkPredCoefsParams p;
  p.n = n; p.sig_stride = sig_stride; p.frame_off = frame_off; p.order = order; p.nb_subfr = nb_subfr; p.subfr_length = subfr_length;
  const size_t fb = sizeof(ob::ObSilkPredCoefs) * (size_t)n, sb = sizeof(float) * (size_t)n * sig_stride;
  ob::ObSilkPredCoefs *d = (ob::ObSilkPredCoefs *)frames;
  float *dr = nullptr, *dx = nullptr;
  int r = OPUSB200_OK;
  if (!on_device) {
    if (cudaMalloc(&d, fb) != cudaSuccess) return OPUSB200_ALLOC_FAIL;
    if (cudaMalloc(&dr, sb) != cudaSuccess || cudaMalloc(&dx, sb) != cudaSuccess) { cudaFree(d); cudaFree(dr); return OPUSB200_ALLOC_FAIL; }
    if (cudaMemcpyAsync(d, frames, fb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess || cudaMemcpyAsync(dr, res_pitch, sb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaMemcpyAsync(dx, x, sb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
      r = OPUSB200_INTERNAL_ERROR;
  }
  if (r == OPUSB200_OK) {
    p.frames = d; p.res_pitch = on_device ? res_pitch : dr; p.x = on_device ? x : dx;
    ob_silk_find_pred_coefs<<<(n + 63) / 64, 64, 0, c->stream>>>(p);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
  }
  if (!on_device) {
    if (r == OPUSB200_OK && cudaMemcpyAsync(frames, d, fb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && r == OPUSB200_OK) r = OPUSB200_INTERNAL_ERROR;
    cudaFree(d); cudaFree(dr); cudaFree(dx);
  }
  return r;
}

// ---- silk_noise_shape_analysis_FLP for n streams (silk_nlsf.cuh: silk_noise_shape_thread) ----
extern "C" __global__ void __launch_bounds__(64) ob_silk_noise_shape_analysis(ob::SilkNoiseShapeParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_noise_shape_thread(p, s);
}
extern "C" int opusb200_silk_noise_shape_analysis(OpusB200Ctx *c, OpusB200SilkNoiseShape *frames, const float *pitch_res, int res_stride, const float *x, int x_stride,
                                                  int x_off, int n, int fs_kHz, int nb_subfr, int shapingLPCOrder, int on_device) {
  static_assert(sizeof(OpusB200SilkNoiseShape) == sizeof(ob::ObSilkNoiseShape), "OpusB200SilkNoiseShape mirrors ObSilkNoiseShape");
  if (!c || !frames || !pitch_res || !x || n <= 0 || (fs_kHz != 8 && fs_kHz != 12 && fs_kHz != 16) || (nb_subfr != 2 && nb_subfr != 4) || shapingLPCOrder < 2 ||
      shapingLPCOrder > 24 || (shapingLPCOrder & 1))
    return OPUSB200_BAD_ARG;
  const int frame = 5 * fs_kHz * nb_subfr, la = 5 * fs_kHz;
  if (res_stride < frame || x_off < la || x_stride < x_off + frame + la) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkNoiseShapeParams p;
  p.n = n; p.res_stride = res_stride; p.x_stride = x_stride; p.x_off = x_off; p.fs_kHz = fs_kHz; p.nb_subfr = nb_subfr; p.shapingLPCOrder = shapingLPCOrder;
  const size_t fb = sizeof(ob::ObSilkNoiseShape) * (size_t)n, rb = sizeof(float) * (size_t)n * res_stride, xb = sizeof(float) * (size_t)n * x_stride;
  ob::ObSilkNoiseShape *d = (ob::ObSilkNoiseShape *)frames;
  float *dr = nullptr, *dx = nullptr;
  int r = OPUSB200_OK;
  if (!on_device) {
    if (cudaMalloc(&d, fb) != cudaSuccess) return OPUSB200_ALLOC_FAIL;
    if (cudaMalloc(&dr, rb) != cudaSuccess || cudaMalloc(&dx, xb) != cudaSuccess) { cudaFree(d); cudaFree(dr); return OPUSB200_ALLOC_FAIL; }
    if (cudaMemcpyAsync(d, frames, fb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess || cudaMemcpyAsync(dr, pitch_res, rb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaMemcpyAsync(dx, x, xb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
      r = OPUSB200_INTERNAL_ERROR;
  }
  if (r == OPUSB200_OK) {
    p.frames = d; p.pitch_res = on_device ? pitch_res : dr; p.x = on_device ? x : dx;
    ob_silk_noise_shape_analysis<<<(n + 63) / 64, 64, 0, c->stream>>>(p);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
  }
  if (!on_device) {
    if (r == OPUSB200_OK && cudaMemcpyAsync(frames, d, fb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && r == OPUSB200_OK) r = OPUSB200_INTERNAL_ERROR;
    cudaFree(d); cudaFree(dr); cudaFree(dx);
  }
  return r;
}

// ---- silk_find_pitch_lags_FLP for n streams (silk_pitch.cuh: silk_find_pitch_lags_thread) ----
#include "silk_pitch.cuh"
extern "C" __global__ void __launch_bounds__(64) ob_silk_find_pitch_lags(ob::SilkPitchParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_find_pitch_lags_thread(p, s);
}
extern "C" int opusb200_silk_find_pitch_lags(OpusB200Ctx *c, OpusB200SilkPitch *frames, float *res, int res_stride, const float *x, int x_stride, int x_off, int n,
                                             int fs_kHz, int nb_subfr, int lpc_order, int complexity, int on_device) {
  static_assert(sizeof(OpusB200SilkPitch) == sizeof(ob::ObSilkPitch), "OpusB200SilkPitch mirrors ObSilkPitch");
  if (!c || !frames || !res || !x || n <= 0 || (fs_kHz != 8 && fs_kHz != 12 && fs_kHz != 16) || (nb_subfr != 2 && nb_subfr != 4) || complexity < 0 || complexity > 2 ||
      (lpc_order != 6 && lpc_order != 8 && lpc_order != 10 && lpc_order != 12 && lpc_order != 16))
    return OPUSB200_BAD_ARG;
  const int buf_len = (22 + 5 * nb_subfr) * fs_kHz;
  if (res_stride < buf_len || x_off < 20 * fs_kHz || x_stride < x_off + (5 * nb_subfr + 2) * fs_kHz) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkPitchParams p;
  p.n = n; p.x_stride = x_stride; p.x_off = x_off; p.res_stride = res_stride; p.fs_kHz = fs_kHz; p.nb_subfr = nb_subfr; p.lpc_order = lpc_order; p.complexity = complexity;
  const size_t fb = sizeof(ob::ObSilkPitch) * (size_t)n, rb = sizeof(float) * (size_t)n * res_stride, xb = sizeof(float) * (size_t)n * x_stride;
  ob::ObSilkPitch *d = (ob::ObSilkPitch *)frames;
  float *dr = res, *dx = nullptr;
  int r = OPUSB200_OK;
  if (!on_device) {
    if (cudaMalloc(&d, fb) != cudaSuccess) return OPUSB200_ALLOC_FAIL;
    dr = nullptr;
    if (cudaMalloc(&dr, rb) != cudaSuccess || cudaMalloc(&dx, xb) != cudaSuccess) { cudaFree(d); cudaFree(dr); return OPUSB200_ALLOC_FAIL; }
    if (cudaMemcpyAsync(d, frames, fb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess || cudaMemcpyAsync(dx, x, xb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
      r = OPUSB200_INTERNAL_ERROR;
  }
  if (r == OPUSB200_OK) {
    p.frames = d; p.res = dr; p.x = on_device ? x : dx;
    ob_silk_find_pitch_lags<<<(n + 63) / 64, 64, 0, c->stream>>>(p);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
  }
  if (!on_device) {
    if (r == OPUSB200_OK && (cudaMemcpyAsync(frames, d, fb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                             cudaMemcpyAsync(res, dr, rb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess))
      r = OPUSB200_INTERNAL_ERROR;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && r == OPUSB200_OK) r = OPUSB200_INTERNAL_ERROR;
    cudaFree(d); cudaFree(dr); cudaFree(dx);
  }
  return r;
}

// ---- silk_VAD_Init / silk_encode_do_VAD_FLP / silk_HP_variable_cutoff for n streams (silk_vad.cuh) ----
#include "silk_vad.cuh"
extern "C" __global__ void __launch_bounds__(128) ob_silk_encode_do_vad(ob::SilkVadParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_encode_do_vad_thread(p, s);
}
extern "C" __global__ void __launch_bounds__(128) ob_silk_hp_variable_cutoff(ob::SilkHpCutoffParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < p.n) ob::silk_hp_variable_cutoff_thread(p, s);
}
extern "C" int opusb200_silk_vad_init(OpusB200SilkVad *frames, int n) {   // silk_VAD_Init: constants only, written on the host into the caller's records
  static_assert(sizeof(OpusB200SilkVad) == sizeof(ob::ObSilkVad), "OpusB200SilkVad mirrors ObSilkVad");
  if (!frames || n <= 0) return OPUSB200_BAD_ARG;
  memset(frames, 0, sizeof(OpusB200SilkVad) * (size_t)n);
  for (int s = 0; s < n; s++) {
    for (int b = 0; b < 4; b++) {
      const int bias = 50 / (b + 1);   // VAD_NOISE_LEVELS_BIAS: approximately pink noise
      frames[s].NoiseLevelBias[b] = bias > 1 ? bias : 1;
      frames[s].NL[b] = 100 * frames[s].NoiseLevelBias[b];
      frames[s].inv_NL[b] = 0x7FFFFFFF / frames[s].NL[b];
      frames[s].NrgRatioSmth_Q8[b] = 100 * 256;   // 20 dB
    }
    frames[s].counter = 15;
  }
  return OPUSB200_OK;
}
extern "C" int opusb200_silk_encode_do_vad(OpusB200Ctx *c, OpusB200SilkVad *frames, const int16_t *pcm, int pcm_stride, const int32_t *activity, int n,
                                           int frame_length, int fs_kHz, int on_device) {
  if (!c || !frames || !pcm || n <= 0 || (fs_kHz != 8 && fs_kHz != 12 && fs_kHz != 16) || (frame_length != 10 * fs_kHz && frame_length != 20 * fs_kHz) ||
      pcm_stride < frame_length)
    return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkVadParams p;
  p.n = n; p.pcm_stride = pcm_stride; p.frame_length = frame_length; p.fs_kHz = fs_kHz;
  const size_t fb = sizeof(ob::ObSilkVad) * (size_t)n, xb = sizeof(int16_t) * (size_t)n * pcm_stride, ab = sizeof(int32_t) * (size_t)n;
  ob::ObSilkVad *d = (ob::ObSilkVad *)frames;
  int16_t *dx = nullptr;
  int32_t *da = nullptr;
  int r = OPUSB200_OK;
  if (!on_device) {
    if (cudaMalloc(&d, fb) != cudaSuccess) return OPUSB200_ALLOC_FAIL;
    if (cudaMalloc(&dx, xb) != cudaSuccess || (activity && cudaMalloc(&da, ab) != cudaSuccess)) { cudaFree(d); cudaFree(dx); return OPUSB200_ALLOC_FAIL; }
    if (cudaMemcpyAsync(d, frames, fb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess || cudaMemcpyAsync(dx, pcm, xb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        (activity && cudaMemcpyAsync(da, activity, ab, cudaMemcpyHostToDevice, c->stream) != cudaSuccess))
      r = OPUSB200_INTERNAL_ERROR;
  }
  if (r == OPUSB200_OK) {
    p.frames = d; p.pcm = on_device ? pcm : dx; p.activity = on_device ? activity : da;
    ob_silk_encode_do_vad<<<(n + 127) / 128, 128, 0, c->stream>>>(p);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
  }
  if (!on_device) {
    if (r == OPUSB200_OK && cudaMemcpyAsync(frames, d, fb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && r == OPUSB200_OK) r = OPUSB200_INTERNAL_ERROR;
    cudaFree(d); cudaFree(dx); cudaFree(da);
  }
  return r;
}
extern "C" int opusb200_silk_hp_variable_cutoff(OpusB200Ctx *c, OpusB200SilkVad *frames, int n, int fs_kHz, int on_device) {
  if (!c || !frames || n <= 0 || (fs_kHz != 8 && fs_kHz != 12 && fs_kHz != 16)) return OPUSB200_BAD_ARG;
  if (!on_device) for (int s = 0; s < n; s++) if (frames[s].prevSignalType == 2 && (frames[s].prevLag <= 0 || frames[s].prevLag > 32767)) return OPUSB200_BAD_ARG;
  OB_CUDA(cudaSetDevice(c->device));
  ob::SilkHpCutoffParams p;
  p.n = n; p.fs_kHz = fs_kHz;
  const size_t fb = sizeof(ob::ObSilkVad) * (size_t)n;
  ob::ObSilkVad *d = (ob::ObSilkVad *)frames;
  if (!on_device) {
    if (cudaMalloc(&d, fb) != cudaSuccess) return OPUSB200_ALLOC_FAIL;
    if (cudaMemcpyAsync(d, frames, fb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { cudaFree(d); return OPUSB200_INTERNAL_ERROR; }
  }
  p.frames = d;
  ob_silk_hp_variable_cutoff<<<(n + 127) / 128, 128, 0, c->stream>>>(p);
  c->launches++;
  int r = cudaGetLastError() == cudaSuccess ? OPUSB200_OK : OPUSB200_INTERNAL_ERROR;
  if (!on_device) {
    if (r == OPUSB200_OK && cudaMemcpyAsync(frames, d, fb, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) r = OPUSB200_INTERNAL_ERROR;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && r == OPUSB200_OK) r = OPUSB200_INTERNAL_ERROR;
    cudaFree(d);
  }
  return r;
}
