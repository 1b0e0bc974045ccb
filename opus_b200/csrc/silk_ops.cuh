// silk_ops.cuh — SILK integer kernels, one THREAD per stream: the
// filters are serial integer recurrences with wrap-around semantics (no scan form), so the
// parallel axis is the batch.  Thread-private arrays live in local memory, which the hardware
// interleaves per lane — a warp's accesses to element k of 32 streams coalesce.
//   ob_silk_lpc_analysis_filter : silk/LPC_analysis_filter.c:49-111
//   ob_silk_resampler_down      : silk/resampler.c:183-219 + resampler_private_AR2.c + resampler_private_down_FIR.c
//                                 (encoder-side integer-ratio down-sampling 1:2, 1:3, 1:4, 1:6)
#pragma once
#include "simt.cuh"

namespace ob {

OB_DEV int32_t silk_smulwb(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * (int16_t)b) >> 16); }
OB_DEV int32_t silk_smlawb(int32_t a, int32_t b, int32_t c) { return (int32_t)(a + (((int64_t)b * (int16_t)c) >> 16)); }
OB_DEV int32_t silk_rshift_round(int32_t a, int s) { return s == 1 ? (a >> 1) + (a & 1) : ((a >> (s - 1)) + 1) >> 1; }
OB_DEV int16_t silk_sat16(int32_t a) { return (int16_t)(a > 32767 ? 32767 : (a < -32768 ? -32768 : a)); }

struct SilkLpcParams {
  int16_t *out;        // [n_items][len]
  const int16_t *in;   // [n_items][len]
  const int16_t *B;    // [n_items][d]  Q12
  int n_items, len, d;
};
OB_DEV void silk_lpc_analysis_filter_thread(const SilkLpcParams &p, int item) {
  const int16_t *in = p.in + (size_t)item * p.len;
  const int16_t *B = p.B + (size_t)item * p.d;
  int16_t *out = p.out + (size_t)item * p.len;
  const int d = p.d;
  int16_t b[24];
#pragma unroll 1
  for (int j = 0; j < d; j++) b[j] = B[j];
#pragma unroll 1
  for (int ix = 0; ix < d; ix++) out[ix] = 0;
#pragma unroll 1
  for (int ix = d; ix < p.len; ix++) {
    const int16_t *q = &in[ix - 1];
    uint32_t acc = 0;
#pragma unroll 1
    for (int j = 0; j < d; j++) acc += (uint32_t)((int32_t)q[-j] * b[j]);
    acc = ((uint32_t)((int32_t)q[1]) << 12) - acc;
    out[ix] = silk_sat16(silk_rshift_round((int32_t)acc, 12));
  }
}

// Per-stream resampler state (HBM).  Mirrors silk_resampler_state_struct (silk/resampler_structs.h:34-50).
struct ObSilkResamplerState {
  int32_t sIIR[2];
  int32_t sFIR[36];
  int16_t delayBuf[48];
};
struct SilkResamplerParams {
  ObSilkResamplerState *state;  // [n_streams]
  const int16_t *in;            // [n_streams][in_len]
  int16_t *out;                 // [n_streams][out_len]
  int n_streams, in_len, out_len;
  int fir_order, fs_in_khz, fs_out_khz, input_delay, inv_ratio_q16, batch;
  int16_t coefs[20];            // 2 AR + up to 18 FIR coefficients
};

OB_DEV int16_t *silk_down_fir(const SilkResamplerParams &p, int32_t *sIIR, int32_t *sFIR, int16_t *out, const int16_t *in, int inLen) {
  int32_t buf[480 + 36];
  const int order = p.fir_order, half = order >> 1;
  int nIn = 0;
#pragma unroll 1
  for (int k = 0; k < order; k++) buf[k] = sFIR[k];
  int32_t s0 = sIIR[0], s1 = sIIR[1];
  const int32_t a0 = p.coefs[0], a1 = p.coefs[1];
#pragma unroll 1
  for (;;) {
    nIn = inLen < p.batch ? inLen : p.batch;
#pragma unroll 1
    for (int k = 0; k < nIn; k++) {
      int32_t o = s0 + (int32_t)((uint32_t)(int32_t)in[k] << 8);
      buf[order + k] = o;
      o = (int32_t)((uint32_t)o << 2);
      s0 = silk_smlawb(s1, o, a0);
      s1 = silk_smulwb(o, a1);
    }
    const int32_t maxi = nIn << 16;
#pragma unroll 1
    for (int32_t idx = 0; idx < maxi; idx += p.inv_ratio_q16) {
      const int32_t *b = buf + (idx >> 16);
      int32_t r = silk_smulwb(b[0] + b[order - 1], p.coefs[2]);
#pragma unroll 1
      for (int k = 1; k < half; k++) r = silk_smlawb(r, b[k] + b[order - 1 - k], p.coefs[2 + k]);
      *out++ = silk_sat16(silk_rshift_round(r, 6));
    }
    in += nIn;
    inLen -= nIn;
    if (inLen > 1) {
#pragma unroll 1
      for (int k = 0; k < order; k++) buf[k] = buf[nIn + k];
    } else break;
  }
#pragma unroll 1
  for (int k = 0; k < order; k++) sFIR[k] = buf[nIn + k];
  sIIR[0] = s0; sIIR[1] = s1;
  return out;
}

OB_DEV void silk_resampler_thread(const SilkResamplerParams &p, int stream) {
  ObSilkResamplerState *S = p.state + stream;
  const int16_t *in = p.in + (size_t)stream * p.in_len;
  int16_t *out = p.out + (size_t)stream * p.out_len;
  int32_t sIIR[2] = {S->sIIR[0], S->sIIR[1]};
  int32_t sFIR[36];
  int16_t delay[48];
#pragma unroll 1
  for (int k = 0; k < 36; k++) sFIR[k] = S->sFIR[k];
  const int nSamples = p.fs_in_khz - p.input_delay;
#pragma unroll 1
  for (int k = 0; k < p.input_delay; k++) delay[k] = S->delayBuf[k];
#pragma unroll 1
  for (int k = 0; k < nSamples; k++) delay[p.input_delay + k] = in[k];
  silk_down_fir(p, sIIR, sFIR, out, delay, p.fs_in_khz);
  silk_down_fir(p, sIIR, sFIR, out + p.fs_out_khz, in + nSamples, p.in_len - p.fs_in_khz);
#pragma unroll 1
  for (int k = 0; k < p.input_delay; k++) S->delayBuf[k] = in[p.in_len - p.input_delay + k];
#pragma unroll 1
  for (int k = 0; k < 36; k++) S->sFIR[k] = sFIR[k];
  S->sIIR[0] = sIIR[0]; S->sIIR[1] = sIIR[1];
}

}  // namespace ob
