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

// ---- float analysis kernels (row a17): double accumulators, one thread per item -----------------
// silk/float/inner_product_FLP.c:35-59 (and energy_FLP.c): the products float*double are exact in
// double; sums keep the reference's grouping of four.
OB_DEV double silk_inner_product_flp(const float *a, const float *b, int n) {
  int i;
  double r = 0.0;
#pragma unroll 1
  for (i = 0; i < n - 3; i += 4)
    r += a[i + 0] * (double)b[i + 0] + a[i + 1] * (double)b[i + 1] + a[i + 2] * (double)b[i + 2] + a[i + 3] * (double)b[i + 3];
#pragma unroll 1
  for (; i < n; i++) r += a[i] * (double)b[i];
  return r;
}
struct SilkBurgParams {
  float *A;         // [n_items][D] out
  float *nrg;       // [n_items] out: residual energy
  const float *x;   // [n_items][nb_subfr*subfr_length]
  float minInvGain;
  int n_items, subfr_length, nb_subfr, D;
};
// silk/float/burg_modified_FLP.c:39-187
OB_DEV void silk_burg_thread(const SilkBurgParams &p, int item) {
  const int subfr_length = p.subfr_length, nb_subfr = p.nb_subfr, D = p.D;
  const float *x = p.x + (size_t)item * nb_subfr * subfr_length;
  float *A = p.A + (size_t)item * D;
  const double minInvGain = p.minInvGain;
  int reached_max_gain = 0;
  double Cf[24], Cl[24], CAf[25], CAb[25], Af[24];
  double C0 = silk_inner_product_flp(x, x, nb_subfr * subfr_length);
#pragma unroll 1
  for (int k = 0; k < 24; k++) Cf[k] = 0.0;
#pragma unroll 1
  for (int s = 0; s < nb_subfr; s++) {
    const float *xp = x + s * subfr_length;
#pragma unroll 1
    for (int n = 1; n < D + 1; n++) Cf[n - 1] += silk_inner_product_flp(xp, xp + n, subfr_length - n);
  }
#pragma unroll 1
  for (int k = 0; k < 24; k++) Cl[k] = Cf[k];
  CAb[0] = CAf[0] = C0 + 1e-5f * C0 + 1e-9f;
  double invGain = 1.0;
  int n;
#pragma unroll 1
  for (n = 0; n < D; n++) {
#pragma unroll 1
    for (int s = 0; s < nb_subfr; s++) {
      const float *xp = x + s * subfr_length;
      double tmp1 = xp[n], tmp2 = xp[subfr_length - n - 1];
#pragma unroll 1
      for (int k = 0; k < n; k++) {
        Cf[k] -= (double)(xp[n] * xp[n - k - 1]);                                // float product, as in the reference
        Cl[k] -= (double)(xp[subfr_length - n - 1] * xp[subfr_length - n + k]);  // float product
        const double Atmp = Af[k];
        tmp1 += xp[n - k - 1] * Atmp;
        tmp2 += xp[subfr_length - n + k] * Atmp;
      }
#pragma unroll 1
      for (int k = 0; k <= n; k++) {
        CAf[k] -= tmp1 * xp[n - k];
        CAb[k] -= tmp2 * xp[subfr_length - n + k - 1];
      }
    }
    double tmp1 = Cf[n], tmp2 = Cl[n];
#pragma unroll 1
    for (int k = 0; k < n; k++) {
      const double Atmp = Af[k];
      tmp1 += Cl[n - k - 1] * Atmp;
      tmp2 += Cf[n - k - 1] * Atmp;
    }
    CAf[n + 1] = tmp1;
    CAb[n + 1] = tmp2;
    double num = CAb[n + 1], nrg_b = CAb[0], nrg_f = CAf[0];
#pragma unroll 1
    for (int k = 0; k < n; k++) {
      const double Atmp = Af[k];
      num += CAb[n - k] * Atmp;
      nrg_b += CAb[k + 1] * Atmp;
      nrg_f += CAf[k + 1] * Atmp;
    }
    double rc = -2.0 * num / (nrg_f + nrg_b);
    tmp1 = invGain * (1.0 - rc * rc);
    if (tmp1 <= minInvGain) {
      rc = sqrt(1.0 - minInvGain / invGain);
      if (num > 0) rc = -rc;
      invGain = minInvGain;
      reached_max_gain = 1;
    } else invGain = tmp1;
#pragma unroll 1
    for (int k = 0; k < (n + 1) >> 1; k++) {
      const double t1 = Af[k], t2 = Af[n - k - 1];
      Af[k] = t1 + rc * t2;
      Af[n - k - 1] = t2 + rc * t1;
    }
    Af[n] = rc;
    if (reached_max_gain) {
#pragma unroll 1
      for (int k = n + 1; k < D; k++) Af[k] = 0.0;
      break;
    }
#pragma unroll 1
    for (int k = 0; k <= n + 1; k++) {
      const double t1 = CAf[k];
      CAf[k] += rc * CAb[n - k + 1];
      CAb[n - k + 1] += rc * t1;
    }
  }
  double nrg_f;
  if (reached_max_gain) {
#pragma unroll 1
    for (int k = 0; k < D; k++) A[k] = (float)(-Af[k]);
#pragma unroll 1
    for (int s = 0; s < nb_subfr; s++) C0 -= silk_inner_product_flp(x + s * subfr_length, x + s * subfr_length, D);
    nrg_f = C0 * invGain;
  } else {
    nrg_f = CAf[0];
    double t1 = 1.0;
#pragma unroll 1
    for (int k = 0; k < D; k++) {
      const double Atmp = Af[k];
      nrg_f += CAf[k + 1] * Atmp;
      t1 += Atmp * Atmp;
      A[k] = (float)(-Atmp);
    }
    nrg_f -= 1e-5f * C0 * t1;
  }
  p.nrg[item] = (float)nrg_f;
}

struct SilkAutocorrParams { float *results; const float *x; int n_items, size, count; };
// silk/float/autocorrelation_FLP.c:36-55; one thread per (item, lag)
OB_DEV void silk_autocorr_thread(const SilkAutocorrParams &p, int idx) {
  const int item = idx / p.count, lag = idx - item * p.count;
  const float *x = p.x + (size_t)item * p.size;
  p.results[idx] = lag < p.size ? (float)silk_inner_product_flp(x, x + lag, p.size - lag) : 0.f;
}

}  // namespace ob
