// silk_ops.cuh — SILK integer kernels, one THREAD per stream: the
// filters are serial integer recurrences with wrap-around semantics (no scan form), so the
// parallel axis is the batch.  Thread-private arrays live in local memory, which the hardware
// interleaves per lane — a warp's accesses to element k of 32 streams coalesce.
//   ob_silk_lpc_analysis_filter : silk/LPC_analysis_filter.c:49-111
//   ob_silk_resampler_down      : silk_resampler (silk/resampler.c:77-219) for every rate pair silk_resampler_init accepts, both directions:
//                                 copy, resampler_private_up2_HQ.c, resampler_private_IIR_FIR.c, resampler_private_AR2.c + _down_FIR.c
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
  int32_t sIIR[6];
  int32_t sFIR[36];    // down-FIR: int32 history; IIR_FIR: the first 8 int16 of this storage (the reference's union)
  int16_t delayBuf[48];
};
enum { OB_RS_COPY = 0, OB_RS_UP2 = 1, OB_RS_IIR_FIR = 2, OB_RS_DOWN = 3 };   // S->resampler_function (silk/resampler.c:74-77)
struct SilkResamplerParams {
  ObSilkResamplerState *state;  // [n_streams]
  const int16_t *in;            // [n_streams][in_len]
  int16_t *out;                 // [n_streams][out_len]
  int n_streams, in_len, out_len;
  int function, fir_order, fir_fracs, fs_in_khz, fs_out_khz, input_delay, inv_ratio_q16, batch;
  int16_t coefs[30];            // 2 AR + up to 27 FIR coefficients (3 phases x 9)
  SilkResamplerParams() : state(nullptr), in(nullptr), out(nullptr), n_streams(0), in_len(0), out_len(0), function(0), fir_order(0), fir_fracs(1),
                          fs_in_khz(0), fs_out_khz(0), input_delay(0), inv_ratio_q16(0), batch(0), coefs{} {}
};
// silk_resampler_init (silk/resampler.c:77-180) as host code: fills the launch parameters for a rate pair; for_enc = 1: 8/12/16/24/48 kHz ->
// 8/12/16 kHz, for_enc = 0: 8/12/16 kHz -> 8/12/16/24/48 kHz.  Returns -1 for a pair the reference rejects.
static inline int silk_resampler_setup(SilkResamplerParams &p, int32_t Fs_in, int32_t Fs_out, int for_enc) {
  static const int8_t delay_enc[5][3] = {{6, 0, 3}, {0, 7, 3}, {0, 1, 10}, {0, 2, 6}, {18, 10, 12}};  // silk/resampler.c:53-61
  static const int8_t delay_dec[3][5] = {{4, 0, 2, 0, 0}, {0, 9, 4, 7, 4}, {0, 3, 12, 7, 7}};          // silk/resampler.c:63-68
  // silk/resampler_rom.c:42-76
  static const int16_t c34[29] = {-20694, -13867, -49, 64, 17, -157, 353, -496, 163, 11047, 22205, -39, 6, 91, -170, 186, 23, -896, 6336, 19928,
                                  -19, -36, 102, -89, -24, 328, -951, 2568, 15909};
  static const int16_t c23[20] = {-14457, -14019, 64, 128, -122, 36, 310, -768, 584, 9267, 17733, 12, 128, 18, -142, 288, -117, -865, 4123, 14459};
  static const int16_t c12[14] = {616, -14323, -10, 39, 58, -46, -84, 120, 184, -315, -541, 1284, 5380, 9024};
  static const int16_t c13[20] = {16102, -15162, -13, 0, 20, 26, 5, -31, -43, -4, 65, 90, 7, -157, -248, -44, 593, 1583, 2612, 3271};
  static const int16_t c14[20] = {22500, -15099, 3, -14, -20, -15, 2, 25, 37, 25, -16, -71, -107, -79, 50, 292, 623, 982, 1288, 1464};
  static const int16_t c16[20] = {27540, -15257, 17, 12, 8, 1, -10, -22, -30, -32, -22, 3, 44, 100, 168, 243, 317, 381, 429, 455};
  const int ri = Fs_in == 8000 ? 0 : Fs_in == 12000 ? 1 : Fs_in == 16000 ? 2 : Fs_in == 24000 ? 3 : Fs_in == 48000 ? 4 : -1;
  const int ro = Fs_out == 8000 ? 0 : Fs_out == 12000 ? 1 : Fs_out == 16000 ? 2 : Fs_out == 24000 ? 3 : Fs_out == 48000 ? 4 : -1;
  if (ri < 0 || ro < 0 || (for_enc ? ro > 2 : ri > 2)) return -1;
  const int16_t *cf = nullptr;
  int order = 0, fracs = 1, ncoef = 0, function, up2x = 0;
  if (Fs_out > Fs_in) {
    if (Fs_out == 2 * Fs_in) function = OB_RS_UP2;
    else { function = OB_RS_IIR_FIR; up2x = 1; }
  } else if (Fs_out < Fs_in) {
    function = OB_RS_DOWN;
    if (Fs_out * 4 == Fs_in * 3) { cf = c34; order = 18; fracs = 3; ncoef = 29; }
    else if (Fs_out * 3 == Fs_in * 2) { cf = c23; order = 18; fracs = 2; ncoef = 20; }
    else if (Fs_out * 2 == Fs_in) { cf = c12; order = 24; ncoef = 14; }
    else if (Fs_out * 3 == Fs_in) { cf = c13; order = 36; ncoef = 20; }
    else if (Fs_out * 4 == Fs_in) { cf = c14; order = 36; ncoef = 20; }
    else if (Fs_out * 6 == Fs_in) { cf = c16; order = 36; ncoef = 20; }
    else return -1;
  } else function = OB_RS_COPY;
  p = SilkResamplerParams();
  p.function = function; p.fir_order = order; p.fir_fracs = fracs;
  p.fs_in_khz = Fs_in / 1000; p.fs_out_khz = Fs_out / 1000;
  p.input_delay = for_enc ? delay_enc[ri][ro] : delay_dec[ri][ro]; p.batch = p.fs_in_khz * 10;
  p.inv_ratio_q16 = (int32_t)((((int64_t)Fs_in << (14 + up2x)) / Fs_out) << 2);
  while ((int32_t)(((int64_t)p.inv_ratio_q16 * Fs_out) >> 16) < (Fs_in << up2x)) p.inv_ratio_q16++;
  for (int i = 0; i < ncoef; i++) p.coefs[i] = cf[i];
  return 0;
}
// silk/resampler_rom.c:84-96 (1/24, 3/24, ... 23/24 sample interpolation), resampler_rom.h:45-46 (all-pass coefficients)
OB_TABLE int16_t ob_silk_frac_fir_12[12][4] = {{189, -600, 617, 30567}, {117, -159, -1070, 29704}, {52, 221, -2392, 28276}, {-4, 529, -3350, 26341},
                                               {-48, 758, -3956, 23973}, {-80, 905, -4235, 21254}, {-99, 972, -4222, 18278}, {-107, 967, -3957, 15143},
                                               {-103, 896, -3487, 11950}, {-91, 773, -2865, 8798}, {-71, 611, -2143, 5784}, {-46, 425, -1375, 2996}};

// AR2 + FIR down-sampling (silk/resampler_private_down_FIR.c:36-193, resampler_private_AR2.c:36-55): symmetric FIR of order 24 / 36 for the
// integer ratios, 18 taps in 2 or 3 interpolation phases for 2:3 and 3:4
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
      int32_t r;
      if (order == 18) {
        const int ind = silk_smulwb(idx & 0xFFFF, p.fir_fracs);
        const int16_t *c0 = &p.coefs[2 + 9 * ind], *c1 = &p.coefs[2 + 9 * (p.fir_fracs - 1 - ind)];
        r = silk_smulwb(b[0], c0[0]);
#pragma unroll 1
        for (int k = 1; k < 9; k++) r = silk_smlawb(r, b[k], c0[k]);
#pragma unroll 1
        for (int k = 0; k < 9; k++) r = silk_smlawb(r, b[17 - k], c1[k]);
      } else {
        r = silk_smulwb(b[0] + b[order - 1], p.coefs[2]);
#pragma unroll 1
        for (int k = 1; k < half; k++) r = silk_smlawb(r, b[k] + b[order - 1 - k], p.coefs[2 + k]);
      }
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
// 2x up-sampling by two chains of three all-pass sections (silk/resampler_private_up2_HQ.c:36-103); ostride = 1
OB_DEV void silk_up2_hq(int32_t *S, int16_t *out, const int16_t *in, int len) {
  int32_t s0 = S[0], s1 = S[1], s2 = S[2], s3 = S[3], s4 = S[4], s5 = S[5];
#pragma unroll 1
  for (int k = 0; k < len; k++) {
    const int32_t in32 = (int32_t)((uint32_t)(int32_t)in[k] << 10);
    int32_t Y, X, o1, o2;
    Y = in32 - s0; X = silk_smulwb(Y, 1746); o1 = s0 + X; s0 = in32 + X;
    Y = o1 - s1; X = silk_smulwb(Y, 14986); o2 = s1 + X; s1 = o1 + X;
    Y = o2 - s2; X = silk_smlawb(Y, Y, 39083 - 65536); o1 = s2 + X; s2 = o2 + X;
    out[2 * k] = silk_sat16(silk_rshift_round(o1, 10));
    Y = in32 - s3; X = silk_smulwb(Y, 6854); o1 = s3 + X; s3 = in32 + X;
    Y = o1 - s4; X = silk_smulwb(Y, 25769); o2 = s4 + X; s4 = o1 + X;
    Y = o2 - s5; X = silk_smlawb(Y, Y, 55542 - 65536); o1 = s5 + X; s5 = o2 + X;
    out[2 * k + 1] = silk_sat16(silk_rshift_round(o1, 10));
  }
  S[0] = s0; S[1] = s1; S[2] = s2; S[3] = s3; S[4] = s4; S[5] = s5;
}
// 2x up-sampling followed by FIR interpolation at 1/12-sample resolution (silk/resampler_private_IIR_FIR.c:35-107); sfir: 8 int16 of history
OB_DEV int16_t *silk_iir_fir(const SilkResamplerParams &p, int32_t *sIIR, int16_t *sfir, int16_t *out, const int16_t *in, int inLen) {
  int16_t buf[2 * 480 + 8];
  int nIn = 0;
#pragma unroll 1
  for (int k = 0; k < 8; k++) buf[k] = sfir[k];
#pragma unroll 1
  for (;;) {
    nIn = inLen < p.batch ? inLen : p.batch;
    silk_up2_hq(sIIR, &buf[8], in, nIn);
    const int32_t maxi = nIn << 17;
#pragma unroll 1
    for (int32_t idx = 0; idx < maxi; idx += p.inv_ratio_q16) {
      const int t = silk_smulwb(idx & 0xFFFF, 12);
      const int16_t *b = &buf[idx >> 16];
      int32_t r = (int32_t)b[0] * ob_silk_frac_fir_12[t][0];
#pragma unroll
      for (int k = 1; k < 4; k++) r += (int32_t)b[k] * ob_silk_frac_fir_12[t][k];
#pragma unroll
      for (int k = 0; k < 4; k++) r += (int32_t)b[4 + k] * ob_silk_frac_fir_12[11 - t][3 - k];
      *out++ = silk_sat16(silk_rshift_round(r, 15));
    }
    in += nIn;
    inLen -= nIn;
    if (inLen > 0) {
#pragma unroll 1
      for (int k = 0; k < 8; k++) buf[k] = buf[(nIn << 1) + k];
    } else break;
  }
#pragma unroll 1
  for (int k = 0; k < 8; k++) sfir[k] = buf[(nIn << 1) + k];
  return out;
}

// silk_resampler() (silk/resampler.c:183-219): the first millisecond comes from the delay buffer, the rest from the input
OB_DEV void silk_resampler_thread(const SilkResamplerParams &p, int stream) {
  ObSilkResamplerState *S = p.state + stream;
  const int16_t *in = p.in + (size_t)stream * p.in_len;
  int16_t *out = p.out + (size_t)stream * p.out_len;
  int16_t delay[48];
  const int nSamples = p.fs_in_khz - p.input_delay;
#pragma unroll 1
  for (int k = 0; k < p.input_delay; k++) delay[k] = S->delayBuf[k];
#pragma unroll 1
  for (int k = 0; k < nSamples; k++) delay[p.input_delay + k] = in[k];
  if (p.function == OB_RS_COPY) {
#pragma unroll 1
    for (int k = 0; k < p.fs_in_khz; k++) out[k] = delay[k];
#pragma unroll 1
    for (int k = 0; k < p.in_len - p.fs_in_khz; k++) out[p.fs_out_khz + k] = in[nSamples + k];
  } else if (p.function == OB_RS_UP2) {
    int32_t sIIR[6];
#pragma unroll
    for (int k = 0; k < 6; k++) sIIR[k] = S->sIIR[k];
    silk_up2_hq(sIIR, out, delay, p.fs_in_khz);
    silk_up2_hq(sIIR, out + p.fs_out_khz, in + nSamples, p.in_len - p.fs_in_khz);
#pragma unroll
    for (int k = 0; k < 6; k++) S->sIIR[k] = sIIR[k];
  } else if (p.function == OB_RS_IIR_FIR) {
    int32_t sIIR[6];
    int16_t sfir[8];
#pragma unroll
    for (int k = 0; k < 6; k++) sIIR[k] = S->sIIR[k];
#pragma unroll 1
    for (int k = 0; k < 8; k++) sfir[k] = ((const int16_t *)S->sFIR)[k];
    silk_iir_fir(p, sIIR, sfir, out, delay, p.fs_in_khz);
    silk_iir_fir(p, sIIR, sfir, out + p.fs_out_khz, in + nSamples, p.in_len - p.fs_in_khz);
#pragma unroll
    for (int k = 0; k < 6; k++) S->sIIR[k] = sIIR[k];
#pragma unroll 1
    for (int k = 0; k < 8; k++) ((int16_t *)S->sFIR)[k] = sfir[k];
  } else {
    int32_t sIIR[2] = {S->sIIR[0], S->sIIR[1]};
    int32_t sFIR[36];
#pragma unroll 1
    for (int k = 0; k < 36; k++) sFIR[k] = S->sFIR[k];
    silk_down_fir(p, sIIR, sFIR, out, delay, p.fs_in_khz);
    silk_down_fir(p, sIIR, sFIR, out + p.fs_out_khz, in + nSamples, p.in_len - p.fs_in_khz);
#pragma unroll 1
    for (int k = 0; k < 36; k++) S->sFIR[k] = sFIR[k];
    S->sIIR[0] = sIIR[0]; S->sIIR[1] = sIIR[1];
  }
#pragma unroll 1
  for (int k = 0; k < p.input_delay; k++) S->delayBuf[k] = in[p.in_len - p.input_delay + k];
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

// ---- gain processing and quantisation of a SILK frame: silk_process_gains_FLP (silk/float/process_gains_FLP.c:35-103), silk_gains_quant
// (silk/gain_quant.c:40-95), silk_lin2log / silk_log2lin (silk/lin2log.c, log2lin.c).  One thread per stream; the record is the stream's view of
// silk_encoder_state_FLP / silk_encoder_control_FLP (what the function reads and writes). ----
struct ObSilkGains {   // == OpusB200SilkGains (include/opus_b200.h)
  float Gains[4], ResNrg[4];
  float LTPredCodGain, input_quality, coding_quality, Lambda;
  int32_t GainsUnq_Q16[4];
  int32_t signalType, quantOffsetType, SNR_dB_Q7, input_tilt_Q15, speech_activity_Q8, condCoding;
  int8_t GainsIndices[4];
  int8_t LastGainIndex, lastGainIndexPrev, pad[2];
};
struct SilkGainsParams { ObSilkGains *frames; int n, nb_subfr, subfr_length, nStatesDelayedDecision; };
OB_DEV int32_t silk_smlawb_(int32_t a, int32_t b, int32_t c) { return a + silk_smulwb(b, c); }
OB_DEV int32_t silk_lin2log(int32_t inLin) {
  const int lz = clz32((uint32_t)inLin), rot = 24 - lz;
  const uint32_t x = (uint32_t)inLin;
  const uint32_t r = rot == 0 ? x : rot < 0 ? ((x << (uint32_t)(-rot)) | (x >> (uint32_t)(32 + rot))) : ((x << (uint32_t)(32 - rot)) | (x >> (uint32_t)rot));   // silk_ROR32
  const int32_t frac_Q7 = (int32_t)(r & 0x7f);
  return silk_smlawb_(frac_Q7, frac_Q7 * (128 - frac_Q7), 179) + ((31 - lz) << 7);
}
OB_DEV int32_t silk_log2lin(int32_t inLog_Q7) {
  if (inLog_Q7 < 0) return 0;
  if (inLog_Q7 >= 3967) return 0x7FFFFFFF;
  int32_t out = 1 << (inLog_Q7 >> 7);
  const int32_t frac_Q7 = inLog_Q7 & 0x7F, poly = silk_smlawb_(frac_Q7, (int32_t)(int16_t)frac_Q7 * (int32_t)(int16_t)(128 - frac_Q7), -174);
  if (inLog_Q7 < 2048) out = out + ((out * poly) >> 7);
  else out = out + (out >> 7) * poly;
  return out;
}
OB_DEV int silk_limit_int(int a, int lo, int hi) { return a < lo ? lo : a > hi ? hi : a; }
OB_DEV void silk_gains_quant(int8_t *ind, int32_t *gain_Q16, int8_t *prev_ind, int conditional, int nb_subfr) {
  const int kOffset = (2 * 128) / 6 + 16 * 128, kScale = (65536 * (64 - 1)) / (((88 - 2) * 128) / 6), kInvScale = (65536 * (((88 - 2) * 128) / 6)) / (64 - 1);
  int8_t prev = *prev_ind;
  for (int k = 0; k < nb_subfr; k++) {
    int8_t v = (int8_t)silk_smulwb(kScale, silk_lin2log(gain_Q16[k]) - kOffset);
    if (v < prev) v++;
    v = (int8_t)silk_limit_int(v, 0, 63);
    if (k == 0 && conditional == 0) {
      v = (int8_t)silk_limit_int(v, prev - 4, 63);
      prev = v;
    } else {
      v = (int8_t)(v - prev);
      const int thr = 2 * 36 - 64 + prev;
      if (v > thr) v = (int8_t)(thr + ((v - thr + 1) >> 1));
      v = (int8_t)silk_limit_int(v, -4, 36);
      if (v > thr) {
        prev = (int8_t)(prev + ((v << 1) - thr));
        prev = prev < 63 ? prev : 63;
      } else prev = (int8_t)(prev + v);
      v = (int8_t)(v + 4);
    }
    ind[k] = v;
    const int32_t lg = silk_smulwb(kInvScale, prev) + kOffset;
    gain_Q16[k] = silk_log2lin(lg < 3967 ? lg : 3967);
  }
  *prev_ind = prev;
}
OB_DEV void silk_process_gains_thread(const SilkGainsParams &p, int s) {
  ObSilkGains g = p.frames[s];
  const int nb = p.nb_subfr;
  int32_t q16[4];
  if (g.signalType == 2) {
    const float sg = 1.0f - 0.5f * (float)(1.0 / (1.0 + exp(-(double)(0.25f * (g.LTPredCodGain - 12.0f)))));
    for (int k = 0; k < nb; k++) g.Gains[k] *= sg;
  }
  const float InvMaxSqrVal = (float)(pow((double)2.0f, (double)(0.33f * (21.0f - g.SNR_dB_Q7 * (1 / 128.0f)))) / p.subfr_length);
  for (int k = 0; k < nb; k++) {
    float gain = g.Gains[k];
    gain = (float)sqrt((double)(gain * gain + g.ResNrg[k] * InvMaxSqrVal));
    g.Gains[k] = gain < 32767.0f ? gain : 32767.0f;
  }
  for (int k = 0; k < nb; k++) { q16[k] = (int32_t)(g.Gains[k] * 65536.0f); g.GainsUnq_Q16[k] = q16[k]; }
  g.lastGainIndexPrev = g.LastGainIndex;
  silk_gains_quant(g.GainsIndices, q16, &g.LastGainIndex, g.condCoding == 2, nb);
  for (int k = 0; k < nb; k++) g.Gains[k] = q16[k] / 65536.0f;
  if (g.signalType == 2) g.quantOffsetType = (g.LTPredCodGain + g.input_tilt_Q15 * (1.0f / 32768.0f) > 1.0f) ? 0 : 1;
  const int qo = (g.signalType >> 1) ? (g.quantOffsetType ? 100 : 32) : (g.quantOffsetType ? 240 : 100);   // silk_Quantization_Offsets_Q10
  const float quant_offset = qo / 1024.0f;
  g.Lambda = 1.2f + -0.05f * p.nStatesDelayedDecision + -0.2f * g.speech_activity_Q8 * (1.0f / 256.0f) + -0.1f * g.input_quality + -0.2f * g.coding_quality +
             0.8f * quant_offset;
  p.frames[s] = g;
}

}  // namespace ob
