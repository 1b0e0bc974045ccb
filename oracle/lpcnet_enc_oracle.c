/* oracle/lpcnet_enc_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Plain-C restatement of the LPCNet feature extraction (SURVEY §8 row f4 / the history analysis of row a20's lpcnet_plc_conceal):
 *   lpcnet_compute_single_frame_features_float, compute_frame_features, frame_analysis, biquad, preemphasis   dnn/lpcnet_enc.c:73-230
 *   forward_transform, inverse_transform, lpcn_compute_band_energy, dct / idct, interp_band_gain,
 *   lpc_from_bands / lpc_from_cepstrum, lpcn_lpc, apply_window, burg_cepstral_analysis                       dnn/freq.c:57-323
 *   compute_pitchdnn                                                                                          dnn/pitchdnn.c:12-62
 *   silk_burg_analysis (dnn/burg.c:97-246) is the same recursion as silk/float/burg_modified_FLP.c: silk_oracle.c's restatement is reused,
 *   with the clamp of the returned energy that only the dnn copy has (burg.c:245).
 *   opus_fft for nfft = 320 (celt/kiss_fft.c:596-613 with the factors 5 x 4 x 4 x 4 of dnn/lpcnet_tables.c), the C builds of
 *   celt_pitch_xcorr / celt_inner_prod / celt_fir (celt/pitch.c:216-262, pitch.h:159-170, celt_lpc.c:93-153: every sum runs in index order).
 * Float arithmetic in the reference's order, including its float / double promotions.  Pinned against the reference's own files compiled with
 * seeded random PitchDNN weights (oracle/_ref/libfargan_ref.so, gen_pitchdnn_weights.py) by tests/test_lpcnet_enc_oracle_vs_ref.py. */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define OB_TABLE static const
#include "../opus_b200/csrc/lpcnet_tables_data.h"

typedef struct {
  const float *bias, *subias;
  const int8_t *weights;
  const float *float_weights;
  const int *weights_idx;
  const float *diag, *scale;
  int nb_inputs, nb_outputs;
} OrcLinear; /* as in nnet_oracle.c */
typedef struct {
  const float *bias, *float_weights;
  int in_channels, out_channels, ktime, kheight;
} OrcConv2d;
void orc_nnet_dense(const OrcLinear *l, float *out, const float *in, int act);
void orc_nnet_gru(const OrcLinear *wi, const OrcLinear *wr, float *state, const float *in);
void orc_nnet_conv2d(const OrcConv2d *c, float *out, float *mem, const float *in, int height, int hstride, int act);
float orc_silk_burg_modified(float *A, const float *x, float minInvGain, int subfr_length, int nb_subfr, int D);
enum { ACT_LINEAR = 0, ACT_SIGMOID = 1, ACT_TANH = 2 }; /* dnn/nnet.h:35-37 */

#define FRAME 160
#define OVERLAP 160
#define WINDOW 320
#define FREQ 161
#define NB_BANDS 18
#define LPC_ORDER 16
#define NB_FEATURES 20
#define NB_TOTAL_FEATURES 36
#define TRAINING_OFFSET 80
#define PITCH_MIN 32
#define PITCH_MAX 256
#define NB_XCORR (PITCH_MAX - PITCH_MIN)
#define PITCH_BUF (PITCH_MAX + 320)
#define IF_MAX_FREQ 30
#define IF_FEATURES (3 * IF_MAX_FREQ - 2)
#define PREEMPH 0.85f

typedef struct { float r, i; } cpx;
static const short eband5ms[NB_BANDS] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 34, 40}; /* freq.c:45-48 */
static const float compensation[NB_BANDS] = {0.8f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.666667f, 0.5f, 0.5f, 0.5f, 0.333333f, 0.25f, 0.25f, 0.2f,
                                             0.166667f, 0.173913f}; /* freq.c:50-52 */

/* ---- the 320-point transform: radix-4 passes with m = 1, 4, 16, then one radix-5 pass with m = 64 ---- */
#define TW320(k) (((const cpx *)ob_lpcnet_fft_tw)[k])
static cpx cmul(cpx a, cpx b) { cpx m; m.r = a.r * b.r - a.i * b.i; m.i = a.r * b.i + a.i * b.r; return m; }
static void radix4_pass(cpx *base, int fstride, int m, int groups, int group_stride) {
  int g, j;
  for (g = 0; g < groups; g++) {
    cpx *F = base + g * group_stride;
    for (j = 0; j < m; j++, F++) {
      if (m == 1) { /* kf_bfly4, m == 1: no twiddles (kiss_fft.c:107-130) */
        cpx s0, s1;
        s0.r = F[0].r - F[2].r; s0.i = F[0].i - F[2].i;
        F[0].r += F[2].r; F[0].i += F[2].i;
        s1.r = F[1].r + F[3].r; s1.i = F[1].i + F[3].i;
        F[2].r = F[0].r - s1.r; F[2].i = F[0].i - s1.i;
        F[0].r += s1.r; F[0].i += s1.i;
        s1.r = F[1].r - F[3].r; s1.i = F[1].i - F[3].i;
        F[1].r = s0.r + s1.i; F[1].i = s0.i - s1.r;
        F[3].r = s0.r - s1.i; F[3].i = s0.i + s1.r;
      } else { /* kiss_fft.c:131-173 */
        const cpx a = cmul(F[m], TW320(j * fstride)), b = cmul(F[2 * m], TW320(2 * j * fstride)), c = cmul(F[3 * m], TW320(3 * j * fstride));
        cpx d, e, f;
        d.r = F->r - b.r; d.i = F->i - b.i;
        F->r += b.r; F->i += b.i;
        e.r = a.r + c.r; e.i = a.i + c.i;
        f.r = a.r - c.r; f.i = a.i - c.i;
        F[2 * m].r = F->r - e.r; F[2 * m].i = F->i - e.i;
        F->r += e.r; F->i += e.i;
        F[m].r = d.r + f.i; F[m].i = d.i - f.r;
        F[3 * m].r = d.r - f.i; F[3 * m].i = d.i + f.r;
      }
    }
  }
}
static void radix5_pass(cpx *F, int fstride, int m) { /* kf_bfly5, kiss_fft.c:236-301 */
  const cpx ya = TW320(fstride * m), yb = TW320(2 * fstride * m);
  int u;
  for (u = 0; u < m; u++) {
    cpx *F0 = F + u, *F1 = F0 + m, *F2 = F0 + 2 * m, *F3 = F0 + 3 * m, *F4 = F0 + 4 * m;
    const cpx s0 = *F0, s1 = cmul(*F1, TW320(u * fstride)), s2 = cmul(*F2, TW320(2 * u * fstride)), s3 = cmul(*F3, TW320(3 * u * fstride)),
              s4 = cmul(*F4, TW320(4 * u * fstride));
    cpx s5, s6, s7, s8, s9, s10, s11, s12;
    s7.r = s1.r + s4.r; s7.i = s1.i + s4.i;
    s10.r = s1.r - s4.r; s10.i = s1.i - s4.i;
    s8.r = s2.r + s3.r; s8.i = s2.i + s3.i;
    s9.r = s2.r - s3.r; s9.i = s2.i - s3.i;
    F0->r = F0->r + (s7.r + s8.r);
    F0->i = F0->i + (s7.i + s8.i);
    s5.r = s0.r + (s7.r * ya.r + s8.r * yb.r);
    s5.i = s0.i + (s7.i * ya.r + s8.i * yb.r);
    s6.r = s10.i * ya.i + s9.i * yb.i;
    s6.i = -(s10.r * ya.i + s9.r * yb.i);
    F1->r = s5.r - s6.r; F1->i = s5.i - s6.i;
    F4->r = s5.r + s6.r; F4->i = s5.i + s6.i;
    s11.r = s0.r + (s7.r * yb.r + s8.r * ya.r);
    s11.i = s0.i + (s7.i * yb.r + s8.i * ya.r);
    s12.r = s9.i * ya.i - s10.i * yb.i;
    s12.i = s10.r * yb.i - s9.r * ya.i;
    F2->r = s11.r + s12.r; F2->i = s11.i + s12.i;
    F3->r = s11.r - s12.r; F3->i = s11.i - s12.i;
  }
}
/* opus_fft_c for the lpcnet kfft: scale by 1/320 into digit-reversed order, then the four passes */
void orc_lpcnet_fft320(cpx *out, const cpx *in) {
  const float scale = 0.0031250000f;
  int i;
  for (i = 0; i < WINDOW; i++) { out[ob_lpcnet_fft_bitrev[i]].r = scale * in[i].r; out[ob_lpcnet_fft_bitrev[i]].i = scale * in[i].i; }
  radix4_pass(out, 80, 1, 80, 4);
  radix4_pass(out, 20, 4, 20, 16);
  radix4_pass(out, 5, 16, 5, 64);
  radix5_pass(out, 1, 64);
}
static void forward_transform(cpx *out, const float *in) { /* freq.c:232-244 */
  cpx x[WINDOW], y[WINDOW];
  int i;
  for (i = 0; i < WINDOW; i++) { x[i].r = in[i]; x[i].i = 0; }
  orc_lpcnet_fft320(y, x);
  for (i = 0; i < FREQ; i++) out[i] = y[i];
}
static void inverse_transform(float *out, const cpx *in) { /* freq.c:246-263 */
  cpx x[WINDOW], y[WINDOW];
  int i;
  for (i = 0; i < FREQ; i++) x[i] = in[i];
  for (; i < WINDOW; i++) { x[i].r = x[WINDOW - i].r; x[i].i = -x[WINDOW - i].i; }
  orc_lpcnet_fft320(y, x);
  out[0] = WINDOW * y[0].r;
  for (i = 1; i < WINDOW; i++) out[i] = WINDOW * y[WINDOW - i].r;
}
static void band_energy(float *bandE, const cpx *X, int inverse) { /* freq.c:55-80 (inverse) and :131-154 */
  float sum[NB_BANDS] = {0};
  int i, j;
  for (i = 0; i < NB_BANDS - 1; i++) {
    const int band_size = (eband5ms[i + 1] - eband5ms[i]) * 4;
    for (j = 0; j < band_size; j++) {
      const float frac = (float)j / band_size;
      const cpx v = X[eband5ms[i] * 4 + j];
      float tmp = v.r * v.r;
      tmp += v.i * v.i;
      if (inverse) tmp = 1.f / (tmp + 1e-9);
      sum[i] += (1 - frac) * tmp;
      sum[i + 1] += frac * tmp;
    }
  }
  sum[0] *= 2;
  sum[NB_BANDS - 1] *= 2;
  for (i = 0; i < NB_BANDS; i++) bandE[i] = sum[i];
}
static void dct18(float *out, const float *in) { /* freq.c:208-218 */
  int i, j;
  for (i = 0; i < NB_BANDS; i++) {
    float sum = 0;
    for (j = 0; j < NB_BANDS; j++) sum += in[j] * ob_lpcnet_dct_table[j * NB_BANDS + i];
    out[i] = sum * sqrt(2. / NB_BANDS);
  }
}
static void idct18(float *out, const float *in) { /* freq.c:220-230 */
  int i, j;
  for (i = 0; i < NB_BANDS; i++) {
    float sum = 0;
    for (j = 0; j < NB_BANDS; j++) sum += in[j] * ob_lpcnet_dct_table[i * NB_BANDS + j];
    out[i] = sum * sqrt(2. / NB_BANDS);
  }
}
static float levinson(float *lpc, const float *ac, int p) { /* lpcn_lpc, freq.c:82-127 (float build: the shifts are no-ops) */
  float error = ac[0];
  int i, j;
  memset(lpc, 0, p * sizeof(float));
  if (ac[0] != 0) {
    for (i = 0; i < p; i++) {
      float rr = 0, r;
      for (j = 0; j < i; j++) rr += lpc[j] * ac[i - j];
      rr += ac[i + 1];
      r = -rr / error;
      lpc[i] = r;
      for (j = 0; j < (i + 1) >> 1; j++) {
        const float t1 = lpc[j], t2 = lpc[i - 1 - j];
        lpc[j] = t1 + r * t2;
        lpc[i - 1 - j] = t2 + r * t1;
      }
      error = error - (r * r) * error;
      if (error < .001f * ac[0]) break;
    }
  }
  return error;
}
float orc_lpc_from_cepstrum(float *lpc, const float *cepstrum) { /* freq.c:303-313 with lpc_from_bands :265-289 and interp_band_gain :192-205 */
  float Ex[NB_BANDS], tmp[NB_BANDS], Xr[FREQ], x_auto[WINDOW], ac[LPC_ORDER + 1];
  cpx X_auto[FREQ];
  int i, j;
  memcpy(tmp, cepstrum, sizeof(tmp));
  tmp[0] += 4;
  idct18(Ex, tmp);
  for (i = 0; i < NB_BANDS; i++) Ex[i] = pow(10.f, Ex[i]) * compensation[i];
  memset(Xr, 0, sizeof(Xr));
  for (i = 0; i < NB_BANDS - 1; i++) {
    const int band_size = (eband5ms[i + 1] - eband5ms[i]) * 4;
    for (j = 0; j < band_size; j++) {
      const float frac = (float)j / band_size;
      Xr[eband5ms[i] * 4 + j] = (1 - frac) * Ex[i] + frac * Ex[i + 1];
    }
  }
  Xr[FREQ - 1] = 0;
  for (i = 0; i < FREQ; i++) { X_auto[i].r = Xr[i]; X_auto[i].i = 0; }
  inverse_transform(x_auto, X_auto);
  for (i = 0; i < LPC_ORDER + 1; i++) ac[i] = x_auto[i];
  ac[0] += ac[0] * 1e-4 + 320 / 12 / 38.;
  for (i = 1; i < LPC_ORDER + 1; i++) ac[i] *= (1 - 6e-5 * i * i);
  return levinson(lpc, ac, LPC_ORDER);
}
static void burg_cepstrum(const float *pcm, float *ceps, int len, int order) { /* compute_burg_cepstrum, freq.c:156-187 */
  float burg_in[FRAME], burg_lpc[LPC_ORDER], x[WINDOW], Eburg[NB_BANDS], Ly[NB_BANDS], g, logMax = -2, follow = -2;
  cpx LPC[FREQ];
  int i;
  for (i = 0; i < len - 1; i++) burg_in[i] = pcm[i + 1] - PREEMPH * pcm[i];
  g = orc_silk_burg_modified(burg_lpc, burg_in, 1e-3, len - 1, 1, order);
  g = g > 0 ? g : 0; /* burg.c:245 */
  g /= len - 2 * (order - 1);
  memset(x, 0, sizeof(x));
  x[0] = 1;
  for (i = 0; i < order; i++) x[i + 1] = -burg_lpc[i] * pow(.995, i + 1);
  forward_transform(LPC, x);
  band_energy(Eburg, LPC, 1);
  for (i = 0; i < NB_BANDS; i++) Eburg[i] *= .45 * g * (1.f / ((float)WINDOW * WINDOW * WINDOW));
  for (i = 0; i < NB_BANDS; i++) {
    double a, b;
    Ly[i] = log10(1e-2 + Eburg[i]);
    a = follow - 2.5;
    a = a > Ly[i] ? a : Ly[i];
    b = logMax - 8;
    Ly[i] = b > a ? b : a;
    logMax = logMax > Ly[i] ? logMax : Ly[i];
    a = follow - 2.5;
    follow = a > Ly[i] ? a : Ly[i];
  }
  dct18(ceps, Ly);
  ceps[0] += -4;
}
void orc_burg_cepstral_analysis(float *ceps, const float *x) { /* freq.c:179-190 */
  int i;
  burg_cepstrum(x, &ceps[0], FRAME / 2, LPC_ORDER);
  burg_cepstrum(&x[FRAME / 2], &ceps[NB_BANDS], FRAME / 2, LPC_ORDER);
  for (i = 0; i < NB_BANDS; i++) {
    const float c0 = ceps[i], c1 = ceps[NB_BANDS + i];
    ceps[i] = .5 * (c0 + c1);
    ceps[NB_BANDS + i] = (c0 - c1);
  }
}

/* ---- the pitch network: dnn/pitchdnn.c:12-62 ---- */
enum { PD_IF1, PD_IF2, PD_DOWN, PD_FINAL, PD_GRU_IN, PD_GRU_REC, PD_COUNT };
typedef struct {
  OrcLinear l[PD_COUNT];
  OrcConv2d c[2];
  float gru_state[64];
  float xcorr_mem1[(NB_XCORR + 2) * 2];
  float xcorr_mem2[(NB_XCORR + 2) * 2 * 8];
} OrcPitchDnn;
/* the layers (pitchdnn.c:12-44): output [192] */
void orc_pitchdnn_net(OrcPitchDnn *st, float *output, const float *if_features, const float *xcorr_features) {
  float if1_out[64], downsampler_in[NB_XCORR + 64], downsampler_out[64];
  float conv1_tmp1[(NB_XCORR + 2) * 8] = {0}, conv1_tmp2[(NB_XCORR + 2) * 8] = {0};
  orc_nnet_dense(&st->l[PD_IF1], if1_out, if_features, ACT_TANH);
  orc_nnet_dense(&st->l[PD_IF2], &downsampler_in[NB_XCORR], if1_out, ACT_TANH);
  memcpy(&conv1_tmp1[1], xcorr_features, NB_XCORR * sizeof(float));
  orc_nnet_conv2d(&st->c[0], &conv1_tmp2[1], st->xcorr_mem1, conv1_tmp1, NB_XCORR, NB_XCORR + 2, ACT_TANH);
  orc_nnet_conv2d(&st->c[1], downsampler_in, st->xcorr_mem2, conv1_tmp2, NB_XCORR, NB_XCORR, ACT_TANH);
  orc_nnet_dense(&st->l[PD_DOWN], downsampler_out, downsampler_in, ACT_TANH);
  orc_nnet_gru(&st->l[PD_GRU_IN], &st->l[PD_GRU_REC], st->gru_state, downsampler_out);
  orc_nnet_dense(&st->l[PD_FINAL], output, st->gru_state, ACT_LINEAR);
}
float orc_pitchdnn_compute(OrcPitchDnn *st, const float *if_features, const float *xcorr_features) {
  float output[192], maxval = -1, sum = 0, count = 0;
  int i, pos = 0;
  orc_pitchdnn_net(st, output, if_features, xcorr_features);
  for (i = 0; i < 180; i++)
    if (output[i] > maxval) { pos = i; maxval = output[i]; }
  for (i = pos - 2 > 0 ? pos - 2 : 0; i <= (pos + 2 < 179 ? pos + 2 : 179); i++) { /* :45-58 */
    const float p = exp(output[i]);
    sum += p * i;
    count += p;
  }
  return (1.f / 60.f) * (sum / count) - 1.5;
}

/* ---- dnn/lpcnet_enc.c ---- */
typedef struct {
  OrcPitchDnn pitchdnn;
  float analysis_mem[OVERLAP];
  float mem_preemph;
  cpx prev_if[IF_MAX_FREQ];
  float if_features[IF_FEATURES];
  float xcorr_features[NB_XCORR];
  float dnn_pitch;
  float pitch_mem[LPC_ORDER];
  float pitch_filt;
  float exc_buf[PITCH_BUF];
  float lp_buf[PITCH_BUF];
  float lp_mem[4];
  float lpc[LPC_ORDER];
  float features[NB_TOTAL_FEATURES];
} OrcLpcnetEnc;
int orc_lpcnet_enc_size(void) { return (int)sizeof(OrcLpcnetEnc); }
void orc_lpcnet_enc_init(OrcLpcnetEnc *st, const OrcLinear *linear /* [6] */, const OrcConv2d *conv /* [2] */) {
  memset(st, 0, sizeof(*st));
  memcpy(st->pitchdnn.l, linear, sizeof(st->pitchdnn.l));
  memcpy(st->pitchdnn.c, conv, sizeof(st->pitchdnn.c));
}
void orc_lpcnet_enc_reset(OrcLpcnetEnc *st) { /* lpcnet_encoder_init(): the state cleared, the model kept */
  memset(st->pitchdnn.gru_state, 0, sizeof(*st) - ((char *)st->pitchdnn.gru_state - (char *)st));
}
static float log10f_ref(float x) { return 0.3010299957f * (float)(1.442695040888963387 * log(x)); } /* celt_log10, lpcnet_enc.c:113, mathops.h:346 */
static float clamp1(float x) { return x < -1.f ? -1.f : x > 1.f ? 1.f : x; }
static float dotf(const float *a, const float *b, int n) { float s = 0; int i; for (i = 0; i < n; i++) s = s + a[i] * b[i]; return s; }

static void frame_features(OrcLpcnetEnc *st, const float *in) { /* compute_frame_features, lpcnet_enc.c:115-205 */
  static const float lp_b[2] = {-0.84946f, 1.f}, lp_a[2] = {-1.54220f, 0.70781f};
  float aligned_in[FRAME], Ly[NB_BANDS], Ex[NB_BANDS], xcorr[NB_XCORR], x[FRAME + LPC_ORDER], win[WINDOW], ener_norm[NB_XCORR];
  float follow = -2, logMax = -2, ener0, xx, yy, xy, frame_corr;
  cpx X[FREQ];
  int i, j, pitch;
  memcpy(aligned_in, &st->analysis_mem[OVERLAP - TRAINING_OFFSET], TRAINING_OFFSET * sizeof(float));
  /* frame_analysis, :73-82 */
  memcpy(win, st->analysis_mem, OVERLAP * sizeof(float));
  memcpy(&win[OVERLAP], in, FRAME * sizeof(float));
  memcpy(st->analysis_mem, &in[FRAME - OVERLAP], OVERLAP * sizeof(float));
  for (i = 0; i < OVERLAP; i++) { win[i] *= ob_lpcnet_half_window[i]; win[WINDOW - 1 - i] *= ob_lpcnet_half_window[i]; }
  forward_transform(X, win);
  band_energy(Ex, X, 0);
  st->if_features[0] = clamp1((1.f / 64) * (10.f * log10f_ref(1e-15f + X[0].r * X[0].r) - 6.f));
  for (i = 1; i < IF_MAX_FREQ; i++) {
    cpx prod;
    float norm_1;
    prod.r = X[i].r * st->prev_if[i].r + X[i].i * st->prev_if[i].i;
    prod.i = X[i].i * st->prev_if[i].r - X[i].r * st->prev_if[i].i;
    norm_1 = 1.f / sqrt(1e-15f + prod.r * prod.r + prod.i * prod.i);
    prod.r *= norm_1;
    prod.i *= norm_1;
    st->if_features[3 * i - 2] = prod.r;
    st->if_features[3 * i - 1] = prod.i;
    st->if_features[3 * i] = clamp1((1.f / 64) * (10.f * log10f_ref(1e-15f + X[i].r * X[i].r + X[i].i * X[i].i) - 6.f));
  }
  memcpy(st->prev_if, X, sizeof(st->prev_if));
  for (i = 0; i < NB_BANDS; i++) {
    float a;
    Ly[i] = log10f_ref(1e-2f + Ex[i]);
    a = follow - 2.5f > Ly[i] ? follow - 2.5f : Ly[i];
    Ly[i] = logMax - 8 > a ? logMax - 8 : a;
    logMax = logMax > Ly[i] ? logMax : Ly[i];
    follow = follow - 2.5f > Ly[i] ? follow - 2.5f : Ly[i];
  }
  dct18(st->features, Ly);
  st->features[0] -= 4;
  orc_lpc_from_cepstrum(st->lpc, st->features);
  for (i = 0; i < LPC_ORDER; i++) st->features[NB_BANDS + 2 + i] = st->lpc[i];
  memmove(st->exc_buf, &st->exc_buf[FRAME], PITCH_MAX * sizeof(float));
  memmove(st->lp_buf, &st->lp_buf[FRAME], PITCH_MAX * sizeof(float));
  memcpy(&aligned_in[TRAINING_OFFSET], in, (FRAME - TRAINING_OFFSET) * sizeof(float));
  memcpy(&x[0], st->pitch_mem, LPC_ORDER * sizeof(float));
  memcpy(&x[LPC_ORDER], aligned_in, FRAME * sizeof(float));
  memcpy(st->pitch_mem, &aligned_in[FRAME - LPC_ORDER], LPC_ORDER * sizeof(float));
  for (i = 0; i < FRAME; i++) { /* celt_fir_c: y[i] = x[i] + sum_j num[ord-1-j] x[i-ord+j], j ascending */
    float s = x[LPC_ORDER + i];
    for (j = 0; j < LPC_ORDER; j++) s = s + st->lpc[LPC_ORDER - 1 - j] * x[i + j];
    st->lp_buf[PITCH_MAX + i] = s;
  }
  for (i = 0; i < FRAME; i++) {
    st->exc_buf[PITCH_MAX + i] = st->lp_buf[PITCH_MAX + i] + .7f * st->pitch_filt;
    st->pitch_filt = st->lp_buf[PITCH_MAX + i];
  }
  { /* biquad, :84-111, in place */
    float mem0 = st->lp_mem[0], mem1 = st->lp_mem[1], *y = &st->lp_buf[PITCH_MAX];
    for (i = 0; i < FRAME; i++) {
      const float xi = y[i], yi = y[i] + mem0, mem00 = mem0;
      mem0 = (lp_b[0] - lp_a[0]) * xi + mem1 - lp_a[0] * mem0;
      mem1 = (lp_b[1] - lp_a[1]) * xi + 1e-30f - lp_a[1] * mem00;
      y[i] = yi;
    }
    st->lp_mem[0] = mem0;
    st->lp_mem[1] = mem1;
  }
  {
    const float *buf = st->exc_buf;
    double ener1;
    for (i = 0; i < NB_XCORR; i++) xcorr[i] = dotf(&buf[PITCH_MAX], &buf[i], FRAME);
    ener0 = dotf(&buf[PITCH_MAX], &buf[PITCH_MAX], FRAME);
    ener1 = dotf(&buf[0], &buf[0], FRAME);
    for (i = 0; i < NB_XCORR; i++) {
      const float ener = 1 + ener0 + ener1;
      st->xcorr_features[i] = 2 * xcorr[i];
      ener_norm[i] = ener;
      ener1 += buf[i + FRAME] * (double)buf[i + FRAME] - buf[i] * (double)buf[i];
    }
    for (i = 0; i < NB_XCORR; i++) st->xcorr_features[i] /= ener_norm[i];
  }
  st->dnn_pitch = orc_pitchdnn_compute(&st->pitchdnn, st->if_features, st->xcorr_features);
  pitch = (int)floor(.5 + 256. / pow(2.f, ((1. / 60.) * ((st->dnn_pitch + 1.5) * 60))));
  xx = dotf(&st->lp_buf[PITCH_MAX], &st->lp_buf[PITCH_MAX], FRAME);
  yy = dotf(&st->lp_buf[PITCH_MAX - pitch], &st->lp_buf[PITCH_MAX - pitch], FRAME);
  xy = dotf(&st->lp_buf[PITCH_MAX], &st->lp_buf[PITCH_MAX - pitch], FRAME);
  frame_corr = xy / sqrt(1 + xx * yy);
  frame_corr = log(1.f + exp(5.f * frame_corr)) / log(1 + exp(5.f));
  st->features[NB_BANDS] = st->dnn_pitch;
  st->features[NB_BANDS + 1] = frame_corr - .5f;
}
/* lpcnet_compute_single_frame_features_float, :219-230 with preemphasis :207-215 */
void orc_lpcnet_enc_features(OrcLpcnetEnc *st, const float *pcm, float *features) {
  float x[FRAME];
  int i;
  for (i = 0; i < FRAME; i++) {
    x[i] = pcm[i] + st->mem_preemph;
    st->mem_preemph = -PREEMPH * pcm[i];
  }
  frame_features(st, x);
  memcpy(features, st->features, NB_TOTAL_FEATURES * sizeof(float));
}
void orc_lpcnet_enc_get(const OrcLpcnetEnc *st, float *if_features, float *xcorr_features, float *lpc, float *misc) {
  memcpy(if_features, st->if_features, sizeof(st->if_features));
  memcpy(xcorr_features, st->xcorr_features, sizeof(st->xcorr_features));
  memcpy(lpc, st->lpc, sizeof(st->lpc));
  misc[0] = st->dnn_pitch; misc[1] = st->pitch_filt; misc[2] = st->mem_preemph;
}
