/* oracle/silk_oracle.c — TEST INFRASTRUCTURE ONLY.
 * Plain-C restatements of the SILK integer kernels on the hot path (a15
 * and a17 follow), bit-exact by construction (32-bit wrap-around arithmetic as in the reference).
 * Pinned against the reference's functions in oracle/_ref by tests/test_silk_oracle_vs_ref.py. */
#include <math.h>
#include <stdint.h>
#include <string.h>

static int32_t smulwb(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * (int16_t)b) >> 16); }        /* silk/macros.h:43 */
static int32_t smlawb(int32_t a, int32_t b, int32_t c) { return (int32_t)(a + (((int64_t)b * (int16_t)c) >> 16)); }
static int32_t rshift_round(int32_t a, int s) { return s == 1 ? (a >> 1) + (a & 1) : ((a >> (s - 1)) + 1) >> 1; }
static int16_t sat16(int32_t a) { return (int16_t)(a > 32767 ? 32767 : (a < -32768 ? -32768 : a)); }

/* silk/LPC_analysis_filter.c:49-111 */
void orc_silk_lpc_analysis_filter(int16_t *out, const int16_t *in, const int16_t *B, int len, int d) {
  int ix, j;
  for (ix = d; ix < len; ix++) {
    const int16_t *p = &in[ix - 1];
    uint32_t acc = (uint32_t)((int32_t)p[0] * B[0]);
    for (j = 1; j < d; j++) acc += (uint32_t)((int32_t)p[-j] * B[j]); /* wrap-around adds */
    acc = ((uint32_t)((int32_t)p[1]) << 12) - acc;
    out[ix] = sat16(rshift_round((int32_t)acc, 12));
  }
  memset(out, 0, d * sizeof(int16_t));
}

/* ---- down-sampling resampler, encoder side (silk/resampler.c:77-219, resampler_private_AR2.c,
 * resampler_private_down_FIR.c) for the integer-ratio cases 1:2, 1:3, 1:4, 1:6 ---- */
typedef struct {
  int32_t sIIR[2];
  int32_t sFIR[36];
  int16_t delayBuf[96];
  int32_t fir_order, fs_in_khz, fs_out_khz, input_delay, inv_ratio_q16, batch;
  const int16_t *coefs;
} OrcResampler;

static const int16_t coefs_1_2[2 + 12] = {616, -14323, -10, 39, 58, -46, -84, 120, 184, -315, -541, 1284, 5380, 9024};
static const int16_t coefs_1_3[2 + 18] = {16102, -15162, -13, 0, 20, 26, 5, -31, -43, -4, 65, 90, 7, -157, -248, -44, 593, 1583, 2612, 3271};
static const int16_t coefs_1_4[2 + 18] = {22500, -15099, 3, -14, -20, -15, 2, 25, 37, 25, -16, -71, -107, -79, 50, 292, 623, 982, 1288, 1464};
static const int16_t coefs_1_6[2 + 18] = {27540, -15257, 17, 12, 8, 1, -10, -22, -30, -32, -22, 3, 44, 100, 168, 243, 317, 381, 429, 455};

int orc_resampler_size(void) { return (int)sizeof(OrcResampler); }

int orc_resampler_init(OrcResampler *S, int fs_in, int fs_out) {
  static const int8_t delay_enc[5][3] = {{6, 0, 3}, {0, 7, 3}, {0, 1, 10}, {0, 2, 6}, {18, 10, 12}};
  int ri = fs_in == 8000 ? 0 : fs_in == 12000 ? 1 : fs_in == 16000 ? 2 : fs_in == 24000 ? 3 : fs_in == 48000 ? 4 : -1;
  int ro = fs_out == 8000 ? 0 : fs_out == 12000 ? 1 : fs_out == 16000 ? 2 : -1;
  memset(S, 0, sizeof(*S));
  if (ri < 0 || ro < 0) return -1;
  S->input_delay = delay_enc[ri][ro];
  S->fs_in_khz = fs_in / 1000;
  S->fs_out_khz = fs_out / 1000;
  S->batch = S->fs_in_khz * 10;
  if (fs_out * 2 == fs_in) { S->fir_order = 24; S->coefs = coefs_1_2; }
  else if (fs_out * 3 == fs_in) { S->fir_order = 36; S->coefs = coefs_1_3; }
  else if (fs_out * 4 == fs_in) { S->fir_order = 36; S->coefs = coefs_1_4; }
  else if (fs_out * 6 == fs_in) { S->fir_order = 36; S->coefs = coefs_1_6; }
  else return -1;
  S->inv_ratio_q16 = (int32_t)((((int64_t)fs_in << 14) / fs_out) << 2);
  while ((int32_t)(((int64_t)S->inv_ratio_q16 * fs_out) >> 16) < fs_in) S->inv_ratio_q16++;
  return 0;
}

static void down_fir(OrcResampler *S, int16_t *out, const int16_t *in, int inLen) {
  int32_t buf[480 + 36];
  const int16_t *fir = &S->coefs[2];
  const int order = S->fir_order, half = order / 2;
  int nIn, k;
  memcpy(buf, S->sFIR, order * sizeof(int32_t));
  for (;;) {
    nIn = inLen < S->batch ? inLen : S->batch;
    for (k = 0; k < nIn; k++) { /* AR2, resampler_private_AR2.c:36-55 */
      int32_t o = S->sIIR[0] + (int32_t)((uint32_t)(int32_t)in[k] << 8);
      buf[order + k] = o;
      o = (int32_t)((uint32_t)o << 2);
      S->sIIR[0] = smlawb(S->sIIR[1], o, S->coefs[0]);
      S->sIIR[1] = smulwb(o, S->coefs[1]);
    }
    {
      int32_t idx, maxi = nIn << 16;
      for (idx = 0; idx < maxi; idx += S->inv_ratio_q16) {
        const int32_t *b = buf + (idx >> 16);
        int32_t r = smulwb(b[0] + b[order - 1], fir[0]);
        for (k = 1; k < half; k++) r = smlawb(r, b[k] + b[order - 1 - k], fir[k]);
        *out++ = sat16(rshift_round(r, 6));
      }
    }
    in += nIn;
    inLen -= nIn;
    if (inLen > 1) memcpy(buf, &buf[nIn], order * sizeof(int32_t));
    else break;
  }
  memcpy(S->sFIR, &buf[nIn], order * sizeof(int32_t));
}

/* silk_resampler(), resampler.c:183-219 */
void orc_resampler_process(OrcResampler *S, int16_t *out, const int16_t *in, int inLen) {
  int nSamples = S->fs_in_khz - S->input_delay;
  memcpy(&S->delayBuf[S->input_delay], in, nSamples * sizeof(int16_t));
  down_fir(S, out, S->delayBuf, S->fs_in_khz);
  down_fir(S, &out[S->fs_out_khz], &in[nSamples], inLen - S->fs_in_khz);
  memcpy(S->delayBuf, &in[inLen - S->input_delay], S->input_delay * sizeof(int16_t));
}

/* ---- float analysis kernels -------------------------------------------- */
/* silk/float/inner_product_FLP.c:35-59, energy_FLP.c:35-62: double accumulation, 4-way grouped */
static double inner_product_flp(const float *a, const float *b, int n) {
  int i;
  double r = 0.0;
  for (i = 0; i < n - 3; i += 4)
    r += a[i + 0] * (double)b[i + 0] + a[i + 1] * (double)b[i + 1] + a[i + 2] * (double)b[i + 2] + a[i + 3] * (double)b[i + 3];
  for (; i < n; i++) r += a[i] * (double)b[i];
  return r;
}
/* silk/float/autocorrelation_FLP.c:36-55 */
void orc_silk_autocorrelation(float *results, const float *x, int size, int count) {
  int i;
  if (count > size) count = size;
  for (i = 0; i < count; i++) results[i] = (float)inner_product_flp(x, x + i, size - i);
}
/* silk/float/burg_modified_FLP.c:39-187 */
float orc_silk_burg_modified(float *A, const float *x, float minInvGain, int subfr_length, int nb_subfr, int D) {
  int k, n, s, reached_max_gain;
  double C0, invGain, num, nrg_f, nrg_b, rc, Atmp, tmp1, tmp2;
  const float *xp;
  double Cf[24], Cl[24], CAf[25], CAb[25], Af[24];
  C0 = inner_product_flp(x, x, nb_subfr * subfr_length);
  memset(Cf, 0, sizeof(Cf));
  for (s = 0; s < nb_subfr; s++) {
    xp = x + s * subfr_length;
    for (n = 1; n < D + 1; n++) Cf[n - 1] += inner_product_flp(xp, xp + n, subfr_length - n);
  }
  memcpy(Cl, Cf, sizeof(Cf));
  CAb[0] = CAf[0] = C0 + 1e-5f * C0 + 1e-9f;
  invGain = 1.0f;
  reached_max_gain = 0;
  for (n = 0; n < D; n++) {
    for (s = 0; s < nb_subfr; s++) {
      xp = x + s * subfr_length;
      tmp1 = xp[n];
      tmp2 = xp[subfr_length - n - 1];
      for (k = 0; k < n; k++) {
        Cf[k] -= xp[n] * xp[n - k - 1];                                  /* float product */
        Cl[k] -= xp[subfr_length - n - 1] * xp[subfr_length - n + k];    /* float product */
        Atmp = Af[k];
        tmp1 += xp[n - k - 1] * Atmp;
        tmp2 += xp[subfr_length - n + k] * Atmp;
      }
      for (k = 0; k <= n; k++) {
        CAf[k] -= tmp1 * xp[n - k];
        CAb[k] -= tmp2 * xp[subfr_length - n + k - 1];
      }
    }
    tmp1 = Cf[n];
    tmp2 = Cl[n];
    for (k = 0; k < n; k++) {
      Atmp = Af[k];
      tmp1 += Cl[n - k - 1] * Atmp;
      tmp2 += Cf[n - k - 1] * Atmp;
    }
    CAf[n + 1] = tmp1;
    CAb[n + 1] = tmp2;
    num = CAb[n + 1];
    nrg_b = CAb[0];
    nrg_f = CAf[0];
    for (k = 0; k < n; k++) {
      Atmp = Af[k];
      num += CAb[n - k] * Atmp;
      nrg_b += CAb[k + 1] * Atmp;
      nrg_f += CAf[k + 1] * Atmp;
    }
    rc = -2.0 * num / (nrg_f + nrg_b);
    tmp1 = invGain * (1.0 - rc * rc);
    if (tmp1 <= minInvGain) {
      rc = sqrt(1.0 - minInvGain / invGain);
      if (num > 0) rc = -rc;
      invGain = minInvGain;
      reached_max_gain = 1;
    } else invGain = tmp1;
    for (k = 0; k < (n + 1) >> 1; k++) {
      tmp1 = Af[k];
      tmp2 = Af[n - k - 1];
      Af[k] = tmp1 + rc * tmp2;
      Af[n - k - 1] = tmp2 + rc * tmp1;
    }
    Af[n] = rc;
    if (reached_max_gain) {
      for (k = n + 1; k < D; k++) Af[k] = 0.0;
      break;
    }
    for (k = 0; k <= n + 1; k++) {
      tmp1 = CAf[k];
      CAf[k] += rc * CAb[n - k + 1];
      CAb[n - k + 1] += rc * tmp1;
    }
  }
  if (reached_max_gain) {
    for (k = 0; k < D; k++) A[k] = (float)(-Af[k]);
    for (s = 0; s < nb_subfr; s++) C0 -= inner_product_flp(x + s * subfr_length, x + s * subfr_length, D);
    nrg_f = C0 * invGain;
  } else {
    nrg_f = CAf[0];
    tmp1 = 1.0;
    for (k = 0; k < D; k++) {
      Atmp = Af[k];
      nrg_f += CAf[k + 1] * Atmp;
      tmp1 += Atmp * Atmp;
      A[k] = (float)(-Atmp);
    }
    nrg_f -= 1e-5f * C0 * tmp1;
  }
  return (float)nrg_f;
}
