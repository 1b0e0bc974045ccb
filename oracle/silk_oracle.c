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

/* ---- silk_resampler: every rate pair silk_resampler_init accepts (silk/resampler.c:77-219): encoder side 8/12/16/24/48 kHz -> 8/12/16 kHz,
 * decoder side 8/12/16 kHz -> 8/12/16/24/48 kHz.  Four functions: copy, the 2x all-pass up-sampler (resampler_private_up2_HQ.c), 2x up-sampling
 * followed by 1/12-sample FIR interpolation (resampler_private_IIR_FIR.c), AR2 + FIR down-sampling with 1, 2 or 3 interpolation phases
 * (resampler_private_AR2.c, resampler_private_down_FIR.c). ---- */
enum { ORC_RS_COPY = 0, ORC_RS_UP2 = 1, ORC_RS_IIR_FIR = 2, ORC_RS_DOWN = 3 };
typedef struct {
  int32_t sIIR[6];
  int32_t sFIR[36];    /* down-FIR: int32 history; IIR_FIR: the first 8 int16 of this storage (silk_resampler_state_struct's union) */
  int16_t delayBuf[96];
  int32_t function, fir_order, fir_fracs, fs_in_khz, fs_out_khz, input_delay, inv_ratio_q16, batch;
  const int16_t *coefs;
} OrcResampler;

/* filter constants: silk/resampler_rom.c:42-96, resampler_rom.h:45-46 */
static const int16_t coefs_3_4[2 + 27] = {-20694, -13867, -49, 64, 17, -157, 353, -496, 163, 11047, 22205, -39, 6, 91, -170, 186, 23, -896, 6336, 19928,
                                          -19, -36, 102, -89, -24, 328, -951, 2568, 15909};
static const int16_t coefs_2_3[2 + 18] = {-14457, -14019, 64, 128, -122, 36, 310, -768, 584, 9267, 17733, 12, 128, 18, -142, 288, -117, -865, 4123, 14459};
static const int16_t coefs_1_2[2 + 12] = {616, -14323, -10, 39, 58, -46, -84, 120, 184, -315, -541, 1284, 5380, 9024};
static const int16_t coefs_1_3[2 + 18] = {16102, -15162, -13, 0, 20, 26, 5, -31, -43, -4, 65, 90, 7, -157, -248, -44, 593, 1583, 2612, 3271};
static const int16_t coefs_1_4[2 + 18] = {22500, -15099, 3, -14, -20, -15, 2, 25, 37, 25, -16, -71, -107, -79, 50, 292, 623, 982, 1288, 1464};
static const int16_t coefs_1_6[2 + 18] = {27540, -15257, 17, 12, 8, 1, -10, -22, -30, -32, -22, 3, 44, 100, 168, 243, 317, 381, 429, 455};
static const int16_t frac_fir_12[12][4] = {{189, -600, 617, 30567}, {117, -159, -1070, 29704}, {52, 221, -2392, 28276}, {-4, 529, -3350, 26341},
                                           {-48, 758, -3956, 23973}, {-80, 905, -4235, 21254}, {-99, 972, -4222, 18278}, {-107, 967, -3957, 15143},
                                           {-103, 896, -3487, 11950}, {-91, 773, -2865, 8798}, {-71, 611, -2143, 5784}, {-46, 425, -1375, 2996}};
static const int16_t up2_hq_0[3] = {1746, 14986, 39083 - 65536}, up2_hq_1[3] = {6854, 25769, 55542 - 65536};

int orc_resampler_size(void) { return (int)sizeof(OrcResampler); }

static int rate_id(int fs) { return fs == 8000 ? 0 : fs == 12000 ? 1 : fs == 16000 ? 2 : fs == 24000 ? 3 : fs == 48000 ? 4 : -1; }

/* silk_resampler_init (resampler.c:77-180) */
int orc_resampler_init2(OrcResampler *S, int fs_in, int fs_out, int for_enc) {
  static const int8_t delay_enc[5][3] = {{6, 0, 3}, {0, 7, 3}, {0, 1, 10}, {0, 2, 6}, {18, 10, 12}};
  static const int8_t delay_dec[3][5] = {{4, 0, 2, 0, 0}, {0, 9, 4, 7, 4}, {0, 3, 12, 7, 7}};
  const int ri = rate_id(fs_in), ro = rate_id(fs_out);
  int up2x = 0;
  memset(S, 0, sizeof(*S));
  if (ri < 0 || ro < 0) return -1;
  if (for_enc) { if (ro > 2) return -1; S->input_delay = delay_enc[ri][ro]; }
  else { if (ri > 2) return -1; S->input_delay = delay_dec[ri][ro]; }
  S->fs_in_khz = fs_in / 1000;
  S->fs_out_khz = fs_out / 1000;
  S->batch = S->fs_in_khz * 10;
  if (fs_out > fs_in) {
    if (fs_out == 2 * fs_in) S->function = ORC_RS_UP2;
    else { S->function = ORC_RS_IIR_FIR; up2x = 1; }
  } else if (fs_out < fs_in) {
    S->function = ORC_RS_DOWN;
    S->fir_fracs = 1;
    if (fs_out * 4 == fs_in * 3) { S->fir_fracs = 3; S->fir_order = 18; S->coefs = coefs_3_4; }
    else if (fs_out * 3 == fs_in * 2) { S->fir_fracs = 2; S->fir_order = 18; S->coefs = coefs_2_3; }
    else if (fs_out * 2 == fs_in) { S->fir_order = 24; S->coefs = coefs_1_2; }
    else if (fs_out * 3 == fs_in) { S->fir_order = 36; S->coefs = coefs_1_3; }
    else if (fs_out * 4 == fs_in) { S->fir_order = 36; S->coefs = coefs_1_4; }
    else if (fs_out * 6 == fs_in) { S->fir_order = 36; S->coefs = coefs_1_6; }
    else return -1;
  } else S->function = ORC_RS_COPY;
  S->inv_ratio_q16 = (int32_t)((((int64_t)fs_in << (14 + up2x)) / fs_out) << 2);
  while ((int32_t)(((int64_t)S->inv_ratio_q16 * fs_out) >> 16) < (fs_in << up2x)) S->inv_ratio_q16++;
  return 0;
}
int orc_resampler_init(OrcResampler *S, int fs_in, int fs_out) { return orc_resampler_init2(S, fs_in, fs_out, 1); }

/* resampler_private_down_FIR.c:36-193 with resampler_private_AR2.c:36-55 */
static void down_fir(OrcResampler *S, int16_t *out, const int16_t *in, int inLen) {
  int32_t buf[480 + 36];
  const int16_t *fir = &S->coefs[2];
  const int order = S->fir_order, half = order / 2;
  int nIn, k;
  memcpy(buf, S->sFIR, order * sizeof(int32_t));
  for (;;) {
    nIn = inLen < S->batch ? inLen : S->batch;
    for (k = 0; k < nIn; k++) {
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
        int32_t r;
        if (order == 18) {   /* interpolating phases: the fractional position picks one of fir_fracs coefficient sets, mirrored for the second half */
          const int ind = smulwb(idx & 0xFFFF, S->fir_fracs);
          const int16_t *c0 = &fir[9 * ind], *c1 = &fir[9 * (S->fir_fracs - 1 - ind)];
          r = smulwb(b[0], c0[0]);
          for (k = 1; k < 9; k++) r = smlawb(r, b[k], c0[k]);
          for (k = 0; k < 9; k++) r = smlawb(r, b[17 - k], c1[k]);
        } else {
          r = smulwb(b[0] + b[order - 1], fir[0]);
          for (k = 1; k < half; k++) r = smlawb(r, b[k] + b[order - 1 - k], fir[k]);
        }
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

/* resampler_private_up2_HQ.c:36-103: three all-pass sections per output phase, state in Q10 */
static void up2_hq(int32_t *S, int16_t *out, const int16_t *in, int len) {
  int k;
  for (k = 0; k < len; k++) {
    const int32_t in32 = (int32_t)((uint32_t)(int32_t)in[k] << 10);
    int32_t Y, X, o1, o2;
    Y = in32 - S[0]; X = smulwb(Y, up2_hq_0[0]); o1 = S[0] + X; S[0] = in32 + X;
    Y = o1 - S[1]; X = smulwb(Y, up2_hq_0[1]); o2 = S[1] + X; S[1] = o1 + X;
    Y = o2 - S[2]; X = smlawb(Y, Y, up2_hq_0[2]); o1 = S[2] + X; S[2] = o2 + X;
    out[2 * k] = sat16(rshift_round(o1, 10));
    Y = in32 - S[3]; X = smulwb(Y, up2_hq_1[0]); o1 = S[3] + X; S[3] = in32 + X;
    Y = o1 - S[4]; X = smulwb(Y, up2_hq_1[1]); o2 = S[4] + X; S[4] = o1 + X;
    Y = o2 - S[5]; X = smlawb(Y, Y, up2_hq_1[2]); o1 = S[5] + X; S[5] = o2 + X;
    out[2 * k + 1] = sat16(rshift_round(o1, 10));
  }
}

/* resampler_private_IIR_FIR.c:35-107 */
static void iir_fir(OrcResampler *S, int16_t *out, const int16_t *in, int inLen) {
  int16_t buf[2 * 480 + 8];
  int16_t *sfir16 = (int16_t *)S->sFIR;
  int nIn, k;
  memcpy(buf, sfir16, 8 * sizeof(int16_t));
  for (;;) {
    int32_t idx, maxi;
    nIn = inLen < S->batch ? inLen : S->batch;
    up2_hq(S->sIIR, &buf[8], in, nIn);
    maxi = nIn << 17;
    for (idx = 0; idx < maxi; idx += S->inv_ratio_q16) {
      const int t = smulwb(idx & 0xFFFF, 12);
      const int16_t *b = &buf[idx >> 16];
      int32_t r = (int32_t)b[0] * frac_fir_12[t][0];
      for (k = 1; k < 4; k++) r += (int32_t)b[k] * frac_fir_12[t][k];
      for (k = 0; k < 4; k++) r += (int32_t)b[4 + k] * frac_fir_12[11 - t][3 - k];
      *out++ = sat16(rshift_round(r, 15));
    }
    in += nIn;
    inLen -= nIn;
    if (inLen > 0) memcpy(buf, &buf[nIn << 1], 8 * sizeof(int16_t));
    else break;
  }
  memcpy(sfir16, &buf[nIn << 1], 8 * sizeof(int16_t));
}

static void rs_run(OrcResampler *S, int16_t *out, const int16_t *in, int inLen) {
  if (S->function == ORC_RS_UP2) up2_hq(S->sIIR, out, in, inLen);
  else if (S->function == ORC_RS_IIR_FIR) iir_fir(S, out, in, inLen);
  else down_fir(S, out, in, inLen);
}
/* silk_resampler(), resampler.c:183-219 */
void orc_resampler_process(OrcResampler *S, int16_t *out, const int16_t *in, int inLen) {
  int nSamples = S->fs_in_khz - S->input_delay;
  memcpy(&S->delayBuf[S->input_delay], in, nSamples * sizeof(int16_t));
  if (S->function == ORC_RS_COPY) {
    memcpy(out, S->delayBuf, S->fs_in_khz * sizeof(int16_t));
    memcpy(&out[S->fs_out_khz], &in[nSamples], (inLen - S->fs_in_khz) * sizeof(int16_t));
  } else {
    rs_run(S, out, S->delayBuf, S->fs_in_khz);
    rs_run(S, &out[S->fs_out_khz], &in[nSamples], inLen - S->fs_in_khz);
  }
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

/* ---- gain processing and quantisation of a SILK frame: silk_process_gains_FLP (silk/float/process_gains_FLP.c:35-103),
 * silk_gains_quant (silk/gain_quant.c:40-95), silk_lin2log / silk_log2lin (silk/lin2log.c, log2lin.c), silk_CLZ_FRAC (silk/Inlines.h:52-62). ---- */
static int32_t orc_ror32(int32_t a32, int rot) { /* silk_ROR32, silk/macros.h */
  const uint32_t x = (uint32_t)a32, r = (uint32_t)rot, m = (uint32_t)-rot;
  if (rot == 0) return a32;
  if (rot < 0) return (int32_t)((x << m) | (x >> (32 - m)));
  return (int32_t)((x << (32 - r)) | (x >> r));
}
int32_t orc_silk_lin2log(int32_t inLin) {
  const int lz = inLin ? __builtin_clz((uint32_t)inLin) : 32;
  const int32_t frac_Q7 = orc_ror32(inLin, 24 - lz) & 0x7f;
  return smlawb(frac_Q7, frac_Q7 * (128 - frac_Q7), 179) + ((31 - lz) << 7);
}
int32_t orc_silk_log2lin(int32_t inLog_Q7) {
  int32_t out, frac_Q7;
  if (inLog_Q7 < 0) return 0;
  if (inLog_Q7 >= 3967) return 0x7FFFFFFF;
  out = 1 << (inLog_Q7 >> 7);
  frac_Q7 = inLog_Q7 & 0x7F;
  if (inLog_Q7 < 2048) out = out + ((out * smlawb(frac_Q7, (int16_t)frac_Q7 * (int32_t)(int16_t)(128 - frac_Q7), -174)) >> 7);
  else out = out + (out >> 7) * smlawb(frac_Q7, (int16_t)frac_Q7 * (int32_t)(int16_t)(128 - frac_Q7), -174);
  return out;
}
#define G_OFFSET ((2 * 128) / 6 + 16 * 128)
#define G_SCALE_Q16 ((65536 * (64 - 1)) / (((88 - 2) * 128) / 6))
#define G_INV_SCALE_Q16 ((65536 * (((88 - 2) * 128) / 6)) / (64 - 1))
static int limit_int(int a, int lo, int hi) { return a < lo ? lo : a > hi ? hi : a; } /* silk_LIMIT with lo <= hi */
void orc_silk_gains_quant(int8_t *ind, int32_t *gain_Q16, int8_t *prev_ind, int conditional, int nb_subfr) {
  int k;
  for (k = 0; k < nb_subfr; k++) {
    ind[k] = (int8_t)smulwb(G_SCALE_Q16, orc_silk_lin2log(gain_Q16[k]) - G_OFFSET);
    if (ind[k] < *prev_ind) ind[k]++;
    ind[k] = (int8_t)limit_int(ind[k], 0, 63);
    if (k == 0 && conditional == 0) {
      ind[k] = (int8_t)limit_int(ind[k], *prev_ind - 4, 63);
      *prev_ind = ind[k];
    } else {
      int thr;
      ind[k] = ind[k] - *prev_ind;
      thr = 2 * 36 - 64 + *prev_ind;
      if (ind[k] > thr) ind[k] = (int8_t)(thr + ((ind[k] - thr + 1) >> 1));
      ind[k] = (int8_t)limit_int(ind[k], -4, 36);
      if (ind[k] > thr) {
        *prev_ind += (ind[k] << 1) - thr;
        *prev_ind = *prev_ind < 63 ? *prev_ind : 63;
      } else *prev_ind += ind[k];
      ind[k] -= -4;
    }
    {
      const int32_t v = smulwb(G_INV_SCALE_Q16, *prev_ind) + G_OFFSET;
      gain_Q16[k] = orc_silk_log2lin(v < 3967 ? v : 3967);
    }
  }
}
typedef struct { /* one stream's view of silk_encoder_state_FLP / silk_encoder_control_FLP: what silk_process_gains_FLP reads and writes */
  float Gains[4], ResNrg[4];
  float LTPredCodGain, input_quality, coding_quality, Lambda;
  int32_t GainsUnq_Q16[4];
  int32_t signalType, quantOffsetType, SNR_dB_Q7, input_tilt_Q15, speech_activity_Q8, condCoding;
  int8_t GainsIndices[4];
  int8_t LastGainIndex, lastGainIndexPrev, pad[2];
} OrcSilkGains;
void orc_silk_process_gains(OrcSilkGains *g, int nb_subfr, int subfr_length, int nStatesDelayedDecision) {
  static const int16_t quant_offsets_Q10[2][2] = {{100, 240}, {32, 100}}; /* silk_Quantization_Offsets_Q10, silk/tables_other.c */
  int32_t pGains_Q16[4];
  float s, InvMaxSqrVal, gain, quant_offset;
  int k;
  if (g->signalType == 2) {
    s = 1.0f - 0.5f * (float)(1.0 / (1.0 + exp(-(0.25f * (g->LTPredCodGain - 12.0f)))));
    for (k = 0; k < nb_subfr; k++) g->Gains[k] *= s;
  }
  InvMaxSqrVal = (float)(pow(2.0f, 0.33f * (21.0f - g->SNR_dB_Q7 * (1 / 128.0f))) / subfr_length);
  for (k = 0; k < nb_subfr; k++) {
    gain = g->Gains[k];
    gain = (float)sqrt(gain * gain + g->ResNrg[k] * InvMaxSqrVal);
    g->Gains[k] = gain < 32767.0f ? gain : 32767.0f;
  }
  for (k = 0; k < nb_subfr; k++) pGains_Q16[k] = (int32_t)(g->Gains[k] * 65536.0f);
  memcpy(g->GainsUnq_Q16, pGains_Q16, nb_subfr * sizeof(int32_t));
  g->lastGainIndexPrev = g->LastGainIndex;
  orc_silk_gains_quant(g->GainsIndices, pGains_Q16, &g->LastGainIndex, g->condCoding == 2, nb_subfr); /* CODE_CONDITIONALLY == 2 */
  for (k = 0; k < nb_subfr; k++) g->Gains[k] = pGains_Q16[k] / 65536.0f;
  if (g->signalType == 2) g->quantOffsetType = (g->LTPredCodGain + g->input_tilt_Q15 * (1.0f / 32768.0f) > 1.0f) ? 0 : 1;
  quant_offset = quant_offsets_Q10[g->signalType >> 1][g->quantOffsetType] / 1024.0f;
  g->Lambda = 1.2f + -0.05f * nStatesDelayedDecision + -0.2f * g->speech_activity_Q8 * (1.0f / 256.0f) + -0.1f * g->input_quality + -0.2f * g->coding_quality +
              0.8f * quant_offset;
}
int orc_silk_gains_size(void) { return (int)sizeof(OrcSilkGains); }

/* ---- noise shaping analysis of a SILK frame: silk_noise_shape_analysis_FLP (silk/float/noise_shape_analysis_FLP.c:36-350) with silk_apply_sine_window_FLP
 * (float/apply_sine_window_FLP.c:38-80), silk_warped_autocorrelation_FLP (float/warped_autocorrelation_FLP.c:35-75), silk_autocorrelation_FLP,
 * silk_schur_FLP (float/schur_FLP.c:34-70), silk_k2a_FLP (float/k2a_FLP.c:35-55), silk_bwexpander_FLP (float/bwexpander_FLP.c:35-50). ---- */
#define NS_MAX_ORDER 24
typedef struct { /* what the function reads and writes in silk_encoder_state_FLP / silk_encoder_control_FLP; == OpusB200SilkNoiseShape */
  int32_t SNR_dB_Q7, input_quality_bands_Q15[2], useCBR, speech_activity_Q8, signalType, warping_Q16;
  float LTPCorr, predGain;
  int32_t pitchL[4];
  float HarmShapeGain_smth, Tilt_smth;                     /* in / out (silk_shape_state_FLP) */
  float input_quality, coding_quality;                     /* out */
  int32_t quantOffsetType;                                 /* out */
  float AR[4 * NS_MAX_ORDER], Gains[4], LF_MA_shp[4], LF_AR_shp[4], HarmShapeGain[4], Tilt[4];   /* out */
} OrcSilkNoiseShape;
int orc_silk_noise_shape_size(void) { return (int)sizeof(OrcSilkNoiseShape); }
static void ns_sine_window(float *w, const float *px, int win_type, int length) {
  const float freq = 3.1415926536f / (length + 1), c = 2.0f - freq * freq;
  float S0 = win_type < 2 ? 0.0f : 1.0f, S1 = win_type < 2 ? freq : 0.5f * c;
  int k;
  for (k = 0; k < length; k += 4) {
    w[k + 0] = px[k + 0] * 0.5f * (S0 + S1);
    w[k + 1] = px[k + 1] * S1;
    S0 = c * S1 - S0;
    w[k + 2] = px[k + 2] * 0.5f * (S1 + S0);
    w[k + 3] = px[k + 3] * S0;
    S1 = c * S0 - S1;
  }
}
static void ns_bwexpander(float *ar, int d, float chirp) {
  float cfac = chirp;
  int i;
  for (i = 0; i < d - 1; i++) { ar[i] *= cfac; cfac *= chirp; }
  ar[d - 1] *= cfac;
}
static void ns_limit(float *coefs, float lambda, float limit, int order, int warped) { /* warped_true2monic_coefs / limit_coefs */
  float tmp, maxabs, chirp, gain = 0;
  int i, iter, ind = 0;
  if (warped) {
    for (i = order - 1; i > 0; i--) coefs[i - 1] -= lambda * coefs[i];
    gain = (1.0f - lambda * lambda) / (1.0f + lambda * coefs[0]);
    for (i = 0; i < order; i++) coefs[i] *= gain;
  }
  for (iter = 0; iter < 10; iter++) {
    maxabs = -1.0f;
    for (i = 0; i < order; i++) {
      tmp = coefs[i] < 0 ? -coefs[i] : coefs[i];
      if (tmp > maxabs) { maxabs = tmp; ind = i; }
    }
    if (maxabs <= limit) return;
    if (warped) {
      for (i = 1; i < order; i++) coefs[i - 1] += lambda * coefs[i];
      gain = 1.0f / gain;
      for (i = 0; i < order; i++) coefs[i] *= gain;
    }
    chirp = 0.99f - (0.8f + 0.1f * iter) * (maxabs - limit) / (maxabs * (ind + 1));
    ns_bwexpander(coefs, order, chirp);
    if (warped) {
      for (i = order - 1; i > 0; i--) coefs[i - 1] -= lambda * coefs[i];
      gain = (1.0f - lambda * lambda) / (1.0f + lambda * coefs[0]);
      for (i = 0; i < order; i++) coefs[i] *= gain;
    }
  }
}
/* pitch_res: the frame's residual from the pitch analysis; x: the frame of the input with la_shape samples in front and shapeWinLength - la_shape -
 * subfr_length behind the last subframe's start */
void orc_silk_noise_shape_analysis(OrcSilkNoiseShape *c, const float *pitch_res, const float *x, int fs_kHz, int nb_subfr, int shapingLPCOrder) {
  const int subfr_length = 5 * fs_kHz, la_shape = 5 * fs_kHz, shapeWinLength = 5 * fs_kHz + 2 * la_shape, order = shapingLPCOrder;
  float SNR_adj_dB, HarmShapeGain, Tilt, nrg, strength, BWExp, gain_mult, gain_add, b, warping, x_windowed[240], auto_corr[NS_MAX_ORDER + 1], rc[NS_MAX_ORDER + 1];
  const float *x_ptr = x - la_shape;
  int k, i, n;
  SNR_adj_dB = c->SNR_dB_Q7 * (1 / 128.0f);
  c->input_quality = 0.5f * (c->input_quality_bands_Q15[0] + c->input_quality_bands_Q15[1]) * (1.0f / 32768.0f);
  c->coding_quality = (float)(1.0 / (1.0 + exp(-(0.25f * (SNR_adj_dB - 20.0f)))));
  if (c->useCBR == 0) {
    b = 1.0f - c->speech_activity_Q8 * (1.0f / 256.0f);
    SNR_adj_dB -= 2.0f * c->coding_quality * (0.5f + 0.5f * c->input_quality) * b * b;
  }
  if (c->signalType == 2) SNR_adj_dB += 2.0f * c->LTPCorr;
  else SNR_adj_dB += (-0.4f * c->SNR_dB_Q7 * (1 / 128.0f) + 6.0f) * (1.0f - c->input_quality);
  if (c->signalType == 2) c->quantOffsetType = 0;
  else {
    const int nSamples = 2 * fs_kHz, nSegs = (5 * nb_subfr) / 2;
    float energy_variation = 0.0f, log_energy_prev = 0.0f, log_energy;
    const float *p = pitch_res;
    for (k = 0; k < nSegs; k++) {
      nrg = (float)nSamples + (float)inner_product_flp(p, p, nSamples);
      log_energy = (float)(3.32192809488736 * log10((double)nrg));
      if (k > 0) energy_variation += log_energy - log_energy_prev < 0 ? -(log_energy - log_energy_prev) : log_energy - log_energy_prev;
      log_energy_prev = log_energy;
      p += nSamples;
    }
    c->quantOffsetType = energy_variation > 0.6f * (nSegs - 1) ? 0 : 1;
  }
  strength = 1e-3f * c->predGain;
  BWExp = 0.94f / (1.0f + strength * strength);
  warping = (float)c->warping_Q16 / 65536.0f + 0.01f * c->coding_quality;
  for (k = 0; k < nb_subfr; k++) {
    const int flat_part = fs_kHz * 3, slope_part = (shapeWinLength - flat_part) / 2;
    float *AR = &c->AR[k * NS_MAX_ORDER];
    ns_sine_window(x_windowed, x_ptr, 1, slope_part);
    memcpy(x_windowed + slope_part, x_ptr + slope_part, flat_part * sizeof(float));
    ns_sine_window(x_windowed + slope_part + flat_part, x_ptr + slope_part + flat_part, 2, slope_part);
    x_ptr += subfr_length;
    if (c->warping_Q16 > 0) {
      double state[NS_MAX_ORDER + 1] = {0}, C[NS_MAX_ORDER + 1] = {0}, tmp1, tmp2;
      for (n = 0; n < shapeWinLength; n++) {
        tmp1 = x_windowed[n];
        for (i = 0; i < order; i += 2) {
          tmp2 = state[i] + warping * state[i + 1] - warping * tmp1;
          state[i] = tmp1;
          C[i] += state[0] * tmp1;
          tmp1 = state[i + 1] + warping * state[i + 2] - warping * tmp2;
          state[i + 1] = tmp2;
          C[i + 1] += state[0] * tmp2;
        }
        state[order] = tmp1;
        C[order] += state[0] * tmp1;
      }
      for (i = 0; i < order + 1; i++) auto_corr[i] = (float)C[i];
    } else {
      for (i = 0; i < order + 1; i++) auto_corr[i] = (float)inner_product_flp(x_windowed, x_windowed + i, shapeWinLength - i);
    }
    auto_corr[0] += auto_corr[0] * 3e-5f + 1.0f;
    { /* silk_schur_FLP */
      double Cs[NS_MAX_ORDER + 1][2] = {{0}}, Ctmp1, Ctmp2, rc_tmp;
      for (i = 0; i <= order; i++) Cs[i][0] = Cs[i][1] = auto_corr[i];
      for (i = 0; i < order; i++) {
        const double den = Cs[0][1] > 1e-9f ? Cs[0][1] : 1e-9f; /* silk_max_float is a macro: the comparison and the result stay double */
        rc_tmp = -Cs[i + 1][0] / den;
        rc[i] = (float)rc_tmp;
        for (n = 0; n < order - i; n++) {
          Ctmp1 = Cs[n + i + 1][0];
          Ctmp2 = Cs[n][1];
          Cs[n + i + 1][0] = Ctmp1 + Ctmp2 * rc_tmp;
          Cs[n][1] = Ctmp2 + Ctmp1 * rc_tmp;
        }
      }
      nrg = (float)Cs[0][1];
    }
    for (i = 0; i < order; i++) { /* silk_k2a_FLP */
      const float rck = rc[i];
      for (n = 0; n < (i + 1) >> 1; n++) {
        const float tmp1 = AR[n], tmp2 = AR[i - n - 1];
        AR[n] = tmp1 + tmp2 * rck;
        AR[i - n - 1] = tmp2 + tmp1 * rck;
      }
      AR[i] = -rck;
    }
    c->Gains[k] = (float)sqrt(nrg);
    if (c->warping_Q16 > 0) { /* warped_gain */
      const float lambda = -warping;
      float gain = AR[order - 1];
      for (i = order - 2; i >= 0; i--) gain = lambda * gain + AR[i];
      c->Gains[k] *= (float)(1.0f / (1.0f - lambda * gain));
    }
    ns_bwexpander(AR, order, BWExp);
    ns_limit(AR, warping, 3.999f, order, c->warping_Q16 > 0);
  }
  gain_mult = (float)pow(2.0f, -0.16f * SNR_adj_dB);
  gain_add = (float)pow(2.0f, 0.16f * 2);
  for (k = 0; k < nb_subfr; k++) { c->Gains[k] *= gain_mult; c->Gains[k] += gain_add; }
  strength = 4.0f * (1.0f + 0.5f * (c->input_quality_bands_Q15[0] * (1.0f / 32768.0f) - 1.0f));
  strength *= c->speech_activity_Q8 * (1.0f / 256.0f);
  if (c->signalType == 2) {
    for (k = 0; k < nb_subfr; k++) {
      b = 0.2f / fs_kHz + 3.0f / c->pitchL[k];
      c->LF_MA_shp[k] = -1.0f + b;
      c->LF_AR_shp[k] = 1.0f - b - b * strength;
    }
    Tilt = -0.25f - (1 - 0.25f) * 0.35f * c->speech_activity_Q8 * (1.0f / 256.0f);
  } else {
    b = 1.3f / fs_kHz;
    c->LF_MA_shp[0] = -1.0f + b;
    c->LF_AR_shp[0] = 1.0f - b - b * strength * 0.6f;
    for (k = 1; k < nb_subfr; k++) { c->LF_MA_shp[k] = c->LF_MA_shp[0]; c->LF_AR_shp[k] = c->LF_AR_shp[0]; }
    Tilt = -0.25f;
  }
  if (c->signalType == 2) {
    HarmShapeGain = 0.3f;
    HarmShapeGain += 0.2f * (1.0f - (1.0f - c->coding_quality) * c->input_quality);
    HarmShapeGain *= (float)sqrt(c->LTPCorr);
  } else HarmShapeGain = 0.0f;
  for (k = 0; k < nb_subfr; k++) {
    c->HarmShapeGain_smth += 0.4f * (HarmShapeGain - c->HarmShapeGain_smth);
    c->HarmShapeGain[k] = c->HarmShapeGain_smth;
    c->Tilt_smth += 0.4f * (Tilt - c->Tilt_smth);
    c->Tilt[k] = c->Tilt_smth;
  }
}

/* ---- pitch analysis of a SILK frame: silk_find_pitch_lags_FLP (silk/float/find_pitch_lags_FLP.c:36-132) and silk_pitch_analysis_core_FLP
 * (silk/float/pitch_analysis_core_FLP.c:66-630: decimation to 8 and 4 kHz through silk_resampler_down2 / _down2_3 (silk/resampler_down2.c:37-74,
 * resampler_down2_3.c:41-103, resampler_private_AR2.c:36-54), the three search stages, the lag codebooks of silk/pitch_est_tables.c), with
 * silk_insertion_sort_decreasing_FLP (silk/float/sort_FLP.c:39-83) and the C build of celt_pitch_xcorr (every lag summed in index order). ---- */
static const int8_t pe_cb2[4][11] = {{0, 2, -1, -1, -1, 0, 0, 1, 1, 0, 1}, {0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0}, {0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0},
                                     {0, -1, 2, 1, 0, 1, 1, 0, 0, -1, -1}};                                  /* silk_CB_lags_stage2 */
static const int8_t pe_cb3[4][34] = {
    {0, 0, 1, -1, 0, 1, -1, 0, -1, 1, -2, 2, -2, -2, 2, -3, 2, 3, -3, -4, 3, -4, 4, 4, -5, 5, -6, -5, 6, -7, 6, 5, 8, -9},
    {0, 0, 1, 0, 0, 0, 0, 0, 0, 0, -1, 1, 0, 0, 1, -1, 0, 1, -1, -1, 1, -1, 2, 1, -1, 2, -2, -2, 2, -2, 2, 2, 3, -3},
    {0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 1, -1, 1, 0, 0, 2, 1, -1, 2, -1, -1, 2, -1, 2, 2, -1, 3, -2, -2, -2, 3},
    {0, 1, 0, 0, 1, 0, 1, -1, 2, -1, 2, -1, 2, 3, -2, 3, -2, -2, 4, 4, -3, 5, -3, -4, 6, -4, 6, 5, -5, 8, -6, -5, -7, 9}}; /* silk_CB_lags_stage3 */
static const int8_t pe_range3[3][4][2] = {{{-5, 8}, {-1, 6}, {-1, 6}, {-4, 10}}, {{-6, 10}, {-2, 6}, {-1, 6}, {-5, 10}}, {{-9, 12}, {-3, 7}, {-2, 7}, {-7, 13}}};
static const int8_t pe_nb3[3] = {16, 24, 34};
static const int8_t pe_cb2_10[2][3] = {{0, 1, 0}, {0, 0, 1}};
static const int8_t pe_cb3_10[2][12] = {{0, 0, 1, -1, 1, -1, 2, -2, 2, -2, 3, -3}, {0, 1, 0, 1, -1, 2, -1, 2, -2, 3, -2, 3}};
static const int8_t pe_range3_10[2][2] = {{-3, 7}, {-2, 7}};
static int16_t pe_sat16(int32_t a) { return (int16_t)(a > 32767 ? 32767 : a < -32768 ? -32768 : a); }
static void pe_down2(int32_t *S, int16_t *out, const int16_t *in, int inLen) {
  int k;
  for (k = 0; k < inLen >> 1; k++) {
    int32_t in32 = (int32_t)in[2 * k] * 1024, Y = in32 - S[0], X = smlawb(Y, Y, 39809 - 65536), out32 = S[0] + X;
    S[0] = in32 + X;
    in32 = (int32_t)in[2 * k + 1] * 1024;
    Y = in32 - S[1];
    X = smulwb(Y, 9872);
    out32 = out32 + S[1];
    out32 = out32 + X;
    S[1] = in32 + X;
    out[k] = pe_sat16(rshift_round(out32, 11));
  }
}
static void pe_down2_3(int32_t *S, int16_t *out, const int16_t *in, int inLen) { /* inLen <= RESAMPLER_MAX_BATCH_SIZE_IN (480): one batch */
  static const int16_t C[6] = {-2797, -6507, 4697, 10739, 1567, 8276}; /* silk_Resampler_2_3_COEFS_LQ */
  int32_t buf[480 + 4], *bp = buf;
  int k, counter = inLen;
  memcpy(buf, S, 4 * sizeof(int32_t));
  for (k = 0; k < inLen; k++) { /* silk_resampler_private_AR2 */
    int32_t out32 = S[4] + (int32_t)in[k] * 256;
    buf[4 + k] = out32;
    out32 = (int32_t)((uint32_t)out32 << 2);
    S[4] = smlawb(S[5], out32, C[0]);
    S[5] = smulwb(out32, C[1]);
  }
  while (counter > 2) {
    int32_t r = smulwb(bp[0], C[2]);
    r = smlawb(r, bp[1], C[3]);
    r = smlawb(r, bp[2], C[5]);
    r = smlawb(r, bp[3], C[4]);
    *out++ = pe_sat16(rshift_round(r, 6));
    r = smulwb(bp[1], C[4]);
    r = smlawb(r, bp[2], C[5]);
    r = smlawb(r, bp[3], C[3]);
    r = smlawb(r, bp[4], C[2]);
    *out++ = pe_sat16(rshift_round(r, 6));
    bp += 3;
    counter -= 3;
  }
  memcpy(S, &buf[inLen], 4 * sizeof(int32_t));
}
static void pe_sort_decreasing(float *a, int *idx, int L, int K) {
  int i, j;
  for (i = 0; i < K; i++) idx[i] = i;
  for (i = 1; i < K; i++) {
    const float value = a[i];
    for (j = i - 1; j >= 0 && value > a[j]; j--) { a[j + 1] = a[j]; idx[j + 1] = idx[j]; }
    a[j + 1] = value;
    idx[j + 1] = i;
  }
  for (i = K; i < L; i++) {
    const float value = a[i];
    if (value > a[K - 1]) {
      for (j = K - 2; j >= 0 && value > a[j]; j--) { a[j + 1] = a[j]; idx[j + 1] = idx[j]; }
      a[j + 1] = value;
      idx[j + 1] = i;
    }
  }
}
static float pe_log2(double x) { return (float)(3.32192809488736 * log10(x)); }
static float pe_dot(const float *a, const float *b, int n) { float s = 0; int i; for (i = 0; i < n; i++) s = s + a[i] * b[i]; return s; } /* celt_pitch_xcorr_c */
/* returns 0 voiced / 1 unvoiced */
int orc_silk_pitch_analysis_core(const float *frame, int *pitch_out, int16_t *lagIndex, int8_t *contourIndex, float *LTPCorr, int prevLag, float search_thres1,
                                 float search_thres2, int Fs_kHz, int complexity, int nb_subfr) {
  float frame_8kHz[40 * 8], frame_4kHz[40 * 4], C[4][(288 >> 1) + 5], xcorr[18 * 4 - 2 * 4 + 1], CC[11], threshold, contour_bias, Cmax, CCmax, CCmax_b, CCmax_new_b,
      CCmax_new, lag_log2, prevLag_log2, delta_lag_log2_sqr, energies_st3[4][34][5], cross_corr_st3[4][34][5];
  int16_t frame_8_FIX[40 * 8], frame_4_FIX[40 * 4], frame_in_FIX[16 * 40], d_comp[(288 >> 1) + 5];
  int32_t filt_state[6];
  const float *target_ptr, *basis_ptr;
  double cross_corr, normalizer, energy, energy_tmp;
  int d_srch[24], length_d_srch, length_d_comp, i, k, d, j, CBimax, CBimax_new, lag, start_lag, end_lag, lag_new, cbk_size, lag_counter, nb_cbk_search;
  const int frame_length = (20 + nb_subfr * 5) * Fs_kHz, frame_length_4kHz = (20 + nb_subfr * 5) * 4, frame_length_8kHz = (20 + nb_subfr * 5) * 8,
            sf_length = 5 * Fs_kHz, sf_length_4kHz = 20, sf_length_8kHz = 40, min_lag = 2 * Fs_kHz, min_lag_4kHz = 8, min_lag_8kHz = 16,
            max_lag = 18 * Fs_kHz - 1, max_lag_4kHz = 72, max_lag_8kHz = 18 * 8 - 1;
  const int8_t *Lag_CB_ptr;
  if (Fs_kHz == 16) {
    for (k = frame_length - 1; k >= 0; k--) frame_in_FIX[k] = pe_sat16((int32_t)lrintf(frame[k]));
    memset(filt_state, 0, 2 * sizeof(int32_t));
    pe_down2(filt_state, frame_8_FIX, frame_in_FIX, frame_length);
    for (k = 0; k < frame_length_8kHz; k++) frame_8kHz[k] = (float)frame_8_FIX[k];
  } else if (Fs_kHz == 12) {
    for (k = frame_length - 1; k >= 0; k--) frame_in_FIX[k] = pe_sat16((int32_t)lrintf(frame[k]));
    memset(filt_state, 0, 6 * sizeof(int32_t));
    pe_down2_3(filt_state, frame_8_FIX, frame_in_FIX, frame_length);
    for (k = 0; k < frame_length_8kHz; k++) frame_8kHz[k] = (float)frame_8_FIX[k];
  } else {
    for (k = frame_length_8kHz - 1; k >= 0; k--) frame_8_FIX[k] = pe_sat16((int32_t)lrintf(frame[k]));
  }
  memset(filt_state, 0, 2 * sizeof(int32_t));
  pe_down2(filt_state, frame_4_FIX, frame_8_FIX, frame_length_8kHz);
  for (k = 0; k < frame_length_4kHz; k++) frame_4kHz[k] = (float)frame_4_FIX[k];
  for (i = frame_length_4kHz - 1; i > 0; i--) { /* silk_ADD_SAT16 on floats: (int32)a + b, saturated, truncated to int16 */
    const float t = (int32_t)frame_4kHz[i] + frame_4kHz[i - 1];
    frame_4kHz[i] = (float)(int16_t)(t > 32767 ? 32767 : t < -32768 ? -32768 : t);
  }
  /* first stage, 4 kHz */
  memset(C, 0, sizeof(float) * nb_subfr * ((288 >> 1) + 5));
  target_ptr = &frame_4kHz[sf_length_4kHz << 2];
  for (k = 0; k < nb_subfr >> 1; k++) {
    basis_ptr = target_ptr - min_lag_4kHz;
    for (i = 0; i < max_lag_4kHz - min_lag_4kHz + 1; i++) xcorr[i] = pe_dot(target_ptr, target_ptr - max_lag_4kHz + i, sf_length_8kHz);
    cross_corr = xcorr[max_lag_4kHz - min_lag_4kHz];
    normalizer = inner_product_flp(target_ptr, target_ptr, sf_length_8kHz) + inner_product_flp(basis_ptr, basis_ptr, sf_length_8kHz) + sf_length_8kHz * 4000.0f;
    C[0][min_lag_4kHz] += (float)(2 * cross_corr / normalizer);
    for (d = min_lag_4kHz + 1; d <= max_lag_4kHz; d++) {
      basis_ptr--;
      cross_corr = xcorr[max_lag_4kHz - d];
      normalizer += basis_ptr[0] * (double)basis_ptr[0] - basis_ptr[sf_length_8kHz] * (double)basis_ptr[sf_length_8kHz];
      C[0][d] += (float)(2 * cross_corr / normalizer);
    }
    target_ptr += sf_length_8kHz;
  }
  for (i = max_lag_4kHz; i >= min_lag_4kHz; i--) C[0][i] -= C[0][i] * i / 4096.0f;
  length_d_srch = 4 + 2 * complexity;
  pe_sort_decreasing(&C[0][min_lag_4kHz], d_srch, max_lag_4kHz - min_lag_4kHz + 1, length_d_srch);
  Cmax = C[0][min_lag_4kHz];
  if (Cmax < 0.2f) {
    memset(pitch_out, 0, nb_subfr * sizeof(int));
    *LTPCorr = 0.0f; *lagIndex = 0; *contourIndex = 0;
    return 1;
  }
  threshold = search_thres1 * Cmax;
  for (i = 0; i < length_d_srch; i++) {
    if (C[0][min_lag_4kHz + i] > threshold) d_srch[i] = (d_srch[i] + min_lag_4kHz) << 1;
    else { length_d_srch = i; break; }
  }
  for (i = min_lag_8kHz - 5; i < max_lag_8kHz + 5; i++) d_comp[i] = 0;
  for (i = 0; i < length_d_srch; i++) d_comp[d_srch[i]] = 1;
  for (i = max_lag_8kHz + 3; i >= min_lag_8kHz; i--) d_comp[i] += d_comp[i - 1] + d_comp[i - 2];
  length_d_srch = 0;
  for (i = min_lag_8kHz; i < max_lag_8kHz + 1; i++) if (d_comp[i + 1] > 0) d_srch[length_d_srch++] = i;
  for (i = max_lag_8kHz + 3; i >= min_lag_8kHz; i--) d_comp[i] += d_comp[i - 1] + d_comp[i - 2] + d_comp[i - 3];
  length_d_comp = 0;
  for (i = min_lag_8kHz; i < max_lag_8kHz + 4; i++) if (d_comp[i] > 0) d_comp[length_d_comp++] = (int16_t)(i - 2);
  /* second stage, 8 kHz */
  memset(C, 0, sizeof(C));
  target_ptr = Fs_kHz == 8 ? &frame[20 * 8] : &frame_8kHz[20 * 8];
  for (k = 0; k < nb_subfr; k++) {
    energy_tmp = inner_product_flp(target_ptr, target_ptr, sf_length_8kHz) + 1.0;
    for (j = 0; j < length_d_comp; j++) {
      d = d_comp[j];
      basis_ptr = target_ptr - d;
      cross_corr = inner_product_flp(basis_ptr, target_ptr, sf_length_8kHz);
      if (cross_corr > 0.0f) {
        energy = inner_product_flp(basis_ptr, basis_ptr, sf_length_8kHz);
        C[k][d] = (float)(2 * cross_corr / (energy + energy_tmp));
      } else C[k][d] = 0.0f;
    }
    target_ptr += sf_length_8kHz;
  }
  CCmax = 0.0f; CCmax_b = -1000.0f; CBimax = 0; lag = -1;
  if (prevLag > 0) {
    if (Fs_kHz == 12) prevLag = (prevLag << 1) / 3;
    else if (Fs_kHz == 16) prevLag = prevLag >> 1;
    prevLag_log2 = pe_log2((float)prevLag);
  } else prevLag_log2 = 0;
  if (nb_subfr == 4) {
    cbk_size = 11; Lag_CB_ptr = &pe_cb2[0][0];
    nb_cbk_search = (Fs_kHz == 8 && complexity > 0) ? 11 : 3;
  } else { cbk_size = 3; Lag_CB_ptr = &pe_cb2_10[0][0]; nb_cbk_search = 3; }
  for (k = 0; k < length_d_srch; k++) {
    d = d_srch[k];
    for (j = 0; j < nb_cbk_search; j++) {
      CC[j] = 0.0f;
      for (i = 0; i < nb_subfr; i++) CC[j] += C[i][d + Lag_CB_ptr[i * cbk_size + j]];
    }
    CCmax_new = -1000.0f; CBimax_new = 0;
    for (i = 0; i < nb_cbk_search; i++) if (CC[i] > CCmax_new) { CCmax_new = CC[i]; CBimax_new = i; }
    lag_log2 = pe_log2((float)d);
    CCmax_new_b = CCmax_new - 0.2f * nb_subfr * lag_log2;
    if (prevLag > 0) {
      delta_lag_log2_sqr = lag_log2 - prevLag_log2;
      delta_lag_log2_sqr *= delta_lag_log2_sqr;
      CCmax_new_b -= 0.2f * nb_subfr * (*LTPCorr) * delta_lag_log2_sqr / (delta_lag_log2_sqr + 0.5f);
    }
    if (CCmax_new_b > CCmax_b && CCmax_new > nb_subfr * search_thres2) { CCmax_b = CCmax_new_b; CCmax = CCmax_new; lag = d; CBimax = CBimax_new; }
  }
  if (lag == -1) {
    memset(pitch_out, 0, 4 * sizeof(int));
    *LTPCorr = 0.0f; *lagIndex = 0; *contourIndex = 0;
    return 1;
  }
  *LTPCorr = (float)(CCmax / nb_subfr);
  if (Fs_kHz > 8) {
    const int8_t *range, *cb3;
    int nb3, cb3_size;
    float scratch[22];
    if (Fs_kHz == 12) lag = rshift_round((int32_t)(int16_t)lag * 3, 1);
    else lag = lag << 1;
    lag = lag < min_lag ? min_lag : lag > max_lag ? max_lag : lag;
    start_lag = lag - 2 > min_lag ? lag - 2 : min_lag;
    end_lag = lag + 2 < max_lag ? lag + 2 : max_lag;
    lag_new = lag; CBimax = 0; CCmax = -1000.0f;
    if (nb_subfr == 4) { range = &pe_range3[complexity][0][0]; cb3 = &pe_cb3[0][0]; nb3 = pe_nb3[complexity]; cb3_size = 34; }
    else { range = &pe_range3_10[0][0]; cb3 = &pe_cb3_10[0][0]; nb3 = 12; cb3_size = 12; }
    target_ptr = &frame[sf_length << 2];
    for (k = 0; k < nb_subfr; k++) { /* silk_P_Ana_calc_corr_st3 / _energy_st3 */
      const int lag_low = range[k * 2], lag_high = range[k * 2 + 1], lag_diff = lag_high - lag_low + 1;
      float xc[22], en[22];
      for (i = 0; i < lag_diff; i++) xc[i] = pe_dot(target_ptr, target_ptr - start_lag - lag_high + i, sf_length);
      lag_counter = 0;
      for (j = lag_low; j <= lag_high; j++) scratch[lag_counter++] = xc[lag_high - j];
      for (i = 0; i < nb3; i++) {
        const int idx = cb3[k * cb3_size + i] - lag_low;
        for (j = 0; j < 5; j++) cross_corr_st3[k][i][j] = scratch[idx + j];
      }
      basis_ptr = target_ptr - (start_lag + lag_low);
      energy = inner_product_flp(basis_ptr, basis_ptr, sf_length) + 1e-3;
      en[0] = (float)energy;
      for (i = 1; i < lag_diff; i++) {
        energy -= basis_ptr[sf_length - i] * (double)basis_ptr[sf_length - i];
        energy += basis_ptr[-i] * (double)basis_ptr[-i];
        en[i] = (float)energy;
      }
      for (i = 0; i < nb3; i++) {
        const int idx = cb3[k * cb3_size + i] - lag_low;
        for (j = 0; j < 5; j++) energies_st3[k][i][j] = en[idx + j];
      }
      target_ptr += sf_length;
    }
    lag_counter = 0;
    contour_bias = 0.05f / lag;
    nb_cbk_search = nb3; cbk_size = cb3_size; Lag_CB_ptr = cb3;
    target_ptr = &frame[20 * Fs_kHz];
    energy_tmp = inner_product_flp(target_ptr, target_ptr, nb_subfr * sf_length) + 1.0;
    for (d = start_lag; d <= end_lag; d++) {
      for (j = 0; j < nb_cbk_search; j++) {
        cross_corr = 0.0;
        energy = energy_tmp;
        for (k = 0; k < nb_subfr; k++) { cross_corr += cross_corr_st3[k][j][lag_counter]; energy += energies_st3[k][j][lag_counter]; }
        if (cross_corr > 0.0) { CCmax_new = (float)(2 * cross_corr / energy); CCmax_new *= 1.0f - contour_bias * j; }
        else CCmax_new = 0.0f;
        if (CCmax_new > CCmax && (d + (int)pe_cb3[0][j]) <= max_lag) { CCmax = CCmax_new; lag_new = d; CBimax = j; }
      }
      lag_counter++;
    }
    for (k = 0; k < nb_subfr; k++) {
      const int v = lag_new + Lag_CB_ptr[k * cbk_size + CBimax], hi = 18 * Fs_kHz;
      pitch_out[k] = v < min_lag ? min_lag : v > hi ? hi : v; /* silk_LIMIT( v, min_lag, PE_MAX_LAG_MS * Fs_kHz ) */
    }
    *lagIndex = (int16_t)(lag_new - min_lag);
    *contourIndex = (int8_t)CBimax;
  } else {
    for (k = 0; k < nb_subfr; k++) {
      const int v = lag + Lag_CB_ptr[k * cbk_size + CBimax];
      pitch_out[k] = v < min_lag_8kHz ? min_lag_8kHz : v > 18 * 8 ? 18 * 8 : v;
    }
    *lagIndex = (int16_t)(lag - min_lag_8kHz);
    *contourIndex = (int8_t)CBimax;
  }
  return 0;
}
typedef struct { /* what silk_find_pitch_lags_FLP reads and writes in the encoder structs; == OpusB200SilkPitch */
  int32_t signalType;                                  /* in: 0 = no voice activity; out: 1 unvoiced / 2 voiced (0 stays) */
  int32_t first_frame_after_reset, speech_activity_Q8, prevSignalType, input_tilt_Q15, prevLag, pitchEstimationThreshold_Q16;
  float LTPCorr;                                       /* in / out */
  float predGain;                                      /* out */
  int32_t pitchL[4];                                   /* out */
  int32_t lagIndex, contourIndex;                      /* out */
} OrcSilkPitch;
int orc_silk_pitch_size(void) { return (int)sizeof(OrcSilkPitch); }
/* x: the frame of the input with ltp_mem_length (20 ms) samples in front and la_pitch (2 ms) behind; res [la_pitch + frame_length + ltp_mem_length] out */
void orc_silk_find_pitch_lags(OrcSilkPitch *c, float *res, const float *x, int fs_kHz, int nb_subfr, int lpc_order, int complexity) {
  const int la_pitch = 2 * fs_kHz, frame_length = 5 * nb_subfr * fs_kHz, ltp_mem_length = 20 * fs_kHz, buf_len = la_pitch + frame_length + ltp_mem_length,
            win_length = (5 * nb_subfr + 4) * fs_kHz;
  const float *x_buf = x - ltp_mem_length, *x_buf_ptr = x_buf + buf_len - win_length;
  float auto_corr[17], A[16], refl_coef[16], Wsig[24 * 16], res_nrg, thrhld;
  double Cs[17][2], rc_tmp;
  int i, n;
  ns_sine_window(Wsig, x_buf_ptr, 1, la_pitch);
  memcpy(Wsig + la_pitch, x_buf_ptr + la_pitch, (win_length - (la_pitch << 1)) * sizeof(float));
  ns_sine_window(Wsig + win_length - la_pitch, x_buf_ptr + win_length - la_pitch, 2, la_pitch);
  for (i = 0; i < lpc_order + 1; i++) auto_corr[i] = (float)inner_product_flp(Wsig, Wsig + i, win_length - i);
  auto_corr[0] += auto_corr[0] * 1e-3f + 1;
  for (i = 0; i <= lpc_order; i++) Cs[i][0] = Cs[i][1] = auto_corr[i];
  for (i = 0; i < lpc_order; i++) { /* silk_schur_FLP */
    rc_tmp = -Cs[i + 1][0] / (Cs[0][1] > 1e-9f ? Cs[0][1] : 1e-9f);
    refl_coef[i] = (float)rc_tmp;
    for (n = 0; n < lpc_order - i; n++) {
      const double Ctmp1 = Cs[n + i + 1][0], Ctmp2 = Cs[n][1];
      Cs[n + i + 1][0] = Ctmp1 + Ctmp2 * rc_tmp;
      Cs[n][1] = Ctmp2 + Ctmp1 * rc_tmp;
    }
  }
  res_nrg = (float)Cs[0][1];
  c->predGain = auto_corr[0] / (res_nrg > 1.0f ? res_nrg : 1.0f);
  for (i = 0; i < lpc_order; i++) { /* silk_k2a_FLP */
    const float rck = refl_coef[i];
    for (n = 0; n < (i + 1) >> 1; n++) {
      const float t1 = A[n], t2 = A[i - n - 1];
      A[n] = t1 + t2 * rck;
      A[i - n - 1] = t2 + t1 * rck;
    }
    A[i] = -rck;
  }
  ns_bwexpander(A, lpc_order, 0.99f);
  for (i = lpc_order; i < buf_len; i++) { /* silk_LPC_analysis_filter_FLP (orders 6 .. 16: the products summed left to right) */
    const float *sp = &x_buf[i - 1];
    float pred = sp[0] * A[0];
    for (n = 1; n < lpc_order; n++) pred = pred + sp[-n] * A[n];
    res[i] = sp[1] - pred;
  }
  memset(res, 0, lpc_order * sizeof(float));
  if (c->signalType != 0 && c->first_frame_after_reset == 0) {
    int pitch[4] = {0, 0, 0, 0};
    int16_t lagIndex = 0;
    int8_t contourIndex = 0;
    thrhld = 0.6f;
    thrhld -= 0.004f * lpc_order;
    thrhld -= 0.1f * c->speech_activity_Q8 * (1.0f / 256.0f);
    thrhld -= 0.15f * (c->prevSignalType >> 1);
    thrhld -= 0.1f * c->input_tilt_Q15 * (1.0f / 32768.0f);
    if (orc_silk_pitch_analysis_core(res, pitch, &lagIndex, &contourIndex, &c->LTPCorr, c->prevLag, c->pitchEstimationThreshold_Q16 / 65536.0f, thrhld, fs_kHz,
                                     complexity, nb_subfr) == 0)
      c->signalType = 2;
    else c->signalType = 1;
    for (i = 0; i < 4; i++) c->pitchL[i] = pitch[i];
    c->lagIndex = lagIndex;
    c->contourIndex = contourIndex;
  } else {
    memset(c->pitchL, 0, sizeof(c->pitchL));
    c->lagIndex = 0; c->contourIndex = 0; c->LTPCorr = 0;
  }
}

/* ---- voice activity detection of a SILK frame: silk_VAD_Init / silk_VAD_GetSA_Q8_c / silk_VAD_GetNoiseLevels (silk/VAD.c:46-360),
 * silk_ana_filt_bank_1 (silk/ana_filt_bank_1.c:39-74), silk_SQRT_APPROX (silk/Inlines.h:67-90), silk_sigm_Q15 (silk/sigm_Q15.c:49-76), the VAD / DTX
 * flags of silk_encode_do_VAD_FLP (silk/float/encode_frame_FLP.c:45-82) and silk_HP_variable_cutoff (silk/HP_variable_cutoff.c:39-77).  The record is
 * silk_VAD_state plus the silk_encoder_state fields those functions read and write. ---- */
typedef struct {
  int32_t AnaState[2], AnaState1[2], AnaState2[2], XnrgSubfr[4], NrgRatioSmth_Q8[4], HPstate, NL[4], inv_NL[4], NoiseLevelBias[4], counter;
  int32_t speech_activity_Q8, input_tilt_Q15, input_quality_bands_Q15[4];
  int32_t signalType, noSpeechCounter, inDTX, VAD_flag;
  int32_t prevSignalType, prevLag, variable_HP_smth1_Q15;
} OrcSilkVad;
static int32_t vad_add_pos_sat32(int32_t a, int32_t b) { const uint32_t s = (uint32_t)a + (uint32_t)b; return (s & 0x80000000u) ? 0x7FFFFFFF : (int32_t)s; }
static int32_t vad_sqrt_approx(int32_t x) {
  int lz; int32_t y, frac_Q7;
  if (x <= 0) return 0;
  lz = __builtin_clz((uint32_t)x);
  frac_Q7 = orc_ror32(x, 24 - lz) & 0x7f;
  y = (lz & 1) ? 32768 : 46214;
  y >>= (lz >> 1);
  return smlawb(y, y, (int16_t)213 * (int32_t)(int16_t)frac_Q7);
}
static int vad_sigm_Q15(int in_Q5) {
  static const int32_t slope_Q10[6] = {237, 153, 73, 30, 12, 7}, pos_Q15[6] = {16384, 23955, 28861, 31213, 32178, 32548}, neg_Q15[6] = {16384, 8812, 3906, 1554, 589, 219};
  if (in_Q5 < 0) {
    in_Q5 = -in_Q5;
    if (in_Q5 >= 6 * 32) return 0;
    return neg_Q15[in_Q5 >> 5] - (int16_t)slope_Q10[in_Q5 >> 5] * (int32_t)(int16_t)(in_Q5 & 0x1F);
  }
  if (in_Q5 >= 6 * 32) return 32767;
  return pos_Q15[in_Q5 >> 5] + (int16_t)slope_Q10[in_Q5 >> 5] * (int32_t)(int16_t)(in_Q5 & 0x1F);
}
static int16_t vad_sat16(int32_t a) { return (int16_t)(a > 32767 ? 32767 : a < -32768 ? -32768 : a); }
static int32_t vad_rshift_round(int32_t a, int s) { return s == 1 ? (a >> 1) + (a & 1) : ((a >> (s - 1)) + 1) >> 1; }
static void vad_ana_filt_bank_1(const int16_t *in, int32_t *S, int16_t *outL, int16_t *outH, int N) {
  const int N2 = N >> 1; int k;
  for (k = 0; k < N2; k++) {
    int32_t in32 = (int32_t)in[2 * k] * 1024, Y, X, out_1, out_2;
    Y = in32 - S[0]; X = smlawb(Y, Y, -24290); out_1 = S[0] + X; S[0] = in32 + X;
    in32 = (int32_t)in[2 * k + 1] * 1024;
    Y = in32 - S[1]; X = smulwb(Y, 5394 << 1); out_2 = S[1] + X; S[1] = in32 + X;
    outL[k] = vad_sat16(vad_rshift_round(out_2 + out_1, 11));
    outH[k] = vad_sat16(vad_rshift_round(out_2 - out_1, 11));
  }
}
void orc_silk_vad_init(OrcSilkVad *c) {
  int b;
  memset(c, 0, sizeof(*c));
  for (b = 0; b < 4; b++) {
    const int32_t bias = 50 / (b + 1);
    c->NoiseLevelBias[b] = bias > 1 ? bias : 1;
    c->NL[b] = 100 * c->NoiseLevelBias[b];
    c->inv_NL[b] = 0x7FFFFFFF / c->NL[b];
    c->NrgRatioSmth_Q8[b] = 100 * 256;
  }
  c->counter = 15;
}
static void vad_noise_levels(const int32_t *pX, OrcSilkVad *c) {
  int k, coef, min_coef;
  if (c->counter < 1000) { min_coef = 32767 / ((c->counter >> 4) + 1); c->counter++; } else min_coef = 0;
  for (k = 0; k < 4; k++) {
    int32_t nl = c->NL[k];
    const int32_t nrg = vad_add_pos_sat32(pX[k], c->NoiseLevelBias[k]), inv_nrg = 0x7FFFFFFF / nrg;
    if (nrg > (int32_t)((uint32_t)nl << 3)) coef = 1024 >> 3;
    else if (nrg < nl) coef = 1024;
    else coef = smulwb((int32_t)(((int64_t)inv_nrg * nl) >> 16), 1024 << 1);
    if (coef < min_coef) coef = min_coef;
    c->inv_NL[k] = smlawb(c->inv_NL[k], inv_nrg - c->inv_NL[k], coef);
    nl = 0x7FFFFFFF / c->inv_NL[k];
    c->NL[k] = nl < 0x00FFFFFF ? nl : 0x00FFFFFF;
  }
}
/* silk_VAD_GetSA_Q8_c: pIn = frame_length samples (inputBuf + 1); frame_length = 10 or 20 ms at fs_kHz. */
void orc_silk_vad_get_sa(OrcSilkVad *c, const int16_t *pIn, int frame_length, int fs_kHz) {
  int16_t X[512 / 2 + 512 / 4 + 512 / 8 + 512 / 8 + 512 / 4];
  int32_t Xnrg[4], ratio_Q8[4], sumSquared = 0, speech_nrg, smooth_coef_Q16;
  int X_offset[4], b, s, i, SA_Q15, pSNR_dB_Q7, input_tilt = 0, SNR_Q7;
  static const int32_t tiltWeights[4] = {30000, 6000, -12000, -12000};
  const int d1 = frame_length >> 1, d2 = frame_length >> 2, d3 = frame_length >> 3;
  int16_t hp_tmp;
  X_offset[0] = 0; X_offset[1] = d3 + d2; X_offset[2] = X_offset[1] + d3; X_offset[3] = X_offset[2] + d2;
  vad_ana_filt_bank_1(pIn, c->AnaState, X, X + X_offset[3], frame_length);
  vad_ana_filt_bank_1(X, c->AnaState1, X, X + X_offset[2], d1);
  vad_ana_filt_bank_1(X, c->AnaState2, X, X + X_offset[1], d2);
  X[d3 - 1] = (int16_t)(X[d3 - 1] >> 1);
  hp_tmp = X[d3 - 1];
  for (i = d3 - 1; i > 0; i--) { X[i - 1] = (int16_t)(X[i - 1] >> 1); X[i] = (int16_t)(X[i] - X[i - 1]); }
  X[0] = (int16_t)(X[0] - (int16_t)c->HPstate);
  c->HPstate = hp_tmp;
  for (b = 0; b < 4; b++) {
    const int sh = 4 - b < 3 ? 4 - b : 3, dl = frame_length >> sh, sl = dl >> 2;
    int off = 0;
    Xnrg[b] = c->XnrgSubfr[b];
    for (s = 0; s < 4; s++) {
      sumSquared = 0;
      for (i = 0; i < sl; i++) { const int32_t x_tmp = X[X_offset[b] + i + off] >> 3; sumSquared += x_tmp * x_tmp; }
      Xnrg[b] = vad_add_pos_sat32(Xnrg[b], s < 3 ? sumSquared : sumSquared >> 1);
      off += sl;
    }
    c->XnrgSubfr[b] = sumSquared;
  }
  vad_noise_levels(Xnrg, c);
  sumSquared = 0;
  for (b = 0; b < 4; b++) {
    speech_nrg = Xnrg[b] - c->NL[b];
    if (speech_nrg > 0) {
      if ((Xnrg[b] & 0xFF800000) == 0) ratio_Q8[b] = (int32_t)((uint32_t)Xnrg[b] << 8) / (c->NL[b] + 1);
      else ratio_Q8[b] = Xnrg[b] / ((c->NL[b] >> 8) + 1);
      SNR_Q7 = orc_silk_lin2log(ratio_Q8[b]) - 8 * 128;
      sumSquared += (int16_t)SNR_Q7 * (int32_t)(int16_t)SNR_Q7;
      if (speech_nrg < (1 << 20)) SNR_Q7 = smulwb((int32_t)((uint32_t)vad_sqrt_approx(speech_nrg) << 6), SNR_Q7);
      input_tilt = smlawb(input_tilt, tiltWeights[b], SNR_Q7);
    } else ratio_Q8[b] = 256;
  }
  sumSquared = sumSquared / 4;
  pSNR_dB_Q7 = (int16_t)(3 * vad_sqrt_approx(sumSquared));
  SA_Q15 = vad_sigm_Q15(smulwb(45000, pSNR_dB_Q7) - 128);
  c->input_tilt_Q15 = (vad_sigm_Q15(input_tilt) - 16384) * 2;
  speech_nrg = 0;
  for (b = 0; b < 4; b++) speech_nrg += (b + 1) * ((Xnrg[b] - c->NL[b]) >> 4);
  if (frame_length == 20 * fs_kHz) speech_nrg >>= 1;
  if (speech_nrg <= 0) SA_Q15 >>= 1;
  else if (speech_nrg < 16384) {
    speech_nrg = (int32_t)((uint32_t)speech_nrg << 16);
    speech_nrg = vad_sqrt_approx(speech_nrg);
    SA_Q15 = smulwb(32768 + speech_nrg, SA_Q15);
  }
  c->speech_activity_Q8 = (SA_Q15 >> 7) < 255 ? (SA_Q15 >> 7) : 255;
  smooth_coef_Q16 = smulwb(4096, smulwb(SA_Q15, SA_Q15));
  if (frame_length == 10 * fs_kHz) smooth_coef_Q16 >>= 1;
  for (b = 0; b < 4; b++) {
    c->NrgRatioSmth_Q8[b] = smlawb(c->NrgRatioSmth_Q8[b], ratio_Q8[b] - c->NrgRatioSmth_Q8[b], smooth_coef_Q16);
    SNR_Q7 = 3 * (orc_silk_lin2log(c->NrgRatioSmth_Q8[b]) - 8 * 128);
    c->input_quality_bands_Q15[b] = vad_sigm_Q15((SNR_Q7 - 16 * 128) >> 4);
  }
}
/* silk_encode_do_VAD_FLP: the VAD, the Opus-level activity override and the VAD / DTX bookkeeping (VAD_flag = VAD_flags[nFramesEncoded]). */
void orc_silk_encode_do_vad(OrcSilkVad *c, const int16_t *pIn, int frame_length, int fs_kHz, int activity) {
  const int thr = 13;   /* SILK_FIX_CONST(SPEECH_ACTIVITY_DTX_THRES = 0.05, 8) */
  orc_silk_vad_get_sa(c, pIn, frame_length, fs_kHz);
  if (activity == 0 && c->speech_activity_Q8 >= thr) c->speech_activity_Q8 = thr - 1;
  if (c->speech_activity_Q8 < thr) {
    c->signalType = 0;
    c->noSpeechCounter++;
    if (c->noSpeechCounter <= 10) c->inDTX = 0;
    else if (c->noSpeechCounter > 20 + 10) { c->noSpeechCounter = 10; c->inDTX = 0; }
    c->VAD_flag = 0;
  } else {
    c->noSpeechCounter = 0; c->inDTX = 0; c->signalType = 1; c->VAD_flag = 1;
  }
}
/* silk_HP_variable_cutoff: the pitch-driven smoother of the high-pass cutoff (the second smoother is opus_encode_native's, src/opus_encoder.c:1974). */
void orc_silk_hp_variable_cutoff(OrcSilkVad *c, int fs_kHz) {
  if (c->prevSignalType == 2) {
    const int32_t pitch_freq_Hz_Q16 = (int32_t)((uint32_t)(fs_kHz * 1000) << 16) / (int16_t)c->prevLag;
    int32_t pitch_freq_log_Q7 = orc_silk_lin2log(pitch_freq_Hz_Q16) - (16 << 7), delta_freq_Q7;
    const int quality_Q15 = c->input_quality_bands_Q15[0];
    pitch_freq_log_Q7 = smlawb(pitch_freq_log_Q7, smulwb((int32_t)((uint32_t)(-quality_Q15) << 2), quality_Q15),
                               pitch_freq_log_Q7 - (orc_silk_lin2log(60 << 16) - (16 << 7)));
    delta_freq_Q7 = pitch_freq_log_Q7 - (c->variable_HP_smth1_Q15 >> 8);
    if (delta_freq_Q7 < 0) delta_freq_Q7 *= 3;
    if (delta_freq_Q7 > 51) delta_freq_Q7 = 51; else if (delta_freq_Q7 < -51) delta_freq_Q7 = -51;   /* SILK_FIX_CONST(0.4, 7) = 51 */
    c->variable_HP_smth1_Q15 = smlawb(c->variable_HP_smth1_Q15, (int16_t)c->speech_activity_Q8 * (int32_t)(int16_t)delta_freq_Q7, 6554);   /* SILK_FIX_CONST(0.1, 16) */
    { const int32_t lo = orc_silk_lin2log(60) << 8, hi = orc_silk_lin2log(100) << 8;
      if (c->variable_HP_smth1_Q15 < lo) c->variable_HP_smth1_Q15 = lo; else if (c->variable_HP_smth1_Q15 > hi) c->variable_HP_smth1_Q15 = hi; }
  }
}
int orc_silk_vad_size(void) { return (int)sizeof(OrcSilkVad); }
