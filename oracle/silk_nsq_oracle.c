/* oracle/silk_nsq_oracle.c — TEST INFRASTRUCTURE ONLY.
 * Plain-C restatement of the SILK noise-shaping quantizers:
 *   silk_NSQ_c          silk/NSQ.c:76-179 (+ quantizer 185-388, scale_states 390-437, NSQ.h:35-97)
 *   silk_NSQ_del_dec_c  silk/NSQ_del_dec.c:113-330 (+ quantizer 336-653, scale_states 655-730)
 * 32-bit wrap-around arithmetic everywhere the reference says _ovflw; pinned bit-exact against the
 * reference's functions in oracle/_ref by tests/test_silk_nsq_oracle_vs_ref.py. */
#include <stdint.h>
#include <string.h>

#define NSQ_MAXFRAME 320
#define NSQ_MAXSUB 80
#define NSQ_LPCBUF 16
#define NSQ_SHAPE 24
#define NSQ_DELAY 40
#define NSQ_VOICED 2

/* frame-invariant configuration: the fields of silk_encoder_state the quantizers read */
typedef struct {
  int32_t fs_khz, nb_subfr, predictLPCOrder, shapingLPCOrder, warping_Q16, nStates;
} OrcNsqCfg;

/* mirrors silk_nsq_state (silk/structs.h:56-69) */
typedef struct {
  int16_t xq[2 * NSQ_MAXFRAME];
  int32_t sLTP_shp_Q14[2 * NSQ_MAXFRAME];
  int32_t sLPC_Q14[NSQ_MAXSUB + NSQ_LPCBUF];
  int32_t sAR2_Q14[NSQ_SHAPE];
  int32_t sLF_AR_shp_Q14, sDiff_shp_Q14, lagPrev, sLTP_buf_idx, sLTP_shp_buf_idx, rand_seed, prev_gain_Q16, rewhite_flag;
} OrcNsqState;

/* per-frame side information + control: what silk_NSQ_wrapper_FLP (silk/float/wrappers_FLP.c:106-170) hands over.
 * Same layout as OpusB200NsqFrame in include/opus_b200.h. */
typedef struct {
  int32_t signalType, quantOffsetType, NLSFInterpCoef_Q2, Seed;
  int32_t HarmShapeGain_Q14[4], Tilt_Q14[4], LF_shp_Q14[4], Gains_Q16[4], pitchL[4];
  int32_t Lambda_Q10, LTP_scale_Q14;
  int16_t PredCoef_Q12[2 * 16];
  int16_t LTPCoef_Q14[5 * 4];
  int16_t AR_Q13[4 * NSQ_SHAPE];
} OrcNsqFrame;

static int32_t smulwb(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * (int16_t)b) >> 16); }
static int32_t smlawb(int32_t a, int32_t b, int32_t c) { return (int32_t)((uint32_t)a + (uint32_t)(int32_t)(((int64_t)b * (int16_t)c) >> 16)); }
static int32_t smlawt(int32_t a, int32_t b, int32_t c) { return (int32_t)((uint32_t)a + (uint32_t)(int32_t)(((int64_t)b * (c >> 16)) >> 16)); }
static int32_t smulww(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * b) >> 16); }
static int32_t smlaww(int32_t a, int32_t b, int32_t c) { return (int32_t)((uint32_t)a + (uint32_t)(int32_t)(((int64_t)b * c) >> 16)); }
static int32_t smulbb(int32_t a, int32_t b) { return (int32_t)(int16_t)a * (int32_t)(int16_t)b; }
static int32_t rshift_round(int32_t a, int s) { return s == 1 ? (a >> 1) + (a & 1) : ((a >> (s - 1)) + 1) >> 1; }
static int32_t sat16(int32_t a) { return a > 32767 ? 32767 : (a < -32768 ? -32768 : a); }
static int32_t addw(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t subw(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static int32_t shlw(int32_t a, int s) { return (int32_t)((uint32_t)a << s); }
static int32_t add_sat32(int32_t a, int32_t b) {
  int64_t s = (int64_t)a + b;
  return s > 2147483647LL ? 2147483647 : (s < -2147483648LL ? (int32_t)(-2147483647 - 1) : (int32_t)s);
}
static int32_t sub_sat32(int32_t a, int32_t b) {
  int64_t s = (int64_t)a - b;
  return s > 2147483647LL ? 2147483647 : (s < -2147483648LL ? (int32_t)(-2147483647 - 1) : (int32_t)s);
}
static int clz32(int32_t x) { return x ? __builtin_clz((uint32_t)x) : 32; }
static int32_t iabs(int32_t a) { return a > 0 ? a : -a; }
static int32_t lshift_sat32(int32_t a, int s) {
  int32_t lo = (int32_t)(-2147483647 - 1) >> s, hi = 2147483647 >> s;
  return shlw(a < lo ? lo : (a > hi ? hi : a), s);
}
/* silk/Inlines.h:93-136 */
static int32_t div32_varq(int32_t a32, int32_t b32, int Qres) {
  int a_headrm = clz32(iabs(a32)) - 1, b_headrm = clz32(iabs(b32)) - 1, lshift;
  int32_t a_nrm = shlw(a32, a_headrm), b_nrm = shlw(b32, b_headrm);
  int32_t b_inv = (2147483647 >> 2) / (b_nrm >> 16);
  int32_t result = smulwb(a_nrm, b_inv);
  a_nrm = subw(a_nrm, shlw((int32_t)(((int64_t)b_nrm * result) >> 32), 3));
  result = smlawb(result, a_nrm, b_inv);
  lshift = 29 + a_headrm - b_headrm - Qres;
  if (lshift < 0) return lshift_sat32(result, -lshift);
  return lshift < 32 ? result >> lshift : 0;
}
/* silk/Inlines.h:139-178 */
static int32_t inverse32_varq(int32_t b32, int Qres) {
  int b_headrm = clz32(iabs(b32)) - 1, lshift;
  int32_t b_nrm = shlw(b32, b_headrm);
  int32_t b_inv = (2147483647 >> 2) / (b_nrm >> 16);
  int32_t result = shlw(b_inv, 16);
  int32_t err_Q32 = shlw(((int32_t)1 << 29) - smulwb(b_nrm, b_inv), 3);
  result = smlaww(result, err_Q32, b_inv);
  lshift = 61 - b_headrm - Qres;
  if (lshift <= 0) return lshift_sat32(result, -lshift);
  return lshift < 32 ? result >> lshift : 0;
}
static int32_t lcg(int32_t seed) { return (int32_t)(907633515u + (uint32_t)seed * 196314165u); } /* silk_RAND */

/* silk/LPC_analysis_filter.c:49-111 (local copy so this file stands alone) */
static void lpc_analysis(int16_t *out, const int16_t *in, const int16_t *B, int len, int d) {
  int ix, j;
  for (ix = d; ix < len; ix++) {
    const int16_t *p = &in[ix - 1];
    uint32_t acc = (uint32_t)((int32_t)p[0] * B[0]);
    for (j = 1; j < d; j++) acc += (uint32_t)((int32_t)p[-j] * B[j]);
    acc = ((uint32_t)((int32_t)p[1]) << 12) - acc;
    out[ix] = (int16_t)sat16(rshift_round((int32_t)acc, 12));
  }
  memset(out, 0, d * sizeof(int16_t));
}

/* silk/NSQ.h:35-67 */
static int32_t short_prediction(const int32_t *buf, const int16_t *a, int order) {
  int32_t out = order >> 1;
  int j;
  for (j = 0; j < order; j++) out = smlawb(out, buf[-j], a[j]);
  return out;
}

/* the two-candidate scalar quantizer shared by both variants (silk/NSQ.c:277-331, NSQ_del_dec.c:456-505):
 * returns candidates q1,q2 (Q10) and their rate terms rd1,rd2 (before adding the squared error) */
static void quant_candidates(int32_t r_Q10, int32_t offset_Q10, int32_t Lambda_Q10, int32_t *q1, int32_t *q2, int32_t *rd1, int32_t *rd2) {
  int32_t q1_Q10 = r_Q10 - offset_Q10, q2_Q10, q1_Q0 = q1_Q10 >> 10;
  if (Lambda_Q10 > 2048) {
    int rdo_offset = Lambda_Q10 / 2 - 512;
    if (q1_Q10 > rdo_offset) q1_Q0 = (q1_Q10 - rdo_offset) >> 10;
    else if (q1_Q10 < -rdo_offset) q1_Q0 = (q1_Q10 + rdo_offset) >> 10;
    else if (q1_Q10 < 0) q1_Q0 = -1;
    else q1_Q0 = 0;
  }
  if (q1_Q0 > 0) {
    q1_Q10 = (q1_Q0 << 10) - 80 + offset_Q10;
    q2_Q10 = q1_Q10 + 1024;
    *rd1 = smulbb(q1_Q10, Lambda_Q10);
    *rd2 = smulbb(q2_Q10, Lambda_Q10);
  } else if (q1_Q0 == 0) {
    q1_Q10 = offset_Q10;
    q2_Q10 = q1_Q10 + (1024 - 80);
    *rd1 = smulbb(q1_Q10, Lambda_Q10);
    *rd2 = smulbb(q2_Q10, Lambda_Q10);
  } else if (q1_Q0 == -1) {
    q2_Q10 = offset_Q10;
    q1_Q10 = q2_Q10 - (1024 - 80);
    *rd1 = smulbb(-q1_Q10, Lambda_Q10);
    *rd2 = smulbb(q2_Q10, Lambda_Q10);
  } else {
    q1_Q10 = shlw(q1_Q0, 10) + 80 + offset_Q10;
    q2_Q10 = q1_Q10 + 1024;
    *rd1 = smulbb(-q1_Q10, Lambda_Q10);
    *rd2 = smulbb(-q2_Q10, Lambda_Q10);
  }
  *q1 = q1_Q10;
  *q2 = q2_Q10;
}

static const int16_t quant_offsets_Q10[2][2] = {{100, 240}, {32, 100}}; /* silk/tables_other.c:77 */

/* ------------------------------------------------------------------------------------------------ */
/* silk/NSQ.c:390-437 */
static void nsq_scale_states(const OrcNsqCfg *c, OrcNsqState *S, const int16_t *x16, int32_t *x_sc_Q10, const int16_t *sLTP,
                             int32_t *sLTP_Q15, int subfr, const OrcNsqFrame *f) {
  const int subfr_length = 5 * c->fs_khz, ltp_mem = 20 * c->fs_khz;
  int i, lag = f->pitchL[subfr];
  int32_t g = f->Gains_Q16[subfr];
  int32_t inv_gain_Q31 = inverse32_varq(g > 1 ? g : 1, 47);
  int32_t inv_gain_Q26 = rshift_round(inv_gain_Q31, 5);
  for (i = 0; i < subfr_length; i++) x_sc_Q10[i] = smulww(x16[i], inv_gain_Q26);
  if (S->rewhite_flag) {
    if (subfr == 0) inv_gain_Q31 = shlw(smulwb(inv_gain_Q31, f->LTP_scale_Q14), 2);
    for (i = S->sLTP_buf_idx - lag - 2; i < S->sLTP_buf_idx; i++) sLTP_Q15[i] = smulwb(inv_gain_Q31, sLTP[i]);
  }
  if (g != S->prev_gain_Q16) {
    int32_t adj = div32_varq(S->prev_gain_Q16, g, 16);
    for (i = S->sLTP_shp_buf_idx - ltp_mem; i < S->sLTP_shp_buf_idx; i++) S->sLTP_shp_Q14[i] = smulww(adj, S->sLTP_shp_Q14[i]);
    if (f->signalType == NSQ_VOICED && S->rewhite_flag == 0)
      for (i = S->sLTP_buf_idx - lag - 2; i < S->sLTP_buf_idx; i++) sLTP_Q15[i] = smulww(adj, sLTP_Q15[i]);
    S->sLF_AR_shp_Q14 = smulww(adj, S->sLF_AR_shp_Q14);
    S->sDiff_shp_Q14 = smulww(adj, S->sDiff_shp_Q14);
    for (i = 0; i < NSQ_LPCBUF; i++) S->sLPC_Q14[i] = smulww(adj, S->sLPC_Q14[i]);
    for (i = 0; i < NSQ_SHAPE; i++) S->sAR2_Q14[i] = smulww(adj, S->sAR2_Q14[i]);
    S->prev_gain_Q16 = g;
  }
}

/* silk/NSQ.h:71-97 */
static int32_t shape_feedback(int32_t diff, int32_t *ar2, const int16_t *coef, int order) {
  int32_t out, t1, t2;
  int j;
  t2 = diff;
  t1 = ar2[0];
  ar2[0] = t2;
  out = order >> 1;
  out = smlawb(out, t2, coef[0]);
  for (j = 2; j < order; j += 2) {
    t2 = ar2[j - 1];
    ar2[j - 1] = t1;
    out = smlawb(out, t1, coef[j - 1]);
    t1 = ar2[j];
    ar2[j] = t2;
    out = smlawb(out, t2, coef[j]);
  }
  ar2[order - 1] = t1;
  out = smlawb(out, t1, coef[order - 1]);
  return shlw(out, 1);
}

/* silk/NSQ.c:185-388 */
static void nsq_subframe(const OrcNsqCfg *c, OrcNsqState *S, int signalType, const int32_t *x_sc_Q10, int8_t *pulses, int16_t *xq,
                         int32_t *sLTP_Q15, const int16_t *a_Q12, const int16_t *b_Q14, const int16_t *AR_shp_Q13, int lag,
                         int32_t HarmPacked, int32_t Tilt_Q14, int32_t LF_shp_Q14, int32_t Gain_Q16, int32_t Lambda_Q10,
                         int32_t offset_Q10) {
  const int length = 5 * c->fs_khz;
  int i;
  int32_t *shp_lag_ptr = &S->sLTP_shp_Q14[S->sLTP_shp_buf_idx - lag + 1];
  int32_t *pred_lag_ptr = &sLTP_Q15[S->sLTP_buf_idx - lag + 2];
  int32_t Gain_Q10 = Gain_Q16 >> 6;
  int32_t *psLPC = &S->sLPC_Q14[NSQ_LPCBUF - 1];
  for (i = 0; i < length; i++) {
    int32_t LPC_pred_Q10, LTP_pred_Q13, n_AR_Q12, n_LF_Q12, n_LTP_Q13, t1, t2, r_Q10, q1, q2, rd1, rd2, rr, exc_Q14, LPC_exc_Q14, xq_Q14, sLF;
    S->rand_seed = lcg(S->rand_seed);
    LPC_pred_Q10 = short_prediction(psLPC, a_Q12, c->predictLPCOrder);
    if (signalType == NSQ_VOICED) {
      int j;
      LTP_pred_Q13 = 2;
      for (j = 0; j < 5; j++) LTP_pred_Q13 = smlawb(LTP_pred_Q13, pred_lag_ptr[-j], b_Q14[j]);
      pred_lag_ptr++;
    } else LTP_pred_Q13 = 0;
    n_AR_Q12 = shape_feedback(S->sDiff_shp_Q14, S->sAR2_Q14, AR_shp_Q13, c->shapingLPCOrder);
    n_AR_Q12 = smlawb(n_AR_Q12, S->sLF_AR_shp_Q14, Tilt_Q14);
    n_LF_Q12 = smulwb(S->sLTP_shp_Q14[S->sLTP_shp_buf_idx - 1], LF_shp_Q14);
    n_LF_Q12 = smlawt(n_LF_Q12, S->sLF_AR_shp_Q14, LF_shp_Q14);
    t1 = subw(shlw(LPC_pred_Q10, 2), n_AR_Q12);
    t1 = subw(t1, n_LF_Q12);
    if (lag > 0) {
      n_LTP_Q13 = smulwb(add_sat32(shp_lag_ptr[0], shp_lag_ptr[-2]), HarmPacked);
      n_LTP_Q13 = smlawt(n_LTP_Q13, shp_lag_ptr[-1], HarmPacked);
      n_LTP_Q13 = shlw(n_LTP_Q13, 1);
      shp_lag_ptr++;
      t2 = LTP_pred_Q13 - n_LTP_Q13;
      t1 = addw(t2, shlw(t1, 1));
      t1 = rshift_round(t1, 3);
    } else t1 = rshift_round(t1, 2);
    r_Q10 = x_sc_Q10[i] - t1;
    if (S->rand_seed < 0) r_Q10 = -r_Q10;
    r_Q10 = r_Q10 > (30 << 10) ? (30 << 10) : (r_Q10 < -(31 << 10) ? -(31 << 10) : r_Q10);
    quant_candidates(r_Q10, offset_Q10, Lambda_Q10, &q1, &q2, &rd1, &rd2);
    rr = r_Q10 - q1;
    rd1 = rd1 + smulbb(rr, rr);
    rr = r_Q10 - q2;
    rd2 = rd2 + smulbb(rr, rr);
    if (rd2 < rd1) q1 = q2;
    pulses[i] = (int8_t)rshift_round(q1, 10);
    exc_Q14 = shlw(q1, 4);
    if (S->rand_seed < 0) exc_Q14 = -exc_Q14;
    LPC_exc_Q14 = exc_Q14 + shlw(LTP_pred_Q13, 1);
    xq_Q14 = addw(LPC_exc_Q14, shlw(LPC_pred_Q10, 4));
    xq[i] = (int16_t)sat16(rshift_round(smulww(xq_Q14, Gain_Q10), 8));
    psLPC++;
    *psLPC = xq_Q14;
    S->sDiff_shp_Q14 = subw(xq_Q14, shlw(x_sc_Q10[i], 4));
    sLF = subw(S->sDiff_shp_Q14, shlw(n_AR_Q12, 2));
    S->sLF_AR_shp_Q14 = sLF;
    S->sLTP_shp_Q14[S->sLTP_shp_buf_idx] = subw(sLF, shlw(n_LF_Q12, 2));
    sLTP_Q15[S->sLTP_buf_idx] = shlw(LPC_exc_Q14, 1);
    S->sLTP_shp_buf_idx++;
    S->sLTP_buf_idx++;
    S->rand_seed = addw(S->rand_seed, pulses[i]);
  }
  memcpy(S->sLPC_Q14, &S->sLPC_Q14[length], NSQ_LPCBUF * sizeof(int32_t));
}

/* silk/NSQ.c:76-179 */
static void nsq_plain(const OrcNsqCfg *c, OrcNsqState *S, OrcNsqFrame *f, const int16_t *x16, int8_t *pulses) {
  const int subfr_length = 5 * c->fs_khz, ltp_mem = 20 * c->fs_khz, frame_length = subfr_length * c->nb_subfr;
  int k, lag = S->lagPrev;
  int interp = f->NLSFInterpCoef_Q2 == 4 ? 0 : 1;
  int32_t sLTP_Q15[2 * NSQ_MAXFRAME], x_sc_Q10[NSQ_MAXSUB];
  int16_t sLTP[2 * NSQ_MAXFRAME];
  int32_t offset_Q10 = quant_offsets_Q10[f->signalType >> 1][f->quantOffsetType];
  int16_t *pxq = &S->xq[ltp_mem];
  S->rand_seed = f->Seed;
  S->sLTP_shp_buf_idx = ltp_mem;
  S->sLTP_buf_idx = ltp_mem;
  for (k = 0; k < c->nb_subfr; k++) {
    const int16_t *A_Q12 = &f->PredCoef_Q12[((k >> 1) | (1 - interp)) * 16];
    int32_t h = f->HarmShapeGain_Q14[k];
    int32_t HarmPacked = (h >> 2) | shlw(h >> 1, 16);
    S->rewhite_flag = 0;
    if (f->signalType == NSQ_VOICED) {
      lag = f->pitchL[k];
      if ((k & (3 - (interp << 1))) == 0) {
        int start_idx = ltp_mem - lag - c->predictLPCOrder - 2;
        lpc_analysis(&sLTP[start_idx], &S->xq[start_idx + k * subfr_length], A_Q12, ltp_mem - start_idx, c->predictLPCOrder);
        S->rewhite_flag = 1;
        S->sLTP_buf_idx = ltp_mem;
      }
    }
    nsq_scale_states(c, S, x16, x_sc_Q10, sLTP, sLTP_Q15, k, f);
    nsq_subframe(c, S, f->signalType, x_sc_Q10, pulses, pxq, sLTP_Q15, A_Q12, &f->LTPCoef_Q14[k * 5], &f->AR_Q13[k * NSQ_SHAPE], lag,
                 HarmPacked, f->Tilt_Q14[k], f->LF_shp_Q14[k], f->Gains_Q16[k], f->Lambda_Q10, offset_Q10);
    x16 += subfr_length;
    pulses += subfr_length;
    pxq += subfr_length;
  }
  S->lagPrev = f->pitchL[c->nb_subfr - 1];
  memmove(S->xq, &S->xq[frame_length], ltp_mem * sizeof(int16_t));
  memmove(S->sLTP_shp_Q14, &S->sLTP_shp_Q14[frame_length], ltp_mem * sizeof(int32_t));
}

/* ------------------------------------------------------------------------------------------------ */
/* delayed-decision variant */
typedef struct {
  int32_t sLPC_Q14[NSQ_MAXSUB + NSQ_LPCBUF];
  int32_t RandState[NSQ_DELAY], Q_Q10[NSQ_DELAY], Xq_Q14[NSQ_DELAY], Pred_Q15[NSQ_DELAY], Shape_Q14[NSQ_DELAY];
  int32_t sAR2_Q14[NSQ_SHAPE];
  int32_t LF_AR_Q14, Diff_Q14, Seed, SeedInit, RD_Q10;
} DelDec; /* silk/NSQ_del_dec.c:36-50; member order matters for the partial state copy */

typedef struct { int32_t Q_Q10, RD_Q10, xq_Q14, LF_AR_Q14, Diff_Q14, sLTP_shp_Q14, LPC_exc_Q14; } Cand;

/* silk/NSQ_del_dec.c:655-730 */
static void dd_scale_states(const OrcNsqCfg *c, OrcNsqState *S, DelDec *dd, const int16_t *x16, int32_t *x_sc_Q10, const int16_t *sLTP,
                            int32_t *sLTP_Q15, int subfr, const OrcNsqFrame *f, int decisionDelay) {
  const int subfr_length = 5 * c->fs_khz, ltp_mem = 20 * c->fs_khz;
  int i, k, lag = f->pitchL[subfr];
  int32_t g = f->Gains_Q16[subfr];
  int32_t inv_gain_Q31 = inverse32_varq(g > 1 ? g : 1, 47);
  int32_t inv_gain_Q26 = rshift_round(inv_gain_Q31, 5);
  for (i = 0; i < subfr_length; i++) x_sc_Q10[i] = smulww(x16[i], inv_gain_Q26);
  if (S->rewhite_flag) {
    if (subfr == 0) inv_gain_Q31 = shlw(smulwb(inv_gain_Q31, f->LTP_scale_Q14), 2);
    for (i = S->sLTP_buf_idx - lag - 2; i < S->sLTP_buf_idx; i++) sLTP_Q15[i] = smulwb(inv_gain_Q31, sLTP[i]);
  }
  if (g != S->prev_gain_Q16) {
    int32_t adj = div32_varq(S->prev_gain_Q16, g, 16);
    for (i = S->sLTP_shp_buf_idx - ltp_mem; i < S->sLTP_shp_buf_idx; i++) S->sLTP_shp_Q14[i] = smulww(adj, S->sLTP_shp_Q14[i]);
    if (f->signalType == NSQ_VOICED && S->rewhite_flag == 0)
      for (i = S->sLTP_buf_idx - lag - 2; i < S->sLTP_buf_idx - decisionDelay; i++) sLTP_Q15[i] = smulww(adj, sLTP_Q15[i]);
    for (k = 0; k < c->nStates; k++) {
      DelDec *d = &dd[k];
      d->LF_AR_Q14 = smulww(adj, d->LF_AR_Q14);
      d->Diff_Q14 = smulww(adj, d->Diff_Q14);
      for (i = 0; i < NSQ_LPCBUF; i++) d->sLPC_Q14[i] = smulww(adj, d->sLPC_Q14[i]);
      for (i = 0; i < NSQ_SHAPE; i++) d->sAR2_Q14[i] = smulww(adj, d->sAR2_Q14[i]);
      for (i = 0; i < NSQ_DELAY; i++) {
        d->Pred_Q15[i] = smulww(adj, d->Pred_Q15[i]);
        d->Shape_Q14[i] = smulww(adj, d->Shape_Q14[i]);
      }
    }
    S->prev_gain_Q16 = g;
  }
}

/* silk/NSQ_del_dec.c:336-653 */
static void dd_subframe(const OrcNsqCfg *c, OrcNsqState *S, DelDec *dd, int signalType, const int32_t *x_Q10, int8_t *pulses, int16_t *xq,
                        int32_t *sLTP_Q15, int32_t *delayedGain_Q10, const int16_t *a_Q12, const int16_t *b_Q14,
                        const int16_t *AR_shp_Q13, int lag, int32_t HarmPacked, int32_t Tilt_Q14, int32_t LF_shp_Q14, int32_t Gain_Q16,
                        int32_t Lambda_Q10, int32_t offset_Q10, int subfr, int *smpl_buf_idx, int decisionDelay) {
  const int length = 5 * c->fs_khz, nStates = c->nStates, order = c->shapingLPCOrder, warp = c->warping_Q16;
  Cand cand[4][2];
  int i, j, k;
  int32_t *shp_lag_ptr = &S->sLTP_shp_Q14[S->sLTP_shp_buf_idx - lag + 1];
  int32_t *pred_lag_ptr = &sLTP_Q15[S->sLTP_buf_idx - lag + 2];
  int32_t Gain_Q10 = Gain_Q16 >> 6;
  for (i = 0; i < length; i++) {
    int32_t LTP_pred_Q14, n_LTP_Q14;
    int winner, rdmin_ind, rdmax_ind, last_idx;
    int32_t rdmin, rdmax, winner_rand;
    if (signalType == NSQ_VOICED) {
      LTP_pred_Q14 = 2;
      for (j = 0; j < 5; j++) LTP_pred_Q14 = smlawb(LTP_pred_Q14, pred_lag_ptr[-j], b_Q14[j]);
      LTP_pred_Q14 = shlw(LTP_pred_Q14, 1);
      pred_lag_ptr++;
    } else LTP_pred_Q14 = 0;
    if (lag > 0) {
      n_LTP_Q14 = smulwb(add_sat32(shp_lag_ptr[0], shp_lag_ptr[-2]), HarmPacked);
      n_LTP_Q14 = smlawt(n_LTP_Q14, shp_lag_ptr[-1], HarmPacked);
      n_LTP_Q14 = LTP_pred_Q14 - shlw(n_LTP_Q14, 2);
      shp_lag_ptr++;
    } else n_LTP_Q14 = 0;
    for (k = 0; k < nStates; k++) {
      DelDec *d = &dd[k];
      Cand *cs = cand[k];
      int32_t LPC_pred_Q14, n_AR_Q14, n_LF_Q14, t1, t2, r_Q10, q1, q2, rd1, rd2, rr, s;
      d->Seed = lcg(d->Seed);
      LPC_pred_Q14 = shlw(short_prediction(&d->sLPC_Q14[NSQ_LPCBUF - 1 + i], a_Q12, c->predictLPCOrder), 4);
      /* warped noise-shaping filter */
      t2 = smlawb(d->Diff_Q14, d->sAR2_Q14[0], warp);
      t1 = smlawb(d->sAR2_Q14[0], subw(d->sAR2_Q14[1], t2), warp);
      d->sAR2_Q14[0] = t2;
      n_AR_Q14 = order >> 1;
      n_AR_Q14 = smlawb(n_AR_Q14, t2, AR_shp_Q13[0]);
      for (j = 2; j < order; j += 2) {
        t2 = smlawb(d->sAR2_Q14[j - 1], subw(d->sAR2_Q14[j], t1), warp);
        d->sAR2_Q14[j - 1] = t1;
        n_AR_Q14 = smlawb(n_AR_Q14, t1, AR_shp_Q13[j - 1]);
        t1 = smlawb(d->sAR2_Q14[j], subw(d->sAR2_Q14[j + 1], t2), warp);
        d->sAR2_Q14[j] = t2;
        n_AR_Q14 = smlawb(n_AR_Q14, t2, AR_shp_Q13[j]);
      }
      d->sAR2_Q14[order - 1] = t1;
      n_AR_Q14 = smlawb(n_AR_Q14, t1, AR_shp_Q13[order - 1]);
      n_AR_Q14 = shlw(n_AR_Q14, 1);
      n_AR_Q14 = smlawb(n_AR_Q14, d->LF_AR_Q14, Tilt_Q14);
      n_AR_Q14 = shlw(n_AR_Q14, 2);
      n_LF_Q14 = smulwb(d->Shape_Q14[*smpl_buf_idx], LF_shp_Q14);
      n_LF_Q14 = smlawt(n_LF_Q14, d->LF_AR_Q14, LF_shp_Q14);
      n_LF_Q14 = shlw(n_LF_Q14, 2);
      t1 = add_sat32(n_AR_Q14, n_LF_Q14);
      t2 = addw(n_LTP_Q14, LPC_pred_Q14);
      t1 = sub_sat32(t2, t1);
      t1 = rshift_round(t1, 4);
      r_Q10 = x_Q10[i] - t1;
      if (d->Seed < 0) r_Q10 = -r_Q10;
      r_Q10 = r_Q10 > (30 << 10) ? (30 << 10) : (r_Q10 < -(31 << 10) ? -(31 << 10) : r_Q10);
      quant_candidates(r_Q10, offset_Q10, Lambda_Q10, &q1, &q2, &rd1, &rd2);
      rr = r_Q10 - q1;
      rd1 = (rd1 + smulbb(rr, rr)) >> 10;
      rr = r_Q10 - q2;
      rd2 = (rd2 + smulbb(rr, rr)) >> 10;
      if (rd1 < rd2) {
        cs[0].RD_Q10 = d->RD_Q10 + rd1; cs[1].RD_Q10 = d->RD_Q10 + rd2; cs[0].Q_Q10 = q1; cs[1].Q_Q10 = q2;
      } else {
        cs[0].RD_Q10 = d->RD_Q10 + rd2; cs[1].RD_Q10 = d->RD_Q10 + rd1; cs[0].Q_Q10 = q2; cs[1].Q_Q10 = q1;
      }
      for (s = 0; s < 2; s++) {
        int32_t exc_Q14 = shlw(cs[s].Q_Q10, 4), LPC_exc_Q14, xq_Q14, sLF;
        if (d->Seed < 0) exc_Q14 = -exc_Q14;
        LPC_exc_Q14 = exc_Q14 + LTP_pred_Q14;
        xq_Q14 = addw(LPC_exc_Q14, LPC_pred_Q14);
        cs[s].Diff_Q14 = subw(xq_Q14, shlw(x_Q10[i], 4));
        sLF = subw(cs[s].Diff_Q14, n_AR_Q14);
        cs[s].sLTP_shp_Q14 = sub_sat32(sLF, n_LF_Q14);
        cs[s].LF_AR_Q14 = sLF;
        cs[s].LPC_exc_Q14 = LPC_exc_Q14;
        cs[s].xq_Q14 = xq_Q14;
      }
    }
    *smpl_buf_idx = (*smpl_buf_idx - 1) % NSQ_DELAY;
    if (*smpl_buf_idx < 0) *smpl_buf_idx += NSQ_DELAY;
    last_idx = (*smpl_buf_idx + decisionDelay) % NSQ_DELAY;
    rdmin = cand[0][0].RD_Q10;
    winner = 0;
    for (k = 1; k < nStates; k++)
      if (cand[k][0].RD_Q10 < rdmin) { rdmin = cand[k][0].RD_Q10; winner = k; }
    winner_rand = dd[winner].RandState[last_idx];
    for (k = 0; k < nStates; k++)
      if (dd[k].RandState[last_idx] != winner_rand) {
        cand[k][0].RD_Q10 += 2147483647 >> 4;
        cand[k][1].RD_Q10 += 2147483647 >> 4;
      }
    rdmax = cand[0][0].RD_Q10;
    rdmin = cand[0][1].RD_Q10;
    rdmax_ind = rdmin_ind = 0;
    for (k = 1; k < nStates; k++) {
      if (cand[k][0].RD_Q10 > rdmax) { rdmax = cand[k][0].RD_Q10; rdmax_ind = k; }
      if (cand[k][1].RD_Q10 < rdmin) { rdmin = cand[k][1].RD_Q10; rdmin_ind = k; }
    }
    if (rdmin < rdmax) {
      memcpy(((int32_t *)&dd[rdmax_ind]) + i, ((int32_t *)&dd[rdmin_ind]) + i, sizeof(DelDec) - i * sizeof(int32_t));
      cand[rdmax_ind][0] = cand[rdmin_ind][1];
    }
    {
      DelDec *d = &dd[winner];
      if (subfr > 0 || i >= decisionDelay) {
        pulses[i - decisionDelay] = (int8_t)rshift_round(d->Q_Q10[last_idx], 10);
        xq[i - decisionDelay] = (int16_t)sat16(rshift_round(smulww(d->Xq_Q14[last_idx], delayedGain_Q10[last_idx]), 8));
        S->sLTP_shp_Q14[S->sLTP_shp_buf_idx - decisionDelay] = d->Shape_Q14[last_idx];
        sLTP_Q15[S->sLTP_buf_idx - decisionDelay] = d->Pred_Q15[last_idx];
      }
    }
    S->sLTP_shp_buf_idx++;
    S->sLTP_buf_idx++;
    for (k = 0; k < nStates; k++) {
      DelDec *d = &dd[k];
      const Cand *s = &cand[k][0];
      d->LF_AR_Q14 = s->LF_AR_Q14;
      d->Diff_Q14 = s->Diff_Q14;
      d->sLPC_Q14[NSQ_LPCBUF + i] = s->xq_Q14;
      d->Xq_Q14[*smpl_buf_idx] = s->xq_Q14;
      d->Q_Q10[*smpl_buf_idx] = s->Q_Q10;
      d->Pred_Q15[*smpl_buf_idx] = shlw(s->LPC_exc_Q14, 1);
      d->Shape_Q14[*smpl_buf_idx] = s->sLTP_shp_Q14;
      d->Seed = addw(d->Seed, rshift_round(s->Q_Q10, 10));
      d->RandState[*smpl_buf_idx] = d->Seed;
      d->RD_Q10 = s->RD_Q10;
    }
    delayedGain_Q10[*smpl_buf_idx] = Gain_Q10;
  }
  for (k = 0; k < nStates; k++) memcpy(dd[k].sLPC_Q14, &dd[k].sLPC_Q14[length], NSQ_LPCBUF * sizeof(int32_t));
}

static int dd_winner(const DelDec *dd, int n) {
  int k, w = 0;
  int32_t m = dd[0].RD_Q10;
  for (k = 1; k < n; k++)
    if (dd[k].RD_Q10 < m) { m = dd[k].RD_Q10; w = k; }
  return w;
}

/* silk/NSQ_del_dec.c:113-330 */
static void nsq_del_dec(const OrcNsqCfg *c, OrcNsqState *S, OrcNsqFrame *f, const int16_t *x16, int8_t *pulses) {
  const int subfr_length = 5 * c->fs_khz, ltp_mem = 20 * c->fs_khz, frame_length = subfr_length * c->nb_subfr;
  int i, k, lag = S->lagPrev, subfr = 0, smpl_buf_idx = 0, last_idx, w;
  int interp = f->NLSFInterpCoef_Q2 == 4 ? 0 : 1;
  int decisionDelay = NSQ_DELAY < subfr_length ? NSQ_DELAY : subfr_length;
  int32_t sLTP_Q15[2 * NSQ_MAXFRAME], x_sc_Q10[NSQ_MAXSUB], delayedGain_Q10[NSQ_DELAY];
  int16_t sLTP[2 * NSQ_MAXFRAME];
  DelDec dd[4];
  int32_t offset_Q10 = quant_offsets_Q10[f->signalType >> 1][f->quantOffsetType], Gain_Q10;
  int16_t *pxq = &S->xq[ltp_mem];
  memset(dd, 0, sizeof(dd));
  for (k = 0; k < c->nStates; k++) {
    DelDec *d = &dd[k];
    d->Seed = (k + f->Seed) & 3;
    d->SeedInit = d->Seed;
    d->LF_AR_Q14 = S->sLF_AR_shp_Q14;
    d->Diff_Q14 = S->sDiff_shp_Q14;
    d->Shape_Q14[0] = S->sLTP_shp_Q14[ltp_mem - 1];
    memcpy(d->sLPC_Q14, S->sLPC_Q14, NSQ_LPCBUF * sizeof(int32_t));
    memcpy(d->sAR2_Q14, S->sAR2_Q14, sizeof(S->sAR2_Q14));
  }
  if (f->signalType == NSQ_VOICED) {
    for (k = 0; k < c->nb_subfr; k++)
      if (f->pitchL[k] - 2 - 1 < decisionDelay) decisionDelay = f->pitchL[k] - 2 - 1;
  } else if (lag > 0) {
    if (lag - 2 - 1 < decisionDelay) decisionDelay = lag - 2 - 1;
  }
  S->sLTP_shp_buf_idx = ltp_mem;
  S->sLTP_buf_idx = ltp_mem;
  for (k = 0; k < c->nb_subfr; k++) {
    const int16_t *A_Q12 = &f->PredCoef_Q12[((k >> 1) | (1 - interp)) * 16];
    int32_t h = f->HarmShapeGain_Q14[k];
    int32_t HarmPacked = (h >> 2) | shlw(h >> 1, 16);
    S->rewhite_flag = 0;
    if (f->signalType == NSQ_VOICED) {
      lag = f->pitchL[k];
      if ((k & (3 - (interp << 1))) == 0) {
        int start_idx;
        if (k == 2) {
          w = dd_winner(dd, c->nStates);
          for (i = 0; i < c->nStates; i++)
            if (i != w) dd[i].RD_Q10 += 2147483647 >> 4;
          last_idx = smpl_buf_idx + decisionDelay;
          for (i = 0; i < decisionDelay; i++) {
            last_idx = (last_idx - 1) % NSQ_DELAY;
            if (last_idx < 0) last_idx += NSQ_DELAY;
            pulses[i - decisionDelay] = (int8_t)rshift_round(dd[w].Q_Q10[last_idx], 10);
            pxq[i - decisionDelay] = (int16_t)sat16(rshift_round(smulww(dd[w].Xq_Q14[last_idx], f->Gains_Q16[1]), 14));
            S->sLTP_shp_Q14[S->sLTP_shp_buf_idx - decisionDelay + i] = dd[w].Shape_Q14[last_idx];
          }
          subfr = 0;
        }
        start_idx = ltp_mem - lag - c->predictLPCOrder - 2;
        lpc_analysis(&sLTP[start_idx], &S->xq[start_idx + k * subfr_length], A_Q12, ltp_mem - start_idx, c->predictLPCOrder);
        S->sLTP_buf_idx = ltp_mem;
        S->rewhite_flag = 1;
      }
    }
    dd_scale_states(c, S, dd, x16, x_sc_Q10, sLTP, sLTP_Q15, k, f, decisionDelay);
    dd_subframe(c, S, dd, f->signalType, x_sc_Q10, pulses, pxq, sLTP_Q15, delayedGain_Q10, A_Q12, &f->LTPCoef_Q14[k * 5],
                &f->AR_Q13[k * NSQ_SHAPE], lag, HarmPacked, f->Tilt_Q14[k], f->LF_shp_Q14[k], f->Gains_Q16[k], f->Lambda_Q10, offset_Q10,
                subfr++, &smpl_buf_idx, decisionDelay);
    x16 += subfr_length;
    pulses += subfr_length;
    pxq += subfr_length;
  }
  w = dd_winner(dd, c->nStates);
  f->Seed = dd[w].SeedInit;
  last_idx = smpl_buf_idx + decisionDelay;
  Gain_Q10 = f->Gains_Q16[c->nb_subfr - 1] >> 6;
  for (i = 0; i < decisionDelay; i++) {
    last_idx = (last_idx - 1) % NSQ_DELAY;
    if (last_idx < 0) last_idx += NSQ_DELAY;
    pulses[i - decisionDelay] = (int8_t)rshift_round(dd[w].Q_Q10[last_idx], 10);
    pxq[i - decisionDelay] = (int16_t)sat16(rshift_round(smulww(dd[w].Xq_Q14[last_idx], Gain_Q10), 8));
    S->sLTP_shp_Q14[S->sLTP_shp_buf_idx - decisionDelay + i] = dd[w].Shape_Q14[last_idx];
  }
  memcpy(S->sLPC_Q14, &dd[w].sLPC_Q14[subfr_length], NSQ_LPCBUF * sizeof(int32_t));
  memcpy(S->sAR2_Q14, dd[w].sAR2_Q14, sizeof(S->sAR2_Q14));
  S->sLF_AR_shp_Q14 = dd[w].LF_AR_Q14;
  S->sDiff_shp_Q14 = dd[w].Diff_Q14;
  S->lagPrev = f->pitchL[c->nb_subfr - 1];
  memmove(S->xq, &S->xq[frame_length], ltp_mem * sizeof(int16_t));
  memmove(S->sLTP_shp_Q14, &S->sLTP_shp_Q14[frame_length], ltp_mem * sizeof(int32_t));
}

/* ---- public oracle entry points ---- */
int orc_nsq_state_size(void) { return (int)sizeof(OrcNsqState); }
int orc_nsq_frame_size(void) { return (int)sizeof(OrcNsqFrame); }
/* silk/init_encoder.c + control_codec.c: zeroed state, prev_gain_Q16 = 1.0, lagPrev = 100 */
void orc_nsq_state_init(OrcNsqState *S) {
  memset(S, 0, sizeof(*S));
  S->prev_gain_Q16 = 65536;
  S->lagPrev = 100;
}
/* the selection rule of silk_NSQ_wrapper_FLP (silk/float/wrappers_FLP.c:163) */
void orc_silk_nsq(const OrcNsqCfg *cfg, OrcNsqState *S, OrcNsqFrame *f, const int16_t *x16, int8_t *pulses) {
  if (cfg->nStates > 1 || cfg->warping_Q16 > 0) nsq_del_dec(cfg, S, f, x16, pulses);
  else nsq_plain(cfg, S, f, x16, pulses);
}
/* flat export of the carried state for comparisons:
 * xq[640] sLTP_shp[640] sLPC[16] sAR2[24] sLF_AR sDiff lagPrev rand_seed prev_gain  (1325 words) */
void orc_nsq_state_export(const OrcNsqState *S, int32_t *o) {
  int i;
  for (i = 0; i < 640; i++) *o++ = S->xq[i];
  for (i = 0; i < 640; i++) *o++ = S->sLTP_shp_Q14[i];
  for (i = 0; i < 16; i++) *o++ = S->sLPC_Q14[i];
  for (i = 0; i < 24; i++) *o++ = S->sAR2_Q14[i];
  *o++ = S->sLF_AR_shp_Q14; *o++ = S->sDiff_shp_Q14; *o++ = S->lagPrev; *o++ = S->rand_seed; *o++ = S->prev_gain_Q16;
}
