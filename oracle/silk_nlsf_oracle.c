/* oracle/silk_nlsf_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Plain-C restatement of the NLSF processing of a SILK frame, integer arithmetic throughout:
 *   silk_process_NLSFs                         silk/process_NLSFs.c:35-107
 *   silk_NLSF_encode                           silk/NLSF_encode.c:38-124
 *   silk_NLSF_VQ                               silk/NLSF_VQ.c:35-76
 *   silk_NLSF_del_dec_quant                    silk/NLSF_del_dec_quant.c:35-215
 *   silk_NLSF_VQ_weights_laroia                silk/NLSF_VQ_weights_laroia.c:41-80
 *   silk_NLSF_decode, _residual_dequant        silk/NLSF_decode.c:35-93
 *   silk_NLSF_unpack                           silk/NLSF_unpack.c:35-54
 *   silk_NLSF_stabilize                        silk/NLSF_stabilize.c:46-142
 *   silk_NLSF2A, silk_NLSF2A_find_poly         silk/NLSF2A.c:45-140
 *   silk_LPC_fit                               silk/LPC_fit.c:36-82
 *   silk_LPC_inverse_pred_gain_c               silk/LPC_inv_pred_gain.c:43-141
 *   silk_bwexpander_32                         silk/bwexpander_32.c:36-50
 *   silk_interpolate                           silk/interpolate.c:35-51
 *   silk_insertion_sort_increasing[_all_values_int16]   silk/sort.c:40-82, 136-154
 *   silk_DIV32_varQ, silk_INVERSE32_varQ       silk/Inlines.h:93-178 (the 64-bit forms of silk/macros.h: OPUS_FAST_INT64)
 * with the codebooks of silk/tables_NLSF_CB_NB_MB.c / tables_NLSF_CB_WB.c and silk/table_LSF_cos.c (data: opus_b200/csrc/silk_tables_data.h,
 * generated from those files).  Pinned against the reference's silk_process_NLSFs by tests/test_silk_nlsf_oracle_vs_ref.py. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OB_TABLE static const
#include "../opus_b200/csrc/silk_tables_data.h"

int32_t orc_silk_lin2log(int32_t inLin); /* silk_oracle.c */

#define MAX_ORDER 16
static int clz32(int32_t x) { return x ? __builtin_clz((uint32_t)x) : 32; }
static int32_t smulbb(int32_t a, int32_t b) { return (int32_t)(int16_t)a * (int32_t)(int16_t)b; }
static int32_t smulwb(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * (int16_t)b) >> 16); }
static int32_t smlawb(int32_t a, int32_t b, int32_t c) { return a + smulwb(b, c); }
static int32_t smmul(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * b) >> 32); }
static int32_t smulww(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * b) >> 16); }
static int32_t rshift_round(int32_t a, int s) { return s == 1 ? (a >> 1) + (a & 1) : ((a >> (s - 1)) + 1) >> 1; }
static int64_t rshift_round64(int64_t a, int s) { return s == 1 ? (a >> 1) + (a & 1) : ((a >> (s - 1)) + 1) >> 1; }
static int32_t lshift_sat32(int32_t a, int s) {
  const int32_t lo = INT32_MIN >> s, hi = INT32_MAX >> s;
  a = a < lo ? lo : a > hi ? hi : a;
  return (int32_t)((uint32_t)a << s);
}
static int32_t sub_sat32(int32_t a, int32_t b) {
  const int64_t d = (int64_t)a - b;
  return d > INT32_MAX ? INT32_MAX : d < INT32_MIN ? INT32_MIN : (int32_t)d;
}
static int32_t div32_varQ(int32_t a32, int32_t b32, int Qres) {
  const int a_headrm = clz32(abs(a32)) - 1, b_headrm = clz32(abs(b32)) - 1;
  int32_t a_nrm = (int32_t)((uint32_t)a32 << a_headrm), b_nrm = (int32_t)((uint32_t)b32 << b_headrm);
  const int32_t b_inv = (INT32_MAX >> 2) / (b_nrm >> 16);
  int32_t result = smulwb(a_nrm, b_inv);
  int lshift;
  a_nrm = (int32_t)((uint32_t)a_nrm - ((uint32_t)smmul(b_nrm, result) << 3));
  result = smlawb(result, a_nrm, b_inv);
  lshift = 29 + a_headrm - b_headrm - Qres;
  if (lshift < 0) return lshift_sat32(result, -lshift);
  return lshift < 32 ? result >> lshift : 0;
}
static int32_t inverse32_varQ(int32_t b32, int Qres) {
  const int b_headrm = clz32(abs(b32)) - 1;
  const int32_t b_nrm = (int32_t)((uint32_t)b32 << b_headrm), b_inv = (INT32_MAX >> 2) / (b_nrm >> 16);
  int32_t result = (int32_t)((uint32_t)b_inv << 16);
  const int32_t err_Q32 = (int32_t)((uint32_t)(((int32_t)1 << 29) - smulwb(b_nrm, b_inv)) << 3);
  int lshift;
  result = (int32_t)(result + (((int64_t)err_Q32 * b_inv) >> 16));
  lshift = 61 - b_headrm - Qres;
  if (lshift <= 0) return lshift_sat32(result, -lshift);
  return lshift < 32 ? result >> lshift : 0;
}

typedef struct { /* silk_NLSF_CB_struct, silk/structs.h:84-97 */
  int nVectors, order, quantStepSize_Q16, invQuantStepSize_Q6;
  const uint8_t *CB1_NLSF_Q8;
  const int16_t *CB1_Wght_Q9;
  const uint8_t *CB1_iCDF, *pred_Q8, *ec_sel, *ec_Rates_Q5;
  const int16_t *deltaMin_Q15;
} NlsfCB;
static void get_cb(NlsfCB *cb, int wb) { /* silk_NLSF_CB_NB_MB / silk_NLSF_CB_WB */
  cb->nVectors = 32;
  cb->order = wb ? 16 : 10;
  cb->quantStepSize_Q16 = wb ? 9830 : 11796;    /* SILK_FIX_CONST(0.15 / 0.18, 16) */
  cb->invQuantStepSize_Q6 = wb ? 427 : 356;     /* SILK_FIX_CONST(1 / 0.15, 6), (1 / 0.18, 6) */
  cb->CB1_NLSF_Q8 = (const uint8_t *)ob_silk_NLSF_CB1_Q8 + (wb ? 320 : 0);
  cb->CB1_Wght_Q9 = (const int16_t *)ob_silk_NLSF_CB1_Wght_Q9 + (wb ? 320 : 0);
  cb->CB1_iCDF = (const uint8_t *)ob_silk_NLSF_CB1_iCDF + (wb ? 64 : 0);
  cb->pred_Q8 = (const uint8_t *)ob_silk_NLSF_PRED_Q8 + (wb ? 18 : 0);
  cb->ec_sel = (const uint8_t *)ob_silk_NLSF_CB2_SELECT + (wb ? 160 : 0);
  cb->ec_Rates_Q5 = (const uint8_t *)ob_silk_NLSF_CB2_BITS_Q5 + (wb ? 72 : 0);
  cb->deltaMin_Q15 = (const int16_t *)ob_silk_NLSF_DELTA_MIN_Q15 + (wb ? 11 : 0);
}

static void sort_increasing(int32_t *a, int *idx, int L, int K) {
  int i, j;
  for (i = 0; i < K; i++) idx[i] = i;
  for (i = 1; i < K; i++) {
    const int32_t value = a[i];
    for (j = i - 1; j >= 0 && value < a[j]; j--) { a[j + 1] = a[j]; idx[j + 1] = idx[j]; }
    a[j + 1] = value;
    idx[j + 1] = i;
  }
  for (i = K; i < L; i++) {
    const int32_t value = a[i];
    if (value < a[K - 1]) {
      for (j = K - 2; j >= 0 && value < a[j]; j--) { a[j + 1] = a[j]; idx[j + 1] = idx[j]; }
      a[j + 1] = value;
      idx[j + 1] = i;
    }
  }
}
void orc_silk_NLSF_stabilize(int16_t *NLSF_Q15, const int16_t *NDeltaMin_Q15, int L) {
  int i, I = 0, k, loops;
  for (loops = 0; loops < 20; loops++) {
    int32_t min_diff = NLSF_Q15[0] - NDeltaMin_Q15[0], diff;
    I = 0;
    for (i = 1; i <= L - 1; i++) {
      diff = NLSF_Q15[i] - (NLSF_Q15[i - 1] + NDeltaMin_Q15[i]);
      if (diff < min_diff) { min_diff = diff; I = i; }
    }
    diff = (1 << 15) - (NLSF_Q15[L - 1] + NDeltaMin_Q15[L]);
    if (diff < min_diff) { min_diff = diff; I = L; }
    if (min_diff >= 0) return;
    if (I == 0) NLSF_Q15[0] = NDeltaMin_Q15[0];
    else if (I == L) NLSF_Q15[L - 1] = (int16_t)((1 << 15) - NDeltaMin_Q15[L]);
    else {
      int32_t min_center = 0, max_center = 1 << 15, c;
      for (k = 0; k < I; k++) min_center += NDeltaMin_Q15[k];
      min_center += NDeltaMin_Q15[I] >> 1;
      for (k = L; k > I; k--) max_center -= NDeltaMin_Q15[k];
      max_center -= NDeltaMin_Q15[I] >> 1;
      c = rshift_round((int32_t)NLSF_Q15[I - 1] + (int32_t)NLSF_Q15[I], 1);
      /* silk_LIMIT_32 orders its bounds itself (silk/SigProc_FIX.h:553-554) */
      if (min_center > max_center) c = c > min_center ? min_center : c < max_center ? max_center : c;
      else c = c > max_center ? max_center : c < min_center ? min_center : c;
      NLSF_Q15[I - 1] = (int16_t)((int16_t)c - (NDeltaMin_Q15[I] >> 1));
      NLSF_Q15[I] = (int16_t)(NLSF_Q15[I - 1] + NDeltaMin_Q15[I]);
    }
  }
  { /* fall back: sort, then force the spacings */
    int j;
    for (i = 1; i < L; i++) {
      const int value = NLSF_Q15[i];
      for (j = i - 1; j >= 0 && value < NLSF_Q15[j]; j--) NLSF_Q15[j + 1] = NLSF_Q15[j];
      NLSF_Q15[j + 1] = (int16_t)value;
    }
    NLSF_Q15[0] = NLSF_Q15[0] > NDeltaMin_Q15[0] ? NLSF_Q15[0] : NDeltaMin_Q15[0];
    for (i = 1; i < L; i++) {
      int32_t s = (int32_t)NLSF_Q15[i - 1] + NDeltaMin_Q15[i];
      s = s > 32767 ? 32767 : s < -32768 ? -32768 : s;
      NLSF_Q15[i] = (int16_t)(NLSF_Q15[i] > s ? NLSF_Q15[i] : s);
    }
    { const int v = (1 << 15) - NDeltaMin_Q15[L]; NLSF_Q15[L - 1] = (int16_t)(NLSF_Q15[L - 1] < v ? NLSF_Q15[L - 1] : v); }
    for (i = L - 2; i >= 0; i--) { const int v = NLSF_Q15[i + 1] - NDeltaMin_Q15[i + 1]; NLSF_Q15[i] = (int16_t)(NLSF_Q15[i] < v ? NLSF_Q15[i] : v); }
  }
}
void orc_silk_NLSF_VQ_weights_laroia(int16_t *pW, const int16_t *pNLSF_Q15, int D) {
  int k;
  int32_t t1, t2;
  t1 = pNLSF_Q15[0] > 1 ? pNLSF_Q15[0] : 1;
  t1 = ((int32_t)1 << 17) / t1;
  t2 = pNLSF_Q15[1] - pNLSF_Q15[0] > 1 ? pNLSF_Q15[1] - pNLSF_Q15[0] : 1;
  t2 = ((int32_t)1 << 17) / t2;
  pW[0] = (int16_t)(t1 + t2 < 32767 ? t1 + t2 : 32767);
  for (k = 1; k < D - 1; k += 2) {
    t1 = pNLSF_Q15[k + 1] - pNLSF_Q15[k] > 1 ? pNLSF_Q15[k + 1] - pNLSF_Q15[k] : 1;
    t1 = ((int32_t)1 << 17) / t1;
    pW[k] = (int16_t)(t1 + t2 < 32767 ? t1 + t2 : 32767);
    t2 = pNLSF_Q15[k + 2] - pNLSF_Q15[k + 1] > 1 ? pNLSF_Q15[k + 2] - pNLSF_Q15[k + 1] : 1;
    t2 = ((int32_t)1 << 17) / t2;
    pW[k + 1] = (int16_t)(t1 + t2 < 32767 ? t1 + t2 : 32767);
  }
  t1 = (1 << 15) - pNLSF_Q15[D - 1] > 1 ? (1 << 15) - pNLSF_Q15[D - 1] : 1;
  t1 = ((int32_t)1 << 17) / t1;
  pW[D - 1] = (int16_t)(t1 + t2 < 32767 ? t1 + t2 : 32767);
}
static void nlsf_unpack(int16_t *ec_ix, uint8_t *pred_Q8, const NlsfCB *cb, int CB1_index) {
  const uint8_t *sel = &cb->ec_sel[CB1_index * cb->order / 2];
  int i;
  for (i = 0; i < cb->order; i += 2) {
    const uint8_t entry = *sel++;
    ec_ix[i] = (int16_t)(((entry >> 1) & 7) * 9);
    pred_Q8[i] = cb->pred_Q8[i + (entry & 1) * (cb->order - 1)];
    ec_ix[i + 1] = (int16_t)(((entry >> 5) & 7) * 9);
    pred_Q8[i + 1] = cb->pred_Q8[i + ((entry >> 4) & 1) * (cb->order - 1) + 1];
  }
}
static void nlsf_decode(int16_t *pNLSF_Q15, const int8_t *NLSFIndices, const NlsfCB *cb) {
  uint8_t pred_Q8[MAX_ORDER];
  int16_t ec_ix[MAX_ORDER], res_Q10[MAX_ORDER];
  const uint8_t *el = &cb->CB1_NLSF_Q8[NLSFIndices[0] * cb->order];
  const int16_t *wq = &cb->CB1_Wght_Q9[NLSFIndices[0] * cb->order];
  int i, out_Q10 = 0;
  nlsf_unpack(ec_ix, pred_Q8, cb, NLSFIndices[0]);
  for (i = cb->order - 1; i >= 0; i--) { /* silk_NLSF_residual_dequant */
    const int pred_Q10 = smulbb(out_Q10, (int16_t)pred_Q8[i]) >> 8;
    out_Q10 = (int)((uint32_t)(int32_t)NLSFIndices[1 + i] << 10);
    if (out_Q10 > 0) out_Q10 = out_Q10 - 102;       /* SILK_FIX_CONST(0.1, 10) */
    else if (out_Q10 < 0) out_Q10 = out_Q10 + 102;
    out_Q10 = smlawb(pred_Q10, out_Q10, cb->quantStepSize_Q16);
    res_Q10[i] = (int16_t)out_Q10;
  }
  for (i = 0; i < cb->order; i++) {
    const int32_t t = ((int32_t)((uint32_t)(int32_t)res_Q10[i] << 14)) / wq[i] + ((int32_t)(int16_t)el[i] << 7);
    pNLSF_Q15[i] = (int16_t)(t < 0 ? 0 : t > 32767 ? 32767 : t);
  }
  orc_silk_NLSF_stabilize(pNLSF_Q15, cb->deltaMin_Q15, cb->order);
}
#define DD_STATES 4
#define AMP 4
#define AMP_EXT 10
static int32_t del_dec_quant(int8_t *indices, const int16_t *x_Q10, const int16_t *w_Q5, const uint8_t *pred_coef_Q8, const int16_t *ec_ix,
                             const uint8_t *ec_rates_Q5, int quant_step_size_Q16, int16_t inv_quant_step_size_Q6, int32_t mu_Q20, int order) {
  int i, j, nStates, ind_tmp, ind_min_max, ind_max_min, in_Q10, res_Q10, pred_Q10, diff_Q10, rate0_Q5, rate1_Q5;
  int16_t out0_Q10, out1_Q10;
  int32_t RD_tmp, min_Q25, min_max_Q25, max_min_Q25;
  int ind_sort[DD_STATES];
  int8_t ind[DD_STATES][MAX_ORDER];
  int16_t prev_out_Q10[2 * DD_STATES];
  int32_t RD_Q25[2 * DD_STATES], RD_min_Q25[DD_STATES], RD_max_Q25[DD_STATES];
  int out0_table[2 * AMP_EXT], out1_table[2 * AMP_EXT];
  for (i = -AMP_EXT; i <= AMP_EXT - 1; i++) {
    out0_Q10 = (int16_t)(i * 1024);
    out1_Q10 = (int16_t)(out0_Q10 + 1024);
    if (i > 0) { out0_Q10 = (int16_t)(out0_Q10 - 102); out1_Q10 = (int16_t)(out1_Q10 - 102); }
    else if (i == 0) out1_Q10 = (int16_t)(out1_Q10 - 102);
    else if (i == -1) out0_Q10 = (int16_t)(out0_Q10 + 102);
    else { out0_Q10 = (int16_t)(out0_Q10 + 102); out1_Q10 = (int16_t)(out1_Q10 + 102); }
    out0_table[i + AMP_EXT] = smulbb(out0_Q10, quant_step_size_Q16) >> 16;
    out1_table[i + AMP_EXT] = smulbb(out1_Q10, quant_step_size_Q16) >> 16;
  }
  nStates = 1;
  RD_Q25[0] = 0;
  prev_out_Q10[0] = 0;
  for (i = order - 1; i >= 0; i--) {
    const uint8_t *rates_Q5 = &ec_rates_Q5[ec_ix[i]];
    in_Q10 = x_Q10[i];
    for (j = 0; j < nStates; j++) {
      pred_Q10 = smulbb((int16_t)pred_coef_Q8[i], prev_out_Q10[j]) >> 8;
      res_Q10 = (int16_t)(in_Q10 - pred_Q10);
      ind_tmp = smulbb(inv_quant_step_size_Q6, res_Q10) >> 16;
      ind_tmp = ind_tmp < -AMP_EXT ? -AMP_EXT : ind_tmp > AMP_EXT - 1 ? AMP_EXT - 1 : ind_tmp;
      ind[j][i] = (int8_t)ind_tmp;
      out0_Q10 = (int16_t)out0_table[ind_tmp + AMP_EXT];
      out1_Q10 = (int16_t)out1_table[ind_tmp + AMP_EXT];
      out0_Q10 = (int16_t)(out0_Q10 + pred_Q10);
      out1_Q10 = (int16_t)(out1_Q10 + pred_Q10);
      prev_out_Q10[j] = out0_Q10;
      prev_out_Q10[j + nStates] = out1_Q10;
      if (ind_tmp + 1 >= AMP) {
        if (ind_tmp + 1 == AMP) { rate0_Q5 = rates_Q5[ind_tmp + AMP]; rate1_Q5 = 280; }
        else { rate0_Q5 = (280 - 43 * AMP) + smulbb(43, ind_tmp); rate1_Q5 = (int16_t)(rate0_Q5 + 43); }
      } else if (ind_tmp <= -AMP) {
        if (ind_tmp == -AMP) { rate0_Q5 = 280; rate1_Q5 = rates_Q5[ind_tmp + 1 + AMP]; }
        else { rate0_Q5 = (280 - 43 * AMP) + smulbb(-43, ind_tmp); rate1_Q5 = (int16_t)(rate0_Q5 - 43); }
      } else { rate0_Q5 = rates_Q5[ind_tmp + AMP]; rate1_Q5 = rates_Q5[ind_tmp + 1 + AMP]; }
      RD_tmp = RD_Q25[j];
      diff_Q10 = (int16_t)(in_Q10 - out0_Q10);
      RD_Q25[j] = (RD_tmp + smulbb(diff_Q10, diff_Q10) * w_Q5[i]) + smulbb(mu_Q20, rate0_Q5);
      diff_Q10 = (int16_t)(in_Q10 - out1_Q10);
      RD_Q25[j + nStates] = (RD_tmp + smulbb(diff_Q10, diff_Q10) * w_Q5[i]) + smulbb(mu_Q20, rate1_Q5);
    }
    if (nStates <= DD_STATES / 2) {
      for (j = 0; j < nStates; j++) ind[j + nStates][i] = (int8_t)(ind[j][i] + 1);
      nStates <<= 1;
      for (j = nStates; j < DD_STATES; j++) ind[j][i] = ind[j - nStates][i];
    } else {
      for (j = 0; j < DD_STATES; j++) {
        if (RD_Q25[j] > RD_Q25[j + DD_STATES]) {
          RD_max_Q25[j] = RD_Q25[j];
          RD_min_Q25[j] = RD_Q25[j + DD_STATES];
          RD_Q25[j] = RD_min_Q25[j];
          RD_Q25[j + DD_STATES] = RD_max_Q25[j];
          out0_Q10 = prev_out_Q10[j];
          prev_out_Q10[j] = prev_out_Q10[j + DD_STATES];
          prev_out_Q10[j + DD_STATES] = out0_Q10;
          ind_sort[j] = j + DD_STATES;
        } else {
          RD_min_Q25[j] = RD_Q25[j];
          RD_max_Q25[j] = RD_Q25[j + DD_STATES];
          ind_sort[j] = j;
        }
      }
      while (1) {
        min_max_Q25 = INT32_MAX; max_min_Q25 = 0; ind_min_max = 0; ind_max_min = 0;
        for (j = 0; j < DD_STATES; j++) {
          if (min_max_Q25 > RD_max_Q25[j]) { min_max_Q25 = RD_max_Q25[j]; ind_min_max = j; }
          if (max_min_Q25 < RD_min_Q25[j]) { max_min_Q25 = RD_min_Q25[j]; ind_max_min = j; }
        }
        if (min_max_Q25 >= max_min_Q25) break;
        ind_sort[ind_max_min] = ind_sort[ind_min_max] ^ DD_STATES;
        RD_Q25[ind_max_min] = RD_Q25[ind_min_max + DD_STATES];
        prev_out_Q10[ind_max_min] = prev_out_Q10[ind_min_max + DD_STATES];
        RD_min_Q25[ind_max_min] = 0;
        RD_max_Q25[ind_min_max] = INT32_MAX;
        memcpy(ind[ind_max_min], ind[ind_min_max], MAX_ORDER * sizeof(int8_t));
      }
      for (j = 0; j < DD_STATES; j++) ind[j][i] = (int8_t)(ind[j][i] + (ind_sort[j] >> 2));
    }
  }
  ind_tmp = 0;
  min_Q25 = INT32_MAX;
  for (j = 0; j < 2 * DD_STATES; j++) if (min_Q25 > RD_Q25[j]) { min_Q25 = RD_Q25[j]; ind_tmp = j; }
  for (j = 0; j < order; j++) indices[j] = ind[ind_tmp & (DD_STATES - 1)][j];
  indices[0] = (int8_t)(indices[0] + (ind_tmp >> 2));
  return min_Q25;
}
int32_t orc_silk_NLSF_encode(int8_t *NLSFIndices, int16_t *pNLSF_Q15, int wb, const int16_t *pW_Q2, int NLSF_mu_Q20, int nSurvivors, int signalType) {
  NlsfCB cb;
  int32_t err_Q24[32], RD_Q25[32];
  int tempIndices1[32], i, s, bestIndex;
  int8_t tempIndices2[32 * MAX_ORDER];
  int16_t res_Q10[MAX_ORDER], NLSF_tmp_Q15[MAX_ORDER], W_adj_Q5[MAX_ORDER], ec_ix[MAX_ORDER];
  uint8_t pred_Q8[MAX_ORDER];
  get_cb(&cb, wb);
  orc_silk_NLSF_stabilize(pNLSF_Q15, cb.deltaMin_Q15, cb.order);
  for (i = 0; i < cb.nVectors; i++) { /* silk_NLSF_VQ */
    const uint8_t *c = &cb.CB1_NLSF_Q8[i * cb.order];
    const int16_t *w = &cb.CB1_Wght_Q9[i * cb.order];
    int32_t sum = 0, pred = 0, d, dw;
    int m;
    for (m = cb.order - 2; m >= 0; m -= 2) {
      d = pNLSF_Q15[m + 1] - ((int32_t)c[m + 1] << 7);
      dw = smulbb(d, w[m + 1]);
      sum = sum + abs(dw - (pred >> 1));
      pred = dw;
      d = pNLSF_Q15[m] - ((int32_t)c[m] << 7);
      dw = smulbb(d, w[m]);
      sum = sum + abs(dw - (pred >> 1));
      pred = dw;
    }
    err_Q24[i] = sum;
  }
  sort_increasing(err_Q24, tempIndices1, cb.nVectors, nSurvivors);
  for (s = 0; s < nSurvivors; s++) {
    const int ind1 = tempIndices1[s];
    const uint8_t *el = &cb.CB1_NLSF_Q8[ind1 * cb.order], *iCDF_ptr;
    const int16_t *wq = &cb.CB1_Wght_Q9[ind1 * cb.order];
    int prob_Q8, bits_q7;
    for (i = 0; i < cb.order; i++) {
      const int32_t W_tmp_Q9 = wq[i];
      NLSF_tmp_Q15[i] = (int16_t)((int16_t)el[i] << 7);
      res_Q10[i] = (int16_t)(smulbb(pNLSF_Q15[i] - NLSF_tmp_Q15[i], W_tmp_Q9) >> 14);
      W_adj_Q5[i] = (int16_t)div32_varQ((int32_t)pW_Q2[i], smulbb(W_tmp_Q9, W_tmp_Q9), 21);
    }
    nlsf_unpack(ec_ix, pred_Q8, &cb, ind1);
    RD_Q25[s] = del_dec_quant(&tempIndices2[s * MAX_ORDER], res_Q10, W_adj_Q5, pred_Q8, ec_ix, cb.ec_Rates_Q5, cb.quantStepSize_Q16,
                              (int16_t)cb.invQuantStepSize_Q6, NLSF_mu_Q20, cb.order);
    iCDF_ptr = &cb.CB1_iCDF[(signalType >> 1) * cb.nVectors];
    prob_Q8 = ind1 == 0 ? 256 - iCDF_ptr[ind1] : iCDF_ptr[ind1 - 1] - iCDF_ptr[ind1];
    bits_q7 = (8 << 7) - orc_silk_lin2log(prob_Q8);
    RD_Q25[s] = RD_Q25[s] + smulbb(bits_q7, NLSF_mu_Q20 >> 2);
  }
  sort_increasing(RD_Q25, &bestIndex, nSurvivors, 1);
  NLSFIndices[0] = (int8_t)tempIndices1[bestIndex];
  memcpy(&NLSFIndices[1], &tempIndices2[bestIndex * MAX_ORDER], cb.order * sizeof(int8_t));
  nlsf_decode(pNLSF_Q15, NLSFIndices, &cb);
  return RD_Q25[0];
}
void orc_silk_NLSF_decode(int16_t *pNLSF_Q15, const int8_t *NLSFIndices, int wb) {
  NlsfCB cb;
  get_cb(&cb, wb);
  nlsf_decode(pNLSF_Q15, NLSFIndices, &cb);
}
static void bwexpander_32(int32_t *ar, int d, int32_t chirp_Q16) {
  const int32_t chirp_minus_one_Q16 = chirp_Q16 - 65536;
  int i;
  for (i = 0; i < d - 1; i++) {
    ar[i] = smulww(chirp_Q16, ar[i]);
    chirp_Q16 += rshift_round(chirp_Q16 * chirp_minus_one_Q16, 16);
  }
  ar[d - 1] = smulww(chirp_Q16, ar[d - 1]);
}
static void lpc_fit(int16_t *a_QOUT, int32_t *a_QIN, int QOUT, int QIN, int d) {
  int i, k, idx = 0;
  for (i = 0; i < 10; i++) {
    int32_t maxabs = 0, chirp_Q16;
    for (k = 0; k < d; k++) {
      const int32_t absval = a_QIN[k] < 0 ? -a_QIN[k] : a_QIN[k];
      if (absval > maxabs) { maxabs = absval; idx = k; }
    }
    maxabs = rshift_round(maxabs, QIN - QOUT);
    if (maxabs > 32767) {
      maxabs = maxabs < 163838 ? maxabs : 163838;
      chirp_Q16 = 65470 - ((int32_t)((uint32_t)(maxabs - 32767) << 14)) / ((maxabs * (idx + 1)) >> 2); /* SILK_FIX_CONST(0.999, 16) = 65470 */
      bwexpander_32(a_QIN, d, chirp_Q16);
    } else break;
  }
  if (i == 10) {
    for (k = 0; k < d; k++) {
      const int32_t v = rshift_round(a_QIN[k], QIN - QOUT);
      a_QOUT[k] = (int16_t)(v > 32767 ? 32767 : v < -32768 ? -32768 : v);
      a_QIN[k] = (int32_t)((uint32_t)(int32_t)a_QOUT[k] << (QIN - QOUT));
    }
  } else {
    for (k = 0; k < d; k++) a_QOUT[k] = (int16_t)rshift_round(a_QIN[k], QIN - QOUT);
  }
}
int32_t orc_silk_LPC_inverse_pred_gain(const int16_t *A_Q12, int order) {
  const int QA = 24;
  const int32_t A_LIMIT = 16773022; /* SILK_FIX_CONST(0.99975, 24) */
  int32_t A_QA[MAX_ORDER + 8], DC_resp = 0, invGain_Q30 = 1 << 30, rc_Q31, rc_mult1_Q30, rc_mult2, tmp1, tmp2;
  int k, n, mult2Q;
  for (k = 0; k < order; k++) { DC_resp += A_Q12[k]; A_QA[k] = (int32_t)((uint32_t)(int32_t)A_Q12[k] << (QA - 12)); }
  if (DC_resp >= 4096) return 0;
  for (k = order - 1; k > 0; k--) {
    if (A_QA[k] > A_LIMIT || A_QA[k] < -A_LIMIT) return 0;
    rc_Q31 = -(int32_t)((uint32_t)A_QA[k] << (31 - QA));
    rc_mult1_Q30 = (1 << 30) - smmul(rc_Q31, rc_Q31);
    invGain_Q30 = (int32_t)((uint32_t)smmul(invGain_Q30, rc_mult1_Q30) << 2);
    if (invGain_Q30 < 107374) return 0; /* SILK_FIX_CONST(1.0f / 1e4f, 30) */
    mult2Q = 32 - clz32(abs(rc_mult1_Q30));
    rc_mult2 = inverse32_varQ(rc_mult1_Q30, mult2Q + 30);
    for (n = 0; n < (k + 1) >> 1; n++) {
      int64_t tmp64;
      tmp1 = A_QA[n];
      tmp2 = A_QA[k - n - 1];
      tmp64 = rshift_round64((int64_t)sub_sat32(tmp1, (int32_t)rshift_round64((int64_t)tmp2 * rc_Q31, 31)) * rc_mult2, mult2Q);
      if (tmp64 > INT32_MAX || tmp64 < INT32_MIN) return 0;
      A_QA[n] = (int32_t)tmp64;
      tmp64 = rshift_round64((int64_t)sub_sat32(tmp2, (int32_t)rshift_round64((int64_t)tmp1 * rc_Q31, 31)) * rc_mult2, mult2Q);
      if (tmp64 > INT32_MAX || tmp64 < INT32_MIN) return 0;
      A_QA[k - n - 1] = (int32_t)tmp64;
    }
  }
  if (A_QA[k] > A_LIMIT || A_QA[k] < -A_LIMIT) return 0;
  rc_Q31 = -(int32_t)((uint32_t)A_QA[0] << (31 - QA));
  rc_mult1_Q30 = (1 << 30) - smmul(rc_Q31, rc_Q31);
  invGain_Q30 = (int32_t)((uint32_t)smmul(invGain_Q30, rc_mult1_Q30) << 2);
  if (invGain_Q30 < 107374) return 0;
  return invGain_Q30;
}
static void find_poly(int32_t *out, const int32_t *cLSF, int dd) {
  const int QA = 16;
  int k, n;
  out[0] = 1 << QA;
  out[1] = -cLSF[0];
  for (k = 1; k < dd; k++) {
    const int32_t ftmp = cLSF[2 * k];
    out[k + 1] = (int32_t)((uint32_t)out[k - 1] << 1) - (int32_t)rshift_round64((int64_t)ftmp * out[k], QA);
    for (n = k; n > 1; n--) out[n] += out[n - 2] - (int32_t)rshift_round64((int64_t)ftmp * out[n - 1], QA);
    out[1] -= ftmp;
  }
}
void orc_silk_NLSF2A(int16_t *a_Q12, const int16_t *NLSF, int d) {
  static const unsigned char ordering16[16] = {0, 15, 8, 7, 4, 11, 12, 3, 2, 13, 10, 5, 6, 9, 14, 1};
  static const unsigned char ordering10[10] = {0, 9, 6, 3, 4, 5, 8, 1, 2, 7};
  const unsigned char *ordering = d == 16 ? ordering16 : ordering10;
  const int QA = 16, dd = d >> 1;
  int32_t cos_LSF_QA[MAX_ORDER + 8], P[MAX_ORDER / 2 + 5], Q[MAX_ORDER / 2 + 5], a32_QA1[MAX_ORDER + 8];
  int k, i;
  for (k = 0; k < d; k++) {
    const int32_t f_int = NLSF[k] >> (15 - 7), f_frac = NLSF[k] - (f_int << (15 - 7));
    const int32_t cos_val = ob_silk_LSFCosTab_Q12[f_int], delta = ob_silk_LSFCosTab_Q12[f_int + 1] - cos_val;
    cos_LSF_QA[ordering[k]] = rshift_round((int32_t)((uint32_t)cos_val << 8) + delta * f_frac, 20 - QA);
  }
  find_poly(P, &cos_LSF_QA[0], dd);
  find_poly(Q, &cos_LSF_QA[1], dd);
  for (k = 0; k < dd; k++) {
    const int32_t Ptmp = P[k + 1] + P[k], Qtmp = Q[k + 1] - Q[k];
    a32_QA1[k] = -Qtmp - Ptmp;
    a32_QA1[d - k - 1] = Qtmp - Ptmp;
  }
  lpc_fit(a_Q12, a32_QA1, 12, QA + 1, d);
  for (i = 0; orc_silk_LPC_inverse_pred_gain(a_Q12, d) == 0 && i < 16; i++) {
    bwexpander_32(a32_QA1, d, 65536 - (int32_t)((uint32_t)2 << i));
    for (k = 0; k < d; k++) a_Q12[k] = (int16_t)rshift_round(a32_QA1[k], QA + 1 - 12);
  }
}
static void interpolate(int16_t *xi, const int16_t *x0, const int16_t *x1, int ifact_Q2, int d) {
  int i;
  for (i = 0; i < d; i++) xi[i] = (int16_t)(x0[i] + (smulbb(x1[i] - x0[i], ifact_Q2) >> 2));
}
typedef struct { /* one stream's view of silk_encoder_state for silk_process_NLSFs; == OpusB200SilkNlsf */
  int16_t NLSF_Q15[16];          /* in: unquantised, out: quantised */
  int16_t prev_NLSFq_Q15[16];
  int16_t PredCoef_Q12[2][16];   /* out */
  int8_t NLSFIndices[17];        /* out */
  int8_t NLSFInterpCoef_Q2, signalType, useInterpolatedNLSFs;
  int32_t speech_activity_Q8, NLSF_MSVQ_Survivors;
} OrcSilkNlsf;
int orc_silk_nlsf_size(void) { return (int)sizeof(OrcSilkNlsf); }
void orc_silk_process_NLSFs(OrcSilkNlsf *f, int wb, int nb_subfr) {
  const int order = wb ? 16 : 10;
  int16_t pNLSF0_temp_Q15[MAX_ORDER], pNLSFW_QW[MAX_ORDER], pNLSFW0_temp_QW[MAX_ORDER];
  int i, doInterpolate, NLSF_mu_Q20;
  NLSF_mu_Q20 = smlawb(3146, -268434, f->speech_activity_Q8); /* SILK_FIX_CONST(0.003, 20), SILK_FIX_CONST(-0.001, 28) */
  if (nb_subfr == 2) NLSF_mu_Q20 = NLSF_mu_Q20 + (NLSF_mu_Q20 >> 1);
  orc_silk_NLSF_VQ_weights_laroia(pNLSFW_QW, f->NLSF_Q15, order);
  doInterpolate = f->useInterpolatedNLSFs == 1 && f->NLSFInterpCoef_Q2 < 4;
  if (doInterpolate) {
    int16_t i_sqr_Q15;
    interpolate(pNLSF0_temp_Q15, f->prev_NLSFq_Q15, f->NLSF_Q15, f->NLSFInterpCoef_Q2, order);
    orc_silk_NLSF_VQ_weights_laroia(pNLSFW0_temp_QW, pNLSF0_temp_Q15, order);
    i_sqr_Q15 = (int16_t)(smulbb(f->NLSFInterpCoef_Q2, f->NLSFInterpCoef_Q2) << 11);
    for (i = 0; i < order; i++) pNLSFW_QW[i] = (int16_t)((pNLSFW_QW[i] >> 1) + (smulbb(pNLSFW0_temp_QW[i], i_sqr_Q15) >> 16));
  }
  orc_silk_NLSF_encode(f->NLSFIndices, f->NLSF_Q15, wb, pNLSFW_QW, NLSF_mu_Q20, f->NLSF_MSVQ_Survivors, f->signalType);
  orc_silk_NLSF2A(f->PredCoef_Q12[1], f->NLSF_Q15, order);
  if (doInterpolate) {
    interpolate(pNLSF0_temp_Q15, f->prev_NLSFq_Q15, f->NLSF_Q15, f->NLSFInterpCoef_Q2, order);
    orc_silk_NLSF2A(f->PredCoef_Q12[0], pNLSF0_temp_Q15, order);
  } else memcpy(f->PredCoef_Q12[0], f->PredCoef_Q12[1], order * sizeof(int16_t));
}

/* ---- LPC estimation of a SILK frame: silk_find_LPC_FLP (silk/float/find_LPC_FLP.c:36-105) with silk_A2NLSF (silk/A2NLSF.c:45-267), the float wrappers
 * silk_A2NLSF_FLP / silk_NLSF2A_FLP (silk/float/wrappers_FLP.c:36-72), silk_LPC_analysis_filter_FLP (orders 10 and 16,
 * silk/float/LPC_analysis_filter_FLP.c:43-139, 214-248) and silk_energy_FLP (silk/float/energy_FLP.c:35-59) on the Burg restatement of silk_oracle.c. ---- */
#include <math.h>
float orc_silk_burg_modified(float *A, const float *x, float minInvGain, int subfr_length, int nb_subfr, int D); /* silk_oracle.c */
static void a2nlsf_trans_poly(int32_t *p, int dd) {
  int k, n;
  for (k = 2; k <= dd; k++) {
    for (n = dd; n > k; n--) p[n - 2] -= p[n];
    p[k - 2] -= (int32_t)((uint32_t)p[k] << 1);
  }
}
static int32_t a2nlsf_eval_poly(const int32_t *p, int32_t x, int dd) {
  const int32_t x_Q16 = (int32_t)((uint32_t)x << 4);
  int32_t y32 = p[dd];
  int n;
  for (n = dd - 1; n >= 0; n--) y32 = (int32_t)(p[n] + (((int64_t)y32 * x_Q16) >> 16)); /* silk_SMLAWW */
  return y32;
}
static void a2nlsf_init(const int32_t *a_Q16, int32_t *P, int32_t *Q, int dd) {
  int k;
  P[dd] = 1 << 16;
  Q[dd] = 1 << 16;
  for (k = 0; k < dd; k++) {
    P[k] = -a_Q16[dd - k - 1] - a_Q16[dd + k];
    Q[k] = -a_Q16[dd - k - 1] + a_Q16[dd + k];
  }
  for (k = dd; k > 0; k--) {
    P[k - 1] -= P[k];
    Q[k - 1] += Q[k];
  }
  a2nlsf_trans_poly(P, dd);
  a2nlsf_trans_poly(Q, dd);
}
void orc_silk_A2NLSF(int16_t *NLSF, int32_t *a_Q16, int d) {
  const int dd = d >> 1;
  int i = 0, k = 1, m, root_ix, ffrac;
  int32_t xlo, xhi, xmid, ylo, yhi, ymid, thr = 0, nom, den, P[MAX_ORDER / 2 + 1], Q[MAX_ORDER / 2 + 1], *p = P;
  a2nlsf_init(a_Q16, P, Q, dd);
  xlo = ob_silk_LSFCosTab_Q12[0];
  ylo = a2nlsf_eval_poly(p, xlo, dd);
  if (ylo < 0) {
    NLSF[0] = 0;
    p = Q;
    ylo = a2nlsf_eval_poly(p, xlo, dd);
    root_ix = 1;
  } else root_ix = 0;
  while (1) {
    xhi = ob_silk_LSFCosTab_Q12[k];
    yhi = a2nlsf_eval_poly(p, xhi, dd);
    if ((ylo <= 0 && yhi >= thr) || (ylo >= 0 && yhi <= -thr)) {
      thr = yhi == 0 ? 1 : 0;
      ffrac = -256;
      for (m = 0; m < 3; m++) {
        xmid = rshift_round(xlo + xhi, 1);
        ymid = a2nlsf_eval_poly(p, xmid, dd);
        if ((ylo <= 0 && ymid >= 0) || (ylo >= 0 && ymid <= 0)) { xhi = xmid; yhi = ymid; }
        else { xlo = xmid; ylo = ymid; ffrac = ffrac + (128 >> m); }
      }
      if (abs(ylo) < 65536) {
        den = ylo - yhi;
        nom = (int32_t)((uint32_t)ylo << (8 - 3)) + (den >> 1);
        if (den != 0) ffrac += nom / den;
      } else ffrac += ylo / ((ylo - yhi) >> (8 - 3));
      { const int32_t v = (int32_t)((uint32_t)k << 8) + ffrac; NLSF[root_ix] = (int16_t)(v < 32767 ? v : 32767); }
      root_ix++;
      if (root_ix >= d) break;
      p = (root_ix & 1) ? Q : P;
      xlo = ob_silk_LSFCosTab_Q12[k - 1];
      ylo = (int32_t)((uint32_t)(int32_t)(1 - (root_ix & 2)) << 12);
    } else {
      k++;
      xlo = xhi;
      ylo = yhi;
      thr = 0;
      if (k > 128) {
        i++;
        if (i > 16) { /* white spectrum */
          NLSF[0] = (int16_t)((1 << 15) / (d + 1));
          for (k = 1; k < d; k++) NLSF[k] = (int16_t)(NLSF[k - 1] + NLSF[0]);
          return;
        }
        bwexpander_32(a_Q16, d, 65536 - (int32_t)((uint32_t)1 << i));
        a2nlsf_init(a_Q16, P, Q, dd);
        p = P;
        xlo = ob_silk_LSFCosTab_Q12[0];
        ylo = a2nlsf_eval_poly(p, xlo, dd);
        if (ylo < 0) {
          NLSF[0] = 0;
          p = Q;
          ylo = a2nlsf_eval_poly(p, xlo, dd);
          root_ix = 1;
        } else root_ix = 0;
        k = 1;
      }
    }
  }
}
static void a2nlsf_flp(int16_t *NLSF_Q15, const float *pAR, int order) {
  int32_t a_fix_Q16[MAX_ORDER];
  int i;
  for (i = 0; i < order; i++) a_fix_Q16[i] = (int32_t)lrintf(pAR[i] * 65536.0f);
  orc_silk_A2NLSF(NLSF_Q15, a_fix_Q16, order);
}
static double energy_flp(const float *data, int n) {
  double r = 0.0;
  int i;
  for (i = 0; i < n - 3; i += 4)
    r += data[i + 0] * (double)data[i + 0] + data[i + 1] * (double)data[i + 1] + data[i + 2] * (double)data[i + 2] + data[i + 3] * (double)data[i + 3];
  for (; i < n; i++) r += data[i] * (double)data[i];
  return r;
}
static void lpc_analysis_filter_flp(float *r_LPC, const float *PredCoef, const float *s, int length, int Order) {
  int ix, j;
  for (ix = Order; ix < length; ix++) {
    const float *s_ptr = &s[ix - 1];
    float LPC_pred = s_ptr[0] * PredCoef[0];
    for (j = 1; j < Order; j++) LPC_pred = LPC_pred + s_ptr[-j] * PredCoef[j];
    r_LPC[ix] = s_ptr[1] - LPC_pred;
  }
  memset(r_LPC, 0, Order * sizeof(float));
}
typedef struct { /* what silk_find_LPC_FLP reads and writes in silk_encoder_state; == OpusB200SilkFindLpc */
  int16_t NLSF_Q15[16];         /* out */
  int16_t prev_NLSFq_Q15[16];
  float minInvGain;
  int32_t useInterpolatedNLSFs, first_frame_after_reset;
  int32_t NLSFInterpCoef_Q2;    /* out */
} OrcSilkFindLpc;
int orc_silk_find_lpc_size(void) { return (int)sizeof(OrcSilkFindLpc); }
/* x: nb_subfr * (subfr_length + order) samples */
void orc_silk_find_LPC(OrcSilkFindLpc *f, const float *x, int order, int nb_subfr, int subfr_length_in) {
  const int subfr_length = subfr_length_in + order;
  float a[MAX_ORDER], a_tmp[MAX_ORDER], res_nrg, res_nrg_2nd, res_nrg_interp, LPC_res[2 * (80 + 16) + 16];
  int16_t NLSF0_Q15[MAX_ORDER], a_fix_Q12[MAX_ORDER];
  int k, i;
  f->NLSFInterpCoef_Q2 = 4;
  res_nrg = orc_silk_burg_modified(a, x, f->minInvGain, subfr_length, nb_subfr, order);
  if (f->useInterpolatedNLSFs && !f->first_frame_after_reset && nb_subfr == 4) {
    res_nrg -= orc_silk_burg_modified(a_tmp, x + 2 * subfr_length, f->minInvGain, subfr_length, 2, order);
    a2nlsf_flp(f->NLSF_Q15, a_tmp, order);
    res_nrg_2nd = 3.402823466e+38f; /* silk_float_MAX = FLT_MAX */
    for (k = 3; k >= 0; k--) {
      interpolate(NLSF0_Q15, f->prev_NLSFq_Q15, f->NLSF_Q15, k, order);
      orc_silk_NLSF2A(a_fix_Q12, NLSF0_Q15, order);
      for (i = 0; i < order; i++) a_tmp[i] = (float)a_fix_Q12[i] * (1.0f / 4096.0f);
      lpc_analysis_filter_flp(LPC_res, a_tmp, x, 2 * subfr_length, order);
      res_nrg_interp = (float)(energy_flp(LPC_res + order, subfr_length - order) + energy_flp(LPC_res + order + subfr_length, subfr_length - order));
      if (res_nrg_interp < res_nrg) {
        res_nrg = res_nrg_interp;
        f->NLSFInterpCoef_Q2 = k;
      } else if (res_nrg_interp > res_nrg_2nd) break;
      res_nrg_2nd = res_nrg_interp;
    }
  }
  if (f->NLSFInterpCoef_Q2 == 4) a2nlsf_flp(f->NLSF_Q15, a, order);
}

/* ---- the prediction analysis of a SILK frame: silk_find_pred_coefs_FLP (silk/float/find_pred_coefs_FLP.c:35-120) = silk_find_LTP_FLP
 * (float/find_LTP_FLP.c:34-65 with corrMatrix_FLP.c:38-95), silk_quant_LTP_gains[_FLP] (quant_LTP_gains.c:34-132, float/wrappers_FLP.c:175-209) with
 * silk_VQ_WMat_EC_c (VQ_WMat_EC.c:35-131), silk_LTP_scale_ctrl_FLP (float/LTP_scale_ctrl_FLP.c:34-58), silk_LTP_analysis_filter_FLP
 * (float/LTP_analysis_filter_FLP.c:34-75), silk_find_LPC_FLP and silk_process_NLSFs[_FLP] (above), silk_residual_energy_FLP
 * (float/residual_energy_FLP.c:91-117). ---- */
int32_t orc_silk_log2lin(int32_t inLog_Q7); /* silk_oracle.c */
typedef struct { /* what silk_find_pred_coefs_FLP reads and writes in silk_encoder_state_FLP / silk_encoder_control_FLP; == OpusB200SilkPredCoefs */
  float Gains[4];
  int32_t pitchL[4];
  float coding_quality;
  int32_t signalType, condCoding, first_frame_after_reset, useInterpolatedNLSFs, PacketLoss_perc, nFramesPerPacket, LBRR_flag, SNR_dB_Q7, speech_activity_Q8,
      NLSF_MSVQ_Survivors;
  int16_t prev_NLSFq_Q15[16];   /* in / out */
  int32_t sum_log_gain_Q7;      /* in / out */
  float LTPCoef[20], LTPredCodGain, LTP_scale, PredCoef[2][16], ResNrg[4];   /* out */
  int8_t LTPIndex[4], PERIndex, LTP_scaleIndex, NLSFInterpCoef_Q2, pad;
  int8_t NLSFIndices[17], pad2[3];
} OrcSilkPredCoefs;
int orc_silk_pred_coefs_size(void) { return (int)sizeof(OrcSilkPredCoefs); }
static int32_t add_pos_sat32(int32_t a, int32_t b) { return (((uint32_t)a + (uint32_t)b) & 0x80000000u) ? INT32_MAX : a + b; }
static void vq_wmat_ec(int8_t *ind, int32_t *res_nrg_Q15, int32_t *rate_dist_Q8, int *gain_Q7, const int32_t *XX, const int32_t *xX, const int8_t *cb_Q7,
                       const uint8_t *cb_gain_Q7, const uint8_t *cl_Q5, int subfr_len, int32_t max_gain_Q7, int L) {
  int32_t neg[5], sum1, sum2;
  int k, j;
  for (j = 0; j < 5; j++) neg[j] = -(int32_t)((uint32_t)xX[j] << 7);
  *rate_dist_Q8 = INT32_MAX;
  *res_nrg_Q15 = INT32_MAX;
  *ind = 0;
  for (k = 0; k < L; k++, cb_Q7 += 5) {
    const int gain_tmp_Q7 = cb_gain_Q7[k];
    const int32_t d = gain_tmp_Q7 - max_gain_Q7, penalty = (int32_t)((uint32_t)(d > 0 ? d : 0) << 11);
    sum1 = 32801; /* SILK_FIX_CONST(1.001, 15) */
    sum2 = neg[0] + XX[1] * cb_Q7[1];
    sum2 = sum2 + XX[2] * cb_Q7[2];
    sum2 = sum2 + XX[3] * cb_Q7[3];
    sum2 = sum2 + XX[4] * cb_Q7[4];
    sum2 = (int32_t)((uint32_t)sum2 << 1);
    sum2 = sum2 + XX[0] * cb_Q7[0];
    sum1 = smlawb(sum1, sum2, cb_Q7[0]);
    sum2 = neg[1] + XX[7] * cb_Q7[2];
    sum2 = sum2 + XX[8] * cb_Q7[3];
    sum2 = sum2 + XX[9] * cb_Q7[4];
    sum2 = (int32_t)((uint32_t)sum2 << 1);
    sum2 = sum2 + XX[6] * cb_Q7[1];
    sum1 = smlawb(sum1, sum2, cb_Q7[1]);
    sum2 = neg[2] + XX[13] * cb_Q7[3];
    sum2 = sum2 + XX[14] * cb_Q7[4];
    sum2 = (int32_t)((uint32_t)sum2 << 1);
    sum2 = sum2 + XX[12] * cb_Q7[2];
    sum1 = smlawb(sum1, sum2, cb_Q7[2]);
    sum2 = neg[3] + XX[19] * cb_Q7[4];
    sum2 = (int32_t)((uint32_t)sum2 << 1);
    sum2 = sum2 + XX[18] * cb_Q7[3];
    sum1 = smlawb(sum1, sum2, cb_Q7[3]);
    sum2 = (int32_t)((uint32_t)neg[4] << 1);
    sum2 = sum2 + XX[24] * cb_Q7[4];
    sum1 = smlawb(sum1, sum2, cb_Q7[4]);
    if (sum1 >= 0) {
      const int32_t bits_res_Q8 = smulbb(subfr_len, orc_silk_lin2log(sum1 + penalty) - (15 << 7)), bits_tot_Q8 = bits_res_Q8 + ((int32_t)cl_Q5[k] << 2);
      if (bits_tot_Q8 <= *rate_dist_Q8) {
        *rate_dist_Q8 = bits_tot_Q8;
        *res_nrg_Q15 = sum1 + penalty;
        *ind = (int8_t)k;
        *gain_Q7 = gain_tmp_Q7;
      }
    }
  }
}
static void quant_LTP_gains(int16_t *B_Q14, int8_t *cbk_index, int8_t *periodicity_index, int32_t *sum_log_gain_Q7, int *pred_gain_dB_Q7, const int32_t *XX_Q17,
                            const int32_t *xX_Q17, int subfr_len, int nb_subfr) {
  static const int cb_off[3] = {0, 8, 24}, cb_size[3] = {8, 16, 32};
  int32_t min_rate_dist_Q7 = INT32_MAX, best_sum_log_gain_Q7 = 0, res_nrg_Q15 = 0;
  int8_t temp_idx[4];
  int j, k;
  for (k = 0; k < 3; k++) {
    const int32_t gain_safety = 51; /* SILK_FIX_CONST(0.4, 7) */
    const uint8_t *cl = (const uint8_t *)ob_silk_LTP_gain_BITS_Q5 + cb_off[k], *cg = (const uint8_t *)ob_silk_LTP_vq_gain + cb_off[k];
    const int8_t *cb = (const int8_t *)ob_silk_LTP_gain_vq + 5 * cb_off[k];
    int32_t rate_dist_Q7 = 0, sum_log_gain_tmp_Q7 = *sum_log_gain_Q7, nrg = 0;
    for (j = 0; j < nb_subfr; j++) {
      const int32_t max_gain_Q7 = orc_silk_log2lin((5333 - sum_log_gain_tmp_Q7) + 896) - gain_safety; /* MAX_SUM_LOG_GAIN_DB / 6 and 7 in Q7 */
      int32_t res_nrg_subfr, rate_dist_subfr, t;
      int gain_Q7 = 0;
      vq_wmat_ec(&temp_idx[j], &res_nrg_subfr, &rate_dist_subfr, &gain_Q7, XX_Q17 + 25 * j, xX_Q17 + 5 * j, cb, cg, cl, subfr_len, max_gain_Q7, cb_size[k]);
      nrg = add_pos_sat32(nrg, res_nrg_subfr);
      rate_dist_Q7 = add_pos_sat32(rate_dist_Q7, rate_dist_subfr);
      t = sum_log_gain_tmp_Q7 + orc_silk_lin2log(gain_safety + gain_Q7) - 896;
      sum_log_gain_tmp_Q7 = t > 0 ? t : 0;
    }
    res_nrg_Q15 = nrg;
    if (rate_dist_Q7 <= min_rate_dist_Q7) {
      min_rate_dist_Q7 = rate_dist_Q7;
      *periodicity_index = (int8_t)k;
      memcpy(cbk_index, temp_idx, nb_subfr);
      best_sum_log_gain_Q7 = sum_log_gain_tmp_Q7;
    }
  }
  {
    const int8_t *cb = (const int8_t *)ob_silk_LTP_gain_vq + 5 * cb_off[*periodicity_index];
    for (j = 0; j < nb_subfr; j++)
      for (k = 0; k < 5; k++) B_Q14[j * 5 + k] = (int16_t)(cb[cbk_index[j] * 5 + k] * 128);
  }
  res_nrg_Q15 = nb_subfr == 2 ? res_nrg_Q15 >> 1 : res_nrg_Q15 >> 2;
  *sum_log_gain_Q7 = best_sum_log_gain_Q7;
  *pred_gain_dB_Q7 = smulbb(-3, orc_silk_lin2log(res_nrg_Q15) - (15 << 7));
}
static double inner_flp(const float *a, const float *b, int n) {
  double r = 0.0;
  int i;
  for (i = 0; i < n - 3; i += 4)
    r += a[i + 0] * (double)b[i + 0] + a[i + 1] * (double)b[i + 1] + a[i + 2] * (double)b[i + 2] + a[i + 3] * (double)b[i + 3];
  for (; i < n; i++) r += a[i] * (double)b[i];
  return r;
}
static void corr_matrix_flp(const float *x, int L, int Order, float *XX) {
  const float *ptr1 = &x[Order - 1], *ptr2 = &x[Order - 2];
  double energy = energy_flp(ptr1, L);
  int j, lag;
  XX[0] = (float)energy;
  for (j = 1; j < Order; j++) {
    energy += ptr1[-j] * ptr1[-j] - ptr1[L - j] * ptr1[L - j];
    XX[j * Order + j] = (float)energy;
  }
  for (lag = 1; lag < Order; lag++) {
    energy = inner_flp(ptr1, ptr2, L);
    XX[lag * Order + 0] = (float)energy;
    XX[0 * Order + lag] = (float)energy;
    for (j = 1; j < Order - lag; j++) {
      energy += ptr1[-j] * ptr2[-j] - ptr1[L - j] * ptr2[L - j];
      XX[(lag + j) * Order + j] = (float)energy;
      XX[j * Order + lag + j] = (float)energy;
    }
    ptr2--;
  }
}
/* res_pitch points at the current frame of the pitch residual (history in front of it), x at the current frame of the input (history in front of it) */
void orc_silk_find_pred_coefs(OrcSilkPredCoefs *c, const float *res_pitch, const float *x, int order, int nb_subfr, int subfr_length) {
  float XXLTP[4 * 25], xXLTP[4 * 5], invGains[4], LPC_in_pre[4 * 16 + 320], minInvGain;
  int16_t NLSF_Q15[16] = {0};
  int i, j, k;
  for (i = 0; i < nb_subfr; i++) invGains[i] = 1.0f / c->Gains[i];
  if (c->signalType == 2) {
    const float *r_ptr = res_pitch;
    int32_t XX_Q17[100], xX_Q17[20];
    int16_t B_Q14[20];
    int pred_gain_dB_Q7, round_loss;
    for (k = 0; k < nb_subfr; k++) { /* silk_find_LTP_FLP */
      const float *lag_ptr = r_ptr - (c->pitchL[k] + 5 / 2);
      float *XX = XXLTP + 25 * k, *xX = xXLTP + 5 * k, xx, temp;
      corr_matrix_flp(lag_ptr, subfr_length, 5, XX);
      for (j = 0; j < 5; j++) xX[j] = (float)inner_flp(&lag_ptr[4 - j], r_ptr, subfr_length);
      xx = (float)energy_flp(r_ptr, subfr_length + 5);
      temp = 0.03f * 0.5f * (XX[0] + XX[24]) + 1.0f;
      temp = 1.0f / (xx > temp ? xx : temp);
      for (j = 0; j < 25; j++) XX[j] *= temp;
      for (j = 0; j < 5; j++) xX[j] *= temp;
      r_ptr += subfr_length;
    }
    for (i = 0; i < nb_subfr * 25; i++) XX_Q17[i] = (int32_t)lrintf(XXLTP[i] * 131072.0f);
    for (i = 0; i < nb_subfr * 5; i++) xX_Q17[i] = (int32_t)lrintf(xXLTP[i] * 131072.0f);
    quant_LTP_gains(B_Q14, c->LTPIndex, &c->PERIndex, &c->sum_log_gain_Q7, &pred_gain_dB_Q7, XX_Q17, xX_Q17, subfr_length, nb_subfr);
    for (i = 0; i < nb_subfr * 5; i++) c->LTPCoef[i] = (float)B_Q14[i] * (1.0f / 16384.0f);
    c->LTPredCodGain = (float)pred_gain_dB_Q7 * (1.0f / 128.0f);
    if (c->condCoding == 0) { /* silk_LTP_scale_ctrl_FLP */
      round_loss = c->PacketLoss_perc * c->nFramesPerPacket;
      if (c->LBRR_flag) round_loss = 2 + smulbb(round_loss, round_loss) / 100;
      c->LTP_scaleIndex = (int8_t)(smulbb((int32_t)c->LTPredCodGain, round_loss) > orc_silk_log2lin(2900 - c->SNR_dB_Q7));
      c->LTP_scaleIndex = (int8_t)(c->LTP_scaleIndex + (smulbb((int32_t)c->LTPredCodGain, round_loss) > orc_silk_log2lin(3900 - c->SNR_dB_Q7)));
    } else c->LTP_scaleIndex = 0;
    { static const int16_t scales[3] = {15565, 12288, 8192}; c->LTP_scale = (float)scales[c->LTP_scaleIndex] / 16384.0f; }
    { /* silk_LTP_analysis_filter_FLP */
      const float *x_ptr = x - order;
      float *res = LPC_in_pre;
      for (k = 0; k < nb_subfr; k++) {
        const float *x_lag_ptr = x_ptr - c->pitchL[k], inv_gain = invGains[k], *B = &c->LTPCoef[k * 5];
        for (i = 0; i < subfr_length + order; i++) {
          res[i] = x_ptr[i];
          for (j = 0; j < 5; j++) res[i] -= B[j] * x_lag_ptr[5 / 2 - j];
          res[i] *= inv_gain;
          x_lag_ptr++;
        }
        res += subfr_length + order;
        x_ptr += subfr_length;
      }
    }
  } else {
    const float *x_ptr = x - order;
    float *pre = LPC_in_pre;
    for (i = 0; i < nb_subfr; i++) {
      for (j = 0; j < subfr_length + order; j++) pre[j] = invGains[i] * x_ptr[j];
      pre += subfr_length + order;
      x_ptr += subfr_length;
    }
    memset(c->LTPCoef, 0, nb_subfr * 5 * sizeof(float));
    c->LTPredCodGain = 0.0f;
    c->sum_log_gain_Q7 = 0;
  }
  if (c->first_frame_after_reset) minInvGain = 1.0f / 1e2f;
  else {
    minInvGain = (float)pow(2, c->LTPredCodGain / 3) / 1e4f;
    minInvGain /= 0.25f + 0.75f * c->coding_quality;
  }
  {
    OrcSilkFindLpc f;
    OrcSilkNlsf q;
    memset(&f, 0, sizeof(f));
    memset(&q, 0, sizeof(q));
    memcpy(f.prev_NLSFq_Q15, c->prev_NLSFq_Q15, sizeof(f.prev_NLSFq_Q15));
    f.minInvGain = minInvGain; f.useInterpolatedNLSFs = c->useInterpolatedNLSFs; f.first_frame_after_reset = c->first_frame_after_reset;
    orc_silk_find_LPC(&f, LPC_in_pre, order, nb_subfr, subfr_length);
    memcpy(NLSF_Q15, f.NLSF_Q15, order * sizeof(int16_t));
    c->NLSFInterpCoef_Q2 = (int8_t)f.NLSFInterpCoef_Q2;
    memcpy(q.NLSF_Q15, NLSF_Q15, sizeof(q.NLSF_Q15));
    memcpy(q.prev_NLSFq_Q15, c->prev_NLSFq_Q15, sizeof(q.prev_NLSFq_Q15));
    q.NLSFInterpCoef_Q2 = c->NLSFInterpCoef_Q2; q.signalType = (int8_t)c->signalType; q.useInterpolatedNLSFs = (int8_t)c->useInterpolatedNLSFs;
    q.speech_activity_Q8 = c->speech_activity_Q8; q.NLSF_MSVQ_Survivors = c->NLSF_MSVQ_Survivors;
    orc_silk_process_NLSFs(&q, order == 16, nb_subfr);
    memcpy(NLSF_Q15, q.NLSF_Q15, order * sizeof(int16_t));
    memcpy(c->NLSFIndices, q.NLSFIndices, 17);
    for (j = 0; j < 2; j++)
      for (i = 0; i < order; i++) c->PredCoef[j][i] = (float)q.PredCoef_Q12[j][i] * (1.0f / 4096.0f);
  }
  { /* silk_residual_energy_FLP */
    const int shift = order + subfr_length;
    float LPC_res[(320 + 4 * 16) / 2];
    lpc_analysis_filter_flp(LPC_res, c->PredCoef[0], LPC_in_pre, 2 * shift, order);
    c->ResNrg[0] = (float)(c->Gains[0] * c->Gains[0] * energy_flp(LPC_res + order, subfr_length));
    c->ResNrg[1] = (float)(c->Gains[1] * c->Gains[1] * energy_flp(LPC_res + order + shift, subfr_length));
    if (nb_subfr == 4) {
      lpc_analysis_filter_flp(LPC_res, c->PredCoef[1], LPC_in_pre + 2 * shift, 2 * shift, order);
      c->ResNrg[2] = (float)(c->Gains[2] * c->Gains[2] * energy_flp(LPC_res + order, subfr_length));
      c->ResNrg[3] = (float)(c->Gains[3] * c->Gains[3] * energy_flp(LPC_res + order + shift, subfr_length));
    }
  }
  memcpy(c->prev_NLSFq_Q15, NLSF_Q15, sizeof(c->prev_NLSFq_Q15));
}
