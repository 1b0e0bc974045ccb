// silk_nlsf.cuh — the NLSF processing of a SILK frame, one THREAD per stream: silk_process_NLSFs (silk/process_NLSFs.c:35-107) with
// silk_NLSF_VQ_weights_laroia (NLSF_VQ_weights_laroia.c:41-80), silk_interpolate (interpolate.c:35-51), silk_NLSF_encode (NLSF_encode.c:38-124: first-stage
// VQ over the 32 codebook vectors, NLSF_VQ.c:35-76; survivors sorted by silk_insertion_sort_increasing, sort.c:40-82; per survivor the delayed-decision
// trellis quantiser of the residual, NLSF_del_dec_quant.c:35-215), silk_NLSF_decode (NLSF_decode.c:35-93), silk_NLSF_unpack (NLSF_unpack.c:35-54),
// silk_NLSF_stabilize (NLSF_stabilize.c:46-142) and silk_NLSF2A (NLSF2A.c:45-140) with silk_LPC_fit (LPC_fit.c:36-82), silk_LPC_inverse_pred_gain
// (LPC_inv_pred_gain.c:43-141) and silk_bwexpander_32 (bwexpander_32.c:36-50).  Everything is serial integer arithmetic on one 10 / 16 coefficient vector
// (a survivor search with data-dependent trellis decisions), so the parallel axis is the batch, as in the other index-producing kernels; the tables are
// the generated silk_tables_data.h.  Integer work: bit-exact.
#pragma once
#include "simt.cuh"
#include "silk_ops.cuh"   // silk_smulwb, silk_lin2log
#include "silk_tables_data.h"

namespace ob {
constexpr int kNlsfMaxOrder = 16, kNlsfStates = 4, kNlsfAmp = 4, kNlsfAmpExt = 10;
struct ObSilkNlsf {   // == OpusB200SilkNlsf (include/opus_b200.h): one stream's view of silk_encoder_state for silk_process_NLSFs
  int16_t NLSF_Q15[16];          // in: unquantised, out: quantised
  int16_t prev_NLSFq_Q15[16];
  int16_t PredCoef_Q12[2][16];   // out
  int8_t NLSFIndices[17];        // out
  int8_t NLSFInterpCoef_Q2, signalType, useInterpolatedNLSFs;
  int32_t speech_activity_Q8, NLSF_MSVQ_Survivors;
};
struct SilkNlsfParams { ObSilkNlsf *frames; int n, wb, nb_subfr; };

OB_DEV int32_t nl_smulbb(int32_t a, int32_t b) { return (int32_t)(int16_t)a * (int32_t)(int16_t)b; }
OB_DEV int32_t nl_smmul(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * b) >> 32); }
OB_DEV int32_t nl_smulww(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * b) >> 16); }
OB_DEV int32_t nl_rr(int32_t a, int s) { return s == 1 ? (a >> 1) + (a & 1) : ((a >> (s - 1)) + 1) >> 1; }
OB_DEV int64_t nl_rr64(int64_t a, int s) { return s == 1 ? (a >> 1) + (a & 1) : ((a >> (s - 1)) + 1) >> 1; }
OB_DEV int32_t nl_abs(int32_t a) { return a < 0 ? -a : a; }
OB_DEV int32_t nl_lshift_sat32(int32_t a, int s) {
  const int32_t lo = INT32_MIN >> s, hi = INT32_MAX >> s;
  a = a < lo ? lo : a > hi ? hi : a;
  return (int32_t)((uint32_t)a << s);
}
OB_DEV int32_t nl_sub_sat32(int32_t a, int32_t b) {
  const int64_t d = (int64_t)a - b;
  return d > INT32_MAX ? INT32_MAX : d < INT32_MIN ? INT32_MIN : (int32_t)d;
}
// silk_DIV32_varQ / silk_INVERSE32_varQ (silk/Inlines.h:93-178)
OB_DEV int32_t nl_div32_varQ(int32_t a32, int32_t b32, int Qres) {
  const int a_headrm = clz32((uint32_t)nl_abs(a32)) - 1, b_headrm = clz32((uint32_t)nl_abs(b32)) - 1;
  int32_t a_nrm = (int32_t)((uint32_t)a32 << a_headrm);
  const int32_t b_nrm = (int32_t)((uint32_t)b32 << b_headrm), b_inv = (INT32_MAX >> 2) / (b_nrm >> 16);
  int32_t result = silk_smulwb(a_nrm, b_inv);
  a_nrm = (int32_t)((uint32_t)a_nrm - ((uint32_t)nl_smmul(b_nrm, result) << 3));
  result = result + silk_smulwb(a_nrm, b_inv);
  const int lshift = 29 + a_headrm - b_headrm - Qres;
  if (lshift < 0) return nl_lshift_sat32(result, -lshift);
  return lshift < 32 ? result >> lshift : 0;
}
OB_DEV int32_t nl_inverse32_varQ(int32_t b32, int Qres) {
  const int b_headrm = clz32((uint32_t)nl_abs(b32)) - 1;
  const int32_t b_nrm = (int32_t)((uint32_t)b32 << b_headrm), b_inv = (INT32_MAX >> 2) / (b_nrm >> 16);
  int32_t result = (int32_t)((uint32_t)b_inv << 16);
  const int32_t err_Q32 = (int32_t)((uint32_t)(((int32_t)1 << 29) - silk_smulwb(b_nrm, b_inv)) << 3);
  result = (int32_t)(result + (((int64_t)err_Q32 * b_inv) >> 16));
  const int lshift = 61 - b_headrm - Qres;
  if (lshift <= 0) return nl_lshift_sat32(result, -lshift);
  return lshift < 32 ? result >> lshift : 0;
}
struct NlsfCB {   // silk_NLSF_CB_struct (silk/structs.h:84-97) of silk_NLSF_CB_NB_MB / silk_NLSF_CB_WB
  int order, quantStepSize_Q16, invQuantStepSize_Q6;
  const unsigned char *CB1_NLSF_Q8;
  const short *CB1_Wght_Q9;
  const unsigned char *CB1_iCDF, *pred_Q8, *ec_sel, *ec_Rates_Q5;
  const short *deltaMin_Q15;
};
OB_DEV NlsfCB nlsf_cb(int wb) {
  NlsfCB cb;
  cb.order = wb ? 16 : 10;
  cb.quantStepSize_Q16 = wb ? 9830 : 11796;    // SILK_FIX_CONST(0.15 / 0.18, 16)
  cb.invQuantStepSize_Q6 = wb ? 427 : 356;     // SILK_FIX_CONST(1 / 0.15, 6), (1 / 0.18, 6)
  cb.CB1_NLSF_Q8 = ob_silk_NLSF_CB1_Q8 + (wb ? 320 : 0);
  cb.CB1_Wght_Q9 = ob_silk_NLSF_CB1_Wght_Q9 + (wb ? 320 : 0);
  cb.CB1_iCDF = ob_silk_NLSF_CB1_iCDF + (wb ? 64 : 0);
  cb.pred_Q8 = ob_silk_NLSF_PRED_Q8 + (wb ? 18 : 0);
  cb.ec_sel = ob_silk_NLSF_CB2_SELECT + (wb ? 160 : 0);
  cb.ec_Rates_Q5 = ob_silk_NLSF_CB2_BITS_Q5 + (wb ? 72 : 0);
  cb.deltaMin_Q15 = ob_silk_NLSF_DELTA_MIN_Q15 + (wb ? 11 : 0);
  return cb;
}
// silk_insertion_sort_increasing: the K smallest of a[0..L) in order, with their indices
OB_DEV void nlsf_sort_increasing(int32_t *a, int *idx, int L, int K) {
  for (int i = 0; i < K; i++) idx[i] = i;
  for (int i = 1; i < K; i++) {
    const int32_t value = a[i];
    int j;
    for (j = i - 1; j >= 0 && value < a[j]; j--) { a[j + 1] = a[j]; idx[j + 1] = idx[j]; }
    a[j + 1] = value;
    idx[j + 1] = i;
  }
  for (int i = K; i < L; i++) {
    const int32_t value = a[i];
    if (value < a[K - 1]) {
      int j;
      for (j = K - 2; j >= 0 && value < a[j]; j--) { a[j + 1] = a[j]; idx[j + 1] = idx[j]; }
      a[j + 1] = value;
      idx[j + 1] = i;
    }
  }
}
OB_DEV void nlsf_stabilize(int16_t *NLSF_Q15, const short *NDeltaMin_Q15, int L) {
  for (int loops = 0; loops < 20; loops++) {
    int32_t min_diff = NLSF_Q15[0] - NDeltaMin_Q15[0], diff;
    int I = 0;
    for (int i = 1; i <= L - 1; i++) {
      diff = NLSF_Q15[i] - (NLSF_Q15[i - 1] + NDeltaMin_Q15[i]);
      if (diff < min_diff) { min_diff = diff; I = i; }
    }
    diff = (1 << 15) - (NLSF_Q15[L - 1] + NDeltaMin_Q15[L]);
    if (diff < min_diff) { min_diff = diff; I = L; }
    if (min_diff >= 0) return;
    if (I == 0) NLSF_Q15[0] = NDeltaMin_Q15[0];
    else if (I == L) NLSF_Q15[L - 1] = (int16_t)((1 << 15) - NDeltaMin_Q15[L]);
    else {
      int32_t min_center = 0, max_center = 1 << 15;
      for (int k = 0; k < I; k++) min_center += NDeltaMin_Q15[k];
      min_center += NDeltaMin_Q15[I] >> 1;
      for (int k = L; k > I; k--) max_center -= NDeltaMin_Q15[k];
      max_center -= NDeltaMin_Q15[I] >> 1;
      int32_t c = nl_rr((int32_t)NLSF_Q15[I - 1] + (int32_t)NLSF_Q15[I], 1);
      if (min_center > max_center) c = c > min_center ? min_center : c < max_center ? max_center : c;   // silk_LIMIT_32 orders its bounds itself
      else c = c > max_center ? max_center : c < min_center ? min_center : c;
      NLSF_Q15[I - 1] = (int16_t)((int16_t)c - (NDeltaMin_Q15[I] >> 1));
      NLSF_Q15[I] = (int16_t)(NLSF_Q15[I - 1] + NDeltaMin_Q15[I]);
    }
  }
  // fall back: sort, then force the spacings (:112-141)
  for (int i = 1; i < L; i++) {
    const int value = NLSF_Q15[i];
    int j;
    for (j = i - 1; j >= 0 && value < NLSF_Q15[j]; j--) NLSF_Q15[j + 1] = NLSF_Q15[j];
    NLSF_Q15[j + 1] = (int16_t)value;
  }
  NLSF_Q15[0] = NLSF_Q15[0] > NDeltaMin_Q15[0] ? NLSF_Q15[0] : NDeltaMin_Q15[0];
  for (int i = 1; i < L; i++) {
    int32_t s = (int32_t)NLSF_Q15[i - 1] + NDeltaMin_Q15[i];
    s = s > 32767 ? 32767 : s < -32768 ? -32768 : s;
    NLSF_Q15[i] = (int16_t)(NLSF_Q15[i] > s ? NLSF_Q15[i] : s);
  }
  { const int v = (1 << 15) - NDeltaMin_Q15[L]; NLSF_Q15[L - 1] = (int16_t)(NLSF_Q15[L - 1] < v ? NLSF_Q15[L - 1] : v); }
  for (int i = L - 2; i >= 0; i--) { const int v = NLSF_Q15[i + 1] - NDeltaMin_Q15[i + 1]; NLSF_Q15[i] = (int16_t)(NLSF_Q15[i] < v ? NLSF_Q15[i] : v); }
}
OB_DEV int32_t nlsf_w(int32_t d) { return ((int32_t)1 << 17) / (d > 1 ? d : 1); }   // 1 << (15 + NLSF_W_Q) over the clamped spacing
OB_DEV void nlsf_weights_laroia(int16_t *pW, const int16_t *x, int D) {
  int32_t t1 = nlsf_w(x[0]), t2 = nlsf_w(x[1] - x[0]);
  pW[0] = (int16_t)(t1 + t2 < 32767 ? t1 + t2 : 32767);
  for (int k = 1; k < D - 1; k += 2) {
    t1 = nlsf_w(x[k + 1] - x[k]);
    pW[k] = (int16_t)(t1 + t2 < 32767 ? t1 + t2 : 32767);
    t2 = nlsf_w(x[k + 2] - x[k + 1]);
    pW[k + 1] = (int16_t)(t1 + t2 < 32767 ? t1 + t2 : 32767);
  }
  t1 = nlsf_w((1 << 15) - x[D - 1]);
  pW[D - 1] = (int16_t)(t1 + t2 < 32767 ? t1 + t2 : 32767);
}
OB_DEV void nlsf_unpack(int16_t *ec_ix, unsigned char *pred_Q8, const NlsfCB &cb, int CB1_index) {
  const unsigned char *sel = &cb.ec_sel[CB1_index * cb.order / 2];
  for (int i = 0; i < cb.order; i += 2) {
    const unsigned char entry = *sel++;
    ec_ix[i] = (int16_t)(((entry >> 1) & 7) * 9);
    pred_Q8[i] = cb.pred_Q8[i + (entry & 1) * (cb.order - 1)];
    ec_ix[i + 1] = (int16_t)(((entry >> 5) & 7) * 9);
    pred_Q8[i + 1] = cb.pred_Q8[i + ((entry >> 4) & 1) * (cb.order - 1) + 1];
  }
}
OB_DEV void nlsf_decode(int16_t *pNLSF_Q15, const int8_t *NLSFIndices, const NlsfCB &cb) {
  unsigned char pred_Q8[kNlsfMaxOrder];
  int16_t ec_ix[kNlsfMaxOrder], res_Q10[kNlsfMaxOrder];
  const unsigned char *el = &cb.CB1_NLSF_Q8[NLSFIndices[0] * cb.order];
  const short *wq = &cb.CB1_Wght_Q9[NLSFIndices[0] * cb.order];
  nlsf_unpack(ec_ix, pred_Q8, cb, NLSFIndices[0]);
  int out_Q10 = 0;
  for (int i = cb.order - 1; i >= 0; i--) {   // silk_NLSF_residual_dequant
    const int pred_Q10 = nl_smulbb(out_Q10, (int16_t)pred_Q8[i]) >> 8;
    out_Q10 = (int)((uint32_t)(int32_t)NLSFIndices[1 + i] << 10);
    if (out_Q10 > 0) out_Q10 -= 102;        // SILK_FIX_CONST(NLSF_QUANT_LEVEL_ADJ, 10)
    else if (out_Q10 < 0) out_Q10 += 102;
    out_Q10 = pred_Q10 + silk_smulwb(out_Q10, cb.quantStepSize_Q16);
    res_Q10[i] = (int16_t)out_Q10;
  }
  for (int i = 0; i < cb.order; i++) {
    const int32_t t = ((int32_t)((uint32_t)(int32_t)res_Q10[i] << 14)) / wq[i] + ((int32_t)(int16_t)el[i] << 7);
    pNLSF_Q15[i] = (int16_t)(t < 0 ? 0 : t > 32767 ? 32767 : t);
  }
  nlsf_stabilize(pNLSF_Q15, cb.deltaMin_Q15, cb.order);
}
OB_DEV int32_t nlsf_del_dec_quant(int8_t *indices, const int16_t *x_Q10, const int16_t *w_Q5, const unsigned char *pred_coef_Q8, const int16_t *ec_ix,
                                    const unsigned char *ec_rates_Q5, int quant_step_size_Q16, int16_t inv_quant_step_size_Q6, int32_t mu_Q20, int order) {
  int ind_sort[kNlsfStates];
  int8_t ind[kNlsfStates][kNlsfMaxOrder];
  int16_t prev_out_Q10[2 * kNlsfStates];
  int32_t RD_Q25[2 * kNlsfStates], RD_min_Q25[kNlsfStates], RD_max_Q25[kNlsfStates];
  int16_t out0_table[2 * kNlsfAmpExt], out1_table[2 * kNlsfAmpExt];
  for (int i = -kNlsfAmpExt; i <= kNlsfAmpExt - 1; i++) {
    int16_t o0 = (int16_t)(i * 1024), o1 = (int16_t)(o0 + 1024);
    if (i > 0) { o0 = (int16_t)(o0 - 102); o1 = (int16_t)(o1 - 102); }
    else if (i == 0) o1 = (int16_t)(o1 - 102);
    else if (i == -1) o0 = (int16_t)(o0 + 102);
    else { o0 = (int16_t)(o0 + 102); o1 = (int16_t)(o1 + 102); }
    out0_table[i + kNlsfAmpExt] = (int16_t)(nl_smulbb(o0, quant_step_size_Q16) >> 16);
    out1_table[i + kNlsfAmpExt] = (int16_t)(nl_smulbb(o1, quant_step_size_Q16) >> 16);
  }
  int nStates = 1;
  RD_Q25[0] = 0;
  prev_out_Q10[0] = 0;
  for (int i = order - 1; i >= 0; i--) {
    const unsigned char *rates_Q5 = &ec_rates_Q5[ec_ix[i]];
    const int in_Q10 = x_Q10[i];
    for (int j = 0; j < nStates; j++) {
      const int pred_Q10 = nl_smulbb((int16_t)pred_coef_Q8[i], prev_out_Q10[j]) >> 8;
      const int res_Q10 = in_Q10 - pred_Q10;
      int ind_tmp = nl_smulbb(inv_quant_step_size_Q6, res_Q10) >> 16;
      ind_tmp = ind_tmp < -kNlsfAmpExt ? -kNlsfAmpExt : ind_tmp > kNlsfAmpExt - 1 ? kNlsfAmpExt - 1 : ind_tmp;
      ind[j][i] = (int8_t)ind_tmp;
      const int16_t out0_Q10 = (int16_t)(out0_table[ind_tmp + kNlsfAmpExt] + pred_Q10), out1_Q10 = (int16_t)(out1_table[ind_tmp + kNlsfAmpExt] + pred_Q10);
      prev_out_Q10[j] = out0_Q10;
      prev_out_Q10[j + nStates] = out1_Q10;
      int rate0_Q5, rate1_Q5;
      if (ind_tmp + 1 >= kNlsfAmp) {
        if (ind_tmp + 1 == kNlsfAmp) { rate0_Q5 = rates_Q5[ind_tmp + kNlsfAmp]; rate1_Q5 = 280; }
        else { rate0_Q5 = (280 - 43 * kNlsfAmp) + nl_smulbb(43, ind_tmp); rate1_Q5 = rate0_Q5 + 43; }
      } else if (ind_tmp <= -kNlsfAmp) {
        if (ind_tmp == -kNlsfAmp) { rate0_Q5 = 280; rate1_Q5 = rates_Q5[ind_tmp + 1 + kNlsfAmp]; }
        else { rate0_Q5 = (280 - 43 * kNlsfAmp) + nl_smulbb(-43, ind_tmp); rate1_Q5 = rate0_Q5 - 43; }
      } else { rate0_Q5 = rates_Q5[ind_tmp + kNlsfAmp]; rate1_Q5 = rates_Q5[ind_tmp + 1 + kNlsfAmp]; }
      const int32_t RD_tmp = RD_Q25[j];
      int diff_Q10 = in_Q10 - out0_Q10;
      RD_Q25[j] = (RD_tmp + nl_smulbb(diff_Q10, diff_Q10) * w_Q5[i]) + nl_smulbb(mu_Q20, rate0_Q5);
      diff_Q10 = in_Q10 - out1_Q10;
      RD_Q25[j + nStates] = (RD_tmp + nl_smulbb(diff_Q10, diff_Q10) * w_Q5[i]) + nl_smulbb(mu_Q20, rate1_Q5);
    }
    if (nStates <= kNlsfStates / 2) {
      for (int j = 0; j < nStates; j++) ind[j + nStates][i] = (int8_t)(ind[j][i] + 1);
      nStates <<= 1;
      for (int j = nStates; j < kNlsfStates; j++) ind[j][i] = ind[j - nStates][i];
    } else {
      for (int j = 0; j < kNlsfStates; j++) {   // sort lower and upper half of RD_Q25, pairwise
        if (RD_Q25[j] > RD_Q25[j + kNlsfStates]) {
          RD_max_Q25[j] = RD_Q25[j];
          RD_min_Q25[j] = RD_Q25[j + kNlsfStates];
          RD_Q25[j] = RD_min_Q25[j];
          RD_Q25[j + kNlsfStates] = RD_max_Q25[j];
          const int16_t t = prev_out_Q10[j];
          prev_out_Q10[j] = prev_out_Q10[j + kNlsfStates];
          prev_out_Q10[j + kNlsfStates] = t;
          ind_sort[j] = j + kNlsfStates;
        } else {
          RD_min_Q25[j] = RD_Q25[j];
          RD_max_Q25[j] = RD_Q25[j + kNlsfStates];
          ind_sort[j] = j;
        }
      }
      while (true) {   // the best of the losing half replaces the worst of the winning half while that improves things
        int32_t min_max_Q25 = INT32_MAX, max_min_Q25 = 0;
        int ind_min_max = 0, ind_max_min = 0;
        for (int j = 0; j < kNlsfStates; j++) {
          if (min_max_Q25 > RD_max_Q25[j]) { min_max_Q25 = RD_max_Q25[j]; ind_min_max = j; }
          if (max_min_Q25 < RD_min_Q25[j]) { max_min_Q25 = RD_min_Q25[j]; ind_max_min = j; }
        }
        if (min_max_Q25 >= max_min_Q25) break;
        ind_sort[ind_max_min] = ind_sort[ind_min_max] ^ kNlsfStates;
        RD_Q25[ind_max_min] = RD_Q25[ind_min_max + kNlsfStates];
        prev_out_Q10[ind_max_min] = prev_out_Q10[ind_min_max + kNlsfStates];
        RD_min_Q25[ind_max_min] = 0;
        RD_max_Q25[ind_min_max] = INT32_MAX;
        for (int k = 0; k < kNlsfMaxOrder; k++) ind[ind_max_min][k] = ind[ind_min_max][k];
      }
      for (int j = 0; j < kNlsfStates; j++) ind[j][i] = (int8_t)(ind[j][i] + (ind_sort[j] >> 2));
    }
  }
  int ind_tmp = 0;
  int32_t min_Q25 = INT32_MAX;
  for (int j = 0; j < 2 * kNlsfStates; j++) if (min_Q25 > RD_Q25[j]) { min_Q25 = RD_Q25[j]; ind_tmp = j; }
  for (int j = 0; j < order; j++) indices[j] = ind[ind_tmp & (kNlsfStates - 1)][j];
  indices[0] = (int8_t)(indices[0] + (ind_tmp >> 2));
  return min_Q25;
}
OB_DEV void nlsf_encode(int8_t *NLSFIndices, int16_t *pNLSF_Q15, const NlsfCB &cb, const int16_t *pW_Q2, int NLSF_mu_Q20, int nSurvivors, int signalType) {
  int32_t err_Q24[32], RD_Q25[32];
  int tempIndices1[32];
  int8_t tempIndices2[32 * kNlsfMaxOrder];
  int16_t res_Q10[kNlsfMaxOrder], W_adj_Q5[kNlsfMaxOrder], ec_ix[kNlsfMaxOrder];
  unsigned char pred_Q8[kNlsfMaxOrder];
  nlsf_stabilize(pNLSF_Q15, cb.deltaMin_Q15, cb.order);
  for (int i = 0; i < 32; i++) {   // silk_NLSF_VQ
    const unsigned char *c = &cb.CB1_NLSF_Q8[i * cb.order];
    const short *w = &cb.CB1_Wght_Q9[i * cb.order];
    int32_t sum = 0, pred = 0;
    for (int m = cb.order - 2; m >= 0; m -= 2) {
      int32_t dw = nl_smulbb(pNLSF_Q15[m + 1] - ((int32_t)c[m + 1] << 7), w[m + 1]);
      sum = sum + nl_abs(dw - (pred >> 1));
      pred = dw;
      dw = nl_smulbb(pNLSF_Q15[m] - ((int32_t)c[m] << 7), w[m]);
      sum = sum + nl_abs(dw - (pred >> 1));
      pred = dw;
    }
    err_Q24[i] = sum;
  }
  nlsf_sort_increasing(err_Q24, tempIndices1, 32, nSurvivors);
  for (int s = 0; s < nSurvivors; s++) {
    const int ind1 = tempIndices1[s];
    const unsigned char *el = &cb.CB1_NLSF_Q8[ind1 * cb.order];
    const short *wq = &cb.CB1_Wght_Q9[ind1 * cb.order];
    for (int i = 0; i < cb.order; i++) {
      const int32_t W_tmp_Q9 = wq[i];
      const int16_t tmp_Q15 = (int16_t)((int16_t)el[i] << 7);
      res_Q10[i] = (int16_t)(nl_smulbb(pNLSF_Q15[i] - tmp_Q15, W_tmp_Q9) >> 14);
      W_adj_Q5[i] = (int16_t)nl_div32_varQ((int32_t)pW_Q2[i], nl_smulbb(W_tmp_Q9, W_tmp_Q9), 21);
    }
    nlsf_unpack(ec_ix, pred_Q8, cb, ind1);
    RD_Q25[s] = nlsf_del_dec_quant(&tempIndices2[s * kNlsfMaxOrder], res_Q10, W_adj_Q5, pred_Q8, ec_ix, cb.ec_Rates_Q5, cb.quantStepSize_Q16,
                                   (int16_t)cb.invQuantStepSize_Q6, NLSF_mu_Q20, cb.order);
    const unsigned char *iCDF_ptr = &cb.CB1_iCDF[(signalType >> 1) * 32];
    const int prob_Q8 = ind1 == 0 ? 256 - iCDF_ptr[ind1] : iCDF_ptr[ind1 - 1] - iCDF_ptr[ind1];
    const int bits_q7 = (8 << 7) - silk_lin2log(prob_Q8);
    RD_Q25[s] = RD_Q25[s] + nl_smulbb(bits_q7, NLSF_mu_Q20 >> 2);
  }
  int bestIndex;
  nlsf_sort_increasing(RD_Q25, &bestIndex, nSurvivors, 1);
  NLSFIndices[0] = (int8_t)tempIndices1[bestIndex];
  for (int i = 0; i < cb.order; i++) NLSFIndices[1 + i] = tempIndices2[bestIndex * kNlsfMaxOrder + i];
  nlsf_decode(pNLSF_Q15, NLSFIndices, cb);
}
OB_DEV void nlsf_bwexpander_32(int32_t *ar, int d, int32_t chirp_Q16) {
  const int32_t chirp_minus_one_Q16 = chirp_Q16 - 65536;
  for (int i = 0; i < d - 1; i++) {
    ar[i] = nl_smulww(chirp_Q16, ar[i]);
    chirp_Q16 += nl_rr(chirp_Q16 * chirp_minus_one_Q16, 16);
  }
  ar[d - 1] = nl_smulww(chirp_Q16, ar[d - 1]);
}
OB_DEV int32_t nlsf_lpc_inverse_pred_gain(const int16_t *A_Q12, int order) {
  const int QA = 24;
  const int32_t A_LIMIT = 16773022;   // SILK_FIX_CONST(0.99975, 24)
  int32_t A_QA[kNlsfMaxOrder], DC_resp = 0, invGain_Q30 = 1 << 30;
  for (int k = 0; k < order; k++) { DC_resp += A_Q12[k]; A_QA[k] = (int32_t)((uint32_t)(int32_t)A_Q12[k] << (QA - 12)); }
  if (DC_resp >= 4096) return 0;
  int k;
  for (k = order - 1; k > 0; k--) {
    if (A_QA[k] > A_LIMIT || A_QA[k] < -A_LIMIT) return 0;
    const int32_t rc_Q31 = -(int32_t)((uint32_t)A_QA[k] << (31 - QA)), rc_mult1_Q30 = (1 << 30) - nl_smmul(rc_Q31, rc_Q31);
    invGain_Q30 = (int32_t)((uint32_t)nl_smmul(invGain_Q30, rc_mult1_Q30) << 2);
    if (invGain_Q30 < 107374) return 0;   // SILK_FIX_CONST(1.0f / MAX_PREDICTION_POWER_GAIN, 30)
    const int mult2Q = 32 - clz32((uint32_t)nl_abs(rc_mult1_Q30));
    const int32_t rc_mult2 = nl_inverse32_varQ(rc_mult1_Q30, mult2Q + 30);
    for (int n = 0; n < (k + 1) >> 1; n++) {
      const int32_t tmp1 = A_QA[n], tmp2 = A_QA[k - n - 1];
      int64_t tmp64 = nl_rr64((int64_t)nl_sub_sat32(tmp1, (int32_t)nl_rr64((int64_t)tmp2 * rc_Q31, 31)) * rc_mult2, mult2Q);
      if (tmp64 > INT32_MAX || tmp64 < INT32_MIN) return 0;
      A_QA[n] = (int32_t)tmp64;
      tmp64 = nl_rr64((int64_t)nl_sub_sat32(tmp2, (int32_t)nl_rr64((int64_t)tmp1 * rc_Q31, 31)) * rc_mult2, mult2Q);
      if (tmp64 > INT32_MAX || tmp64 < INT32_MIN) return 0;
      A_QA[k - n - 1] = (int32_t)tmp64;
    }
  }
  if (A_QA[k] > A_LIMIT || A_QA[k] < -A_LIMIT) return 0;
  const int32_t rc_Q31 = -(int32_t)((uint32_t)A_QA[0] << (31 - QA)), rc_mult1_Q30 = (1 << 30) - nl_smmul(rc_Q31, rc_Q31);
  invGain_Q30 = (int32_t)((uint32_t)nl_smmul(invGain_Q30, rc_mult1_Q30) << 2);
  if (invGain_Q30 < 107374) return 0;
  return invGain_Q30;
}
OB_DEV void nlsf_find_poly(int32_t *out, const int32_t *cLSF, int dd) {
  const int QA = 16;
  out[0] = 1 << QA;
  out[1] = -cLSF[0];
  for (int k = 1; k < dd; k++) {
    const int32_t ftmp = cLSF[2 * k];
    out[k + 1] = (int32_t)((uint32_t)out[k - 1] << 1) - (int32_t)nl_rr64((int64_t)ftmp * out[k], QA);
    for (int n = k; n > 1; n--) out[n] += out[n - 2] - (int32_t)nl_rr64((int64_t)ftmp * out[n - 1], QA);
    out[1] -= ftmp;
  }
}
OB_DEV void nlsf_NLSF2A(int16_t *a_Q12, const int16_t *NLSF, int d) {
  const unsigned char ordering16[16] = {0, 15, 8, 7, 4, 11, 12, 3, 2, 13, 10, 5, 6, 9, 14, 1}, ordering10[10] = {0, 9, 6, 3, 4, 5, 8, 1, 2, 7};
  const int QA = 16, dd = d >> 1;
  int32_t cos_LSF_QA[kNlsfMaxOrder], P[kNlsfMaxOrder / 2 + 1], Q[kNlsfMaxOrder / 2 + 1], a32_QA1[kNlsfMaxOrder];
  for (int k = 0; k < d; k++) {
    const int32_t f_int = NLSF[k] >> (15 - 7), f_frac = NLSF[k] - (f_int << (15 - 7));
    const int32_t cos_val = ob_silk_LSFCosTab_Q12[f_int], delta = ob_silk_LSFCosTab_Q12[f_int + 1] - cos_val;
    cos_LSF_QA[d == 16 ? ordering16[k] : ordering10[k]] = nl_rr((int32_t)((uint32_t)cos_val << 8) + delta * f_frac, 20 - QA);
  }
  nlsf_find_poly(P, &cos_LSF_QA[0], dd);
  nlsf_find_poly(Q, &cos_LSF_QA[1], dd);
  for (int k = 0; k < dd; k++) {
    const int32_t Ptmp = P[k + 1] + P[k], Qtmp = Q[k + 1] - Q[k];
    a32_QA1[k] = -Qtmp - Ptmp;
    a32_QA1[d - k - 1] = Qtmp - Ptmp;
  }
  {   // silk_LPC_fit(a_Q12, a32_QA1, 12, QA + 1, d)
    const int sh = QA + 1 - 12;
    int i, idx = 0;
    for (i = 0; i < 10; i++) {
      int32_t maxabs = 0;
      for (int k = 0; k < d; k++) {
        const int32_t absval = nl_abs(a32_QA1[k]);
        if (absval > maxabs) { maxabs = absval; idx = k; }
      }
      maxabs = nl_rr(maxabs, sh);
      if (maxabs > 32767) {
        maxabs = maxabs < 163838 ? maxabs : 163838;
        const int32_t chirp_Q16 = 65470 - ((int32_t)((uint32_t)(maxabs - 32767) << 14)) / ((maxabs * (idx + 1)) >> 2);   // SILK_FIX_CONST(0.999, 16)
        nlsf_bwexpander_32(a32_QA1, d, chirp_Q16);
      } else break;
    }
    if (i == 10) {
      for (int k = 0; k < d; k++) {
        const int32_t v = nl_rr(a32_QA1[k], sh);
        a_Q12[k] = (int16_t)(v > 32767 ? 32767 : v < -32768 ? -32768 : v);
        a32_QA1[k] = (int32_t)((uint32_t)(int32_t)a_Q12[k] << sh);
      }
    } else for (int k = 0; k < d; k++) a_Q12[k] = (int16_t)nl_rr(a32_QA1[k], sh);
  }
  for (int i = 0; nlsf_lpc_inverse_pred_gain(a_Q12, d) == 0 && i < 16; i++) {
    nlsf_bwexpander_32(a32_QA1, d, 65536 - (int32_t)((uint32_t)2 << i));
    for (int k = 0; k < d; k++) a_Q12[k] = (int16_t)nl_rr(a32_QA1[k], QA + 1 - 12);
  }
}
OB_DEV void nlsf_interpolate(int16_t *xi, const int16_t *x0, const int16_t *x1, int ifact_Q2, int d) {
  for (int i = 0; i < d; i++) xi[i] = (int16_t)(x0[i] + (nl_smulbb(x1[i] - x0[i], ifact_Q2) >> 2));
}
OB_DEV void silk_process_nlsfs_core(ObSilkNlsf &f, int wb, int nb_subfr_) {
  const NlsfCB cb = nlsf_cb(wb);
  const int order = cb.order;
  int16_t tmp0[kNlsfMaxOrder], W[kNlsfMaxOrder], W0[kNlsfMaxOrder];
  int NLSF_mu_Q20 = 3146 + silk_smulwb(-268434, f.speech_activity_Q8);   // SILK_FIX_CONST(0.003, 20), SILK_FIX_CONST(-0.001, 28)
  if (nb_subfr_ == 2) NLSF_mu_Q20 = NLSF_mu_Q20 + (NLSF_mu_Q20 >> 1);
  nlsf_weights_laroia(W, f.NLSF_Q15, order);
  const int doInterpolate = f.useInterpolatedNLSFs == 1 && f.NLSFInterpCoef_Q2 < 4;
  if (doInterpolate) {
    nlsf_interpolate(tmp0, f.prev_NLSFq_Q15, f.NLSF_Q15, f.NLSFInterpCoef_Q2, order);
    nlsf_weights_laroia(W0, tmp0, order);
    const int16_t i_sqr_Q15 = (int16_t)(nl_smulbb(f.NLSFInterpCoef_Q2, f.NLSFInterpCoef_Q2) << 11);
    for (int i = 0; i < order; i++) W[i] = (int16_t)((W[i] >> 1) + (nl_smulbb(W0[i], i_sqr_Q15) >> 16));
  }
  nlsf_encode(f.NLSFIndices, f.NLSF_Q15, cb, W, NLSF_mu_Q20, f.NLSF_MSVQ_Survivors, f.signalType);
  nlsf_NLSF2A(f.PredCoef_Q12[1], f.NLSF_Q15, order);
  if (doInterpolate) {
    nlsf_interpolate(tmp0, f.prev_NLSFq_Q15, f.NLSF_Q15, f.NLSFInterpCoef_Q2, order);
    nlsf_NLSF2A(f.PredCoef_Q12[0], tmp0, order);
  } else for (int i = 0; i < order; i++) f.PredCoef_Q12[0][i] = f.PredCoef_Q12[1][i];
}
OB_DEV void silk_process_nlsfs_thread(const SilkNlsfParams &p, int s) {
  ObSilkNlsf f = p.frames[s];
  silk_process_nlsfs_core(f, p.wb, p.nb_subfr);
  p.frames[s] = f;
}
// ---- LPC estimation of a SILK frame: silk_find_LPC_FLP (silk/float/find_LPC_FLP.c:36-105) — Burg over the frame, and with NLSF interpolation on, Burg over
// the second half, silk_A2NLSF (silk/A2NLSF.c:45-267), then the search over the interpolation factor: interpolated NLSFs -> silk_NLSF2A -> float LPC analysis
// filter over the first half (silk/float/LPC_analysis_filter_FLP.c) -> residual energy (silk/float/energy_FLP.c).  One thread per stream. ----
struct ObSilkFindLpc {   // == OpusB200SilkFindLpc
  int16_t NLSF_Q15[16];         // out
  int16_t prev_NLSFq_Q15[16];
  float minInvGain;
  int32_t useInterpolatedNLSFs, first_frame_after_reset;
  int32_t NLSFInterpCoef_Q2;    // out
};
struct SilkFindLpcParams { ObSilkFindLpc *frames; const float *x; int n, order, nb_subfr, subfr_length; };   // x [n][nb_subfr * (subfr_length + order)]
OB_DEV void a2nlsf_trans_poly(int32_t *p, int dd) {
  for (int k = 2; k <= dd; k++) {
    for (int n = dd; n > k; n--) p[n - 2] -= p[n];
    p[k - 2] -= (int32_t)((uint32_t)p[k] << 1);
  }
}
OB_DEV int32_t a2nlsf_eval_poly(const int32_t *p, int32_t x, int dd) {
  const int32_t x_Q16 = (int32_t)((uint32_t)x << 4);
  int32_t y32 = p[dd];
  for (int n = dd - 1; n >= 0; n--) y32 = (int32_t)(p[n] + (((int64_t)y32 * x_Q16) >> 16));   // silk_SMLAWW
  return y32;
}
OB_DEV void a2nlsf_init(const int32_t *a_Q16, int32_t *P, int32_t *Q, int dd) {
  P[dd] = 1 << 16;
  Q[dd] = 1 << 16;
  for (int k = 0; k < dd; k++) {
    P[k] = -a_Q16[dd - k - 1] - a_Q16[dd + k];
    Q[k] = -a_Q16[dd - k - 1] + a_Q16[dd + k];
  }
  for (int k = dd; k > 0; k--) {
    P[k - 1] -= P[k];
    Q[k - 1] += Q[k];
  }
  a2nlsf_trans_poly(P, dd);
  a2nlsf_trans_poly(Q, dd);
}
OB_DEV void nlsf_A2NLSF(int16_t *NLSF, int32_t *a_Q16, int d) {
  const int dd = d >> 1;
  int i = 0, k = 1, root_ix;
  int32_t xlo, xhi, ylo, yhi, thr = 0, P[kNlsfMaxOrder / 2 + 1], Q[kNlsfMaxOrder / 2 + 1];
  const int32_t *p = P;
  a2nlsf_init(a_Q16, P, Q, dd);
  xlo = ob_silk_LSFCosTab_Q12[0];
  ylo = a2nlsf_eval_poly(p, xlo, dd);
  if (ylo < 0) {
    NLSF[0] = 0;
    p = Q;
    ylo = a2nlsf_eval_poly(p, xlo, dd);
    root_ix = 1;
  } else root_ix = 0;
  while (true) {
    xhi = ob_silk_LSFCosTab_Q12[k];
    yhi = a2nlsf_eval_poly(p, xhi, dd);
    if ((ylo <= 0 && yhi >= thr) || (ylo >= 0 && yhi <= -thr)) {
      thr = yhi == 0 ? 1 : 0;
      int ffrac = -256;
      for (int m = 0; m < 3; m++) {   // BIN_DIV_STEPS_A2NLSF_FIX
        const int32_t xmid = nl_rr(xlo + xhi, 1), ymid = a2nlsf_eval_poly(p, xmid, dd);
        if ((ylo <= 0 && ymid >= 0) || (ylo >= 0 && ymid <= 0)) { xhi = xmid; yhi = ymid; }
        else { xlo = xmid; ylo = ymid; ffrac = ffrac + (128 >> m); }
      }
      if (nl_abs(ylo) < 65536) {
        const int32_t den = ylo - yhi, nom = (int32_t)((uint32_t)ylo << (8 - 3)) + (den >> 1);
        if (den != 0) ffrac += nom / den;
      } else ffrac += ylo / ((ylo - yhi) >> (8 - 3));
      const int32_t v = (int32_t)((uint32_t)k << 8) + ffrac;
      NLSF[root_ix] = (int16_t)(v < 32767 ? v : 32767);
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
      if (k > 128) {   // LSF_COS_TAB_SZ_FIX: not all roots found — expand the bandwidth and start again
        i++;
        if (i > 16) {   // MAX_ITERATIONS_A2NLSF_FIX: white spectrum
          NLSF[0] = (int16_t)((1 << 15) / (d + 1));
          for (k = 1; k < d; k++) NLSF[k] = (int16_t)(NLSF[k - 1] + NLSF[0]);
          return;
        }
        nlsf_bwexpander_32(a_Q16, d, 65536 - (int32_t)((uint32_t)1 << i));
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
OB_DEV void nlsf_a2nlsf_flp(int16_t *NLSF_Q15, const float *pAR, int order) {   // silk_A2NLSF_FLP: silk_float2int rounds to nearest even
  int32_t a_fix_Q16[kNlsfMaxOrder];
  for (int i = 0; i < order; i++) a_fix_Q16[i] = (int32_t)rintf(pAR[i] * 65536.0f);
  nlsf_A2NLSF(NLSF_Q15, a_fix_Q16, order);
}
OB_DEV void silk_find_lpc_core(ObSilkFindLpc &f, const float *x, int order, int nb_subfr, int subfr_length_in) {
  const int subfr_length = subfr_length_in + order;
  float a[kNlsfMaxOrder], a_tmp[kNlsfMaxOrder], nrg[1], LPC_res[2 * (80 + 16)];
  int16_t NLSF0_Q15[kNlsfMaxOrder], a_fix_Q12[kNlsfMaxOrder];
  f.NLSFInterpCoef_Q2 = 4;
  SilkBurgParams b;
  b.A = a; b.nrg = nrg; b.x = x; b.minInvGain = f.minInvGain; b.n_items = 1; b.subfr_length = subfr_length; b.nb_subfr = nb_subfr; b.D = order;
  silk_burg_thread(b, 0);
  float res_nrg = nrg[0];
  if (f.useInterpolatedNLSFs && !f.first_frame_after_reset && nb_subfr == 4) {
    b.A = a_tmp; b.x = x + 2 * subfr_length; b.nb_subfr = 2;
    silk_burg_thread(b, 0);
    res_nrg -= nrg[0];
    nlsf_a2nlsf_flp(f.NLSF_Q15, a_tmp, order);
    float res_nrg_2nd = 3.402823466e+38f;   // silk_float_MAX
    for (int k = 3; k >= 0; k--) {
      nlsf_interpolate(NLSF0_Q15, f.prev_NLSFq_Q15, f.NLSF_Q15, k, order);
      nlsf_NLSF2A(a_fix_Q12, NLSF0_Q15, order);
      for (int i = 0; i < order; i++) a_tmp[i] = (float)a_fix_Q12[i] * (1.0f / 4096.0f);
      for (int ix = order; ix < 2 * subfr_length; ix++) {   // silk_LPC_analysis_filter_FLP: the products summed left to right
        const float *sp = &x[ix - 1];
        float pred = sp[0] * a_tmp[0];
        for (int j = 1; j < order; j++) pred = pred + sp[-j] * a_tmp[j];
        LPC_res[ix] = sp[1] - pred;
      }
      for (int ix = 0; ix < order; ix++) LPC_res[ix] = 0.f;
      const float res_nrg_interp = (float)(silk_inner_product_flp(LPC_res + order, LPC_res + order, subfr_length - order) +
                                           silk_inner_product_flp(LPC_res + order + subfr_length, LPC_res + order + subfr_length, subfr_length - order));
      if (res_nrg_interp < res_nrg) {
        res_nrg = res_nrg_interp;
        f.NLSFInterpCoef_Q2 = k;
      } else if (res_nrg_interp > res_nrg_2nd) break;
      res_nrg_2nd = res_nrg_interp;
    }
  }
  if (f.NLSFInterpCoef_Q2 == 4) nlsf_a2nlsf_flp(f.NLSF_Q15, a, order);
}
OB_DEV void silk_find_lpc_thread(const SilkFindLpcParams &p, int s) {
  ObSilkFindLpc f = p.frames[s];
  silk_find_lpc_core(f, p.x + (size_t)s * p.nb_subfr * (p.subfr_length + p.order), p.order, p.nb_subfr, p.subfr_length);
  p.frames[s] = f;
}
// ---- the prediction analysis of a SILK frame: silk_find_pred_coefs_FLP (silk/float/find_pred_coefs_FLP.c:35-120) — for voiced frames silk_find_LTP_FLP
// (find_LTP_FLP.c:34-65, corrMatrix_FLP.c:38-95), silk_quant_LTP_gains (quant_LTP_gains.c:34-132 with silk_VQ_WMat_EC_c, VQ_WMat_EC.c:35-131, through the
// float wrapper wrappers_FLP.c:175-209), silk_LTP_scale_ctrl_FLP (LTP_scale_ctrl_FLP.c:34-58) and silk_LTP_analysis_filter_FLP (LTP_analysis_filter_FLP.c:34-75);
// then silk_find_LPC_FLP, silk_process_NLSFs and silk_residual_energy_FLP (residual_energy_FLP.c:91-117).  One thread per stream. ----
struct ObSilkPredCoefs {   // == OpusB200SilkPredCoefs
  float Gains[4];
  int32_t pitchL[4];
  float coding_quality;
  int32_t signalType, condCoding, first_frame_after_reset, useInterpolatedNLSFs, PacketLoss_perc, nFramesPerPacket, LBRR_flag, SNR_dB_Q7, speech_activity_Q8,
      NLSF_MSVQ_Survivors;
  int16_t prev_NLSFq_Q15[16];   // in / out
  int32_t sum_log_gain_Q7;      // in / out
  float LTPCoef[20], LTPredCodGain, LTP_scale, PredCoef[2][16], ResNrg[4];   // out
  int8_t LTPIndex[4], PERIndex, LTP_scaleIndex, NLSFInterpCoef_Q2, pad;
  int8_t NLSFIndices[17], pad2[3];
};
// res_pitch / x: [n][sig_stride] with the current frame at frame_off (history in front: at least the longest pitch lag + LTP_ORDER / 2 + order samples)
struct SilkPredCoefsParams { ObSilkPredCoefs *frames; const float *res_pitch, *x; int n, sig_stride, frame_off, order, nb_subfr, subfr_length; };
OB_DEV int32_t nl_add_pos_sat32(int32_t a, int32_t b) { return (((uint32_t)a + (uint32_t)b) & 0x80000000u) ? INT32_MAX : a + b; }
OB_DEV void silk_vq_wmat_ec(int8_t *ind, int32_t *res_nrg_Q15, int32_t *rate_dist_Q8, int *gain_Q7, const int32_t *XX, const int32_t *xX, const signed char *cb_Q7,
                              const unsigned char *cb_gain_Q7, const unsigned char *cl_Q5, int subfr_len, int32_t max_gain_Q7, int L) {
  int32_t neg[5];
  for (int j = 0; j < 5; j++) neg[j] = -(int32_t)((uint32_t)xX[j] << 7);
  *rate_dist_Q8 = INT32_MAX;
  *res_nrg_Q15 = INT32_MAX;
  *ind = 0;
  for (int k = 0; k < L; k++, cb_Q7 += 5) {
    const int gain_tmp_Q7 = cb_gain_Q7[k];
    const int32_t d = gain_tmp_Q7 - max_gain_Q7, penalty = (int32_t)((uint32_t)(d > 0 ? d : 0) << 11);
    const int32_t c0 = cb_Q7[0], c1 = cb_Q7[1], c2 = cb_Q7[2], c3 = cb_Q7[3], c4 = cb_Q7[4];
    int32_t sum1 = 32801, sum2;   // SILK_FIX_CONST(1.001, 15)
    sum2 = neg[0] + XX[1] * c1; sum2 = sum2 + XX[2] * c2; sum2 = sum2 + XX[3] * c3; sum2 = sum2 + XX[4] * c4;
    sum2 = (int32_t)((uint32_t)sum2 << 1); sum2 = sum2 + XX[0] * c0;
    sum1 = sum1 + silk_smulwb(sum2, c0);
    sum2 = neg[1] + XX[7] * c2; sum2 = sum2 + XX[8] * c3; sum2 = sum2 + XX[9] * c4;
    sum2 = (int32_t)((uint32_t)sum2 << 1); sum2 = sum2 + XX[6] * c1;
    sum1 = sum1 + silk_smulwb(sum2, c1);
    sum2 = neg[2] + XX[13] * c3; sum2 = sum2 + XX[14] * c4;
    sum2 = (int32_t)((uint32_t)sum2 << 1); sum2 = sum2 + XX[12] * c2;
    sum1 = sum1 + silk_smulwb(sum2, c2);
    sum2 = neg[3] + XX[19] * c4;
    sum2 = (int32_t)((uint32_t)sum2 << 1); sum2 = sum2 + XX[18] * c3;
    sum1 = sum1 + silk_smulwb(sum2, c3);
    sum2 = (int32_t)((uint32_t)neg[4] << 1); sum2 = sum2 + XX[24] * c4;
    sum1 = sum1 + silk_smulwb(sum2, c4);
    if (sum1 >= 0) {
      const int32_t bits_tot_Q8 = nl_smulbb(subfr_len, silk_lin2log(sum1 + penalty) - (15 << 7)) + ((int32_t)cl_Q5[k] << 2);
      if (bits_tot_Q8 <= *rate_dist_Q8) {
        *rate_dist_Q8 = bits_tot_Q8;
        *res_nrg_Q15 = sum1 + penalty;
        *ind = (int8_t)k;
        *gain_Q7 = gain_tmp_Q7;
      }
    }
  }
}
OB_DEV void silk_quant_ltp_gains(int16_t *B_Q14, int8_t *cbk_index, int8_t *periodicity_index, int32_t *sum_log_gain_Q7, int *pred_gain_dB_Q7,
                                   const int32_t *XX_Q17, const int32_t *xX_Q17, int subfr_len, int nb_subfr) {
  int32_t min_rate_dist_Q7 = INT32_MAX, best_sum_log_gain_Q7 = 0, res_nrg_Q15 = 0;
  int8_t temp_idx[4];
  for (int k = 0; k < 3; k++) {
    const int off = k == 0 ? 0 : k == 1 ? 8 : 24, size = 8 << k;
    const int32_t gain_safety = 51;   // SILK_FIX_CONST(0.4, 7)
    int32_t rate_dist_Q7 = 0, sum_log_gain_tmp_Q7 = *sum_log_gain_Q7, nrg = 0;
    for (int j = 0; j < nb_subfr; j++) {
      const int32_t max_gain_Q7 = silk_log2lin((5333 - sum_log_gain_tmp_Q7) + 896) - gain_safety;   // MAX_SUM_LOG_GAIN_DB / 6 and 7 in Q7
      int32_t res_nrg_subfr, rate_dist_subfr;
      int gain_Q7 = 0;
      silk_vq_wmat_ec(&temp_idx[j], &res_nrg_subfr, &rate_dist_subfr, &gain_Q7, XX_Q17 + 25 * j, xX_Q17 + 5 * j, ob_silk_LTP_gain_vq + 5 * off, ob_silk_LTP_vq_gain + off,
                      ob_silk_LTP_gain_BITS_Q5 + off, subfr_len, max_gain_Q7, size);
      nrg = nl_add_pos_sat32(nrg, res_nrg_subfr);
      rate_dist_Q7 = nl_add_pos_sat32(rate_dist_Q7, rate_dist_subfr);
      const int32_t t = sum_log_gain_tmp_Q7 + silk_lin2log(gain_safety + gain_Q7) - 896;
      sum_log_gain_tmp_Q7 = t > 0 ? t : 0;
    }
    res_nrg_Q15 = nrg;   // the value the last codebook leaves behind is the one the reference reports (quant_LTP_gains.c:122-131)
    if (rate_dist_Q7 <= min_rate_dist_Q7) {
      min_rate_dist_Q7 = rate_dist_Q7;
      *periodicity_index = (int8_t)k;
      for (int j = 0; j < nb_subfr; j++) cbk_index[j] = temp_idx[j];
      best_sum_log_gain_Q7 = sum_log_gain_tmp_Q7;
    }
  }
  const int off = *periodicity_index == 0 ? 0 : *periodicity_index == 1 ? 8 : 24;
  for (int j = 0; j < nb_subfr; j++)
    for (int k = 0; k < 5; k++) B_Q14[j * 5 + k] = (int16_t)(ob_silk_LTP_gain_vq[5 * (off + cbk_index[j]) + k] * 128);
  res_nrg_Q15 = nb_subfr == 2 ? res_nrg_Q15 >> 1 : res_nrg_Q15 >> 2;
  *sum_log_gain_Q7 = best_sum_log_gain_Q7;
  *pred_gain_dB_Q7 = nl_smulbb(-3, silk_lin2log(res_nrg_Q15) - (15 << 7));
}
OB_DEV void silk_corr_matrix_flp(const float *x, int L, float *XX) {   // silk_corrMatrix_FLP, Order = LTP_ORDER
  const float *ptr1 = &x[4], *ptr2 = &x[3];
  double energy = silk_inner_product_flp(ptr1, ptr1, L);
  XX[0] = (float)energy;
  for (int j = 1; j < 5; j++) {
    energy += (double)(ptr1[-j] * ptr1[-j] - ptr1[L - j] * ptr1[L - j]);
    XX[j * 5 + j] = (float)energy;
  }
  for (int lag = 1; lag < 5; lag++) {
    energy = silk_inner_product_flp(ptr1, ptr2, L);
    XX[lag * 5] = (float)energy;
    XX[lag] = (float)energy;
    for (int j = 1; j < 5 - lag; j++) {
      energy += (double)(ptr1[-j] * ptr2[-j] - ptr1[L - j] * ptr2[L - j]);
      XX[(lag + j) * 5 + j] = (float)energy;
      XX[j * 5 + lag + j] = (float)energy;
    }
    ptr2--;
  }
}
OB_DEV void silk_lpc_res_flp(float *r, const float *a, const float *s, int length, int order) {   // silk_LPC_analysis_filter_FLP
  for (int ix = order; ix < length; ix++) {
    const float *sp = &s[ix - 1];
    float pred = sp[0] * a[0];
    for (int j = 1; j < order; j++) pred = pred + sp[-j] * a[j];
    r[ix] = sp[1] - pred;
  }
  for (int ix = 0; ix < order; ix++) r[ix] = 0.f;
}
OB_DEV void silk_find_pred_coefs_thread(const SilkPredCoefsParams &p, int s) {
  ObSilkPredCoefs c = p.frames[s];
  const int order = p.order, nb_subfr = p.nb_subfr, subfr_length = p.subfr_length;
  const float *res_pitch = p.res_pitch + (size_t)s * p.sig_stride + p.frame_off, *x = p.x + (size_t)s * p.sig_stride + p.frame_off;
  float invGains[4], LPC_in_pre[4 * 16 + 320], minInvGain;
  for (int i = 0; i < nb_subfr; i++) invGains[i] = 1.0f / c.Gains[i];
  if (c.signalType == 2) {
    float XXLTP[4 * 25], xXLTP[4 * 5];
    int32_t XX_Q17[100], xX_Q17[20];
    int16_t B_Q14[20];
    const float *r_ptr = res_pitch;
    for (int k = 0; k < nb_subfr; k++) {   // silk_find_LTP_FLP
      const float *lag_ptr = r_ptr - (c.pitchL[k] + 2);
      float *XX = XXLTP + 25 * k, *xX = xXLTP + 5 * k;
      silk_corr_matrix_flp(lag_ptr, subfr_length, XX);
      for (int j = 0; j < 5; j++) xX[j] = (float)silk_inner_product_flp(&lag_ptr[4 - j], r_ptr, subfr_length);
      const float xx = (float)silk_inner_product_flp(r_ptr, r_ptr, subfr_length + 5);
      float temp = 0.03f * 0.5f * (XX[0] + XX[24]) + 1.0f;   // LTP_CORR_INV_MAX
      temp = 1.0f / (xx > temp ? xx : temp);
      for (int j = 0; j < 25; j++) XX[j] *= temp;
      for (int j = 0; j < 5; j++) xX[j] *= temp;
      r_ptr += subfr_length;
    }
    for (int i = 0; i < nb_subfr * 25; i++) XX_Q17[i] = (int32_t)rintf(XXLTP[i] * 131072.0f);
    for (int i = 0; i < nb_subfr * 5; i++) xX_Q17[i] = (int32_t)rintf(xXLTP[i] * 131072.0f);
    int pred_gain_dB_Q7;
    silk_quant_ltp_gains(B_Q14, c.LTPIndex, &c.PERIndex, &c.sum_log_gain_Q7, &pred_gain_dB_Q7, XX_Q17, xX_Q17, subfr_length, nb_subfr);
    for (int i = 0; i < nb_subfr * 5; i++) c.LTPCoef[i] = (float)B_Q14[i] * (1.0f / 16384.0f);
    c.LTPredCodGain = (float)pred_gain_dB_Q7 * (1.0f / 128.0f);
    if (c.condCoding == 0) {   // silk_LTP_scale_ctrl_FLP
      int round_loss = c.PacketLoss_perc * c.nFramesPerPacket;
      if (c.LBRR_flag) round_loss = 2 + nl_smulbb(round_loss, round_loss) / 100;
      const int32_t g = nl_smulbb((int32_t)c.LTPredCodGain, round_loss);
      c.LTP_scaleIndex = (int8_t)((g > silk_log2lin(2900 - c.SNR_dB_Q7)) + (g > silk_log2lin(3900 - c.SNR_dB_Q7)));
    } else c.LTP_scaleIndex = 0;
    c.LTP_scale = (float)(c.LTP_scaleIndex == 0 ? 15565 : c.LTP_scaleIndex == 1 ? 12288 : 8192) / 16384.0f;   // silk_LTPScales_table_Q14
    const float *x_ptr = x - order;
    float *res = LPC_in_pre;
    for (int k = 0; k < nb_subfr; k++) {   // silk_LTP_analysis_filter_FLP
      const float *x_lag_ptr = x_ptr - c.pitchL[k], inv_gain = invGains[k], *B = &c.LTPCoef[k * 5];
      for (int i = 0; i < subfr_length + order; i++) {
        float v = x_ptr[i];
        for (int j = 0; j < 5; j++) v -= B[j] * x_lag_ptr[2 - j];
        res[i] = v * inv_gain;
        x_lag_ptr++;
      }
      res += subfr_length + order;
      x_ptr += subfr_length;
    }
  } else {
    const float *x_ptr = x - order;
    float *pre = LPC_in_pre;
    for (int i = 0; i < nb_subfr; i++) {
      for (int j = 0; j < subfr_length + order; j++) pre[j] = invGains[i] * x_ptr[j];
      pre += subfr_length + order;
      x_ptr += subfr_length;
    }
    for (int i = 0; i < nb_subfr * 5; i++) c.LTPCoef[i] = 0.f;
    c.LTPredCodGain = 0.0f;
    c.sum_log_gain_Q7 = 0;
  }
  if (c.first_frame_after_reset) minInvGain = 1.0f / 1e2f;
  else {
    minInvGain = (float)pow(2.0, (double)(c.LTPredCodGain / 3)) / 1e4f;
    minInvGain /= 0.25f + 0.75f * c.coding_quality;
  }
  ObSilkFindLpc f;
  for (int i = 0; i < 16; i++) { f.NLSF_Q15[i] = 0; f.prev_NLSFq_Q15[i] = c.prev_NLSFq_Q15[i]; }
  f.minInvGain = minInvGain; f.useInterpolatedNLSFs = c.useInterpolatedNLSFs; f.first_frame_after_reset = c.first_frame_after_reset; f.NLSFInterpCoef_Q2 = 4;
  silk_find_lpc_core(f, LPC_in_pre, order, nb_subfr, subfr_length);
  c.NLSFInterpCoef_Q2 = (int8_t)f.NLSFInterpCoef_Q2;
  ObSilkNlsf q;
  for (int i = 0; i < 16; i++) { q.NLSF_Q15[i] = i < order ? f.NLSF_Q15[i] : 0; q.prev_NLSFq_Q15[i] = c.prev_NLSFq_Q15[i]; }
  q.NLSFInterpCoef_Q2 = c.NLSFInterpCoef_Q2; q.signalType = (int8_t)c.signalType; q.useInterpolatedNLSFs = (int8_t)c.useInterpolatedNLSFs;
  q.speech_activity_Q8 = c.speech_activity_Q8; q.NLSF_MSVQ_Survivors = c.NLSF_MSVQ_Survivors;
  for (int i = 0; i < 17; i++) q.NLSFIndices[i] = 0;
  silk_process_nlsfs_core(q, order == 16, nb_subfr);
  for (int i = 0; i < 17; i++) c.NLSFIndices[i] = q.NLSFIndices[i];
  for (int j = 0; j < 2; j++)
    for (int i = 0; i < 16; i++) c.PredCoef[j][i] = i < order ? (float)q.PredCoef_Q12[j][i] * (1.0f / 4096.0f) : 0.f;
  {   // silk_residual_energy_FLP
    const int shift = order + subfr_length;
    float LPC_res[192];
    silk_lpc_res_flp(LPC_res, c.PredCoef[0], LPC_in_pre, 2 * shift, order);
    c.ResNrg[0] = (float)((double)(c.Gains[0] * c.Gains[0]) * silk_inner_product_flp(LPC_res + order, LPC_res + order, subfr_length));
    c.ResNrg[1] = (float)((double)(c.Gains[1] * c.Gains[1]) * silk_inner_product_flp(LPC_res + order + shift, LPC_res + order + shift, subfr_length));
    if (nb_subfr == 4) {
      silk_lpc_res_flp(LPC_res, c.PredCoef[1], LPC_in_pre + 2 * shift, 2 * shift, order);
      c.ResNrg[2] = (float)((double)(c.Gains[2] * c.Gains[2]) * silk_inner_product_flp(LPC_res + order, LPC_res + order, subfr_length));
      c.ResNrg[3] = (float)((double)(c.Gains[3] * c.Gains[3]) * silk_inner_product_flp(LPC_res + order + shift, LPC_res + order + shift, subfr_length));
    }
  }
  for (int i = 0; i < 16; i++) c.prev_NLSFq_Q15[i] = i < order ? q.NLSF_Q15[i] : 0;
  p.frames[s] = c;
}
// ---- noise shaping analysis of a SILK frame: silk_noise_shape_analysis_FLP (silk/float/noise_shape_analysis_FLP.c:36-350) with silk_apply_sine_window_FLP,
// silk_warped_autocorrelation_FLP / silk_autocorrelation_FLP, silk_schur_FLP, silk_k2a_FLP, silk_bwexpander_FLP.  One thread per stream: the warped
// autocorrelation is a serial all-pass recursion over the window (double precision, as in the reference). ----
constexpr int kNsMaxOrder = 24;
struct ObSilkNoiseShape {   // == OpusB200SilkNoiseShape
  int32_t SNR_dB_Q7, input_quality_bands_Q15[2], useCBR, speech_activity_Q8, signalType, warping_Q16;
  float LTPCorr, predGain;
  int32_t pitchL[4];
  float HarmShapeGain_smth, Tilt_smth;                     // in / out
  float input_quality, coding_quality;                     // out
  int32_t quantOffsetType;                                 // out
  float AR[4 * kNsMaxOrder], Gains[4], LF_MA_shp[4], LF_AR_shp[4], HarmShapeGain[4], Tilt[4];   // out
};
// pitch_res [n][res_stride]: the frame's residual from the pitch analysis; x [n][x_stride] with the frame at x_off (la_shape samples in front, look-ahead behind)
struct SilkNoiseShapeParams { ObSilkNoiseShape *frames; const float *pitch_res, *x; int n, res_stride, x_stride, x_off, fs_kHz, nb_subfr, shapingLPCOrder; };
OB_DEV void ns_sine_window(float *w, const float *px, int win_type, int length) {
  const float freq = 3.1415926536f / (length + 1), c = 2.0f - freq * freq;
  float S0 = win_type < 2 ? 0.0f : 1.0f, S1 = win_type < 2 ? freq : 0.5f * c;
  for (int k = 0; k < length; k += 4) {
    w[k + 0] = px[k + 0] * 0.5f * (S0 + S1);
    w[k + 1] = px[k + 1] * S1;
    S0 = c * S1 - S0;
    w[k + 2] = px[k + 2] * 0.5f * (S1 + S0);
    w[k + 3] = px[k + 3] * S0;
    S1 = c * S0 - S1;
  }
}
OB_DEV void ns_bwexpander(float *ar, int d, float chirp) {
  float cfac = chirp;
  for (int i = 0; i < d - 1; i++) { ar[i] *= cfac; cfac *= chirp; }
  ar[d - 1] *= cfac;
}
OB_DEV void ns_limit(float *coefs, float lambda, float limit, int order, int warped) {   // warped_true2monic_coefs / limit_coefs
  float gain = 0.f;
  int ind = 0;
  if (warped) {
    for (int i = order - 1; i > 0; i--) coefs[i - 1] -= lambda * coefs[i];
    gain = (1.0f - lambda * lambda) / (1.0f + lambda * coefs[0]);
    for (int i = 0; i < order; i++) coefs[i] *= gain;
  }
  for (int iter = 0; iter < 10; iter++) {
    float maxabs = -1.0f;
    for (int i = 0; i < order; i++) {
      const float tmp = fabsf(coefs[i]);
      if (tmp > maxabs) { maxabs = tmp; ind = i; }
    }
    if (maxabs <= limit) return;
    if (warped) {
      for (int i = 1; i < order; i++) coefs[i - 1] += lambda * coefs[i];
      gain = 1.0f / gain;
      for (int i = 0; i < order; i++) coefs[i] *= gain;
    }
    const float chirp = 0.99f - (0.8f + 0.1f * iter) * (maxabs - limit) / (maxabs * (ind + 1));
    ns_bwexpander(coefs, order, chirp);
    if (warped) {
      for (int i = order - 1; i > 0; i--) coefs[i - 1] -= lambda * coefs[i];
      gain = (1.0f - lambda * lambda) / (1.0f + lambda * coefs[0]);
      for (int i = 0; i < order; i++) coefs[i] *= gain;
    }
  }
}
OB_DEV void silk_noise_shape_thread(const SilkNoiseShapeParams &p, int s) {
  ObSilkNoiseShape c = p.frames[s];
  const int fs_kHz = p.fs_kHz, nb_subfr = p.nb_subfr, order = p.shapingLPCOrder, subfr_length = 5 * fs_kHz, la_shape = 5 * fs_kHz, shapeWinLength = 15 * fs_kHz;
  const float *pitch_res = p.pitch_res + (size_t)s * p.res_stride, *x_ptr = p.x + (size_t)s * p.x_stride + p.x_off - la_shape;
  float x_windowed[240], auto_corr[kNsMaxOrder + 1], rc[kNsMaxOrder + 1];
  float SNR_adj_dB = c.SNR_dB_Q7 * (1 / 128.0f);
  c.input_quality = 0.5f * (c.input_quality_bands_Q15[0] + c.input_quality_bands_Q15[1]) * (1.0f / 32768.0f);
  c.coding_quality = (float)(1.0 / (1.0 + exp(-(double)(0.25f * (SNR_adj_dB - 20.0f)))));
  if (c.useCBR == 0) {
    const float b = 1.0f - c.speech_activity_Q8 * (1.0f / 256.0f);
    SNR_adj_dB -= 2.0f * c.coding_quality * (0.5f + 0.5f * c.input_quality) * b * b;   // BG_SNR_DECR_dB
  }
  if (c.signalType == 2) SNR_adj_dB += 2.0f * c.LTPCorr;                               // HARM_SNR_INCR_dB
  else SNR_adj_dB += (-0.4f * c.SNR_dB_Q7 * (1 / 128.0f) + 6.0f) * (1.0f - c.input_quality);
  if (c.signalType == 2) c.quantOffsetType = 0;
  else {   // sparseness: fluctuation of the log energy per 2 ms
    const int nSamples = 2 * fs_kHz, nSegs = (5 * nb_subfr) / 2;
    float energy_variation = 0.0f, log_energy_prev = 0.0f;
    for (int k = 0; k < nSegs; k++) {
      const float *q = pitch_res + k * nSamples;
      const float nrg = (float)nSamples + (float)silk_inner_product_flp(q, q, nSamples);
      const float log_energy = (float)(3.32192809488736 * log10((double)nrg));
      if (k > 0) energy_variation += fabsf(log_energy - log_energy_prev);
      log_energy_prev = log_energy;
    }
    c.quantOffsetType = energy_variation > 0.6f * (nSegs - 1) ? 0 : 1;
  }
  float strength = 1e-3f * c.predGain;
  const float BWExp = 0.94f / (1.0f + strength * strength);
  const float warping = (float)c.warping_Q16 / 65536.0f + 0.01f * c.coding_quality;
  for (int k = 0; k < nb_subfr; k++) {
    const int flat_part = fs_kHz * 3, slope_part = (shapeWinLength - flat_part) / 2;
    float *AR = &c.AR[k * kNsMaxOrder];
    ns_sine_window(x_windowed, x_ptr, 1, slope_part);
    for (int i = 0; i < flat_part; i++) x_windowed[slope_part + i] = x_ptr[slope_part + i];
    ns_sine_window(x_windowed + slope_part + flat_part, x_ptr + slope_part + flat_part, 2, slope_part);
    x_ptr += subfr_length;
    if (c.warping_Q16 > 0) {
      double state[kNsMaxOrder + 1], C[kNsMaxOrder + 1];
      for (int i = 0; i <= kNsMaxOrder; i++) { state[i] = 0.0; C[i] = 0.0; }
      for (int n = 0; n < shapeWinLength; n++) {
        double tmp1 = x_windowed[n];
        for (int i = 0; i < order; i += 2) {
          const double tmp2 = state[i] + warping * state[i + 1] - warping * tmp1;
          state[i] = tmp1;
          C[i] += state[0] * tmp1;
          tmp1 = state[i + 1] + warping * state[i + 2] - warping * tmp2;
          state[i + 1] = tmp2;
          C[i + 1] += state[0] * tmp2;
        }
        state[order] = tmp1;
        C[order] += state[0] * tmp1;
      }
      for (int i = 0; i < order + 1; i++) auto_corr[i] = (float)C[i];
    } else {
      for (int i = 0; i < order + 1; i++) auto_corr[i] = (float)silk_inner_product_flp(x_windowed, x_windowed + i, shapeWinLength - i);
    }
    auto_corr[0] += auto_corr[0] * 3e-5f + 1.0f;   // SHAPE_WHITE_NOISE_FRACTION
    float nrg;
    {   // silk_schur_FLP
      double Cs[kNsMaxOrder + 1][2];
      for (int i = 0; i <= order; i++) Cs[i][0] = Cs[i][1] = auto_corr[i];
      for (int i = 0; i < order; i++) {
        const double den = Cs[0][1] > (double)1e-9f ? Cs[0][1] : (double)1e-9f, rc_tmp = -Cs[i + 1][0] / den;
        rc[i] = (float)rc_tmp;
        for (int n = 0; n < order - i; n++) {
          const double Ctmp1 = Cs[n + i + 1][0], Ctmp2 = Cs[n][1];
          Cs[n + i + 1][0] = Ctmp1 + Ctmp2 * rc_tmp;
          Cs[n][1] = Ctmp2 + Ctmp1 * rc_tmp;
        }
      }
      nrg = (float)Cs[0][1];
    }
    for (int i = 0; i < order; i++) {   // silk_k2a_FLP
      const float rck = rc[i];
      for (int n = 0; n < (i + 1) >> 1; n++) {
        const float tmp1 = AR[n], tmp2 = AR[i - n - 1];
        AR[n] = tmp1 + tmp2 * rck;
        AR[i - n - 1] = tmp2 + tmp1 * rck;
      }
      AR[i] = -rck;
    }
    c.Gains[k] = (float)sqrt((double)nrg);
    if (c.warping_Q16 > 0) {   // warped_gain
      const float lambda = -warping;
      float gain = AR[order - 1];
      for (int i = order - 2; i >= 0; i--) gain = lambda * gain + AR[i];
      c.Gains[k] *= 1.0f / (1.0f - lambda * gain);
    }
    ns_bwexpander(AR, order, BWExp);
    ns_limit(AR, warping, 3.999f, order, c.warping_Q16 > 0);
  }
  const float gain_mult = (float)pow((double)2.0f, (double)(-0.16f * SNR_adj_dB)), gain_add = (float)pow((double)2.0f, (double)(0.16f * 2));
  for (int k = 0; k < nb_subfr; k++) { c.Gains[k] *= gain_mult; c.Gains[k] += gain_add; }
  strength = 4.0f * (1.0f + 0.5f * (c.input_quality_bands_Q15[0] * (1.0f / 32768.0f) - 1.0f));
  strength *= c.speech_activity_Q8 * (1.0f / 256.0f);
  float Tilt, HarmShapeGain;
  if (c.signalType == 2) {
    for (int k = 0; k < nb_subfr; k++) {
      const float b = 0.2f / fs_kHz + 3.0f / c.pitchL[k];
      c.LF_MA_shp[k] = -1.0f + b;
      c.LF_AR_shp[k] = 1.0f - b - b * strength;
    }
    Tilt = -0.25f - (1 - 0.25f) * 0.35f * c.speech_activity_Q8 * (1.0f / 256.0f);
    HarmShapeGain = 0.3f;
    HarmShapeGain += 0.2f * (1.0f - (1.0f - c.coding_quality) * c.input_quality);
    HarmShapeGain *= (float)sqrt((double)c.LTPCorr);
  } else {
    const float b = 1.3f / fs_kHz;
    c.LF_MA_shp[0] = -1.0f + b;
    c.LF_AR_shp[0] = 1.0f - b - b * strength * 0.6f;
    for (int k = 1; k < nb_subfr; k++) { c.LF_MA_shp[k] = c.LF_MA_shp[0]; c.LF_AR_shp[k] = c.LF_AR_shp[0]; }
    Tilt = -0.25f;
    HarmShapeGain = 0.0f;
  }
  for (int k = 0; k < nb_subfr; k++) {
    c.HarmShapeGain_smth += 0.4f * (HarmShapeGain - c.HarmShapeGain_smth);
    c.HarmShapeGain[k] = c.HarmShapeGain_smth;
    c.Tilt_smth += 0.4f * (Tilt - c.Tilt_smth);
    c.Tilt[k] = c.Tilt_smth;
  }
  p.frames[s] = c;
}
}  // namespace ob
