// silk_pitch.cuh — the pitch analysis of a SILK frame, one THREAD per stream: silk_find_pitch_lags_FLP (silk/float/find_pitch_lags_FLP.c:36-132:
// windowed autocorrelation, Schur, step-up, bandwidth expansion, LPC residual of history + frame + look-ahead) and silk_pitch_analysis_core_FLP
// (silk/float/pitch_analysis_core_FLP.c:66-630): int16 conversion and decimation to 8 and 4 kHz (silk_resampler_down2, resampler_down2.c:37-74;
// silk_resampler_down2_3, resampler_down2_3.c:41-103 with resampler_private_AR2.c:36-54), stage 1 at 4 kHz (normalised cross-correlation per lag, short-lag
// bias, the best candidates by silk_insertion_sort_decreasing_FLP), stage 2 at 8 kHz over the candidate neighbourhoods with the contour codebook, stage 3
// at the input rate over lag +- 2 and the fine contour codebook (silk/pitch_est_tables.c).  The searches are data-dependent walks over a few hundred
// candidate lags of ONE frame, so the parallel axis is the batch.  Sums in the reference's order (the C build of celt_pitch_xcorr: every lag in index order).
#pragma once
#include "simt.cuh"
#include "silk_nlsf.cuh"   // ns_sine_window, ns_bwexpander, silk_inner_product_flp, integer helpers

namespace ob {
struct ObSilkPitch {   // == OpusB200SilkPitch
  int32_t signalType;   // in: 0 = no voice activity; out: 1 unvoiced / 2 voiced (0 stays)
  int32_t first_frame_after_reset, speech_activity_Q8, prevSignalType, input_tilt_Q15, prevLag, pitchEstimationThreshold_Q16;
  float LTPCorr;        // in / out
  float predGain;       // out
  int32_t pitchL[4];    // out
  int32_t lagIndex, contourIndex;   // out
};
// x [n][x_stride] with the frame at x_off (20 ms of history in front, 2 ms look-ahead behind); res [n][res_stride] out: la_pitch + frame + ltp_mem samples
struct SilkPitchParams { ObSilkPitch *frames; const float *x; float *res; int n, x_stride, x_off, res_stride, fs_kHz, nb_subfr, lpc_order, complexity; };

OB_TABLE signed char ob_pe_cb2[4 * 11] = {0, 2, -1, -1, -1, 0, 0, 1, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0,
                                         0, -1, 2, 1, 0, 1, 1, 0, 0, -1, -1};   // silk_CB_lags_stage2
OB_TABLE signed char ob_pe_cb3[4 * 34] = {
    0, 0, 1, -1, 0, 1, -1, 0, -1, 1, -2, 2, -2, -2, 2, -3, 2, 3, -3, -4, 3, -4, 4, 4, -5, 5, -6, -5, 6, -7, 6, 5, 8, -9,
    0, 0, 1, 0, 0, 0, 0, 0, 0, 0, -1, 1, 0, 0, 1, -1, 0, 1, -1, -1, 1, -1, 2, 1, -1, 2, -2, -2, 2, -2, 2, 2, 3, -3,
    0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 1, -1, 1, 0, 0, 2, 1, -1, 2, -1, -1, 2, -1, 2, 2, -1, 3, -2, -2, -2, 3,
    0, 1, 0, 0, 1, 0, 1, -1, 2, -1, 2, -1, 2, 3, -2, 3, -2, -2, 4, 4, -3, 5, -3, -4, 6, -4, 6, 5, -5, 8, -6, -5, -7, 9};   // silk_CB_lags_stage3
OB_TABLE signed char ob_pe_range3[3 * 4 * 2] = {-5, 8, -1, 6, -1, 6, -4, 10, -6, 10, -2, 6, -1, 6, -5, 10, -9, 12, -3, 7, -2, 7, -7, 13};   // silk_Lag_range_stage3
OB_TABLE signed char ob_pe_cb2_10[2 * 3] = {0, 1, 0, 0, 0, 1};
OB_TABLE signed char ob_pe_cb3_10[2 * 12] = {0, 0, 1, -1, 1, -1, 2, -2, 2, -2, 3, -3, 0, 1, 0, 1, -1, 2, -1, 2, -2, 3, -2, 3};
OB_TABLE signed char ob_pe_range3_10[2 * 2] = {-3, 7, -2, 7};

OB_DEV int16_t pe_sat16(int32_t a) { return (int16_t)(a > 32767 ? 32767 : a < -32768 ? -32768 : a); }
OB_DEV void pe_down2(int32_t *S, int16_t *out, const int16_t *in, int inLen) {
  for (int k = 0; k < inLen >> 1; k++) {
    int32_t in32 = (int32_t)in[2 * k] * 1024, Y = in32 - S[0], X = Y + silk_smulwb(Y, 39809 - 65536), out32 = S[0] + X;
    S[0] = in32 + X;
    in32 = (int32_t)in[2 * k + 1] * 1024;
    Y = in32 - S[1];
    X = silk_smulwb(Y, 9872);
    out32 = out32 + S[1];
    out32 = out32 + X;
    S[1] = in32 + X;
    out[k] = pe_sat16(nl_rr(out32, 11));
  }
}
// silk_resampler_down2_3 for one batch (inLen <= 480); buf: inLen + 4 words of scratch
OB_DEV void pe_down2_3(int32_t *S, int16_t *out, const int16_t *in, int inLen, int32_t *buf) {
  const int32_t C0 = -2797, C1 = -6507, C2 = 4697, C3 = 10739, C4 = 1567, C5 = 8276;   // silk_Resampler_2_3_COEFS_LQ
  for (int k = 0; k < 4; k++) buf[k] = S[k];
  for (int k = 0; k < inLen; k++) {   // silk_resampler_private_AR2
    int32_t out32 = S[4] + (int32_t)in[k] * 256;
    buf[4 + k] = out32;
    out32 = (int32_t)((uint32_t)out32 << 2);
    S[4] = S[5] + silk_smulwb(out32, C0);
    S[5] = silk_smulwb(out32, C1);
  }
  const int32_t *bp = buf;
  for (int counter = inLen; counter > 2; counter -= 3, bp += 3) {
    int32_t r = silk_smulwb(bp[0], C2);
    r = r + silk_smulwb(bp[1], C3);
    r = r + silk_smulwb(bp[2], C5);
    r = r + silk_smulwb(bp[3], C4);
    *out++ = pe_sat16(nl_rr(r, 6));
    r = silk_smulwb(bp[1], C4);
    r = r + silk_smulwb(bp[2], C5);
    r = r + silk_smulwb(bp[3], C3);
    r = r + silk_smulwb(bp[4], C2);
    *out++ = pe_sat16(nl_rr(r, 6));
  }
  for (int k = 0; k < 4; k++) S[k] = buf[inLen + k];
}
OB_DEV void pe_sort_decreasing(float *a, int *idx, int L, int K) {
  for (int i = 0; i < K; i++) idx[i] = i;
  for (int i = 1; i < K; i++) {
    const float value = a[i];
    int j;
    for (j = i - 1; j >= 0 && value > a[j]; j--) { a[j + 1] = a[j]; idx[j + 1] = idx[j]; }
    a[j + 1] = value;
    idx[j + 1] = i;
  }
  for (int i = K; i < L; i++) {
    const float value = a[i];
    if (value > a[K - 1]) {
      int j;
      for (j = K - 2; j >= 0 && value > a[j]; j--) { a[j + 1] = a[j]; idx[j + 1] = idx[j]; }
      a[j + 1] = value;
      idx[j + 1] = i;
    }
  }
}
OB_DEV float pe_log2(double x) { return (float)(3.32192809488736 * log10(x)); }
OB_DEV float pe_dot(const float *a, const float *b, int n) {
  float s = 0;
  for (int i = 0; i < n; i++) s = s + a[i] * b[i];
  return s;
}
// silk_pitch_analysis_core_FLP: 0 voiced, 1 unvoiced
OB_DEV int silk_pitch_analysis_core(const float *frame, int32_t *pitch_out, int32_t *lagIndex, int32_t *contourIndex, float *LTPCorr, int prevLag, float search_thres1,
                                    float search_thres2, int Fs_kHz, int complexity, int nb_subfr) {
  float frame_8kHz[40 * 8], frame_4kHz[40 * 4], C[4][149], xcorr[65], CC[11];
  int16_t frame_8_FIX[40 * 8], frame_4_FIX[40 * 4], d_comp[149];
  int32_t filt_state[6];
  int d_srch[24];
  const int frame_length = (20 + nb_subfr * 5) * Fs_kHz, frame_length_4kHz = (20 + nb_subfr * 5) * 4, frame_length_8kHz = (20 + nb_subfr * 5) * 8, sf_length = 5 * Fs_kHz,
            sf_length_4kHz = 20, sf_length_8kHz = 40, min_lag = 2 * Fs_kHz, min_lag_4kHz = 8, min_lag_8kHz = 16, max_lag = 18 * Fs_kHz - 1, max_lag_4kHz = 72,
            max_lag_8kHz = 18 * 8 - 1;
  if (Fs_kHz == 8) {
    for (int k = frame_length_8kHz - 1; k >= 0; k--) frame_8_FIX[k] = pe_sat16((int32_t)rintf(frame[k]));
  } else {   // the input converted to int16 into the same scratch the 2:3 decimator filters in (32-bit words)
    int32_t scratch[(16 * 40) / 2 + 480 + 4];
    int16_t *in16 = (int16_t *)scratch;
    for (int k = frame_length - 1; k >= 0; k--) in16[k] = pe_sat16((int32_t)rintf(frame[k]));
    for (int k = 0; k < 6; k++) filt_state[k] = 0;
    if (Fs_kHz == 16) pe_down2(filt_state, frame_8_FIX, in16, frame_length);
    else pe_down2_3(filt_state, frame_8_FIX, in16, frame_length, scratch + (16 * 40) / 2);
    for (int k = 0; k < frame_length_8kHz; k++) frame_8kHz[k] = (float)frame_8_FIX[k];
  }
  filt_state[0] = filt_state[1] = 0;
  pe_down2(filt_state, frame_4_FIX, frame_8_FIX, frame_length_8kHz);
  for (int k = 0; k < frame_length_4kHz; k++) frame_4kHz[k] = (float)frame_4_FIX[k];
  for (int i = frame_length_4kHz - 1; i > 0; i--) {   // silk_ADD_SAT16 applied to floats (:168-170): (int32)a + b, saturated, truncated to int16
    const float t = (float)(int32_t)frame_4kHz[i] + frame_4kHz[i - 1];
    frame_4kHz[i] = (float)(int16_t)(t > 32767 ? 32767 : t < -32768 ? -32768 : t);
  }
  // ---- first stage, 4 kHz ----
  for (int k = 0; k < nb_subfr; k++)
    for (int i = 0; i < 149; i++) C[k][i] = 0.f;
  const float *target_ptr = &frame_4kHz[sf_length_4kHz << 2];
  for (int k = 0; k < nb_subfr >> 1; k++) {
    const float *basis_ptr = target_ptr - min_lag_4kHz;
    for (int i = 0; i < max_lag_4kHz - min_lag_4kHz + 1; i++) xcorr[i] = pe_dot(target_ptr, target_ptr - max_lag_4kHz + i, sf_length_8kHz);
    double cross_corr = xcorr[max_lag_4kHz - min_lag_4kHz];
    double normalizer = silk_inner_product_flp(target_ptr, target_ptr, sf_length_8kHz) + silk_inner_product_flp(basis_ptr, basis_ptr, sf_length_8kHz) +
                        (double)(sf_length_8kHz * 4000.0f);
    C[0][min_lag_4kHz] += (float)(2 * cross_corr / normalizer);
    for (int d = min_lag_4kHz + 1; d <= max_lag_4kHz; d++) {
      basis_ptr--;
      cross_corr = xcorr[max_lag_4kHz - d];
      normalizer += basis_ptr[0] * (double)basis_ptr[0] - basis_ptr[sf_length_8kHz] * (double)basis_ptr[sf_length_8kHz];
      C[0][d] += (float)(2 * cross_corr / normalizer);
    }
    target_ptr += sf_length_8kHz;
  }
  for (int i = max_lag_4kHz; i >= min_lag_4kHz; i--) C[0][i] -= C[0][i] * i / 4096.0f;
  int length_d_srch = 4 + 2 * complexity;
  pe_sort_decreasing(&C[0][min_lag_4kHz], d_srch, max_lag_4kHz - min_lag_4kHz + 1, length_d_srch);
  const float Cmax = C[0][min_lag_4kHz];
  if (Cmax < 0.2f) {
    for (int k = 0; k < nb_subfr; k++) pitch_out[k] = 0;
    *LTPCorr = 0.0f; *lagIndex = 0; *contourIndex = 0;
    return 1;
  }
  const float threshold = search_thres1 * Cmax;
  for (int i = 0; i < length_d_srch; i++) {
    if (C[0][min_lag_4kHz + i] > threshold) d_srch[i] = (d_srch[i] + min_lag_4kHz) << 1;
    else { length_d_srch = i; break; }
  }
  for (int i = min_lag_8kHz - 5; i < max_lag_8kHz + 5; i++) d_comp[i] = 0;
  for (int i = 0; i < length_d_srch; i++) d_comp[d_srch[i]] = 1;
  for (int i = max_lag_8kHz + 3; i >= min_lag_8kHz; i--) d_comp[i] = (int16_t)(d_comp[i] + d_comp[i - 1] + d_comp[i - 2]);
  length_d_srch = 0;
  for (int i = min_lag_8kHz; i < max_lag_8kHz + 1; i++) if (d_comp[i + 1] > 0) d_srch[length_d_srch++] = i;
  for (int i = max_lag_8kHz + 3; i >= min_lag_8kHz; i--) d_comp[i] = (int16_t)(d_comp[i] + d_comp[i - 1] + d_comp[i - 2] + d_comp[i - 3]);
  int length_d_comp = 0;
  for (int i = min_lag_8kHz; i < max_lag_8kHz + 4; i++) if (d_comp[i] > 0) d_comp[length_d_comp++] = (int16_t)(i - 2);
  // ---- second stage, 8 kHz ----
  for (int k = 0; k < 4; k++)
    for (int i = 0; i < 149; i++) C[k][i] = 0.f;
  target_ptr = Fs_kHz == 8 ? &frame[20 * 8] : &frame_8kHz[20 * 8];
  for (int k = 0; k < nb_subfr; k++) {
    const double energy_tmp = silk_inner_product_flp(target_ptr, target_ptr, sf_length_8kHz) + 1.0;
    for (int j = 0; j < length_d_comp; j++) {
      const int d = d_comp[j];
      const float *basis_ptr = target_ptr - d;
      const double cross_corr = silk_inner_product_flp(basis_ptr, target_ptr, sf_length_8kHz);
      if (cross_corr > 0.0) C[k][d] = (float)(2 * cross_corr / (silk_inner_product_flp(basis_ptr, basis_ptr, sf_length_8kHz) + energy_tmp));
      else C[k][d] = 0.0f;
    }
    target_ptr += sf_length_8kHz;
  }
  float CCmax = 0.0f, CCmax_b = -1000.0f, prevLag_log2 = 0;
  int CBimax = 0, lag = -1;
  if (prevLag > 0) {
    if (Fs_kHz == 12) prevLag = (prevLag << 1) / 3;
    else if (Fs_kHz == 16) prevLag = prevLag >> 1;
    prevLag_log2 = pe_log2((double)(float)prevLag);
  }
  int cbk_size, nb_cbk_search;
  const signed char *Lag_CB_ptr;
  if (nb_subfr == 4) { cbk_size = 11; Lag_CB_ptr = ob_pe_cb2; nb_cbk_search = (Fs_kHz == 8 && complexity > 0) ? 11 : 3; }
  else { cbk_size = 3; Lag_CB_ptr = ob_pe_cb2_10; nb_cbk_search = 3; }
  for (int k = 0; k < length_d_srch; k++) {
    const int d = d_srch[k];
    for (int j = 0; j < nb_cbk_search; j++) {
      CC[j] = 0.0f;
      for (int i = 0; i < nb_subfr; i++) CC[j] += C[i][d + Lag_CB_ptr[i * cbk_size + j]];
    }
    float CCmax_new = -1000.0f;
    int CBimax_new = 0;
    for (int i = 0; i < nb_cbk_search; i++) if (CC[i] > CCmax_new) { CCmax_new = CC[i]; CBimax_new = i; }
    const float lag_log2 = pe_log2((double)(float)d);
    float CCmax_new_b = CCmax_new - 0.2f * nb_subfr * lag_log2;   // PE_SHORTLAG_BIAS
    if (prevLag > 0) {
      float dl = lag_log2 - prevLag_log2;
      dl *= dl;
      CCmax_new_b -= 0.2f * nb_subfr * (*LTPCorr) * dl / (dl + 0.5f);   // PE_PREVLAG_BIAS
    }
    if (CCmax_new_b > CCmax_b && CCmax_new > nb_subfr * search_thres2) { CCmax_b = CCmax_new_b; CCmax = CCmax_new; lag = d; CBimax = CBimax_new; }
  }
  if (lag == -1) {
    for (int k = 0; k < 4; k++) pitch_out[k] = 0;
    *LTPCorr = 0.0f; *lagIndex = 0; *contourIndex = 0;
    return 1;
  }
  *LTPCorr = (float)(CCmax / nb_subfr);
  if (Fs_kHz > 8) {   // ---- third stage, input rate: lag +- 2 and the fine contour codebook ----
    if (Fs_kHz == 12) lag = nl_rr((int32_t)(int16_t)lag * 3, 1);
    else lag = lag << 1;
    lag = lag < min_lag ? min_lag : lag > max_lag ? max_lag : lag;
    const int start_lag = lag - 2 > min_lag ? lag - 2 : min_lag, end_lag = lag + 2 < max_lag ? lag + 2 : max_lag;
    int lag_new = lag;
    CBimax = 0;
    CCmax = -1000.0f;
    const signed char *range, *cb3;
    int nb3, cb3_size;
    if (nb_subfr == 4) { range = ob_pe_range3 + 8 * complexity; cb3 = ob_pe_cb3; nb3 = complexity == 0 ? 16 : complexity == 1 ? 24 : 34; cb3_size = 34; }
    else { range = ob_pe_range3_10; cb3 = ob_pe_cb3_10; nb3 = 12; cb3_size = 12; }
    float xc[4][22], en[4][22];   // silk_P_Ana_calc_corr_st3 / _energy_st3: per subframe the correlations / energies over its lag range
    target_ptr = &frame[sf_length << 2];
    for (int k = 0; k < nb_subfr; k++) {
      const int lag_low = range[k * 2], lag_high = range[k * 2 + 1], lag_diff = lag_high - lag_low + 1;
      for (int j = lag_low; j <= lag_high; j++) xc[k][j - lag_low] = pe_dot(target_ptr, target_ptr - start_lag - j, sf_length);
      const float *basis_ptr = target_ptr - (start_lag + lag_low);
      double energy = silk_inner_product_flp(basis_ptr, basis_ptr, sf_length) + 1e-3;
      en[k][0] = (float)energy;
      for (int i = 1; i < lag_diff; i++) {
        energy -= basis_ptr[sf_length - i] * (double)basis_ptr[sf_length - i];
        energy += basis_ptr[-i] * (double)basis_ptr[-i];
        en[k][i] = (float)energy;
      }
      target_ptr += sf_length;
    }
    const float contour_bias = 0.05f / lag;   // PE_FLATCONTOUR_BIAS
    target_ptr = &frame[20 * Fs_kHz];
    const double energy_tmp = silk_inner_product_flp(target_ptr, target_ptr, nb_subfr * sf_length) + 1.0;
    int lag_counter = 0;
    for (int d = start_lag; d <= end_lag; d++, lag_counter++) {
      for (int j = 0; j < nb3; j++) {
        double cross_corr = 0.0, energy = energy_tmp;
        for (int k = 0; k < nb_subfr; k++) {
          const int idx = cb3[k * cb3_size + j] - range[k * 2] + lag_counter;
          cross_corr += xc[k][idx];
          energy += en[k][idx];
        }
        float CCmax_new = 0.0f;
        if (cross_corr > 0.0) { CCmax_new = (float)(2 * cross_corr / energy); CCmax_new *= 1.0f - contour_bias * j; }
        if (CCmax_new > CCmax && (d + (int)ob_pe_cb3[j]) <= max_lag) { CCmax = CCmax_new; lag_new = d; CBimax = j; }
      }
    }
    for (int k = 0; k < nb_subfr; k++) {
      const int v = lag_new + cb3[k * cb3_size + CBimax], hi = 18 * Fs_kHz;
      pitch_out[k] = v < min_lag ? min_lag : v > hi ? hi : v;
    }
    *lagIndex = lag_new - min_lag;
    *contourIndex = CBimax;
  } else {
    for (int k = 0; k < nb_subfr; k++) {
      const int v = lag + Lag_CB_ptr[k * cbk_size + CBimax];
      pitch_out[k] = v < min_lag_8kHz ? min_lag_8kHz : v > 18 * 8 ? 18 * 8 : v;
    }
    *lagIndex = lag - min_lag_8kHz;
    *contourIndex = CBimax;
  }
  return 0;
}
OB_DEV void silk_find_pitch_lags_thread(const SilkPitchParams &p, int s) {
  ObSilkPitch c = p.frames[s];
  const int fs_kHz = p.fs_kHz, nb_subfr = p.nb_subfr, lpc_order = p.lpc_order, la_pitch = 2 * fs_kHz, frame_length = 5 * nb_subfr * fs_kHz, ltp_mem_length = 20 * fs_kHz,
            buf_len = la_pitch + frame_length + ltp_mem_length, win_length = (5 * nb_subfr + 4) * fs_kHz;
  const float *x_buf = p.x + (size_t)s * p.x_stride + p.x_off - ltp_mem_length, *x_buf_ptr = x_buf + buf_len - win_length;
  float *res = p.res + (size_t)s * p.res_stride;
  float auto_corr[17], A[16], refl_coef[16];
  {
    float Wsig[24 * 16];
    ns_sine_window(Wsig, x_buf_ptr, 1, la_pitch);
    for (int i = 0; i < win_length - 2 * la_pitch; i++) Wsig[la_pitch + i] = x_buf_ptr[la_pitch + i];
    ns_sine_window(Wsig + win_length - la_pitch, x_buf_ptr + win_length - la_pitch, 2, la_pitch);
    for (int i = 0; i < lpc_order + 1; i++) auto_corr[i] = (float)silk_inner_product_flp(Wsig, Wsig + i, win_length - i);
  }
  auto_corr[0] += auto_corr[0] * 1e-3f + 1;   // FIND_PITCH_WHITE_NOISE_FRACTION
  float res_nrg;
  {   // silk_schur_FLP
    double Cs[17][2];
    for (int i = 0; i <= lpc_order; i++) Cs[i][0] = Cs[i][1] = auto_corr[i];
    for (int i = 0; i < lpc_order; i++) {
      const double rc_tmp = -Cs[i + 1][0] / (Cs[0][1] > (double)1e-9f ? Cs[0][1] : (double)1e-9f);
      refl_coef[i] = (float)rc_tmp;
      for (int n = 0; n < lpc_order - i; n++) {
        const double Ctmp1 = Cs[n + i + 1][0], Ctmp2 = Cs[n][1];
        Cs[n + i + 1][0] = Ctmp1 + Ctmp2 * rc_tmp;
        Cs[n][1] = Ctmp2 + Ctmp1 * rc_tmp;
      }
    }
    res_nrg = (float)Cs[0][1];
  }
  c.predGain = auto_corr[0] / (res_nrg > 1.0f ? res_nrg : 1.0f);
  for (int i = 0; i < lpc_order; i++) {   // silk_k2a_FLP
    const float rck = refl_coef[i];
    for (int n = 0; n < (i + 1) >> 1; n++) {
      const float t1 = A[n], t2 = A[i - n - 1];
      A[n] = t1 + t2 * rck;
      A[i - n - 1] = t2 + t1 * rck;
    }
    A[i] = -rck;
  }
  ns_bwexpander(A, lpc_order, 0.99f);   // FIND_PITCH_BANDWIDTH_EXPANSION
  for (int i = lpc_order; i < buf_len; i++) {   // silk_LPC_analysis_filter_FLP
    const float *sp = &x_buf[i - 1];
    float pred = sp[0] * A[0];
    for (int n = 1; n < lpc_order; n++) pred = pred + sp[-n] * A[n];
    res[i] = sp[1] - pred;
  }
  for (int i = 0; i < lpc_order; i++) res[i] = 0.f;
  if (c.signalType != 0 && c.first_frame_after_reset == 0) {
    float thrhld = 0.6f;
    thrhld -= 0.004f * lpc_order;
    thrhld -= 0.1f * c.speech_activity_Q8 * (1.0f / 256.0f);
    thrhld -= 0.15f * (c.prevSignalType >> 1);
    thrhld -= 0.1f * c.input_tilt_Q15 * (1.0f / 32768.0f);
    for (int k = 0; k < 4; k++) c.pitchL[k] = 0;
    c.signalType = silk_pitch_analysis_core(res, c.pitchL, &c.lagIndex, &c.contourIndex, &c.LTPCorr, c.prevLag, c.pitchEstimationThreshold_Q16 / 65536.0f, thrhld,
                                            fs_kHz, p.complexity, nb_subfr) == 0 ? 2 : 1;
  } else {
    for (int k = 0; k < 4; k++) c.pitchL[k] = 0;
    c.lagIndex = 0; c.contourIndex = 0; c.LTPCorr = 0;
  }
  p.frames[s] = c;
}
}  // namespace ob
