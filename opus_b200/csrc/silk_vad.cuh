// silk_vad.cuh — the voice activity detector of a SILK frame, one THREAD per stream: silk_VAD_GetSA_Q8_c / silk_VAD_GetNoiseLevels
// (silk/VAD.c:82-360) with silk_ana_filt_bank_1 (silk/ana_filt_bank_1.c:39-74), silk_SQRT_APPROX (silk/Inlines.h:67-90), silk_sigm_Q15
// (silk/sigm_Q15.c:49-76), the VAD / DTX bookkeeping of silk_encode_do_VAD_FLP (silk/float/encode_frame_FLP.c:45-82) and silk_HP_variable_cutoff
// (silk/HP_variable_cutoff.c:39-77).  All integer, bit-exact.
//
// The reference runs the three analysis filter banks one after the other through a scratch array of 1.25 frames and then walks the lowest band
// backwards for the differentiator.  Every stage is causal, so here the cascade is STREAMED: each loop step takes 8 input samples (one 16-byte load
// when the stream's row is aligned), pushes them through the 0-8 kHz split (4 low, 4 high), the 0-4 kHz split (2 + 2) and the 0-2 kHz split (1 + 1),
// differentiates the 0-1 kHz sample against the one before, and adds the squares to four per-band accumulators that know their own subframe
// boundaries (the reference's subframe length is decimated_length >> 2, so at 12 kHz the bands drop different tails).  No scratch array, no local
// memory: the whole frame stays in registers and each input byte is read once.
#pragma once
#include "silk_ops.cuh"

namespace ob {

struct ObSilkVad {   // == OpusB200SilkVad (include/opus_b200.h): silk_VAD_state (silk/structs.h:74-85) + the silk_encoder_state fields read / written
  int32_t AnaState[2], AnaState1[2], AnaState2[2], XnrgSubfr[4], NrgRatioSmth_Q8[4], HPstate, NL[4], inv_NL[4], NoiseLevelBias[4], counter;
  int32_t speech_activity_Q8, input_tilt_Q15, input_quality_bands_Q15[4];
  int32_t signalType, noSpeechCounter, inDTX, VAD_flag;
  int32_t prevSignalType, prevLag, variable_HP_smth1_Q15;
};
struct SilkVadParams {
  ObSilkVad *frames;
  const int16_t *pcm;        // [n][pcm_stride], the frame_length samples the reference reads at inputBuf + 1
  const int32_t *activity;   // [n] decision of the Opus-level detector (0 = VAD_NO_ACTIVITY), or null = active everywhere
  int n, pcm_stride, frame_length, fs_kHz;
};

OB_DEV int32_t vad_add_pos_sat32(int32_t a, int32_t b) { const uint32_t s = (uint32_t)a + (uint32_t)b; return (s & 0x80000000u) ? 0x7FFFFFFF : (int32_t)s; }
OB_DEV int32_t vad_sqrt_approx(int32_t x) {
  if (x <= 0) return 0;
  const int lz = clz32((uint32_t)x), rot = 24 - lz;
  const uint32_t u = (uint32_t)x;
  const uint32_t r = rot == 0 ? u : rot < 0 ? ((u << (uint32_t)(-rot)) | (u >> (uint32_t)(32 + rot))) : ((u << (uint32_t)(32 - rot)) | (u >> (uint32_t)rot));
  const int32_t frac_Q7 = (int32_t)(r & 0x7f);
  int32_t y = (lz & 1) ? 32768 : 46214;
  y >>= (lz >> 1);
  return silk_smlawb_(y, y, 213 * frac_Q7);
}
OB_DEV int vad_sigm_Q15(int in_Q5) {
  // slope / positive / negative tables of silk/sigm_Q15.c:35-46, selected by the integer part
  const int neg = in_Q5 < 0, a = neg ? -in_Q5 : in_Q5;
  if (a >= 6 * 32) return neg ? 0 : 32767;
  const int ind = a >> 5;
  const int slope = ind == 0 ? 237 : ind == 1 ? 153 : ind == 2 ? 73 : ind == 3 ? 30 : ind == 4 ? 12 : 7;
  const int pos = ind == 0 ? 16384 : ind == 1 ? 23955 : ind == 2 ? 28861 : ind == 3 ? 31213 : ind == 4 ? 32178 : 32548;
  const int ngv = ind == 0 ? 16384 : ind == 1 ? 8812 : ind == 2 ? 3906 : ind == 3 ? 1554 : ind == 4 ? 589 : 219;
  return neg ? ngv - slope * (a & 0x1F) : pos + slope * (a & 0x1F);
}
// one step of silk_ana_filt_bank_1: an even / odd input pair → one low-band and one high-band sample; S0, S1 = the two all-pass states (Q10)
OB_DEV void vad_split(int32_t &S0, int32_t &S1, int even, int odd, int &lo, int &hi) {
  int32_t in32 = even * 1024, Y = in32 - S0, X = silk_smlawb_(Y, Y, -24290);
  const int32_t out_1 = S0 + X;
  S0 = in32 + X;
  in32 = odd * 1024; Y = in32 - S1; X = silk_smulwb(Y, 5394 << 1);
  const int32_t out_2 = S1 + X;
  S1 = in32 + X;
  lo = silk_sat16(silk_rshift_round(out_2 + out_1, 11));
  hi = silk_sat16(silk_rshift_round(out_2 - out_1, 11));
}
struct alignas(16) VadVec8 { int32_t x, y, z, w; };   // 8 int16 samples, one 16-byte load
struct VadBandAcc { int32_t xnrg, sum, last; int cnt, sub; };
OB_DEV void vad_acc(VadBandAcc &a, int x, int sl) {
  if (a.sub < 4) {   // the reference sums 4 subframes of sl samples; a band's tail past 4 * sl is not looked at
    const int32_t t = x >> 3;
    a.sum += t * t;
    if (++a.cnt == sl) {
      a.xnrg = vad_add_pos_sat32(a.xnrg, a.sub < 3 ? a.sum : a.sum >> 1);   // the last one is the look-ahead subframe, half weight
      a.last = a.sum; a.sum = 0; a.cnt = 0; a.sub++;
    }
  }
}
OB_DEV void vad_noise_levels(const int32_t *pX, ObSilkVad &c) {
  int min_coef = 0;
  if (c.counter < 1000) { min_coef = 32767 / ((c.counter >> 4) + 1); c.counter++; }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int32_t nl = c.NL[k];
    const int32_t nrg = vad_add_pos_sat32(pX[k], c.NoiseLevelBias[k]), inv_nrg = 0x7FFFFFFF / nrg;
    int coef;
    if (nrg > (int32_t)((uint32_t)nl << 3)) coef = 1024 >> 3;
    else if (nrg < nl) coef = 1024;
    else coef = silk_smulwb((int32_t)(((int64_t)inv_nrg * nl) >> 16), 1024 << 1);
    if (coef < min_coef) coef = min_coef;
    c.inv_NL[k] = silk_smlawb_(c.inv_NL[k], inv_nrg - c.inv_NL[k], coef);
    nl = 0x7FFFFFFF / c.inv_NL[k];
    c.NL[k] = nl < 0x00FFFFFF ? nl : 0x00FFFFFF;
  }
}

OB_DEV void silk_encode_do_vad_thread(const SilkVadParams &p, int s) {
  ObSilkVad c = p.frames[s];
  const int16_t *pIn = p.pcm + (size_t)s * p.pcm_stride;
  const int frame_length = p.frame_length, d3 = frame_length >> 3;
  // ---- filter, decimate, differentiate and sum the squares, 8 input samples per step ----
  VadBandAcc b0 = {c.XnrgSubfr[0], 0, 0, 0, 0}, b1 = {c.XnrgSubfr[1], 0, 0, 0, 0}, b2 = {c.XnrgSubfr[2], 0, 0, 0, 0}, b3 = {c.XnrgSubfr[3], 0, 0, 0, 0};
  const int sl0 = d3 >> 2, sl2 = (frame_length >> 2) >> 2, sl3 = (frame_length >> 1) >> 2;
  int32_t A0 = c.AnaState[0], A1 = c.AnaState[1], B0 = c.AnaState1[0], B1 = c.AnaState1[1], C0 = c.AnaState2[0], C1 = c.AnaState2[1];
  int hp = (int16_t)c.HPstate;   // the previous lowest-band sample, already halved
  const bool vec = ((uintptr_t)pIn & 15) == 0;
  for (int g = 0; g < d3; g++) {
    int x[8];
    if (vec) {
      const VadVec8 v = *(const VadVec8 *)(pIn + 8 * g);
      x[0] = (int16_t)(v.x & 0xFFFF); x[1] = v.x >> 16; x[2] = (int16_t)(v.y & 0xFFFF); x[3] = v.y >> 16;
      x[4] = (int16_t)(v.z & 0xFFFF); x[5] = v.z >> 16; x[6] = (int16_t)(v.w & 0xFFFF); x[7] = v.w >> 16;
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) x[i] = pIn[8 * g + i];
    }
    int l[4], h[4], ll[2], lh[2], lll, llh;
#pragma unroll
    for (int i = 0; i < 4; i++) { vad_split(A0, A1, x[2 * i], x[2 * i + 1], l[i], h[i]); vad_acc(b3, h[i], sl3); }     // 4-8 kHz
#pragma unroll
    for (int i = 0; i < 2; i++) { vad_split(B0, B1, l[2 * i], l[2 * i + 1], ll[i], lh[i]); vad_acc(b2, lh[i], sl2); }  // 2-4 kHz
    vad_split(C0, C1, ll[0], ll[1], lll, llh);
    vad_acc(b1, llh, sl0);                                                                                             // 1-2 kHz
    const int half = (int16_t)lll >> 1, y = (int16_t)(half - hp);   // the differentiator on 0-1 kHz, int16 wrap as the reference's in-place update
    hp = half;
    vad_acc(b0, y, sl0);
  }
  c.AnaState[0] = A0; c.AnaState[1] = A1; c.AnaState1[0] = B0; c.AnaState1[1] = B1; c.AnaState2[0] = C0; c.AnaState2[1] = C1;
  c.HPstate = hp;
  c.XnrgSubfr[0] = b0.last; c.XnrgSubfr[1] = b1.last; c.XnrgSubfr[2] = b2.last; c.XnrgSubfr[3] = b3.last;
  int32_t Xnrg[4] = {b0.xnrg, b1.xnrg, b2.xnrg, b3.xnrg};
  // ---- noise levels, signal-to-noise ratios, tilt ----
  vad_noise_levels(Xnrg, c);
  int32_t ratio_Q8[4], sumSquared = 0, input_tilt = 0;
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const int32_t speech_nrg = Xnrg[b] - c.NL[b];
    if (speech_nrg > 0) {
      if ((Xnrg[b] & 0xFF800000) == 0) ratio_Q8[b] = (int32_t)((uint32_t)Xnrg[b] << 8) / (c.NL[b] + 1);
      else ratio_Q8[b] = Xnrg[b] / ((c.NL[b] >> 8) + 1);
      int SNR_Q7 = silk_lin2log(ratio_Q8[b]) - 8 * 128;
      sumSquared += (int32_t)(int16_t)SNR_Q7 * (int32_t)(int16_t)SNR_Q7;
      if (speech_nrg < (1 << 20)) SNR_Q7 = silk_smulwb((int32_t)((uint32_t)vad_sqrt_approx(speech_nrg) << 6), SNR_Q7);
      const int tilt_w = b == 0 ? 30000 : b == 1 ? 6000 : -12000;
      input_tilt = silk_smlawb_(input_tilt, tilt_w, SNR_Q7);
    } else ratio_Q8[b] = 256;
  }
  sumSquared = sumSquared / 4;
  const int pSNR_dB_Q7 = (int16_t)(3 * vad_sqrt_approx(sumSquared));
  int SA_Q15 = vad_sigm_Q15(silk_smulwb(45000, pSNR_dB_Q7) - 128);
  c.input_tilt_Q15 = (vad_sigm_Q15(input_tilt) - 16384) * 2;
  int32_t speech_nrg = 0;
#pragma unroll
  for (int b = 0; b < 4; b++) speech_nrg += (b + 1) * ((Xnrg[b] - c.NL[b]) >> 4);
  if (frame_length == 20 * p.fs_kHz) speech_nrg >>= 1;
  if (speech_nrg <= 0) SA_Q15 >>= 1;
  else if (speech_nrg < 16384) SA_Q15 = silk_smulwb(32768 + vad_sqrt_approx((int32_t)((uint32_t)speech_nrg << 16)), SA_Q15);
  c.speech_activity_Q8 = (SA_Q15 >> 7) < 255 ? (SA_Q15 >> 7) : 255;
  int32_t smooth_coef_Q16 = silk_smulwb(4096, silk_smulwb(SA_Q15, SA_Q15));
  if (frame_length == 10 * p.fs_kHz) smooth_coef_Q16 >>= 1;
#pragma unroll
  for (int b = 0; b < 4; b++) {
    c.NrgRatioSmth_Q8[b] = silk_smlawb_(c.NrgRatioSmth_Q8[b], ratio_Q8[b] - c.NrgRatioSmth_Q8[b], smooth_coef_Q16);
    const int SNR_Q7 = 3 * (silk_lin2log(c.NrgRatioSmth_Q8[b]) - 8 * 128);
    c.input_quality_bands_Q15[b] = vad_sigm_Q15((SNR_Q7 - 16 * 128) >> 4);
  }
  // ---- silk_encode_do_VAD_FLP: the Opus-level override and the VAD / DTX flags ----
  const int thr = 13;   // SILK_FIX_CONST(SPEECH_ACTIVITY_DTX_THRES = 0.05, 8)
  const int activity = p.activity ? p.activity[s] : 1;
  if (activity == 0 && c.speech_activity_Q8 >= thr) c.speech_activity_Q8 = thr - 1;
  if (c.speech_activity_Q8 < thr) {
    c.signalType = 0;   // TYPE_NO_VOICE_ACTIVITY
    c.noSpeechCounter++;
    if (c.noSpeechCounter <= 10) c.inDTX = 0;                                               // NB_SPEECH_FRAMES_BEFORE_DTX
    else if (c.noSpeechCounter > 20 + 10) { c.noSpeechCounter = 10; c.inDTX = 0; }          // MAX_CONSECUTIVE_DTX
    c.VAD_flag = 0;
  } else {
    c.noSpeechCounter = 0; c.inDTX = 0; c.signalType = 1; c.VAD_flag = 1;                   // TYPE_UNVOICED until the pitch analysis says otherwise
  }
  p.frames[s] = c;
}

// silk_HP_variable_cutoff: the first smoother of the high-pass cutoff, driven by the previous frame's pitch lag (the second smoother and the filter
// itself are opus_encode_native's, src/opus_encoder.c:1962-1985)
struct SilkHpCutoffParams { ObSilkVad *frames; int n, fs_kHz; };
OB_DEV void silk_hp_variable_cutoff_thread(const SilkHpCutoffParams &p, int s) {
  ObSilkVad &c = p.frames[s];
  if (c.prevSignalType != 2) return;   // TYPE_VOICED
  const int32_t pitch_freq_Hz_Q16 = (int32_t)((uint32_t)(p.fs_kHz * 1000) << 16) / (int16_t)c.prevLag;
  int32_t pitch_freq_log_Q7 = silk_lin2log(pitch_freq_Hz_Q16) - (16 << 7);
  const int quality_Q15 = c.input_quality_bands_Q15[0];
  pitch_freq_log_Q7 = silk_smlawb_(pitch_freq_log_Q7, silk_smulwb((int32_t)((uint32_t)(-quality_Q15) << 2), quality_Q15),
                                   pitch_freq_log_Q7 - (silk_lin2log(60 << 16) - (16 << 7)));   // VARIABLE_HP_MIN_CUTOFF_HZ
  int32_t delta_freq_Q7 = pitch_freq_log_Q7 - (c.variable_HP_smth1_Q15 >> 8);
  if (delta_freq_Q7 < 0) delta_freq_Q7 *= 3;
  delta_freq_Q7 = delta_freq_Q7 > 51 ? 51 : delta_freq_Q7 < -51 ? -51 : delta_freq_Q7;          // SILK_FIX_CONST(VARIABLE_HP_MAX_DELTA_FREQ = 0.4, 7)
  int32_t v = silk_smlawb_(c.variable_HP_smth1_Q15, (int32_t)(int16_t)c.speech_activity_Q8 * (int32_t)(int16_t)delta_freq_Q7, 6554);   // COEF1 0.1 in Q16
  const int32_t lo = silk_lin2log(60) << 8, hi = silk_lin2log(100) << 8;
  c.variable_HP_smth1_Q15 = v < lo ? lo : v > hi ? hi : v;
}

}  // namespace ob
