// celt_enc_chains.cuh — the per-sample RECURRENCES of the CELT encoder's time-domain analysis, one THREAD per stream.
//
// celt_encode_with_ec runs three kinds of work on the pre-emphasised input before the transform: elementwise / lane-parallel work
// (pre-emphasis, pitch correlations, the comb filter, the MDCT), and sums whose order the reference fixes or that are recurrences by
// nature: transient_analysis (celt/celt_encoder.c:267-466: a 2-pole high-pass, a forward and a backward smoother) and tone_detect's
// LPC sums (:1305-1402).  In a CTA-per-stream kernel the latter are one lane's chain with 31 lanes idle (38 % of the front end's
// instructions at 1/32 lane use, profiles/r01_ncu_celt_enc_analysis_c5.txt).  Here they run as 32 independent streams per warp:
// every lane walks its own stream's frame once, the PCM staged through a transposed shared-memory tile (coalesced rows in, a private
// column per thread), the two per-channel smoother tracks and the channel sum in a transposed scratch in HBM ([index][lane]: every
// access of the warp is one line).  Output: the time-domain part of ObEncFrameInfo; ob_celt_enc_analysis then skips these steps.
#pragma once
#include "celt_ops.cuh"

namespace ob {

#define OB_CH_TS 65           // tile row: 64 floats (+1: conflict-free private rows)
#define OB_CH_LEN 1080        // N + overlap at 20 ms
#define OB_CH_SCRATCH_FLOATS ((OB_CH_LEN + 2 * (OB_CH_LEN / 2)) * 32)   // per block of 32 streams: channel sum + two smoother tracks

// tone_lpc (celt/celt_encoder.c:1305-1360) over the stored channel sum; x(i) = xs[i * 32]
OB_DEV int chain_tone_lpc(const float *xs, int len, int delay, float *lpc) {
  float r00 = 0.f, r01 = 0.f, r02 = 0.f;
#pragma unroll 4
  for (int i = 0; i < len - 2 * delay; i++) {
    const float a = xs[i * 32];
    r00 += a * a;
    r01 += a * xs[(i + delay) * 32];
    r02 += a * xs[(i + 2 * delay) * 32];
  }
  float e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll 1
  for (int i = 0; i < delay; i++) {
    const float a = xs[(len + i - 2 * delay) * 32], b = xs[(len + i - delay) * 32], c = xs[i * 32], d = xs[(i + delay) * 32];
    e11 += a * a - c * c;
    e22 += b * b - d * d;
    e12 += a * b - c * d;
  }
  float r11 = r00 + e11;
  const float r22 = r11 + e22;
  float r12 = r01 + e12;
  {
    const float R00 = r00 + r22, R01 = r01 + r12, R11 = 2 * r11, R02 = 2 * r02, R12 = r12 + r01;
    r00 = R00; r01 = R01; r11 = R11; r02 = R02; r12 = R12;
  }
  const float den = r00 * r11 - r01 * r01;
  if (den < .001f * (r00 * r11)) return 1;
  const float num1 = r02 * r11 - r01 * r12;
  if (num1 >= den) lpc[1] = 1.f;
  else if (num1 <= -den) lpc[1] = -1.f;
  else lpc[1] = f_div(num1, den);
  const float num0 = r00 * r12 - r02 * r01;
  if (.5f * num0 >= den) lpc[0] = 1.999999f;
  else if (.5f * num0 <= -den) lpc[0] = -1.999999f;
  else lpc[0] = f_div(num0, den);
  return 0;
}

// one warp = 32 streams [first_stream, first_stream + 32).  tile: 32 * OB_CH_TS floats of shared memory; scratch: this block's
// OB_CH_SCRATCH_FLOATS floats in HBM.
OB_DEV void enc_chains_warp(const EncAnalysisParams &p, int first_stream, float *tile, float *scratch) {
  const int lane = lane_id();
  const int stream = first_stream + lane;
  const bool live = stream < p.n_streams;
  const int LM = p.LM, N = OB_SHORT_N << LM, CC = p.CC ? p.CC : p.C, up = p.upsample, len = N + OB_OVERLAP, len2 = len / 2;
  const int C = (p.ctl && live) ? p.ctl[stream].C : p.C;
  float *xs = scratch + lane;                                  // channel sum, [len][32]
  float *trk0 = scratch + OB_CH_LEN * 32 + lane;               // smoother track of channel 0, [len2][32]
  float *trk1 = trk0 + (OB_CH_LEN / 2) * 32;
  ObEncAnalysisState *st = live ? p.state + stream : nullptr;
  const float coef0 = 0.85000610f, fd = .0625f;
  // per-channel recurrences
  float hp0[2] = {0.f, 0.f}, hp1[2] = {0.f, 0.f}, mean[2] = {0.f, 0.f}, fwd[2] = {0.f, 0.f}, yprev[2] = {0.f, 0.f};
  float mx_new = 0.f, mx_tail = 0.f;
  EncPcmView pv;
  pv.pcm = nullptr; pv.CC = CC; pv.fade = 0; pv.clear = 0; pv.inc = up; pv.g1 = pv.g2 = 0.f;
  if (p.ctl && live) { const ObOpusFrameCtl &k = p.ctl[stream]; pv.fade = k.fade; pv.clear = k.clear; pv.g1 = k.fade_g1; pv.g2 = k.fade_g2; }
  // one sample of channel c at position i of the analysis window [overlap | frame]
#define OB_CHAIN_STEP(c, i, v)                                                              \
  {                                                                                         \
    const float y_ = (i) < 12 ? 0.f : hp0[c] + (v), m00_ = hp0[c];                          \
    hp0[c] = hp0[c] - (v) + .5f * hp1[c];                                                   \
    hp1[c] = (v)-m00_;                                                                      \
    if ((i)&1) {                                                                            \
      const float x2_ = yprev[c] * yprev[c] + y_ * y_;                                      \
      mean[c] += x2_;                                                                       \
      fwd[c] = x2_ + (1.f - fd) * fwd[c];                                                   \
      ((c) ? trk1 : trk0)[((i) >> 1) * 32] = fd * fwd[c];                                   \
    } else yprev[c] = y_;                                                                   \
  }
  if (live) {
    // the window starts with the UNFILTERED tail of the previous frame (celt/celt_encoder.c:2014)
    const bool pf_path = p.enc_state != nullptr;
    for (int i = 0; i < OB_OVERLAP; i++) {
      float v0, v1 = 0.f;
      if (pf_path) {
        const float *pm = p.prefilter_mem + (size_t)stream * CC * kPfMaxPeriod;
        const int idx = (st->pf_head + kPfMaxPeriod - OB_OVERLAP + i) & (kPfMaxPeriod - 1);
        v0 = pm[idx];
        if (CC == 2) v1 = pm[kPfMaxPeriod + idx];
      } else {
        v0 = st->overlap_mem[0][i];
        if (CC == 2) v1 = st->overlap_mem[1][i];
      }
      OB_CHAIN_STEP(0, i, v0);
      if (CC == 2) OB_CHAIN_STEP(1, i, v1);
      xs[i * 32] = CC == 2 ? v0 + v1 : v0;
    }
  }
  // the frame: N / up input samples per channel, staged 64 floats per stream at a time
  const int total = (N / up) * CC;
  float xprev[2] = {0.f, 0.f};   // 32768 * previous input sample (pre-emphasis memory as coef0 * xprev; the first one comes from the state)
  float m_first[2] = {live ? st->preemph_mem[0] : 0.f, (live && CC == 2) ? st->preemph_mem[1] : 0.f};
  int pos = 0;   // input sample index (per channel) of the next tile element
  for (int base = 0; base < total; base += 64) {
    const int cnt = OB_IMIN(64, total - base);
    for (int r = 0; r < 32; r++) {
      const int sr = first_stream + r;
      if (sr >= p.n_streams) break;
      const float *src = p.pcm + (size_t)sr * total + base;
      for (int j = lane; j < cnt; j += 32) tile[r * OB_CH_TS + j] = src[j];
    }
    wsync();
    if (live) {
      const float *row = tile + lane * OB_CH_TS;
      for (int j = 0; j < cnt; j += CC, pos++) {
        float s0 = row[j], s1 = CC == 2 ? row[j + 1] : 0.f;
        if (pv.clear) { s0 = 0.f; s1 = 0.f; }
        else if (pv.fade) {   // stereo_fade (src/opus_encoder.c:553-580), as enc_pcm_at
          float g = pv.g2;
          if (pos < OB_OVERLAP / pv.inc) {
            float w = ob_window[pos * pv.inc];
            w = w * w;
            g = w * pv.g2 + (1.f - w) * pv.g1;
          }
          float diff = .5f * (s0 - s1);
          diff = g * diff;
          s0 = s0 - diff;
          s1 = s1 + diff;
        }
        {   // celt_maxabs over the fresh part and over the part overlapping the next frame (celt/celt_encoder.c:1963-1966)
          const float a0 = fabsf(s0), a1 = CC == 2 ? fabsf(s1) : 0.f, a = fmaxf(a0, a1);
          if (pos * up < N - OB_OVERLAP) mx_new = fmaxf(mx_new, a);
          else mx_tail = fmaxf(mx_tail, a);
        }
        // celt_preemphasis (celt/celt_encoder.c:557-646): in = x - m, m = coef0 * x; with upsample != 1 the input is zero-stuffed first
        const float x0 = f_mul(32768.f, s0), x1 = f_mul(32768.f, s1);
        for (int u = 0; u < up; u++) {
          const int i = OB_OVERLAP + pos * up + u;
          const float xa = u == 0 ? x0 : 0.f, xb = u == 0 ? x1 : 0.f;
          const float ma = (pos == 0 && u == 0) ? m_first[0] : f_mul(coef0, xprev[0]);
          const float mb = (pos == 0 && u == 0) ? m_first[1] : f_mul(coef0, xprev[1]);
          const float v0 = f_add(xa, -ma), v1 = f_add(xb, -mb);
          xprev[0] = xa; xprev[1] = xb;
          OB_CHAIN_STEP(0, i, v0);
          if (CC == 2) OB_CHAIN_STEP(1, i, v1);
          xs[i * 32] = CC == 2 ? v0 + v1 : v0;
        }
      }
    }
    wsync();
  }
#undef OB_CHAIN_STEP
  if (!live || !p.info) return;
  // ---- transient_analysis: backward smoother, the masking metric (celt/celt_encoder.c:380-436) ----
  int mask_metric = 0, tf_chan = 0;
  if (p.complexity >= 1) {
    for (int c = 0; c < CC; c++) {
      float *trk = c ? trk1 : trk0;
      float mem0 = 0.f, maxE = 0.f;
#pragma unroll 4
      for (int i = len2 - 1; i >= 0; i--) {
        mem0 = trk[i * 32] + 0.875f * mem0;
        trk[i * 32] = 0.125f * mem0;
        maxE = fmaxf(maxE, 0.125f * mem0);
      }
      const float mn = (float)sqrt((double)(mean[c] * maxE) * .5 * len2);
      const float norm = len2 / (1e-15f + mn);
      int unmask = 0;
#pragma unroll 2
      for (int i = 12; i < len2 - 5; i += 4) {
        const double v = floor((double)(64 * norm * (trk[i * 32] + 1e-15f)));
        const int id = (int)(v < 127 ? (v > 0 ? v : 0) : 127);
        unmask += ob_transient_inv_table[id];
      }
      unmask = 64 * unmask * 4 / (6 * (len2 - 17));
      if (c == 0) mask_metric = unmask;
      else if (unmask > mask_metric) { mask_metric = unmask; tf_chan = 1; }
    }
  }
  // ---- tone_detect (celt/celt_encoder.c:1362-1402) ----
  float lpc[2] = {0.f, 0.f};
  int delay = 1;
  int fail = chain_tone_lpc(xs, len, delay, lpc);
  while (delay <= 48000 / 3000 && (fail || (lpc[0] > 1.f && lpc[1] < 0))) {
    delay *= 2;
    fail = chain_tone_lpc(xs, len, delay, lpc);
  }
  ObEncFrameInfo fi;
  if (!fail && lpc[0] * lpc[0] + 3.999999f * lpc[1] < 0) {
    fi.toneishness = -lpc[1];
    fi.tone_freq = (float)(acos((double)(.5f * lpc[0])) / delay);
  } else {
    fi.tone_freq = -1.f;
    fi.toneishness = 0.f;
  }
  const float sample_max = fmaxf(fmaxf(st->overlap_max, mx_new), mx_tail);
  fi.sample_max = sample_max;
  fi.tf_estimate = 0.f; fi.tf_chan = 0; fi.is_transient = 0; fi.patched = 0; fi.pf_on = 0; fi.pitch_index = 15; fi.qg = 0; fi.pf_tapset = 0; fi.gain1 = 0.f;
  fi.pad[0] = fi.pad[1] = fi.pad[2] = 0;
  if (p.complexity >= 1) {   // the cross-channel tail of transient_analysis (celt/celt_encoder.c:437-461)
    fi.tf_chan = tf_chan;
    int is_tr = mask_metric > 200;
    if (fi.toneishness > .98f && fi.tone_freq < 0.026f) { is_tr = 0; mask_metric = 0; }
    const float tf_max = fmaxf(0.f, (float)sqrt((double)(27 * mask_metric)) - 42.f);
    const double a = (double)((float)0.0069 * fminf(163.f, tf_max)) - 0.139;
    fi.tf_estimate = (float)sqrt(a > 0 ? a : 0.0);
    fi.is_transient = is_tr;
  }
  if (CC == 2 && C == 1) fi.tf_chan = 0;   // celt/celt_encoder.c:2095-2096
  fi.toneishness = fminf(fi.toneishness, 1.f - fi.tf_estimate);   // celt/celt_encoder.c:2026
  const int silence = sample_max <= 1.f / (float)(1 << p.lsb_depth);
  fi.short_blocks = fi.is_transient && LM > 0 && !silence;
  p.info[stream] = fi;
  st->overlap_max = mx_tail;
}

}  // namespace ob
