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
// per block of 32 streams, [index][lane]: channel sum (1080), two smoother tracks (2 x 540; later the Syy track of find_best_pitch), the unfiltered
// frame of each channel (2 x 960), the 2:1 decimated search signal before / after whitening (2 x 992; the first doubles as yy_lookup),
// the pitch correlations (512)
#define OB_CH_OFF_TRK (OB_CH_LEN)
#define OB_CH_OFF_FR (OB_CH_OFF_TRK + 2 * (OB_CH_LEN / 2))
#define OB_CH_OFF_XLP (OB_CH_OFF_FR + 2 * OB_MAXN)
#define OB_CH_OFF_XLP2 (OB_CH_OFF_XLP + 992)
#define OB_CH_OFF_XC (OB_CH_OFF_XLP2 + 992)
#define OB_CH_SCRATCH_FLOATS ((OB_CH_OFF_XC + 512) * 32)
#define OB_CH_IN_STRIDE (OB_MAXN + OB_OVERLAP)   // in_buf: [stream][channel][OB_CH_IN_STRIDE]

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

// ---- run_prefilter (celt/celt_encoder.c:1404-1602) for one stream on one thread: pitch_downsample / pitch_search / remove_doubling
// (celt/pitch.c:105-575), the gain rules, comb_filter (celt/celt.c:238-313) and the cancel rule, in the reference's own (serial) order.
// T(buf, i): element i of a transposed per-stream array (the pointer is already offset by the lane).
#define OB_T32(buf, i) (buf)[(i) * 32]
struct ChPre { const float *ring; const float *fr; int head; };   // pre[] = [1024 history samples in a ring | the frame], unfiltered
OB_DEV float ch_at(const ChPre &p, int j) { return j < kPfMaxPeriod ? p.ring[(p.head + j) & (kPfMaxPeriod - 1)] : OB_T32(p.fr, j - kPfMaxPeriod); }
// find_best_pitch (celt/pitch.c:45-103); y(j) = y[j * ystride] (transposed), xcorr transposed
OB_DEV void chain_find_best_pitch(const float *xcorr, const float *y, int ystride, int len, int max_pitch, int *best_pitch) {
  float Syy = 1.f, bn0 = -1.f, bn1 = -1.f, bd0 = 0.f, bd1 = 0.f;
  int bp0 = 0, bp1 = 1;
#pragma unroll 4
  for (int j = 0; j < len; j++) { const float v = OB_T32(y, j * ystride); Syy = Syy + v * v; }
#pragma unroll 2
  for (int i = 0; i < max_pitch; i++) {
    const float xc = OB_T32(xcorr, i);
    if (xc > 0) {
      float x16 = xc;
      x16 *= 1e-12f;
      const float num = x16 * x16;
      if (num * bd1 > bn1 * Syy) {
        if (num * bd0 > bn0 * Syy) { bn1 = bn0; bd1 = bd0; bp1 = bp0; bn0 = num; bd0 = Syy; bp0 = i; }
        else { bn1 = num; bd1 = Syy; bp1 = i; }
      }
    }
    const float a = OB_T32(y, (i + len) * ystride), b = OB_T32(y, i * ystride);
    Syy += a * a - b * b;
    Syy = Syy > 1.f ? Syy : 1.f;
  }
  best_pitch[0] = bp0;
  best_pitch[1] = bp1;
}
OB_DEV float chain_dot(const float *x, const float *y, int n) {   // celt_inner_prod_c: index order
  float s = 0.f;
#pragma unroll 4
  for (int i = 0; i < n; i++) s = s + OB_T32(x, i) * OB_T32(y, i);
  return s;
}
OB_DEV void chain_dual_dot(const float *x, const float *y1, const float *y2, int n, float *xy1, float *xy2) {
  float a = 0.f, b = 0.f;
#pragma unroll 4
  for (int i = 0; i < n; i++) { const float xv = OB_T32(x, i); a = a + xv * OB_T32(y1, i); b = b + xv * OB_T32(y2, i); }
  *xy1 = a; *xy2 = b;
}
// the pitch search of run_prefilter (complexity >= 5): returns gain1 and the pitch
OB_DEVNI float chain_pitch_analysis(const ChPre *pre, int C, int N, float *xlp, float *xlp2, float *xc, float *trk, int prev_period, float prev_gain,
                                    int loss_rate, int *pitch_out) {
  const int max_period = kPfMaxPeriod, min_period = kPfMinPeriod, len = (max_period + N) >> 1;
  // pitch_downsample (celt/pitch.c:140-228)
#pragma unroll 2
  for (int i = 0; i < len; i++) {
    float v = i == 0 ? .25f * ch_at(pre[0], 1) + .5f * ch_at(pre[0], 0)
                     : .25f * ch_at(pre[0], 2 * i - 1) + .25f * ch_at(pre[0], 2 * i + 1) + .5f * ch_at(pre[0], 2 * i);
    if (C == 2)
      v += i == 0 ? .25f * ch_at(pre[1], 1) + .5f * ch_at(pre[1], 0)
                  : .25f * ch_at(pre[1], 2 * i - 1) + .25f * ch_at(pre[1], 2 * i + 1) + .5f * ch_at(pre[1], 2 * i);
    OB_T32(xlp, i) = v;
  }
  float ac[5];
  {   // _celt_autocorr over 5 lags (celt/celt_lpc.c:284-370): xcorr over fastN, then the tail
    const int fastN = len - 4;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll 2
    for (int i = 0; i < fastN; i++) {
      const float x = OB_T32(xlp, i);
      a0 = a0 + x * x;
      a1 = a1 + x * OB_T32(xlp, i + 1);
      a2 = a2 + x * OB_T32(xlp, i + 2);
      a3 = a3 + x * OB_T32(xlp, i + 3);
      a4 = a4 + x * OB_T32(xlp, i + 4);
    }
    ac[0] = a0; ac[1] = a1; ac[2] = a2; ac[3] = a3; ac[4] = a4;
    for (int k = 0; k <= 4; k++) {
      float d = 0.f;
      for (int i = k + fastN; i < len; i++) d = d + OB_T32(xlp, i) * OB_T32(xlp, i - k);
      ac[k] += d;
    }
  }
  float l0, l1, l2, l3, l4;
  {   // _celt_lpc(4), bandwidth expansion, the extra zero
    float lpc[4] = {0.f, 0.f, 0.f, 0.f};
    ac[0] *= 1.0001f;
    for (int i = 1; i <= 4; i++) ac[i] -= ac[i] * (.008f * i) * (.008f * i);
    float error = ac[0];
    if (ac[0] > 1e-10f) {
      for (int i = 0; i < 4; i++) {
        float rr = 0.f;
        for (int j = 0; j < i; j++) rr += lpc[j] * ac[i - j];
        rr += ac[i + 1];
        const float r = -(rr / error);
        lpc[i] = r;
        for (int j = 0; j < (i + 1) >> 1; j++) {
          const float t1 = lpc[j], t2 = lpc[i - 1 - j];
          lpc[j] = t1 + r * t2;
          lpc[i - 1 - j] = t2 + r * t1;
        }
        error = error - (r * r) * error;
        if (error <= .001f * ac[0]) break;
      }
    }
    float tmp = 1.f;
    for (int i = 0; i < 4; i++) { tmp = .9f * tmp; lpc[i] = lpc[i] * tmp; }
    l0 = lpc[0] + .8f; l1 = lpc[1] + .8f * lpc[0]; l2 = lpc[2] + .8f * lpc[1]; l3 = lpc[3] + .8f * lpc[2]; l4 = .8f * lpc[3];
  }
  {   // celt_fir5
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll 2
    for (int i = 0; i < len; i++) {
      const float x = OB_T32(xlp, i);
      float sum = x;
      sum = sum + l0 * m0;
      sum = sum + l1 * m1;
      sum = sum + l2 * m2;
      sum = sum + l3 * m3;
      sum = sum + l4 * m4;
      m4 = m3; m3 = m2; m2 = m1; m1 = m0; m0 = x;
      OB_T32(xlp2, i) = sum;
    }
  }
  // pitch_search(pitch_buf + max_period/2, pitch_buf, N, max_period - 3*min_period) (celt/pitch.c:307-425)
  const float *pb = xlp2, *xl = xlp2 + (max_period >> 1) * 32;
  const int mp = max_period - 3 * min_period, len4 = N >> 2, mp4 = mp >> 2, len2 = N >> 1, mp2 = mp >> 1;
  {   // coarse stage, both signals decimated by 2 again; four lags per pass over j like xcorr_kernel (each sum stays in index order)
    int i = 0;
    for (; i + 3 < mp4; i += 4) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      float y0 = OB_T32(pb, 2 * i), y1 = OB_T32(pb, 2 * (i + 1)), y2 = OB_T32(pb, 2 * (i + 2));
#pragma unroll 4
      for (int j = 0; j < len4; j++) {
        const float x = OB_T32(xl, 2 * j), y3 = OB_T32(pb, 2 * (i + j + 3));
        s0 = s0 + x * y0; s1 = s1 + x * y1; s2 = s2 + x * y2; s3 = s3 + x * y3;
        y0 = y1; y1 = y2; y2 = y3;
      }
      OB_T32(xc, i) = s0; OB_T32(xc, i + 1) = s1; OB_T32(xc, i + 2) = s2; OB_T32(xc, i + 3) = s3;
    }
    for (; i < mp4; i++) {
      float s = 0.f;
      for (int j = 0; j < len4; j++) s = s + OB_T32(xl, 2 * j) * OB_T32(pb, 2 * (i + j));
      OB_T32(xc, i) = s;
    }
  }
  int bp[2];
  chain_find_best_pitch(xc, pb, 2, len4, mp4, bp);
  for (int i = 0; i < mp2; i++) {
    const int d0 = i - 2 * bp[0], d1 = i - 2 * bp[1];
    float v = 0.f;
    if (!((d0 > 2 || d0 < -2) && (d1 > 2 || d1 < -2))) {
      const float sdot = chain_dot(xl, pb + i * 32, len2);
      v = sdot > -1.f ? sdot : -1.f;
    }
    OB_T32(xc, i) = v;
  }
  chain_find_best_pitch(xc, pb, 1, len2, mp2, bp);
  int offset = 0;
  if (bp[0] > 0 && bp[0] < mp2 - 1) {
    const float aa = OB_T32(xc, bp[0] - 1), bb = OB_T32(xc, bp[0]), cc = OB_T32(xc, bp[0] + 1);
    if ((cc - aa) > .7f * (bb - aa)) offset = 1;
    else if ((aa - cc) > .7f * (bb - cc)) offset = -1;
  }
  int pitch_index = max_period - (2 * bp[0] - offset);
  // remove_doubling(pitch_buf, max_period, min_period, N, &pitch_index, prev_period, prev_gain) (celt/pitch.c:438-575)
  float pg;
  {
    const int maxp = max_period / 2, minp = min_period / 2, Nh = N / 2, prevp = prev_period / 2;
    const float *x = pb + maxp * 32;
    float *yyl = xlp;   // (the unwhitened signal is dead)
    (void)trk;
    int T0 = pitch_index / 2;
    if (T0 >= maxp) T0 = maxp - 1;
    const int second_check[16] = {0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2};
    float xx, xy, xy2;
    chain_dual_dot(x, x, x - T0 * 32, Nh, &xx, &xy);
    OB_T32(yyl, 0) = xx;
    float yy = xx;
#pragma unroll 2
    for (int i = 1; i <= maxp; i++) {
      const float a = OB_T32(x, -i), b = OB_T32(x, Nh - i);
      yy = yy + a * a - b * b;
      OB_T32(yyl, i) = yy > 0.f ? yy : 0.f;
    }
    yy = OB_T32(yyl, T0);
    float best_xy = xy, best_yy = yy;
    int T = T0;
    float g, g0;
    g = g0 = xy / (float)sqrt((double)(1 + xx * yy));
    for (int k = 2; k <= 15; k++) {
      const int T1 = (unsigned)(2 * T0 + k) / (unsigned)(2 * k);
      if (T1 < minp) break;
      int T1b;
      if (k == 2) T1b = T1 + T0 > maxp ? T0 : T0 + T1;
      else T1b = (unsigned)(2 * second_check[k] * T0 + k) / (unsigned)(2 * k);
      chain_dual_dot(x, x - T1 * 32, x - T1b * 32, Nh, &xy, &xy2);
      xy = .5f * (xy + xy2);
      yy = .5f * (OB_T32(yyl, T1) + OB_T32(yyl, T1b));
      const float g1 = xy / (float)sqrt((double)(1 + xx * yy));
      float cont;
      const int dT = T1 > prevp ? T1 - prevp : prevp - T1;
      if (dT <= 1) cont = prev_gain;
      else if (dT <= 2 && 5 * k * k < T0) cont = .5f * prev_gain;
      else cont = 0.f;
      float thresh = fmaxf(.3f, .7f * g0 - cont);
      if (T1 < 3 * minp) thresh = fmaxf(.4f, .85f * g0 - cont);
      else if (T1 < 2 * minp) thresh = fmaxf(.5f, .9f * g0 - cont);
      if (g1 > thresh) { best_xy = xy; best_yy = yy; T = T1; g = g1; }
    }
    if (T < minp * 2) {
      const int T1 = T * 5 / 8, T2 = T * 6 / 8;
      chain_dual_dot(x, x - T1 * 32, x - T2 * 32, Nh, &xy, &xy2);
      const float g1 = xy / (float)sqrt((double)(1 + xx * OB_T32(yyl, T1))), g2 = xy2 / (float)sqrt((double)(1 + xx * OB_T32(yyl, T2)));
      if (g1 >= g || g2 >= g) g = 0.f;
    }
    best_xy = best_xy > 0.f ? best_xy : 0.f;
    pg = best_yy <= best_xy ? 1.f : best_xy / (best_yy + 1);
    float c3[3];
    for (int k = 0; k < 3; k++) c3[k] = chain_dot(x, x - (T + k - 1) * 32, Nh);
    int off2 = 0;
    if ((c3[2] - c3[0]) > .7f * (c3[1] - c3[0])) off2 = 1;
    else if ((c3[0] - c3[2]) > .7f * (c3[1] - c3[2])) off2 = -1;
    if (pg > g) pg = g;
    int Tout = 2 * T + off2;
    if (Tout < min_period) Tout = min_period;
    if (Tout > max_period - 2) Tout = max_period - 2;
    pitch_index = Tout;
  }
  float g1 = .7f * pg;
  if (loss_rate > 2) g1 = .5f * g1;
  if (loss_rate > 4) g1 = .5f * g1;
  if (loss_rate > 8) g1 = 0.f;
  *pitch_out = pitch_index;
  return g1;
}
// comb_filter (celt/celt.c:238-313) at one sample: y[i] for x = pre[] + base, i in [0, N)
struct ChComb { int T0, T1, overlap; float g00, g01, g02, g10, g11, g12; int copy, g1zero; };
OB_DEV ChComb chain_comb_setup(int T0, int T1, float g0, float g1, int tapset0, int tapset1, int overlap) {
  ChComb k;
  k.copy = (g0 == 0 && g1 == 0);
  k.T0 = T0 > kPfMinPeriod ? T0 : kPfMinPeriod;
  k.T1 = T1 > kPfMinPeriod ? T1 : kPfMinPeriod;
  const float gt0[3] = {0.3066406250f, 0.4638671875f, 0.7998046875f}, gt1[3] = {0.2170410156f, 0.2680664062f, 0.1000976562f};
  k.g00 = g0 * (tapset0 == 0 ? gt0[0] : tapset0 == 1 ? gt0[1] : gt0[2]); k.g01 = g0 * (tapset0 == 0 ? gt1[0] : tapset0 == 1 ? gt1[1] : gt1[2]);
  k.g02 = g0 * (tapset0 == 0 ? 0.1296386719f : 0.f);
  k.g10 = g1 * (tapset1 == 0 ? gt0[0] : tapset1 == 1 ? gt0[1] : gt0[2]); k.g11 = g1 * (tapset1 == 0 ? gt1[0] : tapset1 == 1 ? gt1[1] : gt1[2]);
  k.g12 = g1 * (tapset1 == 0 ? 0.1296386719f : 0.f);
  if (g0 == g1 && k.T0 == k.T1 && tapset0 == tapset1) overlap = 0;
  k.overlap = overlap;
  k.g1zero = g1 == 0;
  return k;
}
OB_DEV float chain_comb_at(const ChComb &k, const ChPre &x, int base, int i) {
  const int j = base + i;
  const float xi = ch_at(x, j);
  if (k.copy) return xi;
  if (i < k.overlap) {
    const float f = ob_window[i] * ob_window[i];
    return xi + ((1.f - f) * k.g00) * ch_at(x, j - k.T0) + ((1.f - f) * k.g01) * (ch_at(x, j - k.T0 + 1) + ch_at(x, j - k.T0 - 1)) +
           ((1.f - f) * k.g02) * (ch_at(x, j - k.T0 + 2) + ch_at(x, j - k.T0 - 2)) + (f * k.g10) * ch_at(x, j - k.T1) +
           (f * k.g11) * (ch_at(x, j - k.T1 + 1) + ch_at(x, j - k.T1 - 1)) + (f * k.g12) * (ch_at(x, j - k.T1 + 2) + ch_at(x, j - k.T1 - 2));
  }
  if (k.g1zero) return xi;
  return xi + k.g10 * ch_at(x, j - k.T1) + k.g11 * (ch_at(x, j - k.T1 + 1) + ch_at(x, j - k.T1 - 1)) + k.g12 * (ch_at(x, j - k.T1 + 2) + ch_at(x, j - k.T1 - 2));
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
  float *trk0 = scratch + OB_CH_OFF_TRK * 32 + lane;           // smoother track of channel 0, [len2][32]
  float *trk1 = trk0 + (OB_CH_LEN / 2) * 32;
  float *fr0 = scratch + OB_CH_OFF_FR * 32 + lane, *fr1 = fr0 + OB_MAXN * 32;   // the unfiltered frame (pre-emphasised), per channel
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
          fr0[(i - OB_OVERLAP) * 32] = v0;
          if (CC == 2) fr1[(i - OB_OVERLAP) * 32] = v1;
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
  st->overlap_max = mx_tail;
  // ---- run_prefilter (celt/celt_encoder.c:2029-2057, 1404-1602) ----
  if (p.in_buf && p.enc_state) {
    const ObEncState &es = p.enc_state[stream];
    const EncBudget bud = enc_budget(p.max_bytes, p.bitrate, p.vbr, p.constrained_vbr, N, es.vbr_reservoir);
    const int enabled = bud.nbAvailableBytes > 12 * C && !silence && bud.nbCompressedBytes * 8 >= 17 && !p.disable_pf;
    const int prev_period0 = es.prefilter_period, prev_tapset = es.prefilter_tapset, tapset = es.tapset_decision, max_period = kPfMaxPeriod,
              min_period = kPfMinPeriod, overlap = OB_OVERLAP;
    const float prev_gain = es.prefilter_gain;
    ChPre pre[2];
    for (int c = 0; c < CC; c++) {
      pre[c].ring = p.prefilter_mem + ((size_t)stream * CC + c) * kPfMaxPeriod;
      pre[c].fr = c ? fr1 : fr0;
      pre[c].head = st->pf_head;
    }
    int pitch_index = min_period, pf_on = 0, qg = 0;
    float gain1 = 0.f;
    const bool search = enabled && (fi.toneishness > .99f || p.complexity >= 5);
    if (search) {
      if (fi.toneishness > .99f) {   // the tone drives the pitch (celt/celt_encoder.c:1446-1461)
        float tf = fi.tone_freq;
        int multiple = 1;
        if (tf >= 3.1416f) tf = 3.141593f - tf;
        while (tf >= multiple * 0.39f) multiple++;
        if (tf > 0.006148f) {
          pitch_index = (int)floor(.5 + 2.f * 3.14159265358979323846 * multiple / tf);
          if (pitch_index > max_period - 2) pitch_index = max_period - 2;
        } else pitch_index = min_period;
        gain1 = .75f;
      } else {
        gain1 = chain_pitch_analysis(pre, CC, N, scratch + OB_CH_OFF_XLP * 32 + lane, scratch + OB_CH_OFF_XLP2 * 32 + lane, scratch + OB_CH_OFF_XC * 32 + lane,
                                     trk0, prev_period0, prev_gain, p.loss_rate, &pitch_index);
      }
    }
    if (search || prev_gain != 0.f) {   // (otherwise: nothing to search, nothing to fade out — gain1 = 0 fails every threshold)
      if (p.ctl && p.ctl[stream].analysis.valid) gain1 = (float)(gain1 * p.ctl[stream].analysis.max_pitch_ratio);   // celt/celt_encoder.c:1494-1495
      float pf_threshold = .2f;
      {
        const int d = pitch_index > prev_period0 ? pitch_index - prev_period0 : prev_period0 - pitch_index;
        if (d * 10 > pitch_index) {
          pf_threshold += .2f;
          if (fi.tf_estimate > .98f) gain1 = 0.f;
        }
      }
      if (bud.nbAvailableBytes < 25) pf_threshold += .1f;
      if (bud.nbAvailableBytes < 35) pf_threshold += .1f;
      if (prev_gain > .4f) pf_threshold -= .1f;
      if (prev_gain > .55f) pf_threshold -= .1f;
      pf_threshold = fmaxf(pf_threshold, .2f);
      if (gain1 < pf_threshold) { gain1 = 0.f; pf_on = 0; qg = 0; }
      else {
        if (fabsf(gain1 - prev_gain) < .1f) gain1 = prev_gain;
        qg = (int)floor((double)(.5f + gain1 * 32 / 3)) - 1;
        qg = qg < 0 ? 0 : (qg > 7 ? 7 : qg);
        gain1 = 0.09375f * (qg + 1);
        pf_on = 1;
      }
    }
    // apply (celt/celt_encoder.c:1539-1586; offset = shortMdctSize - overlap = 0 in this mode): first the |.| sums of the cancel rule, then the
    // final signal with the gain that survived
    const int prev_period = prev_period0 > min_period ? prev_period0 : min_period;
    int cancel_pitch = 0;
    if (gain1 != 0.f || prev_gain != 0.f) {
      const ChComb k = chain_comb_setup(prev_period, pitch_index, -prev_gain, -gain1, prev_tapset, tapset, overlap);
      float before[2] = {0.f, 0.f}, after[2] = {0.f, 0.f};
      for (int c = 0; c < CC; c++) {
        float b = 0.f, a = 0.f;
#pragma unroll 2
        for (int i = 0; i < N; i++) {
          b += fabsf(OB_T32(pre[c].fr, i));
          a += fabsf(chain_comb_at(k, pre[c], max_period, i));
        }
        before[c] = b; after[c] = a;
      }
      if (CC == 2) {
        const float th0 = (.25f * gain1) * before[0] + .01f * before[1], th1 = (.25f * gain1) * before[1] + .01f * before[0];
        if (after[0] - before[0] > th0 || after[1] - before[1] > th1) cancel_pitch = 1;
        if (before[0] - after[0] < th0 && before[1] - after[1] < th1) cancel_pitch = 1;
      } else if (after[0] > before[0]) cancel_pitch = 1;
    }
    // (both gains zero: the filters are copies, after == before, and cancelling a copy changes nothing: the decision is already "off")
    float g_final = gain1;
    if (cancel_pitch) { g_final = 0.f; gain1 = 0.f; pf_on = 0; qg = 0; }
    {
      // cancelled: the new filter's gain is -0 and only the first `overlap` samples are filtered (the old filter fades out), the rest is a copy
      const ChComb k = chain_comb_setup(prev_period, pitch_index, -prev_gain, cancel_pitch ? -0.f : -g_final, prev_tapset, tapset, overlap);
      for (int c = 0; c < CC; c++) {
        float *dst = p.in_buf + ((size_t)stream * CC + c) * OB_CH_IN_STRIDE;
        float *im = st->overlap_mem[c];
        for (int i = 0; i < overlap; i++) dst[i] = im[i];
        for (int i = 0; i < N; i++) {
          const float y = (cancel_pitch && i >= overlap) ? OB_T32(pre[c].fr, i) : chain_comb_at(k, pre[c], max_period, i);
          dst[overlap + i] = y;
          if (i >= N - overlap) im[i - (N - overlap)] = y;
        }
      }
    }
    // memories (celt/celt_encoder.c:1588-1599): the unfiltered frame replaces the oldest N samples of the ring
    for (int c = 0; c < CC; c++) {
      float *ring = p.prefilter_mem + ((size_t)stream * CC + c) * kPfMaxPeriod;
      for (int i = 0; i < N; i++) ring[(pre[c].head + i) & (max_period - 1)] = OB_T32(pre[c].fr, i);
    }
    st->pf_head = (pre[0].head + N) & (max_period - 1);
    for (int c = 0; c < CC; c++) st->preemph_mem[c] = f_mul(coef0, xprev[c]);
    fi.pf_on = pf_on; fi.pitch_index = pitch_index; fi.qg = qg; fi.pf_tapset = tapset; fi.gain1 = gain1; fi.patched = 0;
  }
  p.info[stream] = fi;
}

}  // namespace ob
