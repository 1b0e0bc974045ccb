// celt_enc_pre.cuh — encoder state shared by the two encoder kernels, the byte-budget rule, and the pitch pre-filter stage of the
// analysis kernel: run_prefilter (celt/celt_encoder.c:1404-1602) = pitch_downsample / pitch_search / remove_doubling
// (celt/pitch.c) + comb_filter (celt/celt.c:238-313), and patch_transient_decision (:472-508).
// Every running sum keeps the reference's index order (one lane per sum); everything elementwise is spread over the CTA.
#pragma once
#include "simt.cuh"
#include "celt_types.h"
#include "celt_vec.cuh"

namespace ob {

struct ObEncState {   // CELTEncoder fields that survive a frame (celt/celt_encoder.c:92-140)
  uint32_t rng;
  int32_t spread_decision, tonal_average, lastCodedBands, hf_average, tapset_decision, prefilter_period, prefilter_tapset,
      consec_transient, intensity;
  float delayedIntra, prefilter_gain, stereo_saving, spec_avg;
  int32_t vbr_reservoir, vbr_drift, vbr_offset, vbr_count;
  float oldBandE[2 * OB_NBANDS], oldLogE[2 * OB_NBANDS], oldLogE2[2 * OB_NBANDS], energyError[2 * OB_NBANDS];
};

// the packet-size rule at the top of celt_encode_with_ec (:1882-1960), needed by both kernels
struct EncBudget { int nbCompressedBytes, nbAvailableBytes, effectiveBytes; int32_t vbr_rate; };
OB_DEV EncBudget enc_budget(int max_bytes, int bitrate, int vbr, int constrained_vbr, int N, int32_t vbr_reservoir) {
  EncBudget b;
  b.nbCompressedBytes = max_bytes < 1275 ? max_bytes : 1275;
  b.vbr_rate = 0;
  if (vbr && bitrate != -1) {
    b.vbr_rate = (bitrate * 6 / (6 * 48000 / N)) << 3;   // bitrate_to_bits (celt/celt.h:151) << BITRES
    b.effectiveBytes = b.vbr_rate >> 6;
  } else {
    if (bitrate != -1) {
      const int lim = (bitrate * N + 4 * 48000) / (8 * 48000);
      b.nbCompressedBytes = b.nbCompressedBytes < lim ? b.nbCompressedBytes : lim;
      if (b.nbCompressedBytes < 2) b.nbCompressedBytes = 2;
    }
    b.effectiveBytes = b.nbCompressedBytes;
  }
  b.nbAvailableBytes = b.nbCompressedBytes;
  if (b.vbr_rate > 0 && constrained_vbr) {
    int32_t max_allowed = (b.vbr_rate + b.vbr_rate - vbr_reservoir) >> 6;
    if (max_allowed < 2) max_allowed = 2;
    if (max_allowed > b.nbAvailableBytes) max_allowed = b.nbAvailableBytes;
    if (max_allowed < b.nbAvailableBytes) { b.nbCompressedBytes = max_allowed; b.nbAvailableBytes = max_allowed; }
  }
  return b;
}

constexpr int kPfMaxPeriod = 1024, kPfMinPeriod = 15;
// Shared-memory scratch of the pre-filter stage.  Only `sc` exists below complexity 5 (the host sizes the allocation);
// the search buffers double as the spectra once the stage is over (buf + xlp2 are contiguous: 2000 floats >= 2 x 960).
struct PfWS {
  float sc[64];       // scalars handed between phases
  float buf[1008];    // xlp (992) -> after the whitening filter: xcorr [0,512) + Syy track [512,1001) -> yy_lookup [0,513)
  float xlp2[992];    // the whitened, 2:1 decimated search signal
};
// The unfiltered history is a ring of 1024 samples per channel in HBM (CELTEncoder::prefilter_mem without the memmove):
// sample j of the reference's pre[] = [history | frame] buffer is ring[(head + j) & 1023] for j < 1024, else frame[j - 1024].
struct PfSig { const float *ring; const float *frame; int head; };
OB_DEV float pf_at(const PfSig &p, int j) { return j < kPfMaxPeriod ? p.ring[(p.head + j) & (kPfMaxPeriod - 1)] : p.frame[j - kPfMaxPeriod]; }
struct PfDecision { int pf_on, pitch_index, qg, tapset; float gain1; };

OB_DEV float pf_dot(const float *x, const float *y, int n) {   // index order; unrolled so the loads run ahead of the serial sum
  float s = 0.f;
#pragma unroll 8
  for (int i = 0; i < n; i++) s = s + x[i] * y[i];
  return s;
}
// find_best_pitch (celt/pitch.c:45-103).  The energy track Syy(i) is a serial chain, but its increments are not: all threads
// square the samples (phase 1), one thread runs the chain and the best-two selection over pipelined loads (phase 2).
// scratch: max_pitch floats.  Every thread of the CTA calls it; the result lands in best_pitch[0..1] (shared memory).
OB_DEVNI void pf_find_best_pitch(const float *__restrict__ xcorr, const float *__restrict__ y, int ystride, int len, int max_pitch,
                                 float *__restrict__ scratch, int *__restrict__ best_pitch, int tid, int nth) {
  for (int i = tid; i < max_pitch; i += nth) {
    const float a = y[(i + len) * ystride], b = y[i * ystride];
    scratch[i] = a * a - b * b;
  }
  bsync();
  if (tid == 0) {
    float Syy = 1.f, bn0 = -1.f, bn1 = -1.f, bd0 = 0.f, bd1 = 0.f;
    int bp0 = 0, bp1 = 1;
#pragma unroll 8
    for (int j = 0; j < len; j++) Syy = Syy + y[j * ystride] * y[j * ystride];
#pragma unroll 4
    for (int i = 0; i < max_pitch; i++) {   // scratch[i] <- Syy as iteration i sees it
      const float d = scratch[i];
      scratch[i] = Syy;
      Syy += d;
      Syy = Syy > 1.f ? Syy : 1.f;
    }
#pragma unroll 4
    for (int i = 0; i < max_pitch; i++) {
      const float xc = xcorr[i], Si = scratch[i];
      if (xc > 0) {
        float x16 = xc;
        x16 *= 1e-12f;
        const float num = x16 * x16;
        if (num * bd1 > bn1 * Si) {
          if (num * bd0 > bn0 * Si) { bn1 = bn0; bd1 = bd0; bp1 = bp0; bn0 = num; bd0 = Si; bp0 = i; }
          else { bn1 = num; bd1 = Si; bp1 = i; }
        }
      }
    }
    best_pitch[0] = bp0;
    best_pitch[1] = bp1;
  }
  bsync();
}
OB_DEV float pf_pitch_gain(float xy, float xx, float yy) { return xy / (float)sqrt((double)(1 + xx * yy)); }

// comb_filter with distinct input/output (celt/celt.c:238-313), elementwise over `tid`; x[k] = pre[base + k]
OB_DEV void pf_comb(float *y, const PfSig &x, int base, int T0, int T1, int N, float g0, float g1, int tapset0, int tapset1, int overlap, int tid,
                    int nth) {
  if (g0 == 0 && g1 == 0) {
    for (int i = tid; i < N; i += nth) y[i] = pf_at(x, base + i);
    return;
  }
  T0 = T0 > kPfMinPeriod ? T0 : kPfMinPeriod;
  T1 = T1 > kPfMinPeriod ? T1 : kPfMinPeriod;
  const float gt0[3] = {0.3066406250f, 0.4638671875f, 0.7998046875f}, gt1[3] = {0.2170410156f, 0.2680664062f, 0.1000976562f};
  const float g00 = g0 * (tapset0 == 0 ? gt0[0] : tapset0 == 1 ? gt0[1] : gt0[2]), g01 = g0 * (tapset0 == 0 ? gt1[0] : tapset0 == 1 ? gt1[1] : gt1[2]),
              g02 = g0 * (tapset0 == 0 ? 0.1296386719f : 0.f);
  const float g10 = g1 * (tapset1 == 0 ? gt0[0] : tapset1 == 1 ? gt0[1] : gt0[2]), g11 = g1 * (tapset1 == 0 ? gt1[0] : tapset1 == 1 ? gt1[1] : gt1[2]),
              g12 = g1 * (tapset1 == 0 ? 0.1296386719f : 0.f);
  if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
  for (int i = tid; i < N; i += nth) {
    const int j = base + i;
    const float xi = pf_at(x, j);
    if (i < overlap) {
      const float f = ob_window[i] * ob_window[i];
      y[i] = xi + ((1.f - f) * g00) * pf_at(x, j - T0) + ((1.f - f) * g01) * (pf_at(x, j - T0 + 1) + pf_at(x, j - T0 - 1)) +
             ((1.f - f) * g02) * (pf_at(x, j - T0 + 2) + pf_at(x, j - T0 - 2)) + (f * g10) * pf_at(x, j - T1) +
             (f * g11) * (pf_at(x, j - T1 + 1) + pf_at(x, j - T1 - 1)) + (f * g12) * (pf_at(x, j - T1 + 2) + pf_at(x, j - T1 - 2));
    } else if (g1 == 0) y[i] = xi;
    else y[i] = xi + g10 * pf_at(x, j - T1) + g11 * (pf_at(x, j - T1 + 1) + pf_at(x, j - T1 - 1)) + g12 * (pf_at(x, j - T1 + 2) + pf_at(x, j - T1 - 2));
  }
}

// The whole pre-filter stage.  in[c]: the channel's [overlap | frame] buffer (ws->in): on return it holds the filtered
// overlap + frame.  fr[c]: N floats of shared scratch (ws->f2) for the unfiltered frame.  ring / in_mem: HBM state.
// All threads of the CTA call it (C warps).
struct PfArgs {
  float *in[2];
  float *fr[2];
  float *ring[2];            // [1024]
  float *in_mem[2];          // [120]
  int head;                  // ring position of the oldest history sample
  int C, N, enabled, complexity, nbAvailableBytes, prefilter_tapset, prev_period, prev_tapset, loss_rate;
  float prev_gain, tf_estimate, tone_freq, toneishness;
  int an_valid = 0;             // st->analysis.valid (CELT_SET_ANALYSIS): gain1 is scaled by max_pitch_ratio (celt/celt_encoder.c:1494-1495)
  float an_max_pitch_ratio = 1.f;
};
OB_DEVNI PfDecision enc_prefilter_stage(PfWS *w, const PfArgs &a) {
  const int tid = warp_in_block() * 32 + lane_id(), nth = warps_per_block() * 32;
  const int N = a.N, C = a.C, overlap = OB_OVERLAP, max_period = kPfMaxPeriod, min_period = kPfMinPeriod;
  if (a.prev_gain == 0.f && !(a.enabled && (a.toneishness > .99f || a.complexity >= 5))) {
    // No search, no tone, nothing to fade out: both comb filters are copies, gain1 = 0 fails every threshold, and the
    // cancel_pitch rule cannot change a copy.  Only the memories move: in = [in_mem | frame], in_mem = frame tail,
    // the ring takes the new frame (celt/celt_encoder.c:1539-1599 with g0 = g1 = 0).
    for (int c = 0; c < C; c++) {
      for (int i = tid; i < overlap; i += nth) a.in[c][i] = a.in_mem[c][i];
      for (int i = tid; i < N; i += nth) a.ring[c][(a.head + i) & (max_period - 1)] = a.in[c][overlap + i];
    }
    bsync();
    for (int c = 0; c < C; c++)
      for (int i = tid; i < overlap; i += nth) a.in_mem[c][i] = a.in[c][N + i];
    bsync();
    PfDecision r;
    r.pf_on = 0; r.pitch_index = min_period; r.qg = 0; r.tapset = a.prefilter_tapset; r.gain1 = 0.f;
    return r;
  }
  PfSig pre[2];
  for (int c = 0; c < C; c++) {
    pre[c].ring = a.ring[c]; pre[c].frame = a.fr[c]; pre[c].head = a.head;
    for (int i = tid; i < N; i += nth) a.fr[c][i] = a.in[c][overlap + i];
  }
  bsync();
  int pitch_index = min_period;
  float gain1 = 0.f;
  if (a.enabled && a.toneishness > .99f) {
    float tf = a.tone_freq;
    int multiple = 1;
    if (tf >= 3.1416f) tf = 3.141593f - tf;
    while (tf >= multiple * 0.39f) multiple++;
    if (tf > 0.006148f) {
      pitch_index = (int)floor(.5 + 2.f * 3.14159265358979323846 * multiple / tf);
      if (pitch_index > max_period - 2) pitch_index = max_period - 2;
    } else pitch_index = min_period;
    gain1 = .75f;
  } else if (a.enabled && a.complexity >= 5) {
    const int len = (max_period + N) >> 1;
    float *xlp = w->buf, *xcorr = w->buf, *track = w->buf + 512, *yyl = w->buf;
    // pitch_downsample (celt/pitch.c:140-228): 2:1 decimation of the channel sum ...
    for (int i = tid; i < len; i += nth) {
      float v = i == 0 ? .25f * pf_at(pre[0], 1) + .5f * pf_at(pre[0], 0)
                       : .25f * pf_at(pre[0], 2 * i - 1) + .25f * pf_at(pre[0], 2 * i + 1) + .5f * pf_at(pre[0], 2 * i);
      if (C == 2)
        v += i == 0 ? .25f * pf_at(pre[1], 1) + .5f * pf_at(pre[1], 0)
                    : .25f * pf_at(pre[1], 2 * i - 1) + .25f * pf_at(pre[1], 2 * i + 1) + .5f * pf_at(pre[1], 2 * i);
      xlp[i] = v;
    }
    bsync();
    // ... _celt_autocorr over 5 lags (celt/celt_lpc.c:284-370): xcorr over fastN then the tail, one lane per lag ...
    if (tid < 5) {
      const int k = tid, fastN = len - 4;
      float ac = pf_dot(xlp, xlp + k, fastN), d = 0.f;
#pragma unroll 1
      for (int i = k + fastN; i < len; i++) d = d + xlp[i] * xlp[i - k];
      w->sc[k] = ac + d;
    }
    bsync();
    // ... _celt_lpc(4), bandwidth expansion, the extra zero (one thread) ...
    if (tid == 0) {
      float ac[5], lpc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < 5; i++) ac[i] = w->sc[i];
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
      for (int i = 0; i < 4; i++) {
        tmp = .9f * tmp;
        lpc[i] = lpc[i] * tmp;
      }
      w->sc[8] = lpc[0] + .8f;
      w->sc[9] = lpc[1] + .8f * lpc[0];
      w->sc[10] = lpc[2] + .8f * lpc[1];
      w->sc[11] = lpc[3] + .8f * lpc[2];
      w->sc[12] = .8f * lpc[3];
    }
    bsync();
    // ... celt_fir5: a 5-tap FIR on the decimated signal
    {
      const float l0 = w->sc[8], l1 = w->sc[9], l2 = w->sc[10], l3 = w->sc[11], l4 = w->sc[12];
      for (int i = tid; i < len; i += nth) {
        float sum = xlp[i];
        sum = sum + l0 * (i >= 1 ? xlp[i - 1] : 0.f);
        sum = sum + l1 * (i >= 2 ? xlp[i - 2] : 0.f);
        sum = sum + l2 * (i >= 3 ? xlp[i - 3] : 0.f);
        sum = sum + l3 * (i >= 4 ? xlp[i - 4] : 0.f);
        sum = sum + l4 * (i >= 5 ? xlp[i - 5] : 0.f);
        w->xlp2[i] = sum;
      }
    }
    bsync();
    // pitch_search(pitch_buf + max_period/2, pitch_buf, N, max_period - 3*min_period) (celt/pitch.c:307-425)
    const float *pb = w->xlp2, *xl = w->xlp2 + (max_period >> 1);
    const int mp = max_period - 3 * min_period, len4 = N >> 2, mp4 = mp >> 2, len2 = N >> 1, mp2 = mp >> 1;
    for (int i = tid; i < mp4; i += nth) {   // coarse: both signals decimated by 2 again; one lane per lag, j in order
      float s = 0.f;
#pragma unroll 8
      for (int j = 0; j < len4; j++) s = s + xl[2 * j] * pb[2 * (i + j)];
      xcorr[i] = s;
    }
    bsync();
    pf_find_best_pitch(xcorr, pb, 2, len4, mp4, track, (int *)w->sc + 16, tid, nth);
    {
      const int b0 = ((int *)w->sc)[16], b1 = ((int *)w->sc)[17];
      for (int i = tid; i < mp2; i += nth) {
        const int d0 = i - 2 * b0, d1 = i - 2 * b1;
        float v = 0.f;
        if (!((d0 > 2 || d0 < -2) && (d1 > 2 || d1 < -2))) {
          const float s = pf_dot(xl, pb + i, len2);
          v = s > -1.f ? s : -1.f;
        }
        xcorr[i] = v;
      }
    }
    bsync();
    pf_find_best_pitch(xcorr, pb, 1, len2, mp2, track, (int *)w->sc + 16, tid, nth);
    if (tid == 0) {
      int bp[2] = {((int *)w->sc)[16], ((int *)w->sc)[17]}, offset = 0;
      if (bp[0] > 0 && bp[0] < mp2 - 1) {
        const float aa = xcorr[bp[0] - 1], bb = xcorr[bp[0]], cc = xcorr[bp[0] + 1];
        if ((cc - aa) > .7f * (bb - aa)) offset = 1;
        else if ((aa - cc) > .7f * (bb - cc)) offset = -1;
      }
      ((int *)w->sc)[18] = max_period - (2 * bp[0] - offset);
    }
    bsync();
    // remove_doubling(pitch_buf, max_period, min_period, N, &pitch_index, prev_period, prev_gain) (celt/pitch.c:438-575)
    {
      const int maxp = max_period / 2, minp = min_period / 2, Nh = N / 2, prevp = a.prev_period / 2;
      const float *x = pb + maxp;
      int T0 = ((int *)w->sc)[18] / 2;
      if (T0 >= maxp) T0 = maxp - 1;
      const int second_check[16] = {0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2};
      // the independent running sums: lanes 0,1: xx, xy;  lanes 2..29: (k, first/second) for k = 2..15;  one more thread
      // (of the second warp when there is one): xx again and the yy_lookup energy track that starts from it
      const int ytid = nth > 32 ? 32 : 30;
      if (tid < 2) w->sc[20 + tid] = pf_dot(x, tid == 0 ? x : x - T0, Nh);
      else if (tid < 30) {
        const int k = 2 + ((tid - 2) >> 1), which = (tid - 2) & 1;
        const int T1 = (unsigned)(2 * T0 + k) / (unsigned)(2 * k);
        int T1b;
        if (k == 2) T1b = T1 + T0 > maxp ? T0 : T0 + T1;
        else T1b = (unsigned)(2 * second_check[k] * T0 + k) / (unsigned)(2 * k);
        w->sc[24 + (tid - 2)] = T1 >= minp ? pf_dot(x, x - (which ? T1b : T1), Nh) : 0.f;
      } else if (tid == ytid) {
        float yy = pf_dot(x, x, Nh);
        float *__restrict__ yl = yyl;
        const float *__restrict__ xr = x;
        yl[0] = yy;
#pragma unroll 4
        for (int i = 1; i <= maxp; i++) {
          yy = yy + xr[-i] * xr[-i] - xr[Nh - i] * xr[Nh - i];
          yl[i] = yy > 0.f ? yy : 0.f;
        }
      }
      bsync();
      if (tid == 0) {
        const float xx = w->sc[20];
        float xy = w->sc[21], yy = yyl[T0];
        float best_xy = xy, best_yy = yy;
        float g, g0;
        int T = T0;
        g = g0 = pf_pitch_gain(xy, xx, yy);
        for (int k = 2; k <= 15; k++) {
          const int T1 = (unsigned)(2 * T0 + k) / (unsigned)(2 * k);
          if (T1 < minp) break;
          int T1b;
          if (k == 2) T1b = T1 + T0 > maxp ? T0 : T0 + T1;
          else T1b = (unsigned)(2 * second_check[k] * T0 + k) / (unsigned)(2 * k);
          xy = .5f * (w->sc[24 + 2 * (k - 2)] + w->sc[24 + 2 * (k - 2) + 1]);
          yy = .5f * (yyl[T1] + yyl[T1b]);
          const float g1 = pf_pitch_gain(xy, xx, yy);
          float cont;
          const int dT = T1 > prevp ? T1 - prevp : prevp - T1;
          if (dT <= 1) cont = a.prev_gain;
          else if (dT <= 2 && 5 * k * k < T0) cont = .5f * a.prev_gain;
          else cont = 0.f;
          float thresh = fmaxf(.3f, .7f * g0 - cont);
          if (T1 < 3 * minp) thresh = fmaxf(.4f, .85f * g0 - cont);
          else if (T1 < 2 * minp) thresh = fmaxf(.5f, .9f * g0 - cont);
          if (g1 > thresh) { best_xy = xy; best_yy = yy; T = T1; g = g1; }
        }
        ((int *)w->sc)[19] = T;
        w->sc[22] = g; w->sc[23] = best_xy; w->sc[52] = best_yy;
      }
      bsync();
      {
        const int T = ((int *)w->sc)[19];
        // the last correlations: T*5/8 and T*6/8 (only used when T < 2*minp) and the three around T
        if (tid < 2) { const int Tq = tid == 0 ? T * 5 / 8 : T * 6 / 8; w->sc[53 + tid] = pf_dot(x, x - Tq, Nh); }
        else if (tid < 5) w->sc[55 + (tid - 2)] = pf_dot(x, x - (T + (tid - 2) - 1), Nh);
      }
      bsync();
      if (tid == 0) {
        const float xx = w->sc[20];
        const int T = ((int *)w->sc)[19];
        float g = w->sc[22], best_xy = w->sc[23];
        const float best_yy = w->sc[52];
        if (T < minp * 2) {
          const int T1 = T * 5 / 8, T2 = T * 6 / 8;
          const float g1 = pf_pitch_gain(w->sc[53], xx, yyl[T1]), g2 = pf_pitch_gain(w->sc[54], xx, yyl[T2]);
          if (g1 >= g || g2 >= g) g = 0.f;
        }
        best_xy = best_xy > 0.f ? best_xy : 0.f;
        float pg = best_yy <= best_xy ? 1.f : best_xy / (best_yy + 1);
        const float c0 = w->sc[55], c1 = w->sc[56], c2 = w->sc[57];
        int offset = 0;
        if ((c2 - c0) > .7f * (c1 - c0)) offset = 1;
        else if ((c0 - c2) > .7f * (c1 - c2)) offset = -1;
        if (pg > g) pg = g;
        int Tout = 2 * T + offset;
        if (Tout < min_period) Tout = min_period;
        if (Tout > max_period - 2) Tout = max_period - 2;
        float g1 = .7f * pg;
        if (a.loss_rate > 2) g1 = .5f * g1;
        if (a.loss_rate > 4) g1 = .5f * g1;
        if (a.loss_rate > 8) g1 = 0.f;
        ((int *)w->sc)[18] = Tout;
        w->sc[22] = g1;
      }
      bsync();
      pitch_index = ((int *)w->sc)[18];
      gain1 = w->sc[22];
    }
  }
  if (a.an_valid) gain1 = (float)(gain1 * a.an_max_pitch_ratio);   // celt/celt_encoder.c:1494-1495
  // gain threshold and quantisation (celt/celt_encoder.c:1499-1536); every thread computes the same scalars
  float pf_threshold = .2f;
  {
    const int d = pitch_index > a.prev_period ? pitch_index - a.prev_period : a.prev_period - pitch_index;
    if (d * 10 > pitch_index) {
      pf_threshold += .2f;
      if (a.tf_estimate > .98f) gain1 = 0.f;
    }
  }
  if (a.nbAvailableBytes < 25) pf_threshold += .1f;
  if (a.nbAvailableBytes < 35) pf_threshold += .1f;
  if (a.prev_gain > .4f) pf_threshold -= .1f;
  if (a.prev_gain > .55f) pf_threshold -= .1f;
  pf_threshold = fmaxf(pf_threshold, .2f);
  int pf_on, qg;
  if (gain1 < pf_threshold) { gain1 = 0.f; pf_on = 0; qg = 0; }
  else {
    if (fabsf(gain1 - a.prev_gain) < .1f) gain1 = a.prev_gain;
    qg = (int)floor((double)(.5f + gain1 * 32 / 3)) - 1;
    qg = qg < 0 ? 0 : (qg > 7 ? 7 : qg);
    gain1 = 0.09375f * (qg + 1);
    pf_on = 1;
  }
  // apply (celt/celt_encoder.c:1539-1586)
  const int prev_period = a.prev_period > min_period ? a.prev_period : min_period, offset = OB_SHORT_N - overlap;
  for (int c = 0; c < C; c++) {
    for (int i = tid; i < overlap; i += nth) a.in[c][i] = a.in_mem[c][i];
    if (offset) pf_comb(a.in[c] + overlap, pre[c], max_period, prev_period, prev_period, offset, -a.prev_gain, -a.prev_gain, a.prev_tapset, a.prev_tapset, 0, tid, nth);
    pf_comb(a.in[c] + overlap + offset, pre[c], max_period + offset, prev_period, pitch_index, N - offset, -a.prev_gain, -gain1, a.prev_tapset,
            a.prefilter_tapset, overlap, tid, nth);
  }
  bsync();
  // |.| sums before / after, in index order: 2C lanes
  if (tid < 2 * C) {
    const int c = tid >> 1, after = tid & 1;
    const float *src = after ? a.in[c] + overlap : a.fr[c];
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < N; i++) s += fabsf(src[i]);
    w->sc[40 + tid] = s;
  }
  bsync();
  int cancel_pitch = 0;
  {
    const float b0 = w->sc[40], a0 = w->sc[41];
    if (C == 2) {
      const float b1 = w->sc[42], a1 = w->sc[43];
      const float th0 = (.25f * gain1) * b0 + .01f * b1, th1 = (.25f * gain1) * b1 + .01f * b0;
      if (a0 - b0 > th0 || a1 - b1 > th1) cancel_pitch = 1;
      if (b0 - a0 < th0 && b1 - a1 < th1) cancel_pitch = 1;
    } else if (a0 > b0) cancel_pitch = 1;
  }
  if (cancel_pitch) {
    for (int c = 0; c < C; c++)
      for (int i = tid; i < N; i += nth) a.in[c][overlap + i] = a.fr[c][i];
    bsync();
    for (int c = 0; c < C; c++)
      pf_comb(a.in[c] + overlap + offset, pre[c], max_period + offset, prev_period, pitch_index, overlap, -a.prev_gain, -0.f, a.prev_tapset,
              a.prefilter_tapset, overlap, tid, nth);
    gain1 = 0.f;
    pf_on = 0;
    qg = 0;
  }
  bsync();
  // memories (celt/celt_encoder.c:1588-1599): the filtered tail for the next MDCT overlap; the unfiltered frame replaces the
  // oldest N samples of the ring (every read of the history is behind the barrier above)
  for (int c = 0; c < C; c++) {
    for (int i = tid; i < overlap; i += nth) a.in_mem[c][i] = a.in[c][N + i];
    for (int i = tid; i < N; i += nth) a.ring[c][(a.head + i) & (max_period - 1)] = a.fr[c][i];
  }
  bsync();
  PfDecision r;
  r.pf_on = pf_on; r.pitch_index = pitch_index; r.qg = qg; r.tapset = a.prefilter_tapset; r.gain1 = gain1;
  return r;
}

// patch_transient_decision (celt/celt_encoder.c:472-508), one thread; newE / oldE are [C][21]
OB_DEVNI int enc_patch_transient(const float *newE, const float *oldE, int C, int end) {
  const int start = 0;
  float spread_old[OB_NBANDS], mean_diff = 0.f;
  if (C == 1) {
    spread_old[start] = oldE[start];
    for (int i = start + 1; i < end; i++) spread_old[i] = fmaxf(spread_old[i - 1] - 1.0f, oldE[i]);
  } else {
    spread_old[start] = fmaxf(oldE[start], oldE[start + OB_NBANDS]);
    for (int i = start + 1; i < end; i++) spread_old[i] = fmaxf(spread_old[i - 1] - 1.0f, fmaxf(oldE[i], oldE[i + OB_NBANDS]));
  }
  for (int i = end - 2; i >= start; i--) spread_old[i] = fmaxf(spread_old[i], spread_old[i + 1] - 1.0f);
  for (int c = 0; c < C; c++)
    for (int i = 2; i < end - 1; i++) {
      const float x1 = fmaxf(0.f, newE[i + c * OB_NBANDS]), x2 = fmaxf(0.f, spread_old[i]);
      mean_diff = mean_diff + fmaxf(0.f, x1 - x2);
    }
  mean_diff = mean_diff / (C * (end - 1 - 2));
  return mean_diff > 1.f;
}

}  // namespace ob
