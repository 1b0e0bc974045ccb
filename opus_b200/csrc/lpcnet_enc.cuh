// lpcnet_enc.cuh — the LPCNet feature extraction, one warp per stream: dnn/lpcnet_enc.c:73-230 (lpcnet_compute_single_frame_features_float,
// compute_frame_features, frame_analysis, biquad, preemphasis), dnn/freq.c:57-323 (forward_transform, inverse_transform, lpcn_compute_band_energy,
// compute_band_energy_inverse, dct / idct, interp_band_gain, lpc_from_cepstrum, lpcn_lpc, apply_window, burg_cepstral_analysis) and the tail of
// compute_pitchdnn (dnn/pitchdnn.c:45-58).  It is what produces the 36 features the DRED encoder codes (dnn/dred_encoder.c:101-102) and what
// lpcnet_plc_conceal runs over the decoded history before the first concealed frame (dnn/lpcnet_plc.c:160-177).
// The frame is cut in two kernels around the pitch network (lpcnet_enc_host.cuh runs its layers in between):
//   lpcnet_enc_analysis_warp   pre-emphasis, windowed 320-point transform, band energies, instantaneous-frequency features, cepstrum, LPC from the
//                              cepstrum (second transform + Levinson), LPC residual, low-pass, the 224 normalised cross-correlations
//   lpcnet_enc_finish_warp     arg-max / local soft-max of the network's 180 outputs, pitch lag, frame correlation, the feature vector
// Arithmetic: every sum in the reference's index order, float / double promotions as the C expressions have them (the doubles are the libm calls
// log / pow / exp / sqrt: the device's differ from glibc by an ulp of a double now and then, which survives the rounding to float only rarely).
#pragma once
#include "simt.cuh"
#include "lpcnet_tables_data.h"
#include "silk_ops.cuh"

namespace ob {
constexpr int kLeFrame = 160, kLeWin = 320, kLeFreq = 161, kLeBands = 18, kLeOrder = 16, kLeFeat = 36, kLePitchMin = 32, kLePitchMax = 256;
constexpr int kLeXcorr = kLePitchMax - kLePitchMin, kLeIfFreq = 30, kLeIf = 3 * kLeIfFreq - 2, kLeTrain = 80, kLeBuf = kLePitchMax + kLeFrame;
constexpr int kLeConvIn = kLeXcorr + 2, kLePitchOut = 192;

OB_TABLE short ob_le_eband[kLeBands] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 34, 40};   // eband5ms, freq.c:45-48
OB_TABLE float ob_le_comp[kLeBands] = {0.8f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 0.666667f, 0.5f, 0.5f, 0.5f, 0.333333f, 0.25f, 0.25f, 0.2f,
                                        0.166667f, 0.173913f};                                              // compensation, freq.c:50-52

// LPCNetEncState (dnn/lpcnet_private.h:25-43) without the pitch network's memories (batched arrays of the host object)
struct ObLpcnetEncState {
  float analysis_mem[kLeFrame];
  float prev_if[2 * kLeIfFreq];
  float pitch_mem[kLeOrder];
  float exc_buf[kLeBuf];   // the reference allocates PITCH_MAX_PERIOD + 320 and uses the first PITCH_MAX_PERIOD + FRAME_SIZE
  float lp_buf[kLeBuf];
  float lpc[kLeOrder];
  float features[kLeFeat];
  float lp_mem[2];
  float mem_preemph, pitch_filt, dnn_pitch, pad;
};
struct LeCpx { float r, i; };
struct LeWS {   // per-warp shared memory
  float win[kLeWin];
  LeCpx fa[kLeWin];
  float x[kLeFrame + kLeOrder];
  float exc[kLeBuf];
  float lp[kLeBuf];
  float Ex[kLeBands], Ly[kLeBands], tmp[kLeBands], feat[kLeFeat], lpc[kLeOrder], ac[kLeOrder + 4];
};
struct LpcnetEncParams {
  ObLpcnetEncState *state;    // [n]
  const void *pcm;            // [n][160] float or int16
  float *if_features;         // [n][88]
  float *conv_in;             // [n][226]: the cross-correlation features at 1..224, zeros at both ends
  const float *pitch_out;     // [n][192] (finish)
  float *features;            // [n][36] (finish)
  int n, pcm_is_int16;
  int pcm_stride;             // elements between two streams' frames (0 = 160)
};

OB_DEV LeCpx le_cmul(LeCpx a, LeCpx b) { LeCpx m; m.r = a.r * b.r - a.i * b.i; m.i = a.r * b.i + a.i * b.r; return m; }
OB_DEV LeCpx le_tw(int k) { LeCpx t; t.r = ob_lpcnet_fft_tw[2 * k]; t.i = ob_lpcnet_fft_tw[2 * k + 1]; return t; }
// kf_bfly4 with twiddles (celt/kiss_fft.c:131-173) on F[0], F[m], F[2m], F[3m]
OB_DEV void le_bfly4(LeCpx *F, int m, int j, int fstride) {
  const LeCpx a = le_cmul(F[m], le_tw(j * fstride)), b = le_cmul(F[2 * m], le_tw(2 * j * fstride)), c = le_cmul(F[3 * m], le_tw(3 * j * fstride));
  LeCpx d, e, f, f0 = F[0];
  d.r = f0.r - b.r; d.i = f0.i - b.i;
  f0.r += b.r; f0.i += b.i;
  e.r = a.r + c.r; e.i = a.i + c.i;
  f.r = a.r - c.r; f.i = a.i - c.i;
  F[2 * m].r = f0.r - e.r; F[2 * m].i = f0.i - e.i;
  f0.r += e.r; f0.i += e.i;
  F[0] = f0;
  F[m].r = d.r + f.i; F[m].i = d.i - f.r;
  F[3 * m].r = d.r - f.i; F[3 * m].i = d.i + f.r;
}
// The 320-point transform of dnn/lpcnet_tables.c's kfft on digit-reversed, pre-scaled data: radix 4 with m = 1, 4, 16, then radix 5 with m = 64
// (opus_fft_impl, celt/kiss_fft.c:562-594); 80 / 80 / 80 / 64 butterflies per pass over the warp's lanes.
OB_DEV void le_fft320(LeCpx *F, int lane) {
  for (int b = lane; b < 80; b += 32) {   // kf_bfly4, m == 1 (kiss_fft.c:107-130)
    LeCpx *G = F + 4 * b, s0, s1, g0 = G[0], g1 = G[1], g2 = G[2], g3 = G[3];
    s0.r = g0.r - g2.r; s0.i = g0.i - g2.i;
    g0.r += g2.r; g0.i += g2.i;
    s1.r = g1.r + g3.r; s1.i = g1.i + g3.i;
    g2.r = g0.r - s1.r; g2.i = g0.i - s1.i;
    g0.r += s1.r; g0.i += s1.i;
    s1.r = g1.r - g3.r; s1.i = g1.i - g3.i;
    g1.r = s0.r + s1.i; g1.i = s0.i - s1.r;
    g3.r = s0.r - s1.i; g3.i = s0.i + s1.r;
    G[0] = g0; G[1] = g1; G[2] = g2; G[3] = g3;
  }
  wsync();
  for (int b = lane; b < 80; b += 32) le_bfly4(F + (b >> 2) * 16 + (b & 3), 4, b & 3, 20);
  wsync();
  for (int b = lane; b < 80; b += 32) le_bfly4(F + (b >> 4) * 64 + (b & 15), 16, b & 15, 5);
  wsync();
  const LeCpx ya = le_tw(64), yb = le_tw(128);
  for (int u = lane; u < 64; u += 32) {   // kf_bfly5 (kiss_fft.c:236-301), fstride 1, m 64
    const LeCpx s0 = F[u], s1 = le_cmul(F[u + 64], le_tw(u)), s2 = le_cmul(F[u + 128], le_tw(2 * u)), s3 = le_cmul(F[u + 192], le_tw(3 * u)),
                s4 = le_cmul(F[u + 256], le_tw(4 * u));
    LeCpx s5, s6, s7, s8, s9, s10, s11, s12, o;
    s7.r = s1.r + s4.r; s7.i = s1.i + s4.i;
    s10.r = s1.r - s4.r; s10.i = s1.i - s4.i;
    s8.r = s2.r + s3.r; s8.i = s2.i + s3.i;
    s9.r = s2.r - s3.r; s9.i = s2.i - s3.i;
    o.r = s0.r + (s7.r + s8.r);
    o.i = s0.i + (s7.i + s8.i);
    F[u] = o;
    s5.r = s0.r + (s7.r * ya.r + s8.r * yb.r);
    s5.i = s0.i + (s7.i * ya.r + s8.i * yb.r);
    s6.r = s10.i * ya.i + s9.i * yb.i;
    s6.i = -(s10.r * ya.i + s9.r * yb.i);
    o.r = s5.r - s6.r; o.i = s5.i - s6.i; F[u + 64] = o;
    o.r = s5.r + s6.r; o.i = s5.i + s6.i; F[u + 256] = o;
    s11.r = s0.r + (s7.r * yb.r + s8.r * ya.r);
    s11.i = s0.i + (s7.i * yb.r + s8.i * ya.r);
    s12.r = s9.i * ya.i - s10.i * yb.i;
    s12.i = s10.r * yb.i - s9.r * ya.i;
    o.r = s11.r + s12.r; o.i = s11.i + s12.i; F[u + 128] = o;
    o.r = s11.r - s12.r; o.i = s11.i - s12.i; F[u + 192] = o;
  }
  wsync();
}
// forward_transform (freq.c:232-244) of a real window: opus_fft_c's scaled digit-reversed copy (kiss_fft.c:596-613), then the passes
OB_DEV void le_forward(LeCpx *F, const float *in, int lane) {
  for (int i = lane; i < kLeWin; i += 32) { LeCpx v; v.r = 0.0031250000f * in[i]; v.i = 0.0031250000f * 0.f; F[ob_lpcnet_fft_bitrev[i]] = v; }
  wsync();
  le_fft320(F, lane);
}
// lpcn_compute_band_energy / compute_band_energy_inverse (freq.c:55-80, 131-154) for band b: the (frac) terms of band b-1, then the (1 - frac) terms of band b
template <bool INVERSE> OB_DEV float le_band_energy(const LeCpx *X, int b) {
  float sum = 0;
  if (b > 0) {
    const int lo = ob_le_eband[b - 1], size = (ob_le_eband[b] - lo) * 4;
    for (int j = 0; j < size; j++) {
      const float frac = (float)j / size;
      const LeCpx v = X[lo * 4 + j];
      float t = v.r * v.r;
      t += v.i * v.i;
      if (INVERSE) t = (float)(1.f / ((double)t + 1e-9));
      sum += frac * t;
    }
  }
  if (b < kLeBands - 1) {
    const int lo = ob_le_eband[b], size = (ob_le_eband[b + 1] - lo) * 4;
    for (int j = 0; j < size; j++) {
      const float frac = (float)j / size;
      const LeCpx v = X[lo * 4 + j];
      float t = v.r * v.r;
      t += v.i * v.i;
      if (INVERSE) t = (float)(1.f / ((double)t + 1e-9));
      sum += (1 - frac) * t;
    }
  }
  if (b == 0 || b == kLeBands - 1) sum *= 2;
  return sum;
}
// dct() / idct() (freq.c:208-230), output i
OB_DEV float le_dct(const float *in, int i) {
  float sum = 0;
  for (int j = 0; j < kLeBands; j++) sum += in[j] * ob_lpcnet_dct_table[j * kLeBands + i];
  return (float)(sum * sqrt(2. / kLeBands));
}
OB_DEV float le_idct(const float *in, int i) {
  float sum = 0;
  for (int j = 0; j < kLeBands; j++) sum += in[j] * ob_lpcnet_dct_table[i * kLeBands + j];
  return (float)(sum * sqrt(2. / kLeBands));
}
OB_DEV float le_log10(float x) { return 0.3010299957f * (float)(1.442695040888963387 * log((double)x)); }   // celt_log10, lpcnet_enc.c:113
OB_DEV float le_clamp1(float x) { return x < -1.f ? -1.f : x > 1.f ? 1.f : x; }
OB_DEV float le_dot(const float *a, const float *b, int n) {   // celt_inner_prod_c / one lag of celt_pitch_xcorr_c: index order
  float s = 0;
#pragma unroll 4
  for (int i = 0; i < n; i++) s = s + a[i] * b[i];
  return s;
}
// lpc_from_cepstrum (freq.c:303-313) with lpc_from_bands (:265-289), interp_band_gain (:192-205), inverse_transform (:246-263), lpcn_lpc (:82-127)
OB_DEV void le_lpc_from_cepstrum(LeWS *ws, const float *cepstrum, float *lpc, int lane) {
  if (lane < kLeBands) ws->tmp[lane] = cepstrum[lane] + (lane == 0 ? 4.f : 0.f);
  wsync();
  if (lane < kLeBands) ws->Ex[lane] = (float)(pow((double)10.f, (double)le_idct(ws->tmp, lane)) * ob_le_comp[lane]);
  wsync();
  // interp_band_gain into win[0..160], the conjugate-symmetric extension is real and even: x[i] = Xr[min(i, 320 - i)]
  for (int i = lane; i < kLeFreq; i += 32) {
    float v = 0;
    if (i < kLeFreq - 1) {
      int b = 0;
      while (b < kLeBands - 2 && ob_le_eband[b + 1] * 4 <= i) b++;
      const int size = (ob_le_eband[b + 1] - ob_le_eband[b]) * 4;
      const float frac = (float)(i - ob_le_eband[b] * 4) / size;
      v = (1 - frac) * ws->Ex[b] + frac * ws->Ex[b + 1];
    }
    ws->win[i] = v;
  }
  wsync();
  for (int i = lane; i < kLeWin; i += 32) {
    LeCpx v;
    v.r = 0.0031250000f * ws->win[i < kLeFreq ? i : kLeWin - i];
    v.i = 0.0031250000f * (i < kLeFreq ? 0.f : -0.f);
    ws->fa[ob_lpcnet_fft_bitrev[i]] = v;
  }
  wsync();
  le_fft320(ws->fa, lane);
  if (lane <= kLeOrder) {
    float a = kLeWin * ws->fa[lane == 0 ? 0 : kLeWin - lane].r;
    if (lane == 0) a = (float)(a + (a * 1e-4 + 320 / 12 / 38.));
    else a = (float)(a * (1 - 6e-5 * lane * lane));
    ws->ac[lane] = a;
  }
  wsync();
  if (lane == 0) {   // lpcn_lpc: the float build's Levinson recursion
    const float *ac = ws->ac;
    float error = ac[0];
    for (int i = 0; i < kLeOrder; i++) lpc[i] = 0;
    if (ac[0] != 0) {
      for (int i = 0; i < kLeOrder; i++) {
        float rr = 0;
        for (int j = 0; j < i; j++) rr += lpc[j] * ac[i - j];
        rr += ac[i + 1];
        const float r = -rr / error;
        lpc[i] = r;
        for (int j = 0; j < (i + 1) >> 1; j++) {
          const float t1 = lpc[j], t2 = lpc[i - 1 - j];
          lpc[j] = t1 + r * t2;
          lpc[i - 1 - j] = t2 + r * t1;
        }
        error = error - (r * r) * error;
        if (error < .001f * ac[0]) break;
      }
    }
  }
  wsync();
}

// everything of compute_frame_features (lpcnet_enc.c:115-205) in front of compute_pitchdnn, for stream s
OB_DEV void lpcnet_enc_analysis_warp(const LpcnetEncParams &p, int s, LeWS *ws) {
  const int lane = lane_id();
  ObLpcnetEncState *st = p.state + s;
  // preemphasis (:207-215) into win[160..319]; the window's first half and aligned_in's first 80 samples come from analysis_mem
  for (int i = lane; i < kLeFrame; i += 32) {
    float cur, prev;
    if (p.pcm_is_int16) {
      const int16_t *q = (const int16_t *)p.pcm + (size_t)s * (p.pcm_stride ? p.pcm_stride : kLeFrame);
      cur = (float)q[i];
      prev = i ? -0.85f * (float)q[i - 1] : st->mem_preemph;
    } else {
      const float *q = (const float *)p.pcm + (size_t)s * (p.pcm_stride ? p.pcm_stride : kLeFrame);
      cur = q[i];
      prev = i ? -0.85f * q[i - 1] : st->mem_preemph;
    }
    ws->win[kLeFrame + i] = cur + prev;
    ws->win[i] = st->analysis_mem[i];
    if (i == kLeFrame - 1) ws->x[0] = -0.85f * cur;   // parked: the new mem_preemph
  }
  for (int i = lane; i < kLeBuf; i += 32) {   // exc_buf / lp_buf moved down by one frame (:159-160)
    ws->exc[i] = i < kLePitchMax ? st->exc_buf[i + kLeFrame] : 0.f;
    ws->lp[i] = i < kLePitchMax ? st->lp_buf[i + kLeFrame] : 0.f;
  }
  wsync();
  const float new_preemph = ws->x[0];
  wsync();
  // x[] = pitch_mem, then aligned_in = the last 80 samples of the previous frame and the first 80 of this one (:131,161-164)
  for (int i = lane; i < kLeFrame + kLeOrder; i += 32) ws->x[i] = i < kLeOrder ? st->pitch_mem[i] : ws->win[kLeTrain + i - kLeOrder];
  wsync();
  for (int i = lane; i < kLeFrame; i += 32) st->analysis_mem[i] = ws->win[kLeFrame + i];
  if (lane < kLeOrder) st->pitch_mem[lane] = ws->x[kLeFrame + lane];
  wsync();
  // apply_window (freq.c:315-321) and the transform
  for (int i = lane; i < kLeFrame; i += 32) {
    ws->win[i] *= ob_lpcnet_half_window[i];
    ws->win[kLeWin - 1 - i] *= ob_lpcnet_half_window[i];
  }
  wsync();
  le_forward(ws->fa, ws->win, lane);
  if (lane < kLeBands) ws->Ex[lane] = le_band_energy<false>(ws->fa, lane);
  // instantaneous-frequency features (:133-145)
  if (lane < kLeIfFreq) {
    const LeCpx X = ws->fa[lane];
    float *f = p.if_features + (size_t)s * kLeIf;
    if (lane == 0) f[0] = le_clamp1((1.f / 64) * (10.f * le_log10(1e-15f + X.r * X.r) - 6.f));
    else {
      const float pr = st->prev_if[2 * lane], pi = st->prev_if[2 * lane + 1];
      LeCpx prod;
      prod.r = X.r * pr + X.i * pi;
      prod.i = X.i * pr - X.r * pi;
      const float norm_1 = (float)(1.f / sqrt((double)(1e-15f + prod.r * prod.r + prod.i * prod.i)));
      f[3 * lane - 2] = prod.r * norm_1;
      f[3 * lane - 1] = prod.i * norm_1;
      f[3 * lane] = le_clamp1((1.f / 64) * (10.f * le_log10(1e-15f + X.r * X.r + X.i * X.i) - 6.f));
    }
    st->prev_if[2 * lane] = X.r;
    st->prev_if[2 * lane + 1] = X.i;
  }
  wsync();
  if (lane < kLeBands) ws->tmp[lane] = le_log10(1e-2f + ws->Ex[lane]);
  wsync();
  if (lane == 0) {   // the follower (:148-155)
    float logMax = -2, follow = -2;
    for (int i = 0; i < kLeBands; i++) {
      float v = ws->tmp[i];
      const float a = follow - 2.5f > v ? follow - 2.5f : v;
      v = logMax - 8 > a ? logMax - 8 : a;
      logMax = logMax > v ? logMax : v;
      follow = follow - 2.5f > v ? follow - 2.5f : v;
      ws->Ly[i] = v;
    }
  }
  wsync();
  if (lane < kLeBands) ws->feat[lane] = le_dct(ws->Ly, lane) - (lane == 0 ? 4.f : 0.f);
  wsync();
  le_lpc_from_cepstrum(ws, ws->feat, ws->lpc, lane);
  if (lane < kLeOrder) { ws->feat[kLeBands + 2 + lane] = ws->lpc[lane]; st->lpc[lane] = ws->lpc[lane]; }
  // celt_fir (celt_lpc.c:93-153, C build): y[i] = x[i] + sum_j lpc[15 - j] x[i - 16 + j], j ascending
  for (int i = lane; i < kLeFrame; i += 32) {
    float v = ws->x[kLeOrder + i];
#pragma unroll
    for (int j = 0; j < kLeOrder; j++) v = v + ws->lpc[kLeOrder - 1 - j] * ws->x[i + j];
    ws->lp[kLePitchMax + i] = v;
  }
  wsync();
  for (int i = lane; i < kLeFrame; i += 32) ws->exc[kLePitchMax + i] = ws->lp[kLePitchMax + i] + .7f * (i ? ws->lp[kLePitchMax + i - 1] : st->pitch_filt);
  wsync();
  if (lane == 0) {   // biquad (:84-111) in place on the new part of lp_buf
    st->pitch_filt = ws->lp[kLePitchMax + kLeFrame - 1];
    const float b0 = -0.84946f, b1 = 1.f, a0 = -1.54220f, a1 = 0.70781f;
    float mem0 = st->lp_mem[0], mem1 = st->lp_mem[1];
    for (int i = 0; i < kLeFrame; i++) {
      const float xi = ws->lp[kLePitchMax + i], yi = xi + mem0, mem00 = mem0;
      mem0 = (b0 - a0) * xi + mem1 - a0 * mem0;
      mem1 = (b1 - a1) * xi + 1e-30f - a1 * mem00;
      ws->lp[kLePitchMax + i] = yi;
    }
    st->lp_mem[0] = mem0;
    st->lp_mem[1] = mem1;
    st->mem_preemph = new_preemph;
  }
  // celt_pitch_xcorr (C build: every lag summed in index order) and the running energy normalisation (:176-196)
  float *xc = ws->win, *en = (float *)ws->fa;
  for (int i = lane; i < kLeXcorr; i += 32) xc[i] = le_dot(&ws->exc[kLePitchMax], &ws->exc[i], kLeFrame);
  float e = 0;
  if (lane == 1) e = le_dot(&ws->exc[kLePitchMax], &ws->exc[kLePitchMax], kLeFrame);
  if (lane == 0) e = le_dot(&ws->exc[0], &ws->exc[0], kLeFrame);
  const float ener0 = shfl(e, 1);
  if (lane == 0) {
    double ener1 = e;
    for (int i = 0; i < kLeXcorr; i++) {
      en[i] = (float)(1 + ener0 + ener1);
      ener1 += ws->exc[i + kLeFrame] * (double)ws->exc[i + kLeFrame] - ws->exc[i] * (double)ws->exc[i];
    }
  }
  wsync();
  float *cin = p.conv_in + (size_t)s * kLeConvIn;
  for (int i = lane; i < kLeXcorr; i += 32) cin[1 + i] = (2 * xc[i]) / en[i];
  for (int i = lane; i < kLeBuf; i += 32) { st->exc_buf[i] = ws->exc[i]; st->lp_buf[i] = ws->lp[i]; }
  if (lane < kLeBands + 2 + kLeOrder) st->features[lane] = ws->feat[lane];
  if (lane + 32 < kLeFeat) st->features[lane + 32] = ws->feat[lane + 32];
}

// the tail of compute_pitchdnn (pitchdnn.c:45-58) and of compute_frame_features (:197-204); lpbuf: 416 floats of per-warp shared memory
OB_DEV void lpcnet_enc_finish_warp(const LpcnetEncParams &p, int s, float *lpbuf) {
  const int lane = lane_id();
  ObLpcnetEncState *st = p.state + s;
  const float *out = p.pitch_out + (size_t)s * kLePitchOut;
  for (int i = lane; i < kLeBuf; i += 32) lpbuf[i] = st->lp_buf[i];
  float best = -1.f;
  int pos = 0;
  for (int i = lane; i < 180; i += 32) {
    const float v = out[i];
    if (v > best) { best = v; pos = i; }
  }
  for (int m = 16; m > 0; m >>= 1) {   // the first index holding the maximum (a scan with `>` keeps the earliest)
    const float ob_ = shfl_xor(best, m);
    const int op = shfl_xor(pos, m);
    if (ob_ > best || (ob_ == best && op < pos)) { best = ob_; pos = op; }
  }
  if (!(best > -1.f)) pos = 0;
  float sum = 0, count = 0;
  for (int i = pos - 2 > 0 ? pos - 2 : 0; i <= (pos + 2 < 179 ? pos + 2 : 179); i++) {
    const float q = (float)exp((double)out[i]);
    sum += q * i;
    count += q;
  }
  const float dnn_pitch = (float)((1.f / 60.f) * (sum / count) - 1.5);
  const int pitch = (int)floor(.5 + 256. / pow((double)2.f, ((1. / 60.) * ((dnn_pitch + 1.5) * 60))));
  wsync();
  const float *x = &lpbuf[kLePitchMax], *y = &lpbuf[kLePitchMax - pitch];
  float d = 0;
  if (lane == 0) d = le_dot(x, x, kLeFrame);
  if (lane == 1) d = le_dot(y, y, kLeFrame);
  if (lane == 2) d = le_dot(x, y, kLeFrame);
  const float xx = shfl(d, 0), yy = shfl(d, 1), xy = shfl(d, 2);
  float frame_corr = (float)(xy / sqrt((double)(1 + xx * yy)));
  frame_corr = (float)(log(1.f + exp((double)(5.f * frame_corr))) / log(1 + exp((double)5.f)));
  if (lane == 0) {
    st->dnn_pitch = dnn_pitch;
    st->features[kLeBands] = dnn_pitch;
    st->features[kLeBands + 1] = frame_corr - .5f;
  }
  wsync();
  float *f = p.features + (size_t)s * kLeFeat;
  for (int i = lane; i < kLeFeat; i += 32) f[i] = i == kLeBands ? dnn_pitch : i == kLeBands + 1 ? frame_corr - .5f : st->features[i];
}

// burg_cepstral_analysis (freq.c:156-190) of one 160-sample frame: ceps [36]; the two 80-sample halves run the Burg recursion on two lanes
struct LeBurgWS { float in[2 * 80]; float lpc[2 * kLeOrder]; float g[2]; float win[kLeWin]; LeCpx fa[kLeWin]; float E[kLeBands], Ly[kLeBands], c[2 * kLeBands]; };
struct LpcnetBurgParams { float *ceps; const float *x; int n; };
OB_DEV void lpcnet_burg_cepstrum_warp(const LpcnetBurgParams &p, int s, LeBurgWS *ws) {
  const int lane = lane_id(), len = kLeFrame / 2;
  const float *x = p.x + (size_t)s * kLeFrame;
  for (int i = lane; i < 2 * len; i += 32) {
    const int h = i / len, k = i - h * len;
    if (k < len - 1) ws->in[h * (len - 1) + k] = x[h * len + k + 1] - 0.85f * x[h * len + k];
  }
  wsync();
  if (lane < 2) {   // silk_burg_analysis (dnn/burg.c:97-246 == silk/float/burg_modified_FLP.c) with the dnn copy's clamp of the returned energy
    SilkBurgParams b;
    b.A = ws->lpc; b.nrg = ws->g; b.x = ws->in; b.minInvGain = 1e-3f; b.n_items = 2; b.subfr_length = len - 1; b.nb_subfr = 1; b.D = kLeOrder;
    silk_burg_thread(b, lane);
  }
  wsync();
  for (int h = 0; h < 2; h++) {
    float g = ws->g[h];
    g = g > 0 ? g : 0;
    g /= len - 2 * (kLeOrder - 1);
    for (int i = lane; i < kLeWin; i += 32)
      ws->win[i] = i == 0 ? 1.f : i <= kLeOrder ? (float)(-ws->lpc[h * kLeOrder + i - 1] * pow(.995, (double)i)) : 0.f;
    wsync();
    le_forward(ws->fa, ws->win, lane);
    if (lane < kLeBands) {
      float e = le_band_energy<true>(ws->fa, lane);
      e = (float)(e * (.45 * g * (1.f / ((float)kLeWin * kLeWin * kLeWin))));
      ws->E[lane] = (float)log10(1e-2 + e);
    }
    wsync();
    if (lane == 0) {
      float logMax = -2, follow = -2;
      for (int i = 0; i < kLeBands; i++) {
        float v = ws->E[i];
        double a = follow - 2.5;
        a = a > v ? a : v;
        const double b = logMax - 8;
        v = (float)(b > a ? b : a);
        logMax = logMax > v ? logMax : v;
        a = follow - 2.5;
        follow = (float)(a > v ? a : v);
        ws->Ly[i] = v;
      }
    }
    wsync();
    if (lane < kLeBands) ws->c[h * kLeBands + lane] = le_dct(ws->Ly, lane) + (lane == 0 ? -4.f : 0.f);
    wsync();
  }
  if (lane < kLeBands) {
    const float c0 = ws->c[lane], c1 = ws->c[kLeBands + lane];
    p.ceps[(size_t)s * 2 * kLeBands + lane] = (float)(.5 * (c0 + c1));
    p.ceps[(size_t)s * 2 * kLeBands + kLeBands + lane] = c0 - c1;
  }
}
// ---- DRED encoder front end: dred_convert_to_16k (dnn/dred_encoder.c:133-213) with filter_df2t (:109-126), one thread per stream.  The input chunk
// (in_len samples per channel at Fs) becomes out_len samples at 16 kHz in int16 range, appended to the stream's input buffer: mono downmix through
// FLOAT2INT16, zero-stuffing by `up`, the 7th-order elliptic low-pass in transposed direct form II (state in registers), decimation by 3. ----
struct DredConvParams {
  const float *pcm;        // [n][pcm_stride]: this chunk starts at pcm_off
  float *input_buffer;     // [n][640]: written at `fill`
  float *resample_mem;     // [n][9]
  int n, Fs, channels, in_len, out_len, fill, pcm_off, pcm_stride;
};
OB_DEV float dred_to_int16(float x) {   // FLOAT2INT16 (celt/float_cast.h:150-156): scale, clamp, round to nearest even
  x = x * 32768.f;
  x = x > -32768 ? x : -32768;
  x = x < 32767 ? x : 32767;
  return rintf(x);
}
OB_DEV void dred_convert_thread(const DredConvParams &p, int s) {
  const float *in = p.pcm + (size_t)s * p.pcm_stride + p.pcm_off;
  float *out = p.input_buffer + (size_t)s * 640 + p.fill;
  const int up = p.Fs == 8000 ? 2 : p.Fs == 12000 ? 4 : p.Fs == 24000 ? 2 : 1;
  float b[8], a[8], b0;
  if (p.Fs == 12000) {        // ellip(7, .2, 70, 5800/24000)
    const float bb[8] = {-0.001017101081f, 0.003673127243f, 0.001009165267f, 0.001009165267f, 0.003673127243f, -0.001017101081f, 0.002033596776f, 0.f};
    const float aa[8] = {-4.930414411612f, 11.291643096504f, -15.322037343815f, 13.216403930898f, -7.220409219553f, 2.310550142771f, -0.334338618782f, 0.f};
    for (int j = 0; j < 8; j++) { b[j] = bb[j]; a[j] = aa[j]; }
    b0 = 0.002033596776f;
  } else if (p.Fs == 8000) {  // ellip(7, .2, 70, 3900/8000)
    const float bb[8] = {0.081670120929f, 0.180401598565f, 0.259391051971f, 0.259391051971f, 0.180401598565f, 0.081670120929f, 0.020109185709f, 0.f};
    const float aa[8] = {-1.393651933659f, 2.609789872676f, -2.403541968806f, 2.056814957331f, -1.148908574570f, 0.473001413788f, -0.110359852412f, 0.f};
    for (int j = 0; j < 8; j++) { b[j] = bb[j]; a[j] = aa[j]; }
    b0 = 0.020109185709f;
  } else {                    // 48 and 24 kHz: ellip(7, .2, 70, 7750/24000)
    const float bb[8] = {0.005873358047f, 0.012980854831f, 0.014531340042f, 0.014531340042f, 0.012980854831f, 0.005873358047f, 0.004523418224f, 0.f};
    const float aa[8] = {-3.878718597768f, 7.748834257468f, -9.653651699533f, 8.007342726666f, -4.379450178552f, 1.463182111810f, -0.231720677804f, 0.f};
    for (int j = 0; j < 8; j++) { b[j] = bb[j]; a[j] = aa[j]; }
    b0 = 0.004523418224f;
  }
  float mem[9];
  for (int j = 0; j < 9; j++) mem[j] = p.resample_mem[(size_t)s * 9 + j];
  const int total = up * p.in_len, step = p.Fs == 8000 || p.Fs == 16000 ? 1 : 3;
  int phase = 0, k = 0, o = 0, ustep = 0;   // ustep counts the zero-stuffing positions, phase the decimation
  for (int i = 0; i < total; i++) {
    float xi = 0.f;
    if (ustep == 0) {
      const float v = p.channels == 1 ? dred_to_int16(in[k]) : dred_to_int16((float)(.5 * (double)(in[2 * k] + in[2 * k + 1])));
      xi = up * v + 1e-30f;
      k++;
    }
    if (++ustep == up) ustep = 0;
    float yi = xi;
    if (p.Fs != 16000) {
      yi = xi * b0 + mem[0];
      const float nyi = -yi;
#pragma unroll
      for (int j = 0; j < 8; j++) mem[j] = mem[j + 1] + b[j] * xi + a[j] * nyi;
    }
    if (phase == 0) { if (o < p.out_len) out[o] = yi; o++; }
    if (++phase == step) phase = 0;
  }
  for (int j = 0; j < 9; j++) p.resample_mem[(size_t)s * 9 + j] = mem[j];
}
}  // namespace ob
