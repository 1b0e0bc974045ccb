// celt_dec_plc.cuh — pitch-based packet-loss concealment of the batched CELT decoder (kernel P, ob_celt_dec_plc).
//
// Replaces, for the streams whose frame is lost, what the vector + synthesis kernels do for a decoded frame
// (celt_decode_lost, celt/celt_decoder.c:813-1079; celt_plc_pitch_search :552-574).  The entropy kernel has already decided
// the kind of concealment and advanced the loss counters (ent_lost_frame); noise PLC goes through the ordinary synthesis path.
//
//   one CTA of 2 warps per 8 streams; it conceals the lost ones among them, one after the other (nothing lost: a 2-load scan):
//     both warps : pitch_downsample + pitch_search over the 2048-sample history ring (first lost frame of a burst only)
//     warp c     : channel c: windowed autocorrelation -> LPC(24) -> excitation (FIR) -> decay estimate -> periodic extrapolation
//                  of N+120 samples -> synthesis IIR -> energy check -> history ring / overhang / de-emphasis -> PCM
//
// Every sum whose value feeds a DECISION (pitch correlations, energies, the IIR recursion) is accumulated in the reference's
// order; only the de-emphasis uses the same warp scan as the synthesis kernel.  The serial pieces (IIR: 24 taps x <= 1080 samples)
// run on one lane: a lost frame is the rare case, and all lost streams of a batch run side by side.
#pragma once
#include "celt_dec_kernels.cuh"
#include "celt_enc_pre.cuh"   // pf_dot, pf_find_best_pitch (shared with the encoder's pitch pre-filter)

namespace ob {

struct DecPlcParams {
  const ObFrameSide *side;  // [n_streams]
  ObDecSynthState *state;   // [n_streams]
  void *pcm;                // [n_streams][pcm_stride] interleaved, float or int16
  int pcm_stride;
  int pcm_is_int16;
  int n_streams;
  int channels;
  int downsample;
  int apply_gain;   // OPUS_SET_GAIN != 0: concealed audio is scaled like decoded audio
  float gain;
  int defer_clip = 0;   // see DecSynthParams::defer_clip
};

constexpr int kPlcMaxPeriod = 1024, kPlcLagMax = 720, kPlcLagMin = 100, kPlcOrd = OB_PLC_LPC_ORDER;
constexpr int kPlcNew = OB_MAXN + OB_OVERLAP;   // most samples one frame extrapolates

struct PlcWS {
  float sc[128];                      // [0,64) per-channel scalars / autocorrelation, [64,112) LPC, [112..) pitch search results
  float lp[kPlcMaxPeriod];            // pitch search: decimated + whitened history; later channel 0's copy of the last period
  float scratch[kPlcMaxPeriod];       // pitch search: raw decimated history, then xcorr + energy track; later channel 1's period
  float exc[2][kPlcOrd + kPlcNew + 8];   // excitation (24 + 1024), later IIR memory + output (24 + N + 120)
  float work[2][kPlcNew + 8];            // windowed copy / FIR output / extrapolated excitation
};

OB_DEVNI void dec_plc_block(const DecPlcParams &p, int stream, unsigned char *smem) {
  const ObFrameSide *side = p.side + stream;
  const int type = OB_PLC_TYPE(side->plc);
  if (side->status <= 0 || type < OB_PLC_PITCH) return;   // uniform over the CTA
  const int lane = lane_id(), warp = warp_in_block(), tid = warp * 32 + lane, nth = warps_per_block() * 32;
  const int CC = p.channels, ds = p.downsample;
  const size_t base = (size_t)stream * p.pcm_stride + (size_t)side->pcm_off * CC;
  if (type == OB_PLC_ZEROS) {  // nothing decoded yet: silence (src/opus_decoder.c:333-341)
    const int total = side->status / ds * CC;
    if (p.pcm_is_int16) { int16_t *o = (int16_t *)p.pcm + base; for (int e = tid; e < total; e += nth) o[e] = 0; }
    else { float *o = (float *)p.pcm + base; for (int e = tid; e < total; e += nth) o[e] = 0.f; }
    return;
  }
  PlcWS *w = (PlcWS *)smem;
  ObDecSynthState *st = p.state + stream;
  const int LM = side->LM, N = OB_SHORT_N << LM, Nd = N / ds;
  const int first = (side->plc & OB_PLC_FIRST) != 0;
  const uint32_t pos = st->pos;
  // decode_mem[c][k] of the reference (before its move) is the ring entry of absolute position pos - 2048 + k
#define PLC_H(c, k) st->hist[c][(pos + (uint32_t)(k)) & (OB_HIST - 1)]
  int *isc = (int *)w->sc;

  // ---- celt_plc_pitch_search (celt_decoder.c:552-574) ----
  if (first) {
    const int len = OB_HIST >> 1;
    float *xlp = w->scratch;
    for (int i = tid; i < len; i += nth) {  // pitch_downsample (celt/pitch.c:140-228)
      float v = i == 0 ? .25f * PLC_H(0, 1) + .5f * PLC_H(0, 0) : .25f * PLC_H(0, 2 * i - 1) + .25f * PLC_H(0, 2 * i + 1) + .5f * PLC_H(0, 2 * i);
      if (CC == 2)
        v += i == 0 ? .25f * PLC_H(1, 1) + .5f * PLC_H(1, 0) : .25f * PLC_H(1, 2 * i - 1) + .25f * PLC_H(1, 2 * i + 1) + .5f * PLC_H(1, 2 * i);
      xlp[i] = v;
    }
    bsync();
    if (tid < 5) {  // _celt_autocorr, 5 lags
      const int k = tid, fastN = len - 4;
      float ac = pf_dot(xlp, xlp + k, fastN), d = 0.f;
#pragma unroll 1
      for (int i = k + fastN; i < len; i++) d = d + xlp[i] * xlp[i - k];
      w->sc[k] = ac + d;
    }
    bsync();
    if (tid == 0) {  // _celt_lpc(4), bandwidth expansion, one more zero
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
    {  // celt_fir5
      const float l0 = w->sc[8], l1 = w->sc[9], l2 = w->sc[10], l3 = w->sc[11], l4 = w->sc[12];
      for (int i = tid; i < len; i += nth) {
        float sum = xlp[i];
        sum = sum + l0 * (i >= 1 ? xlp[i - 1] : 0.f);
        sum = sum + l1 * (i >= 2 ? xlp[i - 2] : 0.f);
        sum = sum + l2 * (i >= 3 ? xlp[i - 3] : 0.f);
        sum = sum + l3 * (i >= 4 ? xlp[i - 4] : 0.f);
        sum = sum + l4 * (i >= 5 ? xlp[i - 5] : 0.f);
        w->lp[i] = sum;
      }
    }
    bsync();
    // pitch_search(lp + 360, lp, 2048 - 720, 720 - 100) (celt/pitch.c:307-425)
    const float *pb = w->lp, *xl = w->lp + (kPlcLagMax >> 1);
    const int slen = OB_HIST - kPlcLagMax, mp = kPlcLagMax - kPlcLagMin;
    const int len4 = slen >> 2, mp4 = mp >> 2, len2 = slen >> 1, mp2 = mp >> 1;
    float *xcorr = w->scratch, *track = w->scratch + 512;
    for (int i = tid; i < mp4; i += nth) {
      float s = 0.f;
#pragma unroll 8
      for (int j = 0; j < len4; j++) s = s + xl[2 * j] * pb[2 * (i + j)];
      xcorr[i] = s;
    }
    bsync();
    pf_find_best_pitch(xcorr, pb, 2, len4, mp4, track, isc + 112, tid, nth);
    {
      const int b0 = isc[112], b1 = isc[113];
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
    pf_find_best_pitch(xcorr, pb, 1, len2, mp2, track, isc + 112, tid, nth);
    if (tid == 0) {
      const int bp0 = isc[112];
      int offset = 0;
      if (bp0 > 0 && bp0 < mp2 - 1) {
        const float aa = xcorr[bp0 - 1], bb = xcorr[bp0], cc = xcorr[bp0 + 1];
        if ((cc - aa) > .7f * (bb - aa)) offset = 1;
        else if ((aa - cc) > .7f * (bb - cc)) offset = -1;
      }
      const int pitch_index = kPlcLagMax - (2 * bp0 - offset);
      isc[116] = pitch_index;
      st->last_pitch_index = pitch_index;
    }
  } else if (tid == 0) isc[116] = st->last_pitch_index;
  bsync();
  const int pitch_index = isc[116];
  const float fade = first ? 1.f : .8f;
  const int exc_length = OB_IMIN(2 * pitch_index, kPlcMaxPeriod);
  const int elen = N + OB_OVERLAP;   // extrapolation_len

  if (warp < CC) {
    const int c = warp;
    float *exc = w->exc[c] + kPlcOrd;           // exc[-24 .. 1024)
    float *work = w->work[c];
    float *lpc = w->sc + 64 + c * kPlcOrd;
    float *per = c == 0 ? w->lp : w->scratch;   // the last pitch period of the decoded signal
    for (int i = lane; i < kPlcMaxPeriod + kPlcOrd; i += 32) exc[i - kPlcOrd] = PLC_H(c, OB_HIST - kPlcMaxPeriod - kPlcOrd + i);
    for (int j = lane; j < pitch_index; j += 32) per[j] = PLC_H(c, OB_HIST - pitch_index + j);
    wsync();
    if (first) {
      // LPC of the last 1024 samples: _celt_autocorr with the MDCT window on both ends (celt/celt_lpc.c:284-370)
      for (int i = lane; i < kPlcMaxPeriod; i += 32) {
        float v = exc[i];
        if (i < OB_OVERLAP) v = v * ob_window[i];
        else if (i >= kPlcMaxPeriod - OB_OVERLAP) v = v * ob_window[kPlcMaxPeriod - 1 - i];
        work[i] = v;
      }
      wsync();
      if (lane <= kPlcOrd) {
        const int k = lane, fastN = kPlcMaxPeriod - kPlcOrd;
        float ac = pf_dot(work, work + k, fastN), d = 0.f;
#pragma unroll 1
        for (int i = k + fastN; i < kPlcMaxPeriod; i++) d = d + work[i] * work[i - k];
        w->sc[c * 32 + k] = ac + d;
      }
      wsync();
      if (lane == 0) {
        float ac[kPlcOrd + 1], l[kPlcOrd];
        for (int i = 0; i <= kPlcOrd; i++) ac[i] = w->sc[c * 32 + i];
        ac[0] *= 1.0001f;                                                            // -40 dB noise floor
        for (int i = 1; i <= kPlcOrd; i++) ac[i] -= ac[i] * (0.008f * 0.008f) * i * i;   // lag window
        for (int i = 0; i < kPlcOrd; i++) l[i] = 0.f;
        float error = ac[0];
        if (ac[0] > 1e-10f) {  // _celt_lpc (celt/celt_lpc.c:37-94)
#pragma unroll 1
          for (int i = 0; i < kPlcOrd; i++) {
            float rr = 0.f;
            for (int j = 0; j < i; j++) rr += l[j] * ac[i - j];
            rr += ac[i + 1];
            const float r = -(rr / error);
            l[i] = r;
            for (int j = 0; j < (i + 1) >> 1; j++) {
              const float t1 = l[j], t2 = l[i - 1 - j];
              l[j] = t1 + r * t2;
              l[i - 1 - j] = t2 + r * t1;
            }
            error = error - (r * r) * error;
            if (error <= .001f * ac[0]) break;
          }
        }
        for (int i = 0; i < kPlcOrd; i++) { lpc[i] = l[i]; st->lpc[c][i] = l[i]; }
      }
    } else if (lane < kPlcOrd) lpc[lane] = st->lpc[c][lane];
    wsync();
    // excitation of the last exc_length samples: celt_fir (celt/celt_lpc.c:142-189), out of place
    {
      const float *x = exc + kPlcMaxPeriod - exc_length;
      for (int i = lane; i < exc_length; i += 32) {
        float sum = x[i];
#pragma unroll
        for (int j = 0; j < kPlcOrd; j++) sum = sum + lpc[kPlcOrd - 1 - j] * x[i + j - kPlcOrd];
        work[i] = sum;
      }
      wsync();
      for (int i = lane; i < exc_length; i += 32) exc[kPlcMaxPeriod - exc_length + i] = work[i];
      wsync();
    }
    // is the excitation decaying? (celt_decoder.c:909-931)  Two serial sums, one lane each.
    float decay;
    {
      const int decay_length = exc_length >> 1;
      float E = 1.f;
      if (lane < 2) {
        const float *e = exc + kPlcMaxPeriod - (lane + 1) * decay_length;
#pragma unroll 4
        for (int i = 0; i < decay_length; i++) E += e[i] * e[i];
      }
      float E1 = shfl(E, 0);
      const float E2 = shfl(E, 1);
      E1 = E1 < E2 ? E1 : E2;
      decay = f_sqrt(f_div(E1, E2));
    }
    // extrapolate the excitation with period pitch_index, each period scaled down once more (celt_decoder.c:938-963)
    {
      const int offs = kPlcMaxPeriod - pitch_index;
      for (int i = lane; i < elen; i += 32) {
        const int k = i / pitch_index, j = i - k * pitch_index;
        float att = fade * decay;
        for (int q = 0; q < k; q++) att = att * decay;
        work[i] = att * exc[offs + j];
      }
    }
    wsync();
    // IIR memory: the 24 samples before the frame (celt_decoder.c:964-970), stored ahead of the output
    float *y = w->exc[c] + kPlcOrd;   // y[-24 .. elen)
    if (lane < kPlcOrd) y[-1 - lane] = PLC_H(c, OB_HIST - 1 - lane);
    wsync();
    int verdict = 0;   // 0 keep, 1 zero, 2 scale
    float ratio = 1.f;
    if (lane == 0) {
      // energy of the signal whose excitation is being copied (celt_decoder.c:957-962)
      float S1 = 0.f;
      {
        int j = 0;
#pragma unroll 4
        for (int i = 0; i < elen; i++, j++) {
          if (j >= pitch_index) j -= pitch_index;
          const float t = per[j];
          S1 += t * t;
        }
      }
      // celt_iir (celt/celt_lpc.c:191-282): y[i] = x[i] - sum_j den[j] * y[i-1-j], in groups of four as the reference unrolls it
      // (the newest terms of outputs 2 and 3 of a group are added last-first)
      float den[kPlcOrd];
#pragma unroll
      for (int j = 0; j < kPlcOrd; j++) den[j] = lpc[j];
      float S2 = 0.f;
#pragma unroll 1
      for (int i = 0; i < elen; i += 4) {
        float s0 = work[i], s1 = work[i + 1], s2 = work[i + 2], s3 = work[i + 3];
#pragma unroll
        for (int j = 0; j < kPlcOrd; j++) {   // rden[j] = den[23-j] against y[i+j-24 + k]
          const float r = den[kPlcOrd - 1 - j];
          s0 = s0 - r * y[i + j - kPlcOrd];
          if (j + 1 < kPlcOrd) s1 = s1 - r * y[i + 1 + j - kPlcOrd];
          if (j + 2 < kPlcOrd) s2 = s2 - r * y[i + 2 + j - kPlcOrd];
          if (j + 3 < kPlcOrd) s3 = s3 - r * y[i + 3 + j - kPlcOrd];
        }
        y[i] = s0;
        s1 = s1 - s0 * den[0];
        y[i + 1] = s1;
        s2 = s2 - s1 * den[0];
        s2 = s2 - s0 * den[1];
        y[i + 2] = s2;
        s3 = s3 - s2 * den[0];
        s3 = s3 - s1 * den[1];
        s3 = s3 - s0 * den[2];
        y[i + 3] = s3;
        S2 += s0 * s0; S2 += s1 * s1; S2 += s2 * s2; S2 += s3 * s3;
      }
      // more energy than the signal it continues (or NaN): attenuate (celt_decoder.c:988-1018)
      if (!(S1 > 0.2f * S2)) verdict = 1;
      else if (S1 < S2) { verdict = 2; ratio = f_sqrt(f_div(S1 + 1.f, S2 + 1.f)); }
    }
    verdict = shfl(verdict, 0);
    ratio = shfl(ratio, 0);
    wsync();
    if (verdict == 1) {
      for (int i = lane; i < elen; i += 32) y[i] = 0.f;
    } else if (verdict == 2) {
      for (int i = lane; i < elen; i += 32) {
        if (i < OB_OVERLAP) {
          const float g = 1.f - ob_window[i] * (1.f - ratio);
          y[i] = g * y[i];
        } else y[i] = ratio * y[i];
      }
    }
    wsync();
    // the frame goes to the history ring, the 120 samples past it wait in `tail` for prefilter_and_fold
    for (int i = lane; i < N; i += 32) st->hist[c][(pos + (uint32_t)i) & (OB_HIST - 1)] = y[i];
    for (int i = lane; i < OB_OVERLAP; i += 32) st->tail[c][i] = y[N + i];
    // de-emphasis, as in the synthesis kernel
    {
      const float coef = 0.85000610f;
      const int seg = (N + 31) >> 5;
      const int j0 = OB_IMIN(lane * seg, N), j1 = OB_IMIN(j0 + seg, N);
      Affine mp = {1.f, 0.f};
      for (int j = j0; j < j1; j++) {
        const float x = y[j] + 1e-30f;
        mp.b = coef * (x + mp.b);
        mp.a = coef * mp.a;
      }
      const Affine inc = group_scan(mp, 32, lane);
      const float pa = shfl_up(inc.a, 1), pb2 = shfl_up(inc.b, 1);
      const float m0 = st->deemph_mem[c];
      float m = lane == 0 ? m0 : pa * m0 + pb2;
      wsync();
      for (int j = j0; j < j1; j++) {
        const float tmp = y[j] + 1e-30f + m;
        m = coef * tmp;
        float v = tmp * (1 / 32768.f);
        if (p.apply_gain) v = v * p.gain;
        y[j] = v;
      }
      const float mlast = shfl(m, (N - 1) / seg);
      if (lane == 0) st->deemph_mem[c] = mlast;
      wsync();
    }
    if (!(side->plc & OB_PLC_PACKET)) {   // a DTX frame inside a packet is soft-clipped with the packet; a lost packet is not
      if (p.defer_clip) {}
      else if (p.pcm_is_int16) soft_clip_channel(y, Nd, ds, &st->declip[c], lane);
      else if (lane == 0 && st->declip[c] != 0.f) st->declip[c] = 0.f;
    }
  }
  bsync();
  if (tid == 0) st->pos = pos + (uint32_t)N;
  {
    const int total = Nd * CC;
    if (p.pcm_is_int16) {
      int16_t *o = (int16_t *)p.pcm + base;
      for (int e = tid; e < total; e += nth) {
        const int j = CC == 2 ? e >> 1 : e, c = CC == 2 ? e & 1 : 0;
        float x = w->exc[c][kPlcOrd + j * ds] * 32768.f;
        x = x > -32768.f ? x : -32768.f;
        x = x < 32767.f ? x : 32767.f;
        o[e] = (int16_t)lrintf(x);
      }
    } else {
      float *o = (float *)p.pcm + base;
      for (int e = tid; e < total; e += nth) {
        const int j = CC == 2 ? e >> 1 : e, c = CC == 2 ? e & 1 : 0;
        o[e] = w->exc[c][kPlcOrd + j * ds];
      }
    }
  }
#undef PLC_H
}

}  // namespace ob
