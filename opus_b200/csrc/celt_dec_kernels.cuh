// celt_dec_kernels.cuh — the kernels of the batched CELT decoder.
//
//   A1 ob_celt_dec_entropy : one THREAD per stream (celt_dec_entropy.cuh).  Packet framing, range
//                            decoding, energies, allocation, band-tree control -> ObScript + ObFrameSide.
//   A2 ob_celt_dec_pvq     : one WARP per stream (celt_dec_pvq.cuh).  Script -> normalised spectrum
//                            X[C][N]: PVQ leaves lane-parallel, then bands in order.
//   B  ob_celt_dec_synth   : one CTA per stream, one warp per output channel (below).  X * band
//                            gains -> inverse MDCT (+TDAC overlap) -> comb post-filter -> de-emphasis
//                            -> PCM.  HBM bound.
//
// Written against simt.cuh so the same source also runs under the test-only host emulator.
#pragma once
#include "celt_dec_entropy.cuh"
#include "celt_dec_pvq.cuh"
#include "celt_fft.cuh"
#include "celt_mdct_fast.cuh"

namespace ob {

// ------------------------------------------------------------------------------------------
// Kernel B: synthesis.  One CTA per stream; warp w produces output channel w.
// ------------------------------------------------------------------------------------------
struct DecSynthParams {
  const ObFrameSide *side;  // [n_streams]
  const float *X;           // [n_streams][2][960]
  ObDecSynthState *state;   // [n_streams]
  void *pcm;                // [n_streams][pcm_stride] interleaved, float or int16
  int pcm_stride;           // elements per stream
  int frame_index;          // frame k of the packet: output goes to element offset k*N*CC of the stream's area
  int pcm_is_int16;
  int n_streams;
  int channels;
  int downsample;           // 48000 / output rate: every downsample-th de-emphasised sample is written (celt_decoder.c:300-405)
  int apply_gain;           // OPUS_SET_GAIN != 0 (src/opus_decoder.c:686-700)
  float gain;
  int no_prefetch = 0;      // A/B aid: load the spectrum where it is used (OPUSB200_SYNTH_NO_PREFETCH)
  int force_smem_fft = 0;   // A/B aid: the general shared-memory FFT for long blocks too (OPUSB200_SYNTH_SMEM_FFT)
  int defer_clip = 0;       // int16 output of a packet that may hold several frames: pcm is a FLOAT staging area and opus_decode()'s soft
                            // clipping runs once over the whole packet afterwards (dec_softclip_packet_block), as in the reference
};

struct SynthWS {  // per warp (= per output channel)
  float freq[OB_MAXN];            // denormalised MDCT coefficients, later the interleave staging area
  float out[64 + OB_MAXN + 64];   // [4..64): previous tail, [64..64+N): IMDCT output (16B aligned start)
  float gain[2 * OB_NBANDS + 2];   // (+2: the struct stays a multiple of 16 bytes, the second warp's arrays keep the 16-byte alignment of the first's)
};
static_assert(sizeof(SynthWS) % 16 == 0, "SynthWS: per-warp instances must keep 16-byte alignment (float4 accesses of the register FFT)");
#define OB_SYN_O 4 /* out[OB_SYN_O + i] <-> reference out_syn[i] ; IMDCT data starts at OB_SYN_O+60 = 64 */

// x[idx] with idx < 0 reaching into the stream's history ring (absolute position pos+idx)
OB_DEV float hist_or_cur(const float *cur, const float *hist, uint32_t pos, int idx) {
  return idx >= 0 ? cur[idx] : hist[(pos + (uint32_t)idx) & (OB_HIST - 1)];
}

// celt/celt.c:238-313 + 166-193, in place (x == y) on cur[istart .. istart+N) where cur[0] is the
// first sample of the frame; samples before the frame come from the history ring `hist`.
OB_DEV void comb_filter_inplace(float *cur, const float *hist, uint32_t pos, int istart, int T0, int T1, int N,
                                float g0, float g1, int tapset0, int tapset1, int lane) {
  if (g0 == 0.f && g1 == 0.f) return;
  T0 = OB_IMAX(T0, 15);
  T1 = OB_IMAX(T1, 15);
  const float gt[9] = {0.3066406250f, 0.2170410156f, 0.1296386719f, 0.4638671875f, 0.2680664062f, 0.f,
                       0.7998046875f, 0.1000976562f, 0.f};
  const float g00 = g0 * gt[tapset0 * 3 + 0], g01 = g0 * gt[tapset0 * 3 + 1], g02 = g0 * gt[tapset0 * 3 + 2];
  const float g10 = g1 * gt[tapset1 * 3 + 0], g11 = g1 * gt[tapset1 * 3 + 1], g12 = g1 * gt[tapset1 * 3 + 2];
  int overlap = OB_OVERLAP;
  if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
  // Outputs at i depend on already-filtered samples no closer than min(T)-2 behind, so blocks of
  // that many samples are independent: process block by block, lanes across a block.
  const int Tmin = OB_IMIN(T0, T1);
  const int iend = istart + N, oend = istart + overlap;
  int i0 = istart;
  wsync();
  while (i0 < oend) {
    int blk = OB_IMIN(Tmin - 2, oend - i0);
    for (int i = i0 + lane; i < i0 + blk; i += 32) {
      float w = ob_window[i - istart];
      float f = w * w;
      float x0 = hist_or_cur(cur, hist, pos, i - T1 + 2), x1 = hist_or_cur(cur, hist, pos, i - T1 + 1);
      float x2 = hist_or_cur(cur, hist, pos, i - T1), x3 = hist_or_cur(cur, hist, pos, i - T1 - 1);
      float x4 = hist_or_cur(cur, hist, pos, i - T1 - 2);
      float a0 = hist_or_cur(cur, hist, pos, i - T0), a1p = hist_or_cur(cur, hist, pos, i - T0 + 1);
      float a1m = hist_or_cur(cur, hist, pos, i - T0 - 1), a2p = hist_or_cur(cur, hist, pos, i - T0 + 2);
      float a2m = hist_or_cur(cur, hist, pos, i - T0 - 2);
      cur[i] = cur[i] + ((1.0f - f) * g00) * a0 + ((1.0f - f) * g01) * (a1p + a1m) + ((1.0f - f) * g02) * (a2p + a2m) +
               (f * g10) * x2 + (f * g11) * (x1 + x3) + (f * g12) * (x0 + x4);
    }
    i0 += blk;
    wsync();
  }
  if (g1 == 0.f) return;
  while (i0 < iend) {
    int blk = OB_IMIN(T1 - 2, iend - i0);
    for (int i = i0 + lane; i < i0 + blk; i += 32) {
      float x0 = hist_or_cur(cur, hist, pos, i - T1 + 2), x1 = hist_or_cur(cur, hist, pos, i - T1 + 1);
      float x2 = hist_or_cur(cur, hist, pos, i - T1), x3 = hist_or_cur(cur, hist, pos, i - T1 - 1);
      float x4 = hist_or_cur(cur, hist, pos, i - T1 - 2);
      cur[i] = cur[i] + g10 * x2 + g11 * (x1 + x3) + g12 * (x0 + x4);
    }
    i0 += blk;
    wsync();
  }
}

// opus_pcm_soft_clip for one channel of one frame (src/opus.c:39-166), which opus_decode() applies before the int16 conversion
// (src/opus_decoder.c:875-878): x[i*stride], i < n, in place.  The whole warp clamps to +-2 and looks for a sample beyond +-1; if there
// is none and no curve is pending from the previous frame, the function is the identity (the common case: one pass over the frame).
// Otherwise lane 0 runs the reference's serial scan.  The reference works on a whole packet: for packets of several frames this
// per-frame form can differ from it where a clipped region crosses a frame boundary.
OB_DEVNI void soft_clip_channel(float *x, int n, int stride, float *declip_mem, int lane) {
  float mx = 0.f;
  for (int i = lane; i < n; i += 32) {
    const float v0 = x[i * stride];
    const float v = v0 > 2.f ? 2.f : v0 < -2.f ? -2.f : v0;
    if (v != v0) x[i * stride] = v;
    const float av = v < 0.f ? -v : v;
    mx = av > mx ? av : mx;
  }
  mx = warp_max(mx);
  float a = *declip_mem;
  if (a == 0.f && mx <= 1.f) return;
  wsync();
  if (lane == 0) {
    int i, curr = 0;
#pragma unroll 1
    for (i = 0; i < n; i++) {  // continue the previous frame's curve up to the first zero crossing
      const float v = x[i * stride];
      if (v * a >= 0) break;
      x[i * stride] = v + a * v * v;
    }
    const float x0 = x[0];
#pragma unroll 1
    while (1) {
#pragma unroll 1
      for (i = curr; i < n; i++)
        if (x[i * stride] > 1 || x[i * stride] < -1) break;
      if (i == n) { a = 0.f; break; }
      int peak_pos = i, start = i, end = i;
      const float xi = x[i * stride];
      float maxval = xi < 0.f ? -xi : xi;
      while (start > 0 && xi * x[(start - 1) * stride] >= 0) start--;
      while (end < n && xi * x[end * stride] >= 0) {
        const float e = x[end * stride], ae = e < 0.f ? -e : e;
        if (ae > maxval) { maxval = ae; peak_pos = end; }
        end++;
      }
      const int special = (start == 0 && xi * x[0] >= 0);
      a = f_div(maxval - 1, maxval * maxval);
      a += a * 2.4e-7f;
      if (xi > 0) a = -a;
#pragma unroll 1
      for (i = start; i < end; i++) { const float v = x[i * stride]; x[i * stride] = v + a * v * v; }
      if (special && peak_pos >= 2) {
        float offset = x0 - x[0];
        const float delta = f_div(offset, (float)peak_pos);
#pragma unroll 1
        for (i = curr; i < peak_pos; i++) {
          offset -= delta;
          float v = x[i * stride] + offset;
          v = v > 1.f ? 1.f : v < -1.f ? -1.f : v;
          x[i * stride] = v;
        }
      }
      curr = end;
      if (curr == n) break;
    }
    *declip_mem = a;
  }
  wsync();
}

OB_DEV void dec_synth_block(const DecSynthParams &p, int stream, unsigned char *smem) {
  const int lane = lane_id();
  const int ch = warp_in_block();  // output channel
  const int CC = p.channels;
  SynthWS *ws = (SynthWS *)smem + ch;
  const ObFrameSide *side = p.side + stream;
  const int status = side->status;
  if (status <= 0) return;  // uniform over the CTA: nothing decoded for this stream in this launch
  const int plc = side->plc;
  if (OB_PLC_TYPE(plc) >= OB_PLC_PITCH) return;  // pitch-based concealment / zeros: ob_celt_dec_plc produces this frame
  ObDecSynthState *st = p.state + stream;
  const int LM = side->LM, C = side->C, transient = side->transient, silence = side->silence;
  const int N = OB_SHORT_N << LM, M = 1 << LM;
  const int nblk = transient ? M : 1, NB = N / nblk, shift = transient ? 3 : 3 - LM;
  int end = side->end_band;
  int bound = silence ? 0 : M * ob_ebands[end];
  const int ds = p.downsample, Nd = N / ds;
  if (ds != 1) bound = OB_IMIN(bound, Nd);  // celt_decoder.c:451-452
  // ---- the spectrum: this lane's loads (30 values at N = 960) are issued before the band gains are computed, so that the HBM latency runs under
  //      the 42 double-precision exp() of the gains (OPUSB200_SYNTH_NO_PREFETCH: load where the value is used, for A/B runs) ----
  const float *X0 = p.X + (size_t)stream * 2 * OB_MAXN;
  const float *X1 = X0 + N;
  const int srcc = (C == 2 && CC == 2) ? ch : 0;
  const float *Xs = srcc ? X1 : X0;
  const bool downmix = CC == 1 && C == 2;
  constexpr int kPer = OB_MAXN / 32;
  float xv[kPer];
  if (!p.no_prefetch) {
#pragma unroll
    for (int k = 0; k < kPer; k++) {
      const int j = lane + 32 * k;
      xv[k] = j < bound ? Xs[j] : 0.f;
    }
  }
  // ---- band gains (denormalise_bands, celt/bands.c:188-257) ----
  for (int i = lane; i < 2 * OB_NBANDS; i += 32) {
    int band = i >= OB_NBANDS ? i - OB_NBANDS : i;
    float lg = side->bandE[i] + ob_emeans[band];
    ws->gain[i] = exp2_ref(32.f < lg ? 32.f : lg);
  }
  wsync();
  // ---- apply gains -> freq ----
  {
    const float *gs = ws->gain + srcc * OB_NBANDS;
    int band = 0;
#pragma unroll
    for (int k = 0; k < kPer; k++) {
      const int j = lane + 32 * k;
      if (j < N) {
        float v = 0.f;
        if (j < bound) {
          int u = j >> LM;
          while (ob_ebands[band + 1] <= u) band++;
          v = (p.no_prefetch ? Xs[j] : xv[k]) * gs[band];
          if (downmix) {  // celt_decoder.c:478-481
            float v2 = X1[j] * ws->gain[OB_NBANDS + band];
            v = (.5f * v) + (.5f * v2);
          }
        }
        ws->freq[j] = v;
      }
    }
  }
  // previous tail -> out[0..60)
  float *out = ws->out + OB_SYN_O;
  if (plc & OB_PLC_FOLD) {
    // prefilter_and_fold (celt_decoder.c:576-625): the previous frame was extrapolated in the time domain, 120 samples past its end.
    // Undo the post-filter on that overhang (comb filter with the gains negated, constant part only: overlap = 0) and fold it
    // with the window, so that it looks like the tail an MDCT frame would have left.
    const float *hist = st->hist[ch];
    const uint32_t pos = st->pos;
    for (int i = lane; i < OB_OVERLAP; i += 32) out[i] = st->tail[ch][i];
    wsync();
    const float g1 = -side->pf_g[1];
    const int T1 = OB_IMAX(side->pf_T[1], 15), tap = side->pf_tap[1];
    const float gt[9] = {0.3066406250f, 0.2170410156f, 0.1296386719f, 0.4638671875f, 0.2680664062f, 0.f,
                         0.7998046875f, 0.1000976562f, 0.f};
    const float g10 = g1 * gt[tap * 3 + 0], g11 = g1 * gt[tap * 3 + 1], g12 = g1 * gt[tap * 3 + 2];
    for (int i = lane; i < OB_OVERLAP; i += 32) {
      float v = out[i];
      if (g1 != 0.f) {
        const float x0 = hist_or_cur(out, hist, pos, i - T1 + 2), x1 = hist_or_cur(out, hist, pos, i - T1 + 1);
        const float x2 = hist_or_cur(out, hist, pos, i - T1), x3 = hist_or_cur(out, hist, pos, i - T1 - 1);
        const float x4 = hist_or_cur(out, hist, pos, i - T1 - 2);
        v = v + g10 * x2 + g11 * (x1 + x3) + g12 * (x0 + x4);
      }
      out[OB_OVERLAP + i] = v;  // etmp, in the still unused IMDCT area
    }
    wsync();
    for (int i = lane; i < OB_OVERLAP / 2; i += 32)
      out[i] = ob_window[i] * out[OB_OVERLAP + OB_OVERLAP - 1 - i] + ob_window[OB_OVERLAP - i - 1] * out[OB_OVERLAP + i];
  } else {
    for (int i = lane; i < OB_OVERLAP / 2; i += 32) out[i] = st->tail[ch][i];
  }
  wsync();
  // ---- IMDCT ----
  // long block (one 960-point transform): the register FFT of celt_mdct_fast.cuh, in place in the overlap-add area (two passes over shared memory
  // instead of ten, no run-time butterfly indices); the arithmetic per butterfly is the reference's either way.  OPUSB200_SYNTH_SMEM_FFT keeps the
  // general path for A/B runs.
  if (nblk == 1 && shift == 0 && !p.force_smem_fft) {
    const Mdct480Global tabs;
    imdct480_warp(ws->freq, out + OB_OVERLAP / 2, &tabs, lane);
    // TDAC mirror (celt/mdct.c:372-388)
    for (int i = lane; i < OB_OVERLAP / 2; i += 32) {
      float *xp1 = out + OB_OVERLAP - 1 - i, *yp1 = out + i;
      float x1 = *xp1, x2 = *yp1;
      float w1 = ob_window[i], w2 = ob_window[OB_OVERLAP - 1 - i];
      *yp1 = x2 * w2 - x1 * w1;
      *xp1 = x2 * w1 + x1 * w2;
    }
    wsync();
  } else {
    // pre-rotation into digit-reversed order (celt/mdct.c:304-331)
    const int N2 = NB, N4 = NB >> 1;  // mdct size is 2*NB
    const float *trig = ob_mdct_trig;
    for (int i = 0; i < shift; i++) trig += (1920 >> i) >> 1;
    const short *bitrev = bitrev_table(shift);
    for (int q = lane; q < nblk * N4; q += 32) {
      const int b = q / N4, i = q - b * N4;
      const float *in = ws->freq + b;
      float x1 = in[2 * i * nblk], x2 = in[nblk * (N2 - 1 - 2 * i)];
      float t0 = trig[i], t1 = trig[N4 + i];
      float yr = x2 * t0 + x1 * t1;
      float yi = x1 * t0 - x2 * t1;
      int rev = bitrev[i];
      float *yp = out + OB_OVERLAP / 2 + b * NB;
      yp[2 * rev + 1] = yr;
      yp[2 * rev] = yi;
    }
    fft_stages((cpx *)(out + OB_OVERLAP / 2), shift, nblk, lane);
    // post-rotation, both ends at once (celt/mdct.c:335-369)
    for (int q = lane; q < nblk * (N4 >> 1); q += 32) {
      const int b = q / (N4 >> 1), i = q - b * (N4 >> 1);
      float *yp0 = out + OB_OVERLAP / 2 + b * NB + 2 * i;
      float *yp1 = out + OB_OVERLAP / 2 + b * NB + N2 - 2 - 2 * i;
      float re = yp0[1], im = yp0[0];
      float t0 = trig[i], t1 = trig[N4 + i];
      float yr = re * t0 + im * t1;
      float yi = re * t1 - im * t0;
      re = yp1[1]; im = yp1[0];
      yp0[0] = yr;
      yp1[1] = yi;
      t0 = trig[N4 - i - 1]; t1 = trig[N2 - i - 1];
      yr = re * t0 + im * t1;
      yi = re * t1 - im * t0;
      yp1[0] = yr;
      yp0[1] = yi;
    }
    wsync();
    // TDAC mirror (celt/mdct.c:372-388)
    for (int q = lane; q < nblk * (OB_OVERLAP / 2); q += 32) {
      const int b = q / (OB_OVERLAP / 2), i = q - b * (OB_OVERLAP / 2);
      float *xp1 = out + b * NB + OB_OVERLAP - 1 - i, *yp1 = out + b * NB + i;
      float x1 = *xp1, x2 = *yp1;
      float w1 = ob_window[i], w2 = ob_window[OB_OVERLAP - 1 - i];
      *yp1 = x2 * w2 - x1 * w1;
      *xp1 = x2 * w1 + x1 * w2;
    }
    wsync();
  }
  // new tail for the next frame
  for (int i = lane; i < OB_OVERLAP / 2; i += 32) st->tail[ch][i] = out[N + i];
  // ---- comb post-filter (celt_decoder.c:1540-1552) ----
  const uint32_t pos = st->pos;
  float *hist = st->hist[ch];
  comb_filter_inplace(out, hist, pos, 0, side->pf_T[0], side->pf_T[1], OB_SHORT_N, side->pf_g[0], side->pf_g[1],
                      side->pf_tap[0], side->pf_tap[1], lane);
  if (LM != 0)
    comb_filter_inplace(out, hist, pos, OB_SHORT_N, side->pf_T[1], side->pf_T[2], N - OB_SHORT_N, side->pf_g[1],
                        side->pf_g[2], side->pf_tap[1], side->pf_tap[2], lane);
  wsync();
  // history ring <- filtered samples
  for (int i = lane; i < N; i += 32) hist[(pos + (uint32_t)i) & (OB_HIST - 1)] = out[i];
  // ---- de-emphasis: tmp_j = x_j + 1e-30 + m ; m = coef*tmp_j  (celt_decoder.c:288-408) as a
  // warp scan: lane l owns samples [l*seg, (l+1)*seg) ----
  {
    const float coef = 0.85000610f;
    const int seg = (N + 31) >> 5;
    const int j0 = OB_IMIN(lane * seg, N), j1 = OB_IMIN(j0 + seg, N);
    Affine mp = {1.f, 0.f};  // maps incoming m to outgoing m over my segment
    for (int j = j0; j < j1; j++) {
      float x = out[j] + 1e-30f;
      mp.b = coef * (x + mp.b);
      mp.a = coef * mp.a;
    }
    Affine inc = group_scan(mp, 32, lane);
    float pa = shfl_up(inc.a, 1), pb = shfl_up(inc.b, 1);
    const float m0 = st->deemph_mem[ch];
    float m = lane == 0 ? m0 : pa * m0 + pb;
    wsync();
    for (int j = j0; j < j1; j++) {
      float tmp = out[j] + 1e-30f + m;
      m = coef * tmp;
      float v = tmp * (1 / 32768.f);
      if (p.apply_gain) v = v * p.gain;
      out[j] = v;
    }
    float mlast = shfl(m, (N - 1) / seg);  // lane holding the last sample
    if (lane == 0) st->deemph_mem[ch] = mlast;
    wsync();
  }
  // ---- opus_decode()'s soft clipping of the int16 output; opus_decode_float() clears the memory (src/opus_decoder.c:875-878) ----
  if (!(plc & OB_PLC_PACKET)) {   // a lost packet's frames are returned before this step (src/opus_decoder.c:764-778)
    if (p.defer_clip) {}   // clipped per packet by dec_softclip_packet_block
    else if (p.pcm_is_int16) soft_clip_channel(out, Nd, ds, &st->declip[ch], lane);
    else if (lane == 0 && st->declip[ch] != 0.f) st->declip[ch] = 0.f;
  }
  // ---- interleave through shared memory, then fully coalesced stores ----
  bsync();
  if (ch == 0 && lane == 0) st->pos = pos + (uint32_t)N;
  {
    const size_t base = (size_t)stream * p.pcm_stride + (size_t)side->pcm_off * CC;
    const int total = Nd * CC;
    const int tid = ch * 32 + lane, nth = CC * 32;
    const SynthWS *w0 = (const SynthWS *)smem;
    if (p.pcm_is_int16) {
      int16_t *o = (int16_t *)p.pcm + base;
      for (int e = tid; e < total; e += nth) {
        int j = CC == 2 ? e >> 1 : e, c = CC == 2 ? e & 1 : 0;
        float x = w0[c].out[OB_SYN_O + j * ds] * 32768.f;
        x = x > -32768.f ? x : -32768.f;
        x = x < 32767.f ? x : 32767.f;
        o[e] = (int16_t)lrintf(x);
      }
    } else {
      float *o = (float *)p.pcm + base;
      for (int e = tid; e < total; e += nth) {
        int j = CC == 2 ? e >> 1 : e, c = CC == 2 ? e & 1 : 0;
        o[e] = w0[c].out[OB_SYN_O + j * ds];
      }
    }
  }
}

// opus_decode()'s soft clipping + int16 conversion over a whole packet (src/opus_decoder.c:875-878, src/opus.c:39-166), for packets that
// may hold several frames: the synthesis / concealment kernels left the packet's float PCM in `stage`.  One CTA per stream, one warp per
// channel for the (mostly skipped) clipping scan, then every thread converts and stores.
struct DecClipParams {
  float *stage;                 // [n_streams][stage_stride] interleaved float PCM of the packet
  int stage_stride;
  int16_t *pcm;                 // [n_streams][pcm_stride]
  int pcm_stride;
  const int32_t *out_samples;   // [n_streams] samples per channel the packet produced (<= 0: nothing)
  const int32_t *pkt_len;       // [n_streams]; 0 = lost packet: returned without soft clipping (src/opus_decoder.c:764-778)
  ObDecSynthState *state;
  int n_streams, channels;
};
OB_DEV void dec_softclip_packet_block(const DecClipParams &p, int stream) {
  const int lane = lane_id(), ch = warp_in_block(), CC = p.channels;
  const int n = p.out_samples[stream];
  if (n <= 0) return;
  float *x = p.stage + (size_t)stream * p.stage_stride;
  if (p.pkt_len[stream] != 0) soft_clip_channel(x + ch, n, CC, &p.state[stream].declip[ch], lane);
  bsync();
  int16_t *o = p.pcm + (size_t)stream * p.pcm_stride;
  for (int e = ch * 32 + lane; e < n * CC; e += CC * 32) {
    float v = x[e] * 32768.f;
    v = v > -32768.f ? v : -32768.f;
    v = v < 32767.f ? v : 32767.f;
    o[e] = (int16_t)lrintf(v);
  }
}

}  // namespace ob
