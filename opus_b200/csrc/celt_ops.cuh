// celt_ops.cuh — encoder-side CELT kernel operations as batched kernels (SURVEY §8 rows a1, a6,
// a9), each with the reference's exact float operation order so results are bit-identical to the
// reference's C functions (compile with -fmad=false):
//   ob_mdct_forward  : clt_mdct_forward_c via compute_mdcts (celt/mdct.c:122-264, celt_encoder.c:511-554)
//                      one warp per channel-frame; fold+window and pre-rotation fused, FFT in shared memory
//   ob_pitch_xcorr   : celt_pitch_xcorr_c (celt/pitch.c:230-305); one warp per item, a lane per lag,
//                      every lag accumulated sequentially like the reference
//   ob_exp_rotation  : exp_rotation (celt/vq.c:75-147), both directions; one thread per vector
//   ob_pvq_search    : op_pvq_search_c (celt/vq.c:205-374); one thread per vector, the 32 vectors of a
//                      warp transposed in shared memory ([element][lane], conflict-free because all
//                      lanes scan the same element index)
// These are the building blocks of the fused encoder (next round) and the B3-level parity surface.
#pragma once
#include "celt_fft.cuh"
#include "celt_mdct_fast.cuh"
#include "celt_vec.cuh"
#include "celt_enc_pre.cuh"
#include "opus_enc_layer.cuh"

namespace ob {

struct MdctFwdParams {
  const float *in;   // [n_items][in_stride]: B*N + overlap samples per item
  float *out;        // [n_items][out_stride]: B*N coefficients per item
  int n_items, in_stride, out_stride, LM, transient;
};
struct MdctWS { float in[OB_MAXN + OB_OVERLAP + 4]; float f2[2 * 15 * kF480PadSmall + 2]; };   // f2: 480 complex points: 960 floats for the shared-memory FFT, 990 in the
                                                                                              // padded layout (33 per group of 32) of the register FFT

// forward MDCT of the B*N+overlap samples already in ws->in; coefficients go to `o` (HBM, or ws->in itself: the time
// samples are dead once the pre-rotation has run)
// fast: the long block (LM = 3, one 960-point transform) takes the register FFT of celt_mdct_fast.cuh (tables from global memory); the arithmetic per
// butterfly is the same, the results are bit-identical (the general kernel ob_mdct_forward passes false and stays the A/B reference)
OB_DEV void mdct_forward_from_ws(MdctWS *ws, int LM, int transient, float *o, int lane, bool fast = true) {
  if (fast && LM == 3 && !transient) {
    const Mdct480Global tabs;
    cpx *F = (cpx *)ws->f2;
    mdct480_prerotate<kF480PadSmall>(ws->in, F, &tabs, lane);
    wsync();
    if (lane < 15) f480_group_stages<false>(F + lane * kF480PadSmall, &tabs);
    wsync();
    f480_column_stages_postrotate<kF480PadSmall>(F, &tabs, o, lane);
    return;
  }
  const int nblk = transient ? 1 << LM : 1, NB = transient ? OB_SHORT_N : OB_SHORT_N << LM;
  const int shift = transient ? 3 : 3 - LM;
  const int N2 = NB, N4 = NB >> 1, nfft = N4, overlap = OB_OVERLAP;
  const float *trig = ob_mdct_trig;
  for (int i = 0; i < shift; i++) trig += (1920 >> i) >> 1;
  const float scale = f_div(1.f, (float)nfft);
  const short *bitrev = bitrev_table(shift);
  cpx *f2 = (cpx *)ws->f2;
  const int q1 = (overlap + 3) >> 2;
  // window + fold + pre-rotation, one (block, i) pair per lane step (celt/mdct.c:157-226)
  for (int q = lane; q < nblk * N4; q += 32) {
    const int b = q / N4, i = q - b * N4;
    const float *in = ws->in + b * NB;
    const float *xp1 = in + (overlap >> 1) + 2 * i, *xp2 = in + N2 - 1 + (overlap >> 1) - 2 * i;
    float re, im;
    if (i < q1) {
      const float w1 = ob_window[(overlap >> 1) + 2 * i], w2 = ob_window[(overlap >> 1) - 1 - 2 * i];
      re = xp1[N2] * w2 + (*xp2) * w1;
      im = (*xp1) * w1 - xp2[-N2] * w2;
    } else if (i < N4 - q1) {
      re = *xp2;
      im = *xp1;
    } else {
      const int k = i - (N4 - q1);
      const float w1 = ob_window[2 * k], w2 = ob_window[overlap - 1 - 2 * k];
      re = -(xp1[-N2] * w1) + (*xp2) * w2;
      im = (*xp1) * w2 + xp2[N2] * w1;
    }
    const float t0 = trig[i], t1 = trig[N4 + i];
    const float yr = re * t0 - im * t1;
    const float yi = im * t0 + re * t1;
    cpx v;
    v.r = yr * scale;
    v.i = yi * scale;
    f2[b * nfft + bitrev[i]] = v;
  }
  fft_stages(f2, shift, nblk, lane);
  // post-rotation (celt/mdct.c:236-262); out index = b + nblk*k
  for (int q = lane; q < nblk * N4; q += 32) {
    const int b = q / N4, i = q - b * N4;
    const cpx fp = f2[b * nfft + i];
    const float t0 = trig[i], t1 = trig[N4 + i];
    o[b + nblk * (2 * i)] = fp.i * t1 - fp.r * t0;
    o[b + nblk * (N2 - 1 - 2 * i)] = fp.r * t1 + fp.i * t0;
  }
}
OB_DEV void mdct_forward_warp(const MdctFwdParams &p, int item, unsigned char *smem_warp) {
  const int lane = lane_id();
  MdctWS *ws = (MdctWS *)smem_warp;
  const int n_in = (OB_SHORT_N << p.LM) + OB_OVERLAP;
  const float *g = p.in + (size_t)item * p.in_stride;
  for (int i = lane; i < n_in; i += 32) ws->in[i] = g[i];
  wsync();
  mdct_forward_from_ws(ws, p.LM, p.transient, p.out + (size_t)item * p.out_stride, lane, false);
}

// ---- encoder front end, one warp per (stream, channel): pre-emphasis -> MDCT -> band energies -> log2 -> normalised bands
//      (celt_preemphasis celt/celt_encoder.c:557, compute_mdcts :511, compute_band_energies celt/bands.c:151,
//       amp2Log2 celt/quant_bands.c:553, normalise_bands celt/bands.c:169), the path of celt_encode_with_ec when the pitch
//      pre-filter is off.  The spectrum never leaves shared memory; HBM sees the PCM once and the normalised bands once. ----
struct ObEncAnalysisState { float preemph_mem[2]; float overlap_mem[2][OB_OVERLAP]; float overlap_max; int32_t pf_head; };   // pf_head: ring position of prefilter_mem
// what the frame encoder needs from the time domain (celt_encode_with_ec :1963-1970, :2017)
struct ObEncFrameInfo {
  float tone_freq, toneishness, sample_max, tf_estimate;
  int32_t short_blocks, is_transient, tf_chan, patched;   // patched: patch_transient_decision switched to short blocks (bandLogE2 differs)
  int32_t pf_on, pitch_index, qg, pf_tapset;             // run_prefilter's decision
  float gain1;
  int32_t pad[3];
};
struct EncAnalysisParams {
  int end = OB_NBANDS;         // CELT_SET_END_BAND
  int CC = 0;                  // input channels when they differ from the coded channels C (CELT_SET_CHANNELS: mono coding of a stereo input); 0 = C
  int upsample = 1;            // 48000 / input rate (celt_encoder_init, celt/celt_encoder.c:255): pcm holds N/upsample samples per channel
  ObEncAnalysisState *state;   // [n_streams]
  const float *pcm;            // [n_streams][N*C] interleaved, +-1.0 full scale
  const int32_t *short_blocks; // [n_streams] transient decision per stream, or null (= long blocks)
  float *X;                    // [n_streams][C*N]
  float *bandE, *bandLogE;     // [n_streams][C*21]
  ObEncFrameInfo *info;        // [n_streams] or null
  int n_streams, C, LM;
  int complexity, lsb_depth;   // transient_analysis runs from complexity 1 (celt/celt_encoder.c:2020); lsb_depth defines silence
  // pitch pre-filter / patch_transient_decision (complexity >= 5): they read the frame coder's state and the byte budget
  const ObEncState *enc_state = nullptr; // [n_streams] or null (transform only)
  float *prefilter_mem = nullptr;        // [n_streams][C][1024] unfiltered pre-emphasised history
  float *bandLogE2 = nullptr;            // [n_streams][C*21]: the long-block log energies when the frame was patched to short blocks
  int max_bytes = 0, bitrate = -1, vbr = 0, constrained_vbr = 0, disable_pf = 0, loss_rate = 0;
  int first_stream = 0;                  // this launch covers streams [first_stream, first_stream + gridDim.x)
  // Opus layer (opus_encode): per-stream coded channels / end band / AnalysisInfo, and the stereo width fade + guard applied to the
  // dc-rejected input as it is read (src/opus_encoder.c:2011-2022, 2329-2349).  Arrays are then strided by the INPUT channel count.
  const ObOpusFrameCtl *ctl = nullptr;
  // the transient / tone recurrences already ran in ob_celt_enc_chains (celt_enc_chains.cuh): p.info[stream] holds their results
  int chains_done = 0;
  // ... and so did run_prefilter: in_buf [n][CC][N + overlap (stride 1080)] holds the filtered window the transforms start from, every
  // time-domain memory is up to date (null: the analysis kernel pre-emphasises and filters itself)
  float *in_buf = nullptr;
};
// the encoder's input sample (i, c) of one stream: the caller's PCM, or with the Opus layer in front the dc-rejected signal after
// stereo_fade (src/opus_encoder.c:553-580) and the float guard
struct EncPcmView {
  const float *pcm;   // this stream's interleaved frame
  int CC, fade, clear, inc;
  float g1, g2;
};
OB_DEV float enc_pcm_at(const EncPcmView &v, int i, int c) {
  if (v.clear) return 0.f;
  const float x = v.pcm[(size_t)v.CC * i + c];
  if (!v.fade) return x;
  const float L = v.pcm[(size_t)2 * i], R = v.pcm[(size_t)2 * i + 1];
  float g = v.g2;
  if (i < OB_OVERLAP / v.inc) {
    float w = ob_window[i * v.inc];
    w = w * w;
    g = w * v.g2 + (1.f - w) * v.g1;
  }
  float diff = .5f * (L - R);
  diff = g * diff;
  return c == 0 ? L - diff : R + diff;
}
// tone_lpc (celt/celt_encoder.c:1305-1360) on x = in0 (+ in1): the six running sums are independent sequential chains,
// one lane each (lanes 0-2: r00 r01 r02, lanes 3-5: the three edge corrections); lane 0 solves the 2x2 system.
OB_DEV int tone_lpc_warp(const float *in0, const float *in1, int len, int delay, float *lpc, int lane) {
  float acc = 0.f;
  if (lane < 3) {
    const int d = lane * delay;
#pragma unroll 8
    for (int i = 0; i < len - 2 * delay; i++) {   // unrolled so that the shared-memory loads run ahead of the serial sum
      const float a = in1 ? in0[i] + in1[i] : in0[i], b = in1 ? in0[i + d] + in1[i + d] : in0[i + d];
      acc += a * b;
    }
  } else if (lane < 6) {
    const int o1 = lane == 4 ? len - delay : len - 2 * delay, o2 = lane == 5 ? len - delay : o1;
    const int p1 = lane == 4 ? delay : 0, p2 = lane == 5 ? delay : p1;
#pragma unroll 1
    for (int i = 0; i < delay; i++) {
      const float a = in1 ? in0[o1 + i] + in1[o1 + i] : in0[o1 + i], b = in1 ? in0[o2 + i] + in1[o2 + i] : in0[o2 + i];
      const float c = in1 ? in0[p1 + i] + in1[p1 + i] : in0[p1 + i], d = in1 ? in0[p2 + i] + in1[p2 + i] : in0[p2 + i];
      acc += a * b - c * d;
    }
  }
  float r00 = shfl(acc, 0), r01 = shfl(acc, 1), r02 = shfl(acc, 2);
  const float e11 = shfl(acc, 3), e22 = shfl(acc, 4), e12 = shfl(acc, 5);
  float r11 = r00 + e11;
  float r22 = r11 + e22;
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
// transient_analysis for one channel (celt/celt_encoder.c:267-466): three first-order recurrences over the frame, serial by
// nature (and by the float order the reference fixes) -> one lane; returns the channel's `unmask` metric.  tmp: len/2 floats.
OB_TABLE unsigned char ob_transient_inv_table[128] = {
    255, 255, 156, 110, 86, 70, 59, 51, 45, 40, 37, 33, 31, 28, 26, 25, 23, 22, 21, 20, 19, 18, 17, 16, 16, 15, 15, 14, 13, 13, 12, 12,
    12,  12,  11,  11,  11, 10, 10, 10, 9,  9,  9,  9,  9,  9,  8,  8,  8,  8,  8,  7,  7,  7,  7,  7,  7,  6,  6,  6,  6,  6,  6,  6,
    6,   6,   6,   6,   6,  6,  6,  6,  6,  5,  5,  5,  5,  5,  5,  5,  5,  5,  5,  5,  5,  4,  4,  4,  4,  4,  4,  4,  4,  4,  4,  4,
    4,   4,   4,   4,   4,  4,  4,  4,  4,  4,  4,  4,  4,  4,  3,  3,  3,  3,  3,  3,  3,  3,  3,  3,  3,  3,  3,  3,  3,  3,  3,  2,
};
OB_DEVNI int transient_unmask_warp(const float *__restrict__ in, float *__restrict__ tmp, int len, int lane) {
  const int len2 = len / 2;
  const float forward_decay = .0625f;
  float norm = 0.f;
  if (lane == 0) {
    // the high-pass filter (two samples per step) feeds the forward smoother directly: only len/2 values are stored
    float hp0 = 0.f, hp1 = 0.f, mean = 0.f, mem0 = 0.f;
#pragma unroll 4
    for (int i = 0; i < len2; i++) {
      float y[2];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const float x = in[2 * i + k], mem00 = hp0;
        y[k] = 2 * i + k < 12 ? 0.f : hp0 + x;
        hp0 = hp0 - x + .5f * hp1;
        hp1 = x - mem00;
      }
      const float x2 = y[0] * y[0] + y[1] * y[1];
      mean += x2;
      mem0 = x2 + (1.f - forward_decay) * mem0;
      tmp[i] = forward_decay * mem0;
    }
    mem0 = 0.f;
    float maxE = 0.f;
#pragma unroll 4
    for (int i = len2 - 1; i >= 0; i--) {
      mem0 = tmp[i] + 0.875f * mem0;
      tmp[i] = 0.125f * mem0;
      maxE = fmaxf(maxE, 0.125f * mem0);
    }
    mean = (float)sqrt((double)(mean * maxE) * .5 * len2);
    norm = len2 / (1e-15f + mean);
  }
  norm = shfl(norm, 0);
  wsync();
  // the table pass is an integer sum: any order gives the reference's value, so the warp shares it
  int unmask = 0;
  for (int i = 12 + 4 * lane; i < len2 - 5; i += 4 * 32) {
    const double v = floor((double)(64 * norm * (tmp[i] + 1e-15f)));
    const int id = (int)(v < 127 ? (v > 0 ? v : 0) : 127);
    unmask += ob_transient_inv_table[id];
  }
  for (int o = 16; o; o >>= 1) unmask += shfl_xor(unmask, o);
  return 64 * unmask * 4 / (6 * (len2 - 17));
}
// compute_mdcts with upsample != 1 (celt/celt_encoder.c:544-556): scale the part of the spectrum that exists at the input rate, clear the rest
OB_DEV void enc_mdct_upsample(float *freq, int N, int up, int lane) {
  if (up == 1) return;
  const int bound = N / up;
  const float g = (float)up;
  for (int j = lane; j < N; j += 32) freq[j] = j < bound ? f_mul(freq[j], g) : 0.f;
  wsync();
}
// compute_mdcts' tail (celt/celt_encoder.c:539-556) for the whole CTA: mono coding of a stereo input averages the two spectra into
// channel 0's, then the upsample fix.  f0 / f1: the two warps' spectra (f1 unused unless CC == 2 && C == 1).
OB_DEV void enc_mdct_tail(float *freq, float *f0, const float *f1, int N, int up, int C, int CC, int c, int lane) {
  if (CC == 2 && C == 1) {
    bsync();   // both spectra complete
    if (c == 0) {
      for (int j = lane; j < N; j += 32) f0[j] = f_add(f_mul(.5f, f0[j]), f_mul(.5f, f1[j]));
      wsync();
    }
  }
  if (c < C) enc_mdct_upsample(freq, N, up, lane);
}
// one CTA per stream, one warp per channel
OB_DEV void enc_analysis_block(const EncAnalysisParams &p, int stream, unsigned char *smem) {
  const int lane = lane_id(), c = warp_in_block();
  MdctWS *ws = (MdctWS *)smem + c;
  const int LM = p.LM, N = OB_SHORT_N << LM, M = 1 << LM, CC = p.CC ? p.CC : p.C;   // C coded channels (CELT_SET_CHANNELS), CC input channels: one warp per INPUT channel
  const int C = p.ctl ? p.ctl[stream].C : p.C, SC = p.ctl ? CC : p.C, end_band = p.ctl ? p.ctl[stream].end : p.end;   // SC: channel stride of the per-stream arrays
  int transient = p.short_blocks ? (p.short_blocks[stream] != 0) : 0;
  ObEncAnalysisState &st = p.state[stream];
  const int up = p.upsample;
  EncPcmView pv;
  pv.pcm = p.pcm + (size_t)stream * (N / up) * CC; pv.CC = CC; pv.fade = 0; pv.clear = 0; pv.inc = up; pv.g1 = pv.g2 = 0.f;
  if (p.ctl) { const ObOpusFrameCtl &k = p.ctl[stream]; pv.fade = k.fade; pv.clear = k.clear; pv.g1 = k.fade_g1; pv.g2 = k.fade_g2; }
  // pre-emphasis is a 2-tap FIR on the input: in[i] = x[i] - 0.85*x[i-1]  (m = coef0*x is carried, not the output)
  const float coef0 = 0.85000610f;
  float mx_new = 0.f, mx_tail = 0.f;   // celt_maxabs over the fresh part and over the part overlapping the next frame
  const bool pf_path = p.enc_state != nullptr && p.info != nullptr;   // the frame-encoder configuration (celt_enc_host.cuh)
  float *pmem = pf_path ? p.prefilter_mem + ((size_t)stream * CC + c) * kPfMaxPeriod : nullptr;
  // the analysis window starts with the UNFILTERED tail of the previous frame (celt/celt_encoder.c:2014); without a pre-filter
  // history that is the same as the MDCT overlap memory
  if (p.in_buf) {   // ob_celt_enc_chains left the pre-emphasised, pre-filtered window
    const float *src = p.in_buf + ((size_t)stream * CC + c) * (OB_MAXN + OB_OVERLAP);
    for (int i = lane; i < N + OB_OVERLAP; i += 32) ws->in[i] = src[i];
  } else {
  for (int i = lane; i < OB_OVERLAP; i += 32) ws->in[i] = pf_path ? pmem[(st.pf_head + kPfMaxPeriod - OB_OVERLAP + i) & (kPfMaxPeriod - 1)] : st.overlap_mem[c][i];
  if (up == 1) {
    for (int i = lane; i < N; i += 32) {
      const float s = enc_pcm_at(pv, i, c);
      const float x = f_mul(32768.f, s);
      const float m = i == 0 ? st.preemph_mem[c] : f_mul(coef0, f_mul(32768.f, enc_pcm_at(pv, i - 1, c)));
      ws->in[OB_OVERLAP + i] = f_add(x, -m);
      if (i < N - OB_OVERLAP) mx_new = fmaxf(mx_new, fabsf(s));
      else mx_tail = fmaxf(mx_tail, fabsf(s));
    }
  } else {
    // input below 48 kHz: the samples are zero-stuffed to 48 kHz ahead of the pre-emphasis (celt/celt_encoder.c:583-589)
    for (int i = lane; i < N; i += 32) {
      const bool on = i % up == 0, prev_on = (i - 1) % up == 0;
      const float s = on ? enc_pcm_at(pv, i / up, c) : 0.f;
      const float x = on ? f_mul(32768.f, s) : 0.f;
      const float m = i == 0 ? st.preemph_mem[c] : f_mul(coef0, prev_on ? f_mul(32768.f, enc_pcm_at(pv, (i - 1) / up, c)) : 0.f);
      ws->in[OB_OVERLAP + i] = f_add(x, -m);
      if (on) {
        if (i < N - OB_OVERLAP) mx_new = fmaxf(mx_new, fabsf(s));
        else mx_tail = fmaxf(mx_tail, fabsf(s));
      }
    }
  }
  }
  mx_new = warp_max(mx_new);
  mx_tail = warp_max(mx_tail);
  wsync();
  if (lane == 0 && !p.in_buf) st.preemph_mem[c] = up == 1 ? f_mul(coef0, f_mul(32768.f, enc_pcm_at(pv, N - 1, c))) : f_mul(coef0, 0.f);
  if (!pf_path) for (int i = lane; i < OB_OVERLAP; i += 32) st.overlap_mem[c][i] = ws->in[N + i];
  // cross-channel part: maxima and the tone detector need both channels' time signals
  float *xch = (float *)((MdctWS *)smem + CC);  // [2][2] maxima exchange, [4..5] transient metric per channel, [6] decision
  if (lane == 0) { xch[c * 2] = mx_new; xch[c * 2 + 1] = mx_tail; }
  if (p.chains_done) {
    if (c == 0 && lane == 0) ((int *)xch)[6] = p.info[stream].short_blocks;
  } else if (p.complexity >= 1 && p.info) {   // transient_analysis: lane 0 of each channel's warp walks its channel (f2 is free until the MDCT)
    const int um = transient_unmask_warp(ws->in, ws->f2, N + OB_OVERLAP, lane);
    if (lane == 0) ((int *)xch)[4 + c] = um;
  }
  bsync();
  if (c == 0 && p.info && !p.chains_done) {
    float a = xch[0], t = xch[1];
    if (CC == 2) { a = fmaxf(a, xch[2]); t = fmaxf(t, xch[3]); }
    const float sample_max = fmaxf(fmaxf(st.overlap_max, a), t);
    const float *in0 = ((MdctWS *)smem)->in, *in1 = CC == 2 ? ((MdctWS *)smem + 1)->in : nullptr;
    // tone_detect (celt/celt_encoder.c:1362-1402)
    float lpc[2] = {0.f, 0.f};
    int delay = 1;
    int fail = tone_lpc_warp(in0, in1, N + OB_OVERLAP, delay, lpc, lane);
    while (delay <= 48000 / 3000 && (fail || (lpc[0] > 1.f && lpc[1] < 0))) {
      delay *= 2;
      fail = tone_lpc_warp(in0, in1, N + OB_OVERLAP, delay, lpc, lane);
    }
    if (lane == 0) {
      ObEncFrameInfo fi;
      if (!fail && lpc[0] * lpc[0] + 3.999999f * lpc[1] < 0) {
        fi.toneishness = -lpc[1];
        fi.tone_freq = (float)(acos((double)(.5f * lpc[0])) / delay);
      } else {
        fi.tone_freq = -1.f;
        fi.toneishness = 0.f;
      }
      fi.sample_max = sample_max;
      fi.tf_estimate = 0.f; fi.tf_chan = 0; fi.is_transient = transient; fi.patched = 0; fi.pf_on = 0; fi.pitch_index = 15; fi.qg = 0; fi.pf_tapset = 0; fi.gain1 = 0.f;
      fi.pad[0] = fi.pad[1] = fi.pad[2] = 0;
      if (p.complexity >= 1) {   // the cross-channel tail of transient_analysis (celt/celt_encoder.c:437-461)
        int mask_metric = ((int *)xch)[4];
        if (CC == 2 && ((int *)xch)[5] > mask_metric) { mask_metric = ((int *)xch)[5]; fi.tf_chan = 1; }
        int is_tr = mask_metric > 200;
        if (fi.toneishness > .98f && fi.tone_freq < 0.026f) { is_tr = 0; mask_metric = 0; }
        const float tf_max = fmaxf(0.f, (float)sqrt((double)(27 * mask_metric)) - 42.f);
        const double a = (double)((float)0.0069 * fminf(163.f, tf_max)) - 0.139;
        fi.tf_estimate = (float)sqrt(a > 0 ? a : 0.0);
        fi.is_transient = is_tr;
      }
      if (CC == 2 && C == 1) fi.tf_chan = 0;   // celt/celt_encoder.c:2095-2096
      fi.toneishness = fminf(fi.toneishness, 1.f - fi.tf_estimate);   // celt/celt_encoder.c:2026, before a patched frame resets tf_estimate
      // short blocks need LM > 0 and 3 bits of room after the two header flags: always there unless the frame is silent
      // (celt/celt_encoder.c:2060-2067; the frame coder re-checks with the real range-coder position)
      const int silence = sample_max <= 1.f / (float)(1 << p.lsb_depth);
      fi.short_blocks = fi.is_transient && LM > 0 && !silence;
      p.info[stream] = fi;
      st.overlap_max = t;
      ((int *)xch)[6] = fi.short_blocks;
    }
  }
  bsync();   // the tone detector reads ws->in of both channels; the MDCT below overwrites it
  if (p.info) transient = ((int *)xch)[6];
  float *freq = ws->in;
  PfWS *pw = (PfWS *)((unsigned char *)smem + (size_t)CC * sizeof(MdctWS) + 32);
  if (pf_path && p.in_buf) {
    if (p.complexity >= 5) freq = pw->buf + c * OB_MAXN;   // the window stays for patch_transient_decision's second transform
    bsync();
  } else if (pf_path) {   // run_prefilter (celt/celt_encoder.c:2029-2057): every thread of the CTA takes part
    const ObEncState &es = p.enc_state[stream];
    ObEncFrameInfo fi = p.info[stream];
    const EncBudget bud = enc_budget(p.max_bytes, p.bitrate, p.vbr, p.constrained_vbr, N, es.vbr_reservoir);
    const int silence = fi.sample_max <= 1.f / (float)(1 << p.lsb_depth);
    PfArgs a;
    for (int cc = 0; cc < CC; cc++) {
      a.in[cc] = ((MdctWS *)smem + cc)->in;
      a.fr[cc] = ((MdctWS *)smem + cc)->f2;
      a.ring[cc] = p.prefilter_mem + ((size_t)stream * CC + cc) * kPfMaxPeriod;
      a.in_mem[cc] = st.overlap_mem[cc];
    }
    a.head = st.pf_head;
    a.C = CC; a.N = N; a.complexity = p.complexity; a.nbAvailableBytes = bud.nbAvailableBytes; a.prefilter_tapset = es.tapset_decision;
    a.prev_period = es.prefilter_period; a.prev_tapset = es.prefilter_tapset; a.prev_gain = es.prefilter_gain; a.loss_rate = p.loss_rate;
    a.tf_estimate = fi.tf_estimate; a.tone_freq = fi.tone_freq; a.toneishness = fminf(fi.toneishness, 1.f - fi.tf_estimate);
    a.enabled = bud.nbAvailableBytes > 12 * C && !silence && bud.nbCompressedBytes * 8 >= 17 && !p.disable_pf;
    if (p.ctl) { a.an_valid = p.ctl[stream].analysis.valid; a.an_max_pitch_ratio = p.ctl[stream].analysis.max_pitch_ratio; }
    const PfDecision d = enc_prefilter_stage(pw, a);
    if (c == 0 && lane == 0) {
      fi.pf_on = d.pf_on; fi.pitch_index = d.pitch_index; fi.qg = d.qg; fi.pf_tapset = d.tapset; fi.gain1 = d.gain1; fi.patched = 0;
      p.info[stream] = fi;
      st.pf_head = (a.head + N) & (kPfMaxPeriod - 1);
    }
    // from complexity 5 the spectrum goes to the (now dead) search buffers and the time signal stays:
    // patch_transient_decision may ask for a second transform
    if (p.complexity >= 5) freq = pw->buf + c * OB_MAXN;
    bsync();
  }
  // both warps' spectrum buffers (for the mono downmix)
  float *freq0 = (pf_path && p.complexity >= 5) ? pw->buf : ((MdctWS *)smem)->in;
  const float *freq1 = (pf_path && p.complexity >= 5) ? pw->buf + OB_MAXN : ((MdctWS *)smem + (CC - 1))->in;
  if (pf_path && p.complexity >= 8 && transient) {
    // secondMdct (celt/celt_encoder.c:2077-2090): the long-block log energies of a transient frame, for dynalloc_analysis
    mdct_forward_from_ws(ws, LM, 0, freq, lane);
    wsync();
    enc_mdct_tail(freq, freq0, freq1, N, up, C, CC, c, lane);
    if (c < C && lane < OB_NBANDS) {
      float sum = 0.f;
      const int j0 = M * ob_ebands[lane], j1 = M * ob_ebands[lane + 1];
#pragma unroll 4
      for (int j = j0; j < j1; j++) sum = f_add(sum, f_mul(freq[j], freq[j]));
      const float E0 = f_sqrt(f_add(1e-27f, sum));
      // (the frame kernel adds the LM/2 compensation, as for a patched frame)
      p.bandLogE2[(size_t)stream * SC * OB_NBANDS + c * OB_NBANDS + lane] = f_add((float)(1.442695040888963387 * log((double)E0)), -ob_emeans[lane]);
    }
    wsync();
  }
  if (CC == 2 && C == 1) bsync();   // (the second-MDCT energies above read channel 0's spectrum buffer)
  mdct_forward_from_ws(ws, LM, transient, freq, lane);
  wsync();
  enc_mdct_tail(freq, freq0, freq1, N, up, C, CC, c, lane);
  if (pf_path && p.complexity >= 5 && LM > 0 && !transient) {
    // patch_transient_decision (celt/celt_encoder.c:2238-2257) needs both channels' log energies: first pass of the band sums
    float E0 = 0.f;
    if (c < C && lane < OB_NBANDS) {
      float sum = 0.f;
      const int j0 = M * ob_ebands[lane], j1 = M * ob_ebands[lane + 1];
#pragma unroll 4
      for (int j = j0; j < j1; j++) sum = f_add(sum, f_mul(freq[j], freq[j]));
      E0 = f_sqrt(f_add(1e-27f, sum));
      pw->sc[c * OB_NBANDS + lane] = f_add((float)(1.442695040888963387 * log((double)E0)), -ob_emeans[lane]);
    }
    bsync();
    if (c == 0 && lane == 0) {
      const ObEncFrameInfo fi = p.info[stream];
      const int silence = fi.sample_max <= 1.f / (float)(1 << p.lsb_depth);
      int patch = 0;
      if (!silence && !fi.is_transient) patch = enc_patch_transient(pw->sc, p.enc_state[stream].oldBandE, C, end_band);
      ((int *)xch)[6] = patch;
    }
    bsync();
    if (((int *)xch)[6]) {
      if (c < C && lane < OB_NBANDS && p.bandLogE2) p.bandLogE2[(size_t)stream * SC * OB_NBANDS + c * OB_NBANDS + lane] = pw->sc[c * OB_NBANDS + lane];   // the frame kernel adds the LM/2 compensation
      if (c == 0 && lane == 0) {
        ObEncFrameInfo fi = p.info[stream];
        fi.is_transient = 1; fi.short_blocks = 1; fi.patched = 1; fi.tf_estimate = .2f;
        p.info[stream] = fi;
      }
      transient = 1;
      mdct_forward_from_ws(ws, LM, 1, freq, lane);
      wsync();
      enc_mdct_tail(freq, freq0, freq1, N, up, C, CC, c, lane);
    }
  }
  if (c >= C) return;   // the second input channel of a mono-coded stream has nothing left to do (no barrier below)
  // band energies: one lane per band, summed in index order like celt_inner_prod_c
  float E = 0.f, g = 0.f;
  if (lane < OB_NBANDS) {
    float sum = 0.f;
    const int j0 = M * ob_ebands[lane], j1 = M * ob_ebands[lane + 1];
#pragma unroll 4
    for (int j = j0; j < j1; j++) sum = f_add(sum, f_mul(freq[j], freq[j]));
    E = f_sqrt(f_add(1e-27f, sum));
    g = f_div(1.f, f_add(1e-27f, E));
    const size_t o = (size_t)stream * SC * OB_NBANDS + c * OB_NBANDS + lane;
    p.bandE[o] = E;
    p.bandLogE[o] = f_add((float)(1.442695040888963387 * log((double)E)), -ob_emeans[lane]);
  }
  float *X = p.X + (size_t)stream * SC * N + (size_t)c * N;
#pragma unroll 1
  for (int b = 0; b < OB_NBANDS; b++) {
    const float gb = shfl(g, b);
    const int j1 = M * ob_ebands[b + 1];
    for (int j = M * ob_ebands[b] + lane; j < j1; j += 32) X[j] = f_mul(freq[j], gb);
  }
}

struct XcorrParams {
  const float *x;  // [n_items][len]
  const float *y;  // [n_items][len + max_pitch]
  float *xcorr;    // [n_items][max_pitch]
  int n_items, len, max_pitch;
};
OB_DEV void pitch_xcorr_warp(const XcorrParams &p, int item, unsigned char *smem_warp) {
  const int lane = lane_id();
  float *sx = (float *)smem_warp, *sy = sx + ((p.len + 3) & ~3);
  const float *gx = p.x + (size_t)item * p.len, *gy = p.y + (size_t)item * (p.len + p.max_pitch);
  for (int i = lane; i < p.len; i += 32) sx[i] = gx[i];
  for (int i = lane; i < p.len + p.max_pitch; i += 32) sy[i] = gy[i];
  wsync();
  float *o = p.xcorr + (size_t)item * p.max_pitch;
  for (int lag = lane; lag < p.max_pitch; lag += 32) {
    float sum = 0.f;
#pragma unroll 4
    for (int j = 0; j < p.len; j++) sum = sum + sx[j] * sy[lag + j];  // sequential, like xcorr_kernel_c
    o[lag] = sum;
  }
}

// ---- thread-per-vector operations on a transposed shared-memory tile: element j of lane l at
// tile[j * 32 + l] ----
struct VqParams {
  float *X;        // [n_items][N] in/out
  int *iy;         // [n_items][N] out (search)
  float *yy;       // [n_items] out (search)
  const int *K;    // [n_items]
  int n_items, N, dir, stride, spread;
};
#define OB_T(tile, j) (tile)[(j) * 32 + lane]

OB_DEV void rotate_seq_t(float *tile, int lane, int base, int len, int stride, float c, float s) {
  const float ms = -s;
  for (int i = 0; i < len - stride; i++) {
    float x1 = OB_T(tile, base + i), x2 = OB_T(tile, base + i + stride);
    OB_T(tile, base + i + stride) = c * x2 + s * x1;
    OB_T(tile, base + i) = c * x1 + ms * x2;
  }
  for (int i = len - 2 * stride - 1; i >= 0; i--) {
    float x1 = OB_T(tile, base + i), x2 = OB_T(tile, base + i + stride);
    OB_T(tile, base + i + stride) = c * x2 + s * x1;
    OB_T(tile, base + i) = c * x1 + ms * x2;
  }
}
OB_DEV void exp_rotation_t(float *tile, int lane, int len, int dir, int stride, int K, int spread) {
  if (2 * K >= len || spread == 0) return;
  const int factor = spread == 1 ? 15 : spread == 2 ? 10 : 5;
  float gain = f_div((float)len, (float)(len + factor * K));
  float theta = .5f * (gain * gain);
  float c = cos_norm_ref(theta), s = cos_norm_ref(1.0f - theta);
  int stride2 = 0;
  if (len >= 8 * stride) {
    stride2 = 1;
    while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++;
  }
  len = (unsigned)len / (unsigned)stride;
  for (int i = 0; i < stride; i++) {
    if (dir < 0) {
      if (stride2) rotate_seq_t(tile, lane, i * len, len, stride2, s, c);
      rotate_seq_t(tile, lane, i * len, len, 1, c, s);
    } else {
      rotate_seq_t(tile, lane, i * len, len, 1, c, -s);
      if (stride2) rotate_seq_t(tile, lane, i * len, len, stride2, s, -c);
    }
  }
}

// load / store a warp's 32 vectors between HBM ([item][N]) and the transposed tile, coalesced
OB_DEV void tile_load(float *tile, const float *g, int first_item, int n_items, int N, int lane) {
  for (int v = 0; v < 32; v++) {
    const int item = first_item + v;
    if (item >= n_items) break;
    for (int j = lane; j < N; j += 32) tile[j * 32 + v] = g[(size_t)item * N + j];
  }
  wsync();
}
OB_DEV void tile_store(const float *tile, float *g, int first_item, int n_items, int N, int lane) {
  wsync();
  for (int v = 0; v < 32; v++) {
    const int item = first_item + v;
    if (item >= n_items) break;
    for (int j = lane; j < N; j += 32) g[(size_t)item * N + j] = tile[j * 32 + v];
  }
}

OB_DEV void exp_rotation_warp(const VqParams &p, int warp, unsigned char *smem_warp) {
  const int lane = lane_id();
  float *tile = (float *)smem_warp;
  const int first = warp * 32, item = first + lane;
  tile_load(tile, p.X, first, p.n_items, p.N, lane);
  if (item < p.n_items) exp_rotation_t(tile, lane, p.N, p.dir, p.stride, p.K[item], p.spread);
  tile_store(tile, p.X, first, p.n_items, p.N, lane);
}

// op_pvq_search_c, one thread per vector; tiles: X (|X| on return), y (2*pulses as float), sign bytes
OB_DEV void pvq_search_warp(const VqParams &p, int warp, unsigned char *smem_warp) {
  const int lane = lane_id();
  const int N = p.N;
  float *tx = (float *)smem_warp, *ty = tx + N * 32;
  unsigned char *ts = (unsigned char *)(ty + N * 32);
  const int first = warp * 32, item = first + lane;
  tile_load(tx, p.X, first, p.n_items, N, lane);
  if (item < p.n_items) {
    const int K = p.K[item];
    float sum = 0.f, xy = 0.f, yy = 0.f;
    for (int j = 0; j < N; j++) {
      float x = OB_T(tx, j);
      OB_T(ts, j) = x < 0;
      OB_T(tx, j) = fabsf(x);
      OB_T(ty, j) = 0.f;
    }
    int pulsesLeft = K;
    if (K > (N >> 1)) {
      for (int j = 0; j < N; j++) sum += OB_T(tx, j);
      if (!(sum > 1e-15f && sum < 64.f)) {
        OB_T(tx, 0) = 1.f;
        for (int j = 1; j < N; j++) OB_T(tx, j) = 0.f;
        sum = 1.f;
      }
      const float rcp = (K + 0.8f) * f_div(1.f, sum);
      for (int j = 0; j < N; j++) {
        const float x = OB_T(tx, j);
        const int iyj = (int)floorf(rcp * x);
        float yj = (float)iyj;
        yy = yy + yj * yj;
        xy = xy + x * yj;
        OB_T(ty, j) = yj * 2;
        pulsesLeft -= iyj;
      }
    }
    if (pulsesLeft > N + 3) {
      float tmp = (float)pulsesLeft;
      yy = yy + tmp * tmp;
      yy = yy + tmp * OB_T(ty, 0);
      OB_T(ty, 0) += 2.f * tmp;  // iy[0] += pulsesLeft (y holds 2*iy)
      pulsesLeft = 0;
    }
    for (int i = 0; i < pulsesLeft; i++) {
      int best_id = 0;
      yy = yy + 1;
      float Rxy = xy + OB_T(tx, 0), Ryy = yy + OB_T(ty, 0);
      Rxy = Rxy * Rxy;
      float best_den = Ryy, best_num = Rxy;
      for (int j = 1; j < N; j++) {
        Rxy = xy + OB_T(tx, j);
        Ryy = yy + OB_T(ty, j);
        Rxy = Rxy * Rxy;
        if (best_den * Rxy > Ryy * best_num) { best_den = Ryy; best_num = Rxy; best_id = j; }
      }
      xy = xy + OB_T(tx, best_id);
      yy = yy + OB_T(ty, best_id);
      OB_T(ty, best_id) += 2.f;
    }
    p.yy[item] = yy;
  }
  wsync();
  // iy out (coalesced), X out
  for (int v = 0; v < 32; v++) {
    const int it = first + v;
    if (it >= p.n_items) break;
    for (int j = lane; j < N; j += 32) {
      int m = (int)(ty[j * 32 + v] * .5f);
      p.iy[(size_t)it * N + j] = ts[j * 32 + v] ? -m : m;
    }
  }
  tile_store(tx, p.X, first, p.n_items, N, lane);
}

}  // namespace ob
